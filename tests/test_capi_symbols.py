"""CPU: the C-ABI library loads and exports every symbol include/vss_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_in_header():
    src = open(os.path.join(ROOT, "include", "vss_b200.h")).read()
    return sorted(set(re.findall(r"VSS_API[^;(]*?\b(vss_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from rsoccer_isaac_cleanrl_b200 import _lib
    assert _declared_in_header() == _lib.declared_symbols()


def test_library_exports_every_symbol():
    from rsoccer_isaac_cleanrl_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = C.CDLL(_lib.LIB_PATH)
    for name in _declared_in_header():
        assert hasattr(lib, name), name
    assert b"sm_100a" in _lib.load_library().vss_version()


def test_params_struct_matches_oracle_restatement():
    """vss_default_params (product) against the oracle's independent restatement of the constants."""
    from rsoccer_isaac_cleanrl_b200 import _lib
    from oracle import vss_oracle as orc
    a, b = _lib.default_params(), orc.default_params()
    assert [f for f, _ in a._fields_] == [f for f, _ in b._fields_]
    for f, _ in a._fields_:
        assert getattr(a, f) == pytest.approx(getattr(b, f), rel=1e-6), f


def test_philox_host_hook_known_answers():
    import numpy as np
    from rsoccer_isaac_cleanrl_b200 import _lib
    lib = _lib.load_library()
    ctr = np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], np.uint32)
    key = np.array([0xA4093822, 0x299F31D0], np.uint32)
    out = np.zeros(4, np.uint32)
    lib.vss_philox4x32_10(ctr.ctypes.data, key.ctypes.data, out.ctypes.data)
    assert [int(x) for x in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rsoccer_isaac_cleanrl_b200 import _lib
    lib = _lib.load_library()
    h = C.c_void_p()
    p = _lib.default_params()
    rc = lib.vss_create(C.byref(h), C.byref(p), 16, 0, 0, 1)
    assert rc == -3 and b"no CPU fallback" in lib.vss_last_error()
    import rsoccer_isaac_cleanrl_b200 as R
    with pytest.raises(RuntimeError):
        R.Engine(16, "cuda:0")
    with pytest.raises(RuntimeError):
        R.Engine(16, "cpu")


def test_argument_validation_without_a_device():
    """Bad arguments are rejected before any CUDA call: empty batches, null pointers, bad views."""
    from rsoccer_isaac_cleanrl_b200 import _lib
    lib = _lib.load_library()
    p = _lib.default_params()
    h = C.c_void_p()
    assert lib.vss_create(C.byref(h), C.byref(p), 0, 0, 0, 1) == -1 and b"num_envs" in lib.vss_last_error()
    assert lib.vss_create(None, C.byref(p), 16, 0, 0, 1) == -1
    p.substeps = 1000
    assert lib.vss_create(C.byref(h), C.byref(p), 16, 0, 0, 1) == -1 and b"substeps" in lib.vss_last_error()
    assert lib.vss_step(None, None, None, None, None, None, None, None, None) == -1
    assert lib.vss_gae(None, None, None, None, None, None, None, 128, 16, 0.99, 0.95, None) == -1
    assert lib.vss_gemm_bf16_tn(None, 8, None, 8, None, 8, 128, 128, 64, 0, None, None, 0, 1, 0, None) == -1
    assert lib.vss_destroy(None) == 0
    assert lib.vss_num_envs(None) == 0
