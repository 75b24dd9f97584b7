"""Times the tcgen05 GEMM on the PPO update shapes (minibatch 131072) with CUDA events, and beside every
shape the library path on the same operands: cuBLAS bf16 GEMM through torch (F.linear / matmul) followed by
the separate element-wise kernels a non-fused implementation needs (bias + tanh, tanh', fp32 accumulate).
usage: python profiles/gemm_bench.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.engine import (EPI_ATOMIC_F32, EPI_BIAS_TANH_BF16, EPI_DTANH_BF16,  # noqa: E402
                                               gemm_bf16)
from rsoccer_isaac_cleanrl_b200.tc_mlp import _splits  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
M = 131072
dev = "cuda"


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


rows = []
for (K, N) in [(64, 256), (256, 512), (512, 512), (512, 256)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: gemm_bf16(a, w, out, EPI_BIAS_TANH_BF16, bias=b))
    rows.append(("fwd  bias+tanh", M, N, K, t))
for (K, N) in [(256, 512), (512, 512), (512, 256)]:   # dgrad: K = n_out, N = k_in
    dz = torch.randn(M, K, device=dev).to(torch.bfloat16)
    wt = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    y = torch.tanh(torch.randn(M, N, device=dev)).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: gemm_bf16(dz, wt, out, EPI_DTANH_BF16, aux=y))
    rows.append(("dgrad tanh'   ", M, N, K, t))
for (n_out, k_in) in [(256, 64), (512, 256), (512, 512), (256, 512)]:   # wgrad: M_gemm = n_out, N_gemm = k_in, K = batch
    dz = torch.randn(M, n_out, device=dev).to(torch.bfloat16)
    x = torch.randn(M, k_in, device=dev).to(torch.bfloat16)
    dw = torch.zeros(n_out, k_in, device=dev)
    splits = _splits(n_out, k_in, M)   # the PPO loop's own choice (tc_mlp.py)
    t = timeit(lambda: gemm_bf16(dz, x, dw, EPI_ATOMIC_F32, splits=splits, mn_major=True))
    rows.append(("wgrad split-K ", n_out, k_in, M, t))
# ---- the library path on the same shapes: cuBLAS bf16 + separate element-wise kernels
import torch.nn.functional as F  # noqa: E402

lib = {}
for (K, N) in [(64, 256), (256, 512), (512, 512), (512, 256)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.zeros(N, device=dev, dtype=torch.bfloat16)
    lib[("fwd  bias+tanh", M, N, K)] = (timeit(lambda: torch.tanh(F.linear(a, w, b))), timeit(lambda: F.linear(a, w)))
for (K, N) in [(256, 512), (512, 512), (512, 256)]:
    dz = torch.randn(M, K, device=dev).to(torch.bfloat16)
    wt = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    y = torch.tanh(torch.randn(M, N, device=dev)).to(torch.bfloat16)
    lib[("dgrad tanh'   ", M, N, K)] = (timeit(lambda: F.linear(dz, wt) * (1 - y * y)), timeit(lambda: F.linear(dz, wt)))
for (n_out, k_in) in [(256, 64), (512, 256), (512, 512), (256, 512)]:
    dz = torch.randn(M, n_out, device=dev).to(torch.bfloat16)
    x = torch.randn(M, k_in, device=dev).to(torch.bfloat16)
    lib[("wgrad split-K ", n_out, k_in, M)] = (timeit(lambda: (dz.t() @ x).float()), timeit(lambda: dz.t() @ x))
tot_f = tot_t = 0.0
for name, m, n, k, t in rows:
    fl = 2.0 * m * n * k
    tot_f += fl; tot_t += t
    lt, lg = lib[(name, m, n, k)]
    print(f"{name} M={m:7d} N={n:4d} K={k:7d}  {t * 1e6:8.1f} us  {fl / t / 1e12:7.1f} TFLOP/s   | cuBLAS bf16 + "
          f"elementwise {lt * 1e6:8.1f} us ({fl / lt / 1e12:6.1f}), GEMM alone {lg * 1e6:8.1f} us ({fl / lg / 1e12:6.1f})")
lib_t = sum(v[0] for v in lib.values())
print(f"one MLP fwd+bwd (hidden layers): {tot_t * 1e3:.3f} ms, {tot_f / tot_t / 1e12:.1f} TFLOP/s   | library path "
      f"{lib_t * 1e3:.3f} ms, {tot_f / lib_t / 1e12:.1f} TFLOP/s")
