"""The PPO gradient all-reduce over NVLink peer memory (csrc/peer_reduce.cu, include/vss_b200.h).

`PeerGradients` owns one rank's share of a `vss_peer` group: a device buffer that holds the rank's flat
gradient and that the other ranks of the node map through CUDA IPC. `allreduce(out)` launches ONE kernel
that synchronises with the peers through flag words in those buffers and writes the rank-ordered sum of all
gradients into `out` — no host involvement, capturable in a CUDA graph. One node only (CUDA IPC); the
caller falls back to NCCL when the group cannot be formed.
"""
import ctypes as C

import torch

from . import _lib


class _DeviceMemory:
    """A raw device pointer as a `__cuda_array_interface__` provider (float32 vector)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _check(lib, rc):
    if rc != 0:
        raise RuntimeError(f"libvss_b200 peer error {rc}: {lib.vss_peer_last_error().decode()}")


class PeerGradients:
    def __init__(self, num_floats, device, rank, world):
        import torch.distributed as dist
        self.lib = _lib.load_library()
        self.device = torch.device(device)
        self.rank, self.world = int(rank), int(world)
        self._h = C.c_void_p()
        # Every step that can fail locally is followed by an exchange of the outcome, so that either all ranks go on
        # or all ranks raise (a rank that left early would leave the others waiting in the next collective).
        err, handle = None, C.create_string_buffer(64)
        try:
            _check(self.lib, self.lib.vss_peer_create(C.byref(self._h), self.device.index, self.rank, self.world,
                                                      int(num_floats)))
            _check(self.lib, self.lib.vss_peer_ipc_handle(self._h, handle))
        except Exception as e:  # noqa: BLE001
            err = f"rank {self.rank}: {e}"
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (err, handle.raw))
        self._raise_if_any([g[0] for g in gathered])
        try:
            _check(self.lib, self.lib.vss_peer_connect(self._h, b"".join(g[1] for g in gathered)))
        except Exception as e:  # noqa: BLE001
            err = f"rank {self.rank}: {e}"
        dist.all_gather_object(gathered, err)   # (also: every rank has mapped every buffer before anyone signals into them)
        self._raise_if_any(gathered)
        n = int(self.lib.vss_peer_num_floats(self._h))
        self._mem = _DeviceMemory(self.lib.vss_peer_buffer(self._h), n)
        self.buffer = torch.as_tensor(self._mem, device=self.device)   # the rank's gradient buffer (n floats, zeroed)

    def _raise_if_any(self, errors):
        bad = [e for e in errors if e]
        if bad:
            self.close()
            raise RuntimeError("peer-memory group could not be formed: " + "; ".join(bad))

    def allreduce(self, out):
        """out[i] = sum over ranks of buffer_rank[i] (all ranks get bit-identical sums). `out`: f32 CUDA tensor of at
        least `buffer.numel()` elements, 16-byte aligned. Collective: every rank calls it the same number of times."""
        if out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous() or out.numel() < self.buffer.numel():
            raise ValueError("PeerGradients.allreduce: out must be a contiguous f32 CUDA tensor as large as the buffer")
        _check(self.lib, self.lib.vss_peer_allreduce(self._h, out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.buffer = None
            self.lib.vss_peer_destroy(self._h)
            self._h = C.c_void_p()
