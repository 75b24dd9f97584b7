"""Split-K sweep of the wgrad GEMMs (dW = dZ^T h, split-K with fp32 atomics) at minibatch 131072.
usage: python profiles/wgrad_splits_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.engine import EPI_ATOMIC_F32, gemm_bf16  # noqa: E402

M, dev, iters = 131072, "cuda", 30


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
    return e0.elapsed_time(e1) / iters * 1e3


for (n_out, k_in) in [(256, 64), (512, 256), (512, 512), (256, 512)]:
    dz = torch.randn(M, n_out, device=dev).to(torch.bfloat16)
    x = torch.randn(M, k_in, device=dev).to(torch.bfloat16)
    dw = torch.zeros(n_out, k_in, device=dev)
    tiles = (n_out // 128) * max(1, k_in // (128 if k_in % 128 == 0 else 64))
    base = max(1, 296 // tiles)
    res = []
    for s in sorted({base, max(1, base // 2), max(1, base // 3), max(1, base // 4), max(1, (base * 3) // 4), max(1, base // 6), max(1, base // 8)}):
        res.append("%d:%.1f" % (s, timeit(lambda: gemm_bf16(dz, x, dw, EPI_ATOMIC_F32, splits=s, mn_major=True))))
    print(f"wgrad {n_out}x{k_in} (tiles {tiles}, default splits {base}) us by splits:", " ".join(res), flush=True)
