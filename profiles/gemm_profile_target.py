"""Profiling target: every tcgen05 GEMM shape of one PPO minibatch (131072 rows), launched twice.
   ncu --set full -k regex:k_gemm -o out python profiles/gemm_profile_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.engine import EPI_ATOMIC_F32, EPI_BIAS_TANH_BF16, EPI_DTANH_BF16, gemm_bf16  # noqa: E402
from rsoccer_isaac_cleanrl_b200.tc_mlp import _splits  # noqa: E402

M, dev = 131072, "cuda"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for (K, N) in [(64, 256), (256, 512), (512, 512), (512, 256)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        gemm_bf16(a, w, out, EPI_BIAS_TANH_BF16, bias=b)
for (K, N) in [(256, 512), (512, 512), (512, 256)]:
    dz = torch.randn(M, K, device=dev).to(torch.bfloat16)
    wt = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    y = torch.tanh(torch.randn(M, N, device=dev)).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        gemm_bf16(dz, wt, out, EPI_DTANH_BF16, aux=y)
for (n_out, k_in) in [(256, 64), (512, 256), (512, 512), (256, 512)]:
    dz = torch.randn(M, n_out, device=dev).to(torch.bfloat16)
    x = torch.randn(M, k_in, device=dev).to(torch.bfloat16)
    dw = torch.zeros(n_out, k_in, device=dev)
    for _ in range(reps):
        gemm_bf16(dz, x, dw, EPI_ATOMIC_F32, splits=_splits(n_out, k_in, M), mn_major=True)
torch.cuda.synchronize()
print("ok")
