"""Tuning run: `SingleAgent.step_host` (vss_step_view_host) at 2^20 fields for different numbers of field ranges.
   python profiles/e2e_ranges_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.envs import VSS, SingleAgent, load_cfg  # noqa: E402

n = 1 << 20
cfg = load_cfg()
cfg["env"]["numEnvs"] = n
task = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=1)
task.reset_buf.zero_()
view = SingleAgent(task)
pa = [torch.rand((n, 2)).mul_(2).sub_(1).pin_memory() for _ in range(2)]
for chunks in (1, 2, 4, 8, 12, 16, 24, 32, 64):
    view.HOST_CHUNKS = chunks
    for i in range(3):
        view.step_host(pa[i & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        view.step_host(pa[i & 1])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"{chunks:3d} ranges: {ms:.3f} ms per step, {n / ms * 1e3:.3e} env-steps/s, {n * 112 / ms / 1e6:.1f} GB/s device->host")
