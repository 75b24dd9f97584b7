"""Small workload that runs every step-kernel variant (compute-sanitizer is closed on the GPU pool, so the
out-of-bounds check that is enforced is tests/test_gpu_api.py::test_outputs_stay_inside_their_buffers):
full contract, three views with side outputs, field ranges, 8 / 16 / 32 fields per warp, barrier
and no-barrier launch shapes, reset, GAE.
usage: [compute-sanitizer --tool memcheck] python profiles/sanitizer_smoke.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.engine import gae  # noqa: E402
from rsoccer_isaac_cleanrl_b200.envs import CMA, DMA, VSS, SingleAgent, load_cfg  # noqa: E402

for n in (100, 3000, 9000, 40000, 160000):
    cfg = load_cfg(); cfg["env"]["numEnvs"] = n
    envs = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=n)
    st = envs.engine.get_state()
    st[58, :n] = torch.randint(390, 400, (n,), device="cuda", dtype=torch.int32).view(torch.float32)
    envs.engine.set_state(st)
    for i in range(3):
        envs.step(torch.rand((n, 2, 3, 2), device="cuda") * 2 - 1)
    for cls in (SingleAgent, CMA, DMA):
        v = cls(VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=n + 1))
        v.enable_episode_stats() if hasattr(v, "enable_episode_stats") else None
        nv = v.num_view_envs
        x16 = torch.zeros((nv, 64), device="cuda", dtype=torch.bfloat16)
        df, tf = torch.zeros(nv, device="cuda"), torch.zeros(nv, device="cuda")
        for i in range(3):
            v.step(torch.rand((nv, v.ACT_DIM), device="cuda") * 2 - 1, obs_bf16_out=x16, done_f_out=df, timeout_f_out=tf)
        v.HOST_CHUNKS, v.HOST_CHUNK_MIN_FIELDS = 3, 1
        v.step_host((torch.rand((nv, v.ACT_DIM)) * 2 - 1).pin_memory())
    torch.cuda.synchronize()
    print("ok", n, flush=True)
for T, N in ((9, 1000), (128, 4096), (16, 200000)):
    a = [torch.randn((T, N), device="cuda") for _ in range(3)]
    d = (torch.rand((T, N), device="cuda") < 0.05).float()
    gae(*a, d, d * 0.5, 0.99, 0.95)
torch.cuda.synchronize()
print("done")
