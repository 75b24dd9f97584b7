"""Profiling target: N full-contract steps at a given batch size and launch shape.
   python profiles/step_small_profile.py <fields> <warps per tile, 0 = automatic> [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.envs import VSS, load_cfg  # noqa: E402

n, w = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cfg = load_cfg()
cfg["env"]["numEnvs"] = n
task = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=1)
task.reset_buf.zero_()
st = task.engine.get_state()
st[58, :n] = torch.randint(0, 400, (n,), device="cuda", dtype=torch.int32).view(torch.float32)
task.engine.set_state(st)
task.engine.warps_per_tile = w
acts = [torch.rand((n, 2, 3, 2), device="cuda") * 2 - 1 for _ in range(4)]
for i in range(steps):
    task.step(acts[i & 3])
torch.cuda.synchronize()
print("ok", n, task.engine.warps_per_tile)
