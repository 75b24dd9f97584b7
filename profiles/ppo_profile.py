"""Kernel-time breakdown of one PPO update (rollout and update phases) with torch.profiler.
usage: python profiles/ppo_profile.py [env_id] [num_envs] [backend]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200 import ppo  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "sa"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
backend = sys.argv[3] if len(sys.argv) > 3 else "auto"
args = ppo.parse_args(["--env-id", env_id, "--num-envs", str(n), "--quiet", "--mlp-backend", backend,
                       "--total-timesteps", str(n * 128 * 3)])
# warm-up run (allocations, cuBLAS handles), then profile a 1-update run
ppo.train(args)
args.total_timesteps = n * 128 * 1
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st = ppo.train(args)
    torch.cuda.synchronize()
print(f"rollout {st['rollout_s']:.3f}s update {st['update_s']:.3f}s")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
