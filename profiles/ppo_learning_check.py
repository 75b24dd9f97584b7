"""Functional check: PPO on the B200 engine learns — with the tcgen05 bf16 MLPs AND with the fp32 torch
MLPs (`--mlp-backend torch`, the reference's arithmetic), same seeds, so the return curves can be laid side
by side. Trains for a fixed sample budget, prints the episodic return along the way, then plays the trained
agent (as blue) against the `zero` and `ou` teams with the goal-only reward (the reference's evaluation,
ppo…:389-423, play.py:131-164) and appends one JSON line to gpurun_out/learning_check.jsonl.
usage: python profiles/ppo_learning_check.py [env_id sa|cma|dma] [backend tc|torch] [seed] [total_timesteps] [num_envs]"""
import json
import os
import re
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rsoccer_isaac_cleanrl_b200 import play, ppo  # noqa: E402
from rsoccer_isaac_cleanrl_b200.envs import VSS, load_cfg  # noqa: E402

env_id = sys.argv[1] if len(sys.argv) > 1 else "sa"
backend = sys.argv[2] if len(sys.argv) > 2 else "tc"
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
total = int(sys.argv[4]) if len(sys.argv) > 4 else 150_000_000
n = int(sys.argv[5]) if len(sys.argv) > 5 else (4095 if env_id == "dma" else 4096)
args = ppo.parse_args(["--env-id", env_id, "--num-envs", str(n), "--total-timesteps", str(total), "--seed", str(seed),
                       "--mlp-backend", backend, "--tensorboard", "false"])
hist = []
t0 = time.time()
st = ppo.train(args, log=hist.append)
wall = time.time() - t0
k = max(1, len(hist) // 10)
for line in hist[::k] + [hist[-1]]:
    print(line)
curve = [(int(re.search(r"step (\d+)", h).group(1)), float(re.search(r"ep_ret\(last\) (-?[\d.]+)", h).group(1)))
         for h in hist if "ep_ret" in h]
print(f"[{env_id} {backend} seed {seed}] trained {st['global_step']} samples in {wall:.1f} s "
      f"({st['global_step'] / wall / 1e6:.2f} M samples/s overall), sanitised fields {st['sanitised_fields']}")
path = f"/tmp/ppo_{env_id}_{backend}_{seed}.pt"
torch.save(st["agent"].state_dict(), path)
del st
cfg = load_cfg()
cfg["env"]["numEnvs"] = 1065
envs = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=123)
envs.w_goal, envs.w_grad, envs.w_move, envs.w_energy = 1.0, 0.0, 0.0, 0.0
blue = play.get_team(f"ppo-{env_id}", path, mlp_backend=backend)
res = {"env_id": env_id, "backend": backend, "seed": seed, "samples": total, "train_wall_s": wall,
       "ep_ret_curve": curve[::max(1, len(curve) // 24)] + curve[-1:], "ep_ret_last": curve[-1][1]}
for opp in ("zero", "ou"):
    score, length = play.play_matches(envs, blue, play.get_team(opp), 3000)
    res[f"score_vs_{opp}"], res[f"length_vs_{opp}"] = score, length
    print(f"trained ppo-{env_id} ({backend}) as blue vs {opp}: score {score:+.3f}, mean length {length:.1f}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "learning_check.jsonl"), "a") as f:
    f.write(json.dumps(res) + "\n")
