"""Functional check: PPO on the B200 engine learns. Trains ppo-sa for a fixed sample budget and
reports the episodic return early vs late, then plays the trained agent (as blue) against the `ou`
and `zero` teams with the goal-only reward (as the reference's evaluation, ppo…:389-423).
usage: python profiles/ppo_learning_check.py [total_timesteps] [num_envs]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200 import play, ppo  # noqa: E402
from rsoccer_isaac_cleanrl_b200.envs import VSS, load_cfg  # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 150_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
args = ppo.parse_args(["--env-id", "sa", "--num-envs", str(n), "--total-timesteps", str(total), "--quiet", "--seed", "1"])
hist = []


def log(msg):
    hist.append(msg)


args.quiet = False
t0 = time.time()
st = ppo.train(args, log=log)
wall = time.time() - t0
k = max(1, len(hist) // 12)
for line in hist[::k] + [hist[-1]]:
    print(line)
print(f"trained {st['global_step']} samples in {wall:.1f} s ({st['global_step'] / wall / 1e6:.2f} M samples/s overall)")
path = "/tmp/ppo_sa_agent.pt"
torch.save(st["agent"].state_dict(), path)
del st
cfg = load_cfg()
cfg["env"]["numEnvs"] = 1065
envs = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=123)
envs.w_goal, envs.w_grad, envs.w_move, envs.w_energy = 1.0, 0.0, 0.0, 0.0
blue = play.get_team("ppo-sa", path)
for opp in ("zero", "ou"):
    score, length = play.play_matches(envs, blue, play.get_team(opp), 3000)
    base, blen = play.play_matches(envs, play.get_team("ou"), play.get_team(opp), 3000)
    print(f"trained ppo-sa (blue) vs {opp}: score {score:+.3f}, mean length {length:.1f}   |   ou vs {opp}: {base:+.3f}, {blen:.1f}")
