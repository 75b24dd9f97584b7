"""Find the first non-finite tensor during PPO training WITHOUT host syncs (flags stay on the device
until the end). usage: python profiles/nan_debug.py [timesteps] [backend] [graph] [seed]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200 import ppo  # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
backend = sys.argv[2] if len(sys.argv) > 2 else "auto"
graph = sys.argv[3] if len(sys.argv) > 3 else "true"
seed = sys.argv[4] if len(sys.argv) > 4 else "1"
args = ppo.parse_args(["--env-id", "sa", "--num-envs", "4096", "--total-timesteps", str(total), "--quiet",
                       "--seed", seed, "--mlp-backend", backend, "--cuda-graph", graph])
KEYS = ("obs", "term_obs", "actions", "logprobs", "rewards", "values", "next_values", "advantages", "returns",
        "flat_grad", "flat")
n_upd = total // (4096 * 128) + 2
flags = torch.ones((n_upd, len(KEYS) + 1), device="cuda", dtype=torch.bool)
extra = torch.zeros((n_upd, 4), device="cuda")


def hook(update, t):
    for i, k in enumerate(KEYS):
        flags[update, i] = torch.isfinite(t[k]).all()
    st = t["env"].engine.get_state()
    flags[update, len(KEYS)] = torch.isfinite(st[:58, :4096]).all()
    extra[update, 0] = t["agent"].actor_logstd.detach().min()
    extra[update, 1] = t["stats"]["approx_kl"]
    extra[update, 2] = t["flat_grad"].abs().max()
    extra[update, 3] = t["advantages"].abs().max()


ppo.train(args, hook=hook)
f = flags.cpu()
e = extra.cpu()
bad_rows = (~f).any(1).nonzero().flatten().tolist()
print("updates with a non-finite tensor:", bad_rows[:10], "..." if len(bad_rows) > 10 else "")
if bad_rows:
    u = bad_rows[0]
    for uu in range(max(1, u - 3), min(u + 2, n_upd)):
        print(f"update {uu}:", {k: bool(f[uu, i]) for i, k in enumerate(KEYS + ('env_state',))},
              "logstd_min %.3f kl %.4f |grad|max %.3g |adv|max %.3g" % tuple(e[uu].tolist()))
else:
    print("no non-finite values; last: logstd_min %.3f kl %.4f |grad|max %.3g |adv|max %.3g" % tuple(e[n_upd - 3].tolist()))
