"""Stress: many fields stepped for many steps with (a) random actions or (b) a ball-chasing controller
on all six robots (heavy contact: scrums around the ball, pushing along walls and into goals).
Reports non-finite state words. usage: python profiles/nan_hunt.py [envs] [steps] [chase|random]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.envs import VSS, load_cfg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
mode = sys.argv[3] if len(sys.argv) > 3 else "chase"
cfg = load_cfg()
cfg["env"]["numEnvs"] = n
envs = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=7)
g = torch.Generator(device="cuda").manual_seed(3)
acts = [torch.rand((n, 2, 3, 2), device="cuda", generator=g) * 2.4 - 1.2 for _ in range(8)]
obs = envs.reset()["obs"]
worst = 0
for t in range(steps):
    if mode == "chase":
        rel = obs[..., 0:2] - obs[..., 4:6]                       # ball - own position (view frame)
        c, s = obs[..., 8], obs[..., 9]
        fwd = rel[..., 0] * c + rel[..., 1] * s
        lat = -rel[..., 0] * s + rel[..., 1] * c
        turn = torch.atan2(lat, fwd)
        a = torch.stack([1.0 - 1.5 * turn, 1.0 + 1.5 * turn], -1) + 0.2 * acts[t & 7]
    else:
        a = acts[t & 7]
    obs = envs.step(a.contiguous())[0]["obs"]
    if (t + 1) % 100 == 0:
        st = envs.engine.get_state()[:58, :n]
        bad = (~torch.isfinite(st)).any(0)
        nb = int(bad.sum())
        worst = max(worst, nb)
        if nb or (t + 1) % 500 == 0:
            print(f"step {t + 1}: non-finite fields {nb}, max |state word| {float(st[:, ~bad].abs().max()):.3f}", flush=True)
        if nb:
            i = int(bad.nonzero()[0])
            print("  field", i, "words", [round(float(x), 4) for x in st[:, i].tolist()])
            break
print("mode", mode, "env-steps:", n * (t + 1), "max fields with non-finite state:", worst)
