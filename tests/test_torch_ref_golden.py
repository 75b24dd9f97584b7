"""CPU: the torch restatements `bench.py` times as "what the reference runs" (oracle/torch_ref.py)
against vectors produced by executing the reference's own functions (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import torch_ref as tr


def test_obs_rewards_dones_match_the_reference_jit_functions():
    g = np.load(os.path.join(GOLDEN, "jit_functions.npz"))
    t = lambda k: torch.from_numpy(g[k])
    s = dict(ball_pos=t("ball_pos"), ball_vel=t("ball_vel"), prev_ball_pos=t("prev_ball_pos"), r_pos=t("r_pos"),
             prev_r_pos=t("prev_r_pos"), r_vel=t("r_vel"), quats=t("quats"), r_w=t("r_w"), acts=t("acts"),
             reset_buf=t("reset_buf"), progress=t("progress"))
    obs, term_obs, rew, reset = tr.obs_rewards_dones(s, w=(10.0, 2.0, 3.0, 0.5))
    np.testing.assert_allclose(obs.numpy(), g["obs"], rtol=1e-6, atol=1e-6)
    assert torch.equal(obs, term_obs)
    assert np.array_equal(reset.numpy(), g["dones"])
    assert np.array_equal(rew[..., 0].numpy(), g["goal_rew"].astype(np.float32) * 10.0)
    np.testing.assert_allclose(rew[..., 1].numpy(), g["grad_rew"] * 2.0, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rew[..., 2].numpy(), g["move_rew"] * 3.0, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rew[..., 3].numpy(), g["energy_rew"] * 0.5, rtol=1e-6, atol=1e-7)


def test_gae_loop_matches_the_reference_loop():
    z = np.load(os.path.join(GOLDEN, "gae.npz"))
    for name in "abc":
        a = [torch.from_numpy(z[f"{name}_{k}"]) for k in ("rewards", "values", "next_values", "next_dones", "next_timeouts")]
        adv, ret = tr.gae_loop(*a)
        np.testing.assert_allclose(adv.numpy(), z[f"{name}_advantages"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(ret.numpy(), z[f"{name}_returns"], rtol=1e-6, atol=1e-6)
