"""CPU: the oracle restatement against golden vectors produced by the reference's own code
(tests/golden/make_golden.py). This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import vss_oracle as orc

from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "jit_functions.npz"))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        got = orc.philox4x32_10(ctr, key)
        assert [int(x) for x in got] == want


def test_obs_matches_reference_jit(g):
    obs = orc.compute_obs(g["ball_pos"], g["ball_vel"], g["r_pos"], g["r_vel"], g["r_rot"], g["r_w"][..., 0],
                          g["acts"])
    ref = g["obs"]
    assert obs.shape == ref.shape == (96, 2, 3, 52)
    # slots holding cos/sin of the yaw: the reference goes quat -> atan2 -> cos/sin in fp32
    trig = np.zeros(52, bool)
    for base, width in ((4, 9), (13, 9), (22, 9), (31, 7), (38, 7), (45, 7)):
        trig[base + 4:base + 6] = True
    # everything else is a signed copy: bit-exact including the sign of zero
    assert np.array_equal(obs[..., ~trig].view(np.uint32), ref[..., ~trig].view(np.uint32))
    np.testing.assert_allclose(obs[..., trig], ref[..., trig], rtol=0, atol=1e-6)


def test_rewards_and_dones_match_reference_jit(g):
    goal = orc.compute_goal_rew(g["ball_pos"])
    assert goal.dtype == np.int64 and np.array_equal(goal, g["goal_rew"])
    assert set(np.unique(goal)) == {-1, 0, 1}
    dones = orc.compute_dones(g["ball_pos"], g["progress"])
    assert np.array_equal(dones, g["dones"])
    grad = orc.compute_grad_rew(g["prev_ball_pos"], g["ball_pos"])
    move = orc.compute_move_rew(g["prev_r_pos"], g["r_pos"], g["prev_ball_pos"], g["ball_pos"])
    energy = orc.compute_energy_rew(g["acts"])
    # fp32; differences of O(1) norms -> 1e-5 relative plus a few ulp of the norms
    np.testing.assert_allclose(grad, g["grad_rew"], rtol=1e-5, atol=5e-7)
    np.testing.assert_allclose(move, g["move_rew"], rtol=1e-5, atol=5e-7)
    np.testing.assert_allclose(energy, g["energy_rew"], rtol=1e-6, atol=0)


def test_gae_matches_reference_loop():
    z = np.load(os.path.join(GOLDEN, "gae.npz"))
    for name in "abc":
        adv, ret = orc.gae(z[f"{name}_rewards"], z[f"{name}_values"], z[f"{name}_next_values"],
                           z[f"{name}_next_dones"], z[f"{name}_next_timeouts"], 0.99, 0.95)
        np.testing.assert_allclose(adv, z[f"{name}_advantages"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(ret, z[f"{name}_returns"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name,view", [("sa", orc.VIEW_SA), ("cma", orc.VIEW_CMA), ("dma", orc.VIEW_DMA)])
def test_views_match_reference_wrappers(name, view):
    z = np.load(os.path.join(GOLDEN, "wrappers.npz"))
    n, T = int(z["n"]), int(z["T"])
    nv = n * 3 if name == "dma" else n
    ep_ret, ep_len = np.zeros((nv, 4), np.float32), np.zeros((nv,), np.int32)
    for t in range(T):
        # the action buffer the reference stepped the task with (after OU + policy overwrite)
        abuf = np.ascontiguousarray(z[f"{name}{t}_stepped_actions"], np.float32)
        # policy-controlled slots hold the policy action (wrappers.py:103,135,165)
        pa = z[f"{name}{t}_policy_action"]
        if name == "sa":
            assert np.array_equal(abuf[:, 0, 0, :], pa)
        else:
            assert np.array_equal(abuf[:, 0].reshape(n, 6), pa.reshape(n, 6))
        out = orc.view_outputs(view, z[f"in{t}_obs"], z[f"in{t}_term_obs"], z[f"in{t}_rew"], z[f"in{t}_reset"],
                               z[f"in{t}_timeout"].astype(np.uint8), z[f"in{t}_progress_f"], abuf, ep_ret, ep_len)
        assert np.array_equal(out["obs"], z[f"{name}{t}_obs"].reshape(nv, 52))
        assert np.array_equal(out["term_obs"], z[f"{name}{t}_term_obs"].reshape(nv, 52))
        assert np.array_equal(out["done"], z[f"{name}{t}_done"])
        assert np.array_equal(out["timeout"].astype(bool), z[f"{name}{t}_timeout"])
        assert np.array_equal(out["progress"], z[f"{name}{t}_progress"])
        np.testing.assert_allclose(out["rews"], z[f"{name}{t}_rews"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(out["reward"], z[f"{name}{t}_reward"], rtol=1e-6, atol=1e-6)
        assert np.array_equal(abuf, z[f"{name}{t}_action_buf_after"])
        np.testing.assert_allclose(out["ret_ret"], z[f"{name}{t}_ret_ret"], rtol=1e-6, atol=1e-6)
        assert np.array_equal(out["ret_len"], z[f"{name}{t}_ret_len"])
        np.testing.assert_allclose(ep_ret, z[f"{name}{t}_ep_ret"], rtol=1e-6, atol=1e-6)
        assert np.array_equal(ep_len, z[f"{name}{t}_ep_len"])
