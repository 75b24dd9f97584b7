"""CPU: the product's per-lane step code (host build, tests/emu) against the oracle."""
import os

import pytest

import parity_checks as pc
from backends import EmuBackend
from conftest import GOLDEN
from oracle import vss_oracle as orc


def test_reset_matches_oracle():
    pc.check_reset(EmuBackend)


def test_injected_step_is_exact():
    assert pc.check_injected(EmuBackend) > 0


def test_injected_step_energy_weight():
    pc.check_injected(EmuBackend, n=100, steps=2, w_energy=0.25, w_goal=1.0, w_grad=0.0)


def test_golden_rewards_and_obs():
    pc.check_golden_injected(EmuBackend, os.path.join(GOLDEN, "jit_functions.npz"))


def test_rollout_tracks_oracle_physics():
    r = pc.check_rollout(EmuBackend)
    print(r)


@pytest.mark.parametrize("view", [orc.VIEW_SA, orc.VIEW_CMA, orc.VIEW_DMA])
def test_views(view):
    pc.check_views(EmuBackend, view)


def test_degenerate_contact_normal_regression():
    pc.check_degenerate_contact(EmuBackend, os.path.join(GOLDEN, "degenerate_ball_on_box_corner.npz"))


def test_nonfinite_state_guard():
    pc.check_nonfinite_guard(EmuBackend)


def test_contact_heavy_rollout_stays_finite():
    pc.check_chase_stress(EmuBackend)


def test_wall_and_goal_post_contacts_track_oracle():
    print("flips", pc.check_wall_stress(EmuBackend))


def test_robot_pair_contacts_track_oracle():
    print("flips", pc.check_pair_stress(EmuBackend))


def test_physics_model_embodies_the_scene_spec():
    pc.check_physics_spec(EmuBackend)


def test_reset_reject_rate_matches_the_reference_rule():
    """envs/vss.py:281-299; SURVEY App. D: P(reject) = 0.1799 per draw."""
    rate = pc.check_reset_reject_rate(EmuBackend, n=40000)
    assert 0.17 < rate < 0.19


@pytest.mark.parametrize("view", [orc.VIEW_SA, orc.VIEW_DMA])
def test_ou_noise_moments(view):
    """envs/wrappers.py:5-19: innovation N(0, 0.15)."""
    pc.check_ou_moments(EmuBackend, view, n=8192, steps=5)


@pytest.mark.parametrize("view", [orc.VIEW_SA, orc.VIEW_CMA, orc.VIEW_DMA])
def test_packed_host_rows(view):
    """include/vss_b200.h vss_set_step_packed: 52 bf16 obs | f32 reward | u8 done | u8 timeout."""
    pc.check_packed_rows(EmuBackend, view)


# ---- the k_step_cta structure (one tile shared by W warps, bodies dealt to the warps)
@pytest.mark.parametrize("wpt", [2, 3, 5, 8, (8, 8), (4, 16)])
def test_cta_structure_matches_oracle(wpt):
    from backends import with_wpt
    B = with_wpt(EmuBackend, *wpt) if isinstance(wpt, tuple) else with_wpt(EmuBackend, wpt)
    pc.check_reset(B, n=300)
    assert pc.check_injected(B, n=300, steps=3) > 0
    pc.check_rollout(B, n=500, steps=12)
    for view in (orc.VIEW_SA, orc.VIEW_CMA, orc.VIEW_DMA):
        pc.check_views(B, view, n=200, steps=6)
        pc.check_packed_rows(B, view, n=200, steps=3)


@pytest.mark.parametrize("wpt", [2, 8, (8, 8), (3, 16)])
def test_cta_structure_is_bit_identical_to_the_lane_structure(wpt):
    """Same per-body code in the same order per field: every output word equal, over a contact-heavy rollout."""
    import numpy as np
    from backends import with_wpt
    n = 777
    B = with_wpt(EmuBackend, *wpt) if isinstance(wpt, tuple) else with_wpt(EmuBackend, wpt)
    a, b = EmuBackend(n, seed=3, goff=9), B(n, seed=3, goff=9)
    rb_a, rb_b = np.ones(n, np.int64), np.ones(n, np.int64)
    oa, ob = a.reset_dones(rb_a), b.reset_dones(rb_b)
    assert np.array_equal(pc.bits(oa), pc.bits(ob))
    rb_a[:] = 0; rb_b[:] = 0
    rng = np.random.default_rng(1)
    pc.stage_interesting_state(a, np.random.default_rng(5)); pc.stage_interesting_state(b, np.random.default_rng(5))
    obs = oa
    for t in range(25):
        act = pc.chase_actions(obs, rng.uniform(-1.2, 1.2, (n, 2, 3, 2)))
        xa, xb = a.step(act, rb_a), b.step(act, rb_b)
        for k in ("obs", "term_obs", "rew", "timeout", "progress_f"):
            assert np.array_equal(xa[k].view(np.uint8), xb[k].view(np.uint8)), (t, k)
        assert np.array_equal(rb_a, rb_b)
        assert np.array_equal(a.get_state().view(np.uint32), b.get_state().view(np.uint32)), t
        obs = xa["obs"]
    # the views (OU noise, statistics) too
    abuf_a, abuf_b = np.zeros((n, 2, 3, 2), np.float32), np.zeros((n, 2, 3, 2), np.float32)
    for view in (orc.VIEW_SA, orc.VIEW_CMA, orc.VIEW_DMA):
        nv, adim = (3 * n if view == orc.VIEW_DMA else n), (6 if view == orc.VIEW_CMA else 2)
        era, erb = np.zeros((nv, 4), np.float32), np.zeros((nv, 4), np.float32)
        ela, elb = np.zeros(nv, np.int32), np.zeros(nv, np.int32)
        for t in range(4):
            pa = rng.uniform(-1.2, 1.2, (nv, adim)).astype(np.float32)
            xa = a.step_view(view, pa, abuf_a, rb_a, era, ela, packed=True)
            xb = b.step_view(view, pa, abuf_b, rb_b, erb, elb, packed=True)
            for k in xa:
                assert np.array_equal(xa[k].view(np.uint8), xb[k].view(np.uint8)), (view, t, k)
            assert np.array_equal(abuf_a.view(np.uint32), abuf_b.view(np.uint32)) and np.array_equal(rb_a, rb_b)
            assert np.array_equal(era, erb) and np.array_equal(ela, elb)
