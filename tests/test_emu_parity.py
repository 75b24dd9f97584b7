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
