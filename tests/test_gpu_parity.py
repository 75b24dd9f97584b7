"""GPU (B200): libvss_b200.so through the C-ABI against the oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest

import parity_checks as pc
from conftest import GOLDEN
from oracle import vss_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from backends import GpuBackend
    return GpuBackend


def test_reset_matches_oracle(Gpu):
    pc.check_reset(Gpu)


def test_injected_step_is_exact(Gpu):
    assert pc.check_injected(Gpu) > 0


def test_injected_step_energy_weight(Gpu):
    pc.check_injected(Gpu, n=100, steps=2, w_energy=0.25, w_goal=1.0, w_grad=0.0)


def test_golden_rewards_and_obs(Gpu):
    pc.check_golden_injected(Gpu, os.path.join(GOLDEN, "jit_functions.npz"))


def test_rollout_tracks_oracle_physics(Gpu):
    print(pc.check_rollout(Gpu))


@pytest.mark.parametrize("n", [1, 31, 33, 4097])
def test_ragged_sizes(Gpu, n):
    be, p = pc.make_backend_pair(Gpu, n, 3, 10)
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rng = np.random.default_rng(n)
    for t in range(5):
        st = pc.oracle_from_backend(be)
        rb_ref = rb.copy()
        actions = rng.uniform(-1, 1, (n, 2, 3, 2)).astype(np.float32)
        out = be.step(actions, rb)
        ref = orc.step(p, 3, 10, st, actions, rb_ref)
        assert pc.compare_full_step(out, ref, rb, rb_ref, n, f"ragged {n} step {t}", exact_physics=False) <= 1


@pytest.mark.parametrize("view", [orc.VIEW_SA, orc.VIEW_CMA, orc.VIEW_DMA])
def test_views(Gpu, view):
    pc.check_views(Gpu, view)


def test_gpu_count_invariance(Gpu):
    """Sharding by global field id: two engines over [0,n) and [n,2n) equal one engine over [0,2n)."""
    n = 96
    whole = Gpu(2 * n, seed=11, goff=0)
    lo, hi = Gpu(n, seed=11, goff=0), Gpu(n, seed=11, goff=n)
    ones2, ones = np.ones(2 * n, np.int64), np.ones(n, np.int64)
    ow = whole.reset_dones(ones2)
    ol, oh = lo.reset_dones(ones), hi.reset_dones(ones)
    assert np.array_equal(pc.bits(ow[:n]), pc.bits(ol)) and np.array_equal(pc.bits(ow[n:]), pc.bits(oh))
    rng = np.random.default_rng(0)
    rbw, rbl, rbh = np.zeros(2 * n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
    for t in range(30):
        a = rng.uniform(-1, 1, (2 * n, 2, 3, 2)).astype(np.float32)
        w, l, h = whole.step(a, rbw), lo.step(a[:n], rbl), hi.step(a[n:], rbh)
        for k in ("obs", "term_obs", "rew", "timeout", "progress_f"):
            assert np.array_equal(w[k][:n], l[k]) and np.array_equal(w[k][n:], h[k]), (t, k)
        assert np.array_equal(rbw[:n], rbl) and np.array_equal(rbw[n:], rbh)


def test_gae_matches_reference_loop_and_oracle():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsoccer_isaac_cleanrl_b200.engine import gae
    z = np.load(os.path.join(GOLDEN, "gae.npz"))
    for name in "abc":
        args = [torch.from_numpy(z[f"{name}_{k}"]).cuda() for k in
                ("rewards", "values", "next_values", "next_dones", "next_timeouts")]
        adv, ret = gae(*args, 0.99, 0.95)
        # golden: the reference's python loop executed on CPU torch
        np.testing.assert_allclose(adv.cpu().numpy(), z[f"{name}_advantages"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(ret.cpu().numpy(), z[f"{name}_returns"], rtol=1e-6, atol=1e-6)
        # oracle: same IEEE op sequence -> bit-exact
        oadv, oret = orc.gae(*[z[f"{name}_{k}"] for k in
                               ("rewards", "values", "next_values", "next_dones", "next_timeouts")], 0.99, 0.95)
        assert np.array_equal(pc.bits(adv.cpu().numpy()), pc.bits(oadv))
        assert np.array_equal(pc.bits(ret.cpu().numpy()), pc.bits(oret))
    # BASELINE size (T=128, N=65536): bit-exact against the oracle
    rng = np.random.default_rng(1)
    T, N = 128, 65536
    a = [rng.normal(size=(T, N)).astype(np.float32) for _ in range(3)]
    d = (rng.uniform(size=(T, N)) < 0.01).astype(np.float32)
    to = ((rng.uniform(size=(T, N)) < 0.5) * d).astype(np.float32)
    adv, ret = gae(*[torch.from_numpy(x).cuda() for x in (*a, d, to)], 0.99, 0.95)
    oadv, oret = orc.gae(*a, d, to, 0.99, 0.95)
    assert np.array_equal(pc.bits(adv.cpu().numpy()), pc.bits(oadv))
    assert np.array_equal(pc.bits(ret.cpu().numpy()), pc.bits(oret))


@pytest.mark.parametrize("T,N", [(1, 5), (7, 33), (8, 129), (9, 1000), (23, 4097), (128, 150000), (40, 262144)])
def test_gae_ragged_shapes_bit_exact(T, N):
    """T below / equal to / not a multiple of the 8-step load group (remainder loop), both load
    schedules (prefetching below 131072 columns, plain above): bit-exact against the oracle."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsoccer_isaac_cleanrl_b200.engine import gae
    rng = np.random.default_rng(T * 1000 + N)
    a = [rng.normal(size=(T, N)).astype(np.float32) for _ in range(3)]
    d = (rng.uniform(size=(T, N)) < 0.05).astype(np.float32)
    to = ((rng.uniform(size=(T, N)) < 0.5) * d).astype(np.float32)
    adv, ret = gae(*[torch.from_numpy(x).cuda() for x in (*a, d, to)], 0.99, 0.95)
    oadv, oret = orc.gae(*a, d, to, 0.99, 0.95)
    assert np.array_equal(pc.bits(adv.cpu().numpy()), pc.bits(oadv))
    assert np.array_equal(pc.bits(ret.cpu().numpy()), pc.bits(oret))


def test_full_size_properties(Gpu):
    """N = 1,048,576 fields (top of the BASELINE sweep): size-independent properties."""
    n = 1 << 20
    be, p = pc.make_backend_pair(Gpu, n, 5, 0)
    rb = np.ones(n, np.int64)
    obs0 = be.reset_dones(rb)
    rb[:] = 0
    rng = np.random.default_rng(2)
    a = rng.uniform(-1, 1, (n, 2, 3, 2)).astype(np.float32)
    out = be.step(a, rb)
    assert np.isfinite(out["obs"]).all() and np.isfinite(out["rew"]).all()
    # yellow rows are the 180-degree mirror of the blue view of the same state (vss.py:560-574)
    o = out["term_obs"]
    assert np.array_equal(pc.bits(o[:, 1, :, 0:4]), pc.bits(-o[:, 0, :, 0:4]))
    # blue robot i's own block == yellow's opponent block for that robot, mirrored
    for i in range(3):
        own = o[:, 0, 0, 4 + 9 * i:4 + 9 * i + 7]      # blue robot i seen by blue robot 0 (perm identity)
        opp = o[:, 1, 0, 31 + 7 * i:31 + 7 * i + 7]    # the same robot seen by yellow robot 0
        m = np.array([-1, -1, -1, -1, -1, -1, 1], np.float32)
        assert np.array_equal(pc.bits(opp), pc.bits(own * m))
    # kept fields: obs == term_obs; progress advanced by exactly one; a sample matches the oracle
    keep = rb == 0
    assert np.array_equal(pc.bits(out["obs"][keep]), pc.bits(out["term_obs"][keep]))
    assert np.all(out["progress_f"] == 1.0)
    # energy reward off at the default weights; goal reward is in {-10, 0, 10}
    assert np.all(out["rew"][..., 3] == 0) and set(np.unique(out["rew"][..., 0])) <= {-10.0, 0.0, 10.0}
    assert np.array_equal(out["rew"][:, 0, :, 1], -out["rew"][:, 1, :, 1])


def test_degenerate_contact_normal_regression(Gpu):
    pc.check_degenerate_contact(Gpu, os.path.join(GOLDEN, "degenerate_ball_on_box_corner.npz"))


def test_nonfinite_state_guard(Gpu):
    pc.check_nonfinite_guard(Gpu)


def test_contact_heavy_rollout_stays_finite(Gpu):
    pc.check_chase_stress(Gpu, n=65536, steps=300)


def test_wall_and_goal_post_contacts_track_oracle(Gpu):
    pc.check_wall_stress(Gpu, n=8192, steps=6)


def test_robot_pair_contacts_track_oracle(Gpu):
    pc.check_pair_stress(Gpu, n=8192, steps=6)


def test_physics_model_embodies_the_scene_spec(Gpu):
    pc.check_physics_spec(Gpu)
