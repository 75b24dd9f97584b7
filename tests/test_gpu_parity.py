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


# ----------------------------------------------------------------------------------------------------
# The launch shapes the benchmark times (csrc/vss_step.cu::launch_shape), against the oracle:
#   4 096 fields   -> 8 fields per warp,  2 warps per CTA, no barriers      (BASELINE configs[1] ppo-sa)
#   16 384, 21 845 -> 16 fields per warp, 2 warps per CTA                    (configs[2] cma; configs[3] dma fields)
#   65 536         -> 32 fields per warp, 2 warps per CTA                    (configs[3], sweep)
#   100 003        -> 4 warps per CTA + CTA-wide barriers, ragged against the 128-field CTAs, one wave
#   131 072        -> the same shape with the first-wave stagger (> 888 CTAs): what BENCH / SCALE time at 2^20
# ----------------------------------------------------------------------------------------------------
TIMED_SIZES = [4096, 16384, 21845, 65536, 100003, 131072]


@pytest.mark.parametrize("n", TIMED_SIZES)
def test_timed_launch_shapes_rollout_tracks_oracle(Gpu, n):
    r = pc.check_rollout(Gpu, n=n, steps=4, seed=n % 97)
    assert r["dones"] > 0 and r["timeouts"] > 0 and r["goals"] > 0


@pytest.mark.parametrize("n", TIMED_SIZES)
def test_timed_launch_shapes_injected_step_is_exact(Gpu, n):
    assert pc.check_injected(Gpu, n=n, steps=2, seed=n % 89) > 0


@pytest.mark.parametrize("view,n", [(orc.VIEW_SA, 4096), (orc.VIEW_CMA, 16384), (orc.VIEW_DMA, 21845),
                                    (orc.VIEW_SA, 65536), (orc.VIEW_SA, 100003), (orc.VIEW_SA, 131072),
                                    (orc.VIEW_CMA, 131072), (orc.VIEW_DMA, 131072)])
def test_timed_launch_shapes_views_track_oracle(Gpu, view, n):
    pc.check_views(Gpu, view, n=n, steps=3, seed=5 + n % 7)


def test_timed_launch_shape_wall_and_pair_contacts(Gpu):
    """The contact stress checks on the 4-warp, barrier-synchronised, staggered shape."""
    pc.check_wall_stress(Gpu, n=131072, steps=3)
    pc.check_pair_stress(Gpu, n=131072, steps=3)


def test_step_ranges_at_full_size_match_oracle():
    """2^20 fields (the benchmark's size) stepped through `SingleAgent.step_host`, i.e. as 8 field ranges
    (vss_set_step_range) on two streams: every range against the oracle's view step from the same state,
    and the whole against one un-chunked launch of a twin engine."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rsoccer_isaac_cleanrl_b200.envs import VSS, SingleAgent, load_cfg
    n, seed = 1 << 20, 31
    cfg = load_cfg()
    cfg["env"]["numEnvs"] = n
    task = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=seed)
    twin = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=seed)
    view, view2 = SingleAgent(task), SingleAgent(twin)
    assert n >= view.HOST_CHUNK_MIN_FIELDS and view.HOST_CHUNKS > 1
    task.reset_buf.zero_(); twin.reset_buf.zero_()
    rng = np.random.default_rng(0)
    st = task.engine.get_state()
    st[58, :n] = torch.from_numpy(rng.integers(380, 400, n).astype(np.int32)).cuda().view(torch.float32)
    task.engine.set_state(st); twin.engine.set_state(st)
    del st
    p = orc.default_params()
    abuf = np.zeros((n, 2, 3, 2), np.float32)
    rb = np.zeros(n, np.int64)
    for t in range(2):
        ost = orc.State.from_soa(task.engine.get_state().cpu().numpy(), n)
        pa = torch.from_numpy(rng.uniform(-1.2, 1.2, (n, 2)).astype(np.float32)).pin_memory()
        step_index = task.engine.step_count
        assert step_index == t
        obs_h, rew_h, done_h = view.step_host(pa)
        o2, r2, d2, _ = view2.step(pa.cuda())
        # (a) chunked == un-chunked, bit for bit (same kernels, same RNG keys)
        assert torch.equal(view._obs, o2["obs"]) and torch.equal(view._reward, r2) and torch.equal(view._done, d2)
        assert torch.equal(view.action_buf, view2.action_buf)
        assert torch.equal(task.engine.get_state(), twin.engine.get_state())
        # (b) what the caller receives on the host == the device buffers
        host = view.host_outputs_as_f32(obs_h, rew_h, done_h)
        assert torch.equal(host[1], view._reward.cpu()) and torch.equal(host[2], view._done.cpu())
        np.testing.assert_allclose(host[0].numpy(), view._obs.cpu().numpy(), rtol=2 ** -8, atol=0)
        # (c) against the oracle
        rb_ref = rb.copy()
        ref = orc.step_view(p, seed, 0, step_index, ost, orc.VIEW_SA, pa.numpy(), abuf, rb_ref)
        rb[...] = task.reset_buf.cpu().numpy()
        same = rb == rb_ref
        tol = lambda a, b: np.abs(a - b) <= 3 * (pc.PHYS_ATOL + pc.PHYS_RTOL * np.abs(b))
        good = same & tol(view._term_obs.cpu().numpy(), ref["term_obs"]).all(1) & tol(view._reward.cpu().numpy(), ref["reward"])
        good &= (view._done.cpu().numpy() == ref["done"]) & (view._timeout_u8.cpu().numpy() == ref["timeout"])
        assert (~good).sum() <= pc.FLIP_FRACTION * n, int((~good).sum())
        assert rb.sum() > 0
        both = good & (rb != 0)   # fields reset by both: the fresh observation agrees
        pc.assert_obs_equal(view._obs.cpu().numpy()[both], ref["obs"][both], "obs after reset", trig_atol=1e-6)
        np.testing.assert_allclose(view.action_buf.cpu().numpy()[same], abuf[same], rtol=0, atol=2e-6)
        abuf[...] = view.action_buf.cpu().numpy()
    assert task.engine.step_count == 2


def test_reset_reject_rate_matches_the_reference_rule(Gpu):
    """envs/vss.py:281-299; SURVEY App. D: P(reject) = 0.1799 per draw."""
    rate = pc.check_reset_reject_rate(Gpu, n=200000)
    assert abs(rate - 0.18) < 0.005, rate


@pytest.mark.parametrize("view", [orc.VIEW_SA, orc.VIEW_CMA, orc.VIEW_DMA])
def test_ou_noise_moments(Gpu, view):
    """envs/wrappers.py:5-19: the in-kernel N(0, 0.15) innovation of the opponents' action buffer."""
    r = pc.check_ou_moments(Gpu, view, n=65536, steps=6)
    assert abs(r["std"] - 0.15) < 5e-4


@pytest.mark.parametrize("view,n", [(orc.VIEW_SA, 500), (orc.VIEW_CMA, 4097), (orc.VIEW_DMA, 21845), (orc.VIEW_SA, 131072)])
def test_packed_host_rows(Gpu, view, n):
    """include/vss_b200.h vss_set_step_packed: 52 bf16 obs | f32 reward | u8 done | u8 timeout."""
    pc.check_packed_rows(Gpu, view, n=n, steps=4)


# ---- launch shapes: k_step (one warp per tile) and k_step_cta (2..8 warps share a tile)
def test_every_launch_shape_is_bit_identical(Gpu):
    """include/vss_b200.h vss_set_step_warps_per_tile: every shape computes the same words (full contract
    and the three views), on a size that is ragged against the 32-field tiles, over contact-heavy steps."""
    from backends import with_wpt
    n = 4097
    shapes = [1, 2, 4, 7, 8, (8, 16), (8, 8), (3, 8), (1, 32), (1, 8)]   # warps per tile or (warps, fields) per tile
    bes = [(with_wpt(Gpu, *w) if isinstance(w, tuple) else with_wpt(Gpu, w))(n, seed=3, goff=9) for w in shapes]
    assert [b.eng.warps_per_tile for b in bes] == [w[0] if isinstance(w, tuple) else w for w in shapes]
    assert [b.eng.fields_per_tile for b in bes[5:]] == [16, 8, 8, 32, 8]
    rbs = [np.ones(n, np.int64) for _ in bes]
    obs = [b.reset_dones(rb) for b, rb in zip(bes, rbs)]
    for rb in rbs:
        rb[:] = 0
    for b in bes:
        pc.stage_interesting_state(b, np.random.default_rng(5))
    rng = np.random.default_rng(1)
    o = obs[0]
    for t in range(12):
        act = pc.chase_actions(o, rng.uniform(-1.2, 1.2, (n, 2, 3, 2)))
        outs = [b.step(act, rb) for b, rb in zip(bes, rbs)]
        for w, x, rb in zip(shapes[1:], outs[1:], rbs[1:]):
            for k in ("obs", "term_obs", "rew", "timeout", "progress_f"):
                assert np.array_equal(outs[0][k].view(np.uint8), x[k].view(np.uint8)), (t, w, k)
            assert np.array_equal(rbs[0], rb)
        o = outs[0]["obs"]
    for view in (orc.VIEW_SA, orc.VIEW_CMA, orc.VIEW_DMA):
        nv, adim = (3 * n if view == orc.VIEW_DMA else n), (6 if view == orc.VIEW_CMA else 2)
        abufs = [np.zeros((n, 2, 3, 2), np.float32) for _ in bes]
        ers, els = [np.zeros((nv, 4), np.float32) for _ in bes], [np.zeros(nv, np.int32) for _ in bes]
        for b in bes:
            b.step_count = 0
        for t in range(3):
            pa = rng.uniform(-1.2, 1.2, (nv, adim)).astype(np.float32)
            outs = [b.step_view(view, pa, ab, rb, er, el, packed=True) for b, ab, rb, er, el in zip(bes, abufs, rbs, ers, els)]
            for w, x, ab in zip(shapes[1:], outs[1:], abufs[1:]):
                for k in x:
                    assert np.array_equal(outs[0][k].view(np.uint8), x[k].view(np.uint8)), (view, t, w, k)
                assert np.array_equal(abufs[0].view(np.uint32), ab.view(np.uint32))
        assert all(b.step_count == 3 for b in bes)


@pytest.mark.parametrize("wpt,n", [(7, 4096), (4, 21845), (2, 65536), (8, 1000), (2, 131072), ((8, 8), 4096),
                                   ((8, 16), 16384), ((5, 8), 1001)])
def test_cta_launch_shapes_track_oracle(Gpu, wpt, n):
    from backends import with_wpt
    B = with_wpt(Gpu, *wpt) if isinstance(wpt, tuple) else with_wpt(Gpu, wpt)
    r = pc.check_rollout(B, n=n, steps=4, seed=n % 97)
    assert r["dones"] > 0 and r["timeouts"] > 0 and r["goals"] > 0
    assert pc.check_injected(B, n=min(n, 20000), steps=2) > 0
    pc.check_views(B, orc.VIEW_DMA, n=min(n, 20000), steps=3)
