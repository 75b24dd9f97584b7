"""Parity checks shared by the CPU (emu backend) and GPU (C-ABI backend) test files.

Tolerances (stated once, used everywhere):
  * integer / index / flag outputs (reset, timeout, progress, done, episode lengths): exact.
  * observations: every slot is a signed copy of a state word -> compared as raw bits,
    including the sign of zero; the cos/sin slots of freshly reset robots differ between
    glibc and CUDA libm -> atol 1e-6 there.
  * rewards: the CUDA side issues the same IEEE fp32 operation sequence as the oracle
    (no FMA contraction) -> bit-exact given identical states.
  * physics (float32 kernel vs float64 oracle of the same model): |a-b| <= 2e-5 + 1e-4|b|
    per control step; a contact test that flips on a last-bit difference changes a velocity
    discontinuously, so up to 0.5 % of the fields may exceed it ("branch flips").
"""
import numpy as np

from oracle import vss_oracle as orc

TRIG = np.zeros(52, bool)
for _base in (4, 13, 22, 31, 38, 45):
    TRIG[_base + 4:_base + 6] = True

PHYS_ATOL, PHYS_RTOL, FLIP_FRACTION = 2e-5, 1e-4, 0.005


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bits_equal(a, b, what):
    a, b = bits(a), bits(b)
    bad = np.argwhere(a != b)
    assert bad.size == 0, f"{what}: {len(bad)} words differ, first at {bad[:5].tolist()}"


def assert_obs_equal(a, b, what, trig_atol=0.0):
    """(…,52) observations: non-trig slots bitwise; trig slots bitwise or within trig_atol."""
    assert a.shape == b.shape, (a.shape, b.shape)
    assert_bits_equal(a[..., ~TRIG], b[..., ~TRIG], what + " (copy slots)")
    if trig_atol == 0.0:
        assert_bits_equal(a[..., TRIG], b[..., TRIG], what + " (cos/sin slots)")
    else:
        np.testing.assert_allclose(a[..., TRIG], b[..., TRIG], rtol=0, atol=trig_atol, err_msg=what)


def oracle_from_backend(be):
    return orc.State.from_soa(be.get_state(), be.n)


def make_backend_pair(Backend, n, seed, goff, **kw):
    p = orc.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    be = Backend(n, seed=seed, goff=goff, params=p)
    return be, p


# --------------------------------------------------------------------------- reset
def check_reset(Backend, n=1000, seed=123, goff=5000):
    be, p = make_backend_pair(Backend, n, seed, goff)
    ones = np.ones(n, np.int64)
    obs = be.reset_dones(ones)
    st = orc.State(n)
    orc.reset_dones(p, seed, goff, st, ones)
    got = orc.State.from_soa(be.get_state(), n)
    assert_bits_equal(got.ball_pos, st.ball_pos, "reset ball_pos")
    assert_bits_equal(got.ball_vel, st.ball_vel, "reset ball_vel")
    assert_bits_equal(got.r_pos, st.r_pos, "reset r_pos")
    assert_bits_equal(got.r_vel, st.r_vel, "reset r_vel")
    assert_bits_equal(got.r_w, st.r_w, "reset r_w")
    assert_bits_equal(got.r_act, st.r_act, "reset r_act")
    np.testing.assert_allclose(got.r_rot, st.r_rot, rtol=0, atol=1e-6)
    assert np.array_equal(got.episode, st.episode) and np.all(got.episode == 1)
    assert np.array_equal(got.progress, st.progress)
    ref_obs = orc.compute_obs(st.ball_pos, st.ball_vel, st.r_pos, st.r_vel, st.r_rot, st.r_w, st.r_act)
    assert_obs_equal(obs, ref_obs, "reset obs", trig_atol=1e-6)
    # distribution facts of envs/vss.py:267-327
    ent = np.concatenate([got.ball_pos[:, None, :], got.r_pos.reshape(n, 6, 2)], 1)
    d = np.linalg.norm(ent[:, :, None, :] - ent[:, None, :, :], axis=-1) + np.eye(7) * 10
    assert d.min() >= 0.07 - 1e-6
    assert np.abs(ent[..., 0]).max() <= 1.36 / 2 and np.abs(ent[..., 1]).max() <= 1.16 / 2
    assert np.abs(got.ball_vel).max() <= 0.5
    np.testing.assert_allclose(np.linalg.norm(got.r_rot, axis=-1), 1.0, atol=1e-6)
    # a second masked reset only touches flagged fields and advances their episode counter
    mask = (np.arange(n) % 3 == 0).astype(np.int64)
    before = be.get_state()
    be.reset_dones(mask)
    after = be.get_state()
    keep = mask == 0
    assert np.array_equal(before.view(np.uint32)[:, :n][:, keep], after.view(np.uint32)[:, :n][:, keep])
    assert np.all(after.view(np.uint32)[59, :n][~keep] == 2)
    orc.reset_dones(p, seed, goff, st, mask)
    assert_bits_equal(orc.State.from_soa(after, n).r_pos, st.r_pos, "second reset r_pos")
    return be, p


# --------------------------------------------------------------------------- full step
def stage_interesting_state(be, rng):
    """After a reset, push some fields towards timeouts, goals and contacts."""
    n = be.n
    s = be.get_state()
    prog = s.view(np.int32)[58]
    prog[:n] = rng.integers(0, 390, n)
    k = min(n, 48)
    prog[:k] = 396 + (np.arange(k) % 4)                 # timeouts within a few steps
    g = slice(k, min(n, 2 * k))
    m = s[:, g].shape[1]
    s[0, g] = np.where(np.arange(m) % 2 == 0, 0.72, -0.72)      # ball near a goal mouth, moving in
    s[1, g] = rng.uniform(-0.15, 0.15, m)
    s[2, g] = np.where(np.arange(m) % 2 == 0, 1.0, -1.0) * rng.uniform(0.3, 1.5, m)
    s[3, g] = rng.uniform(-0.2, 0.2, m)
    c = slice(min(n, 2 * k), min(n, 3 * k))                      # robots packed around the ball
    m = s[:, c].shape[1]
    for r in range(6):
        s[4 + 9 * r, c] = s[0, c] + rng.uniform(-0.09, 0.09, m)
        s[5 + 9 * r, c] = s[1, c] + rng.uniform(-0.09, 0.09, m)
        s[6 + 9 * r, c] = rng.uniform(-1, 1, m)
        s[7 + 9 * r, c] = rng.uniform(-1, 1, m)
    w = slice(min(n, 3 * k), min(n, 4 * k))                      # robots and ball against the walls
    m = s[:, w].shape[1]
    s[0, w] = rng.uniform(-0.7, 0.7, m); s[1, w] = np.where(np.arange(m) % 2 == 0, 0.62, -0.62)
    s[3, w] = np.where(np.arange(m) % 2 == 0, 1.0, -1.0)
    for r in range(6):
        s[4 + 9 * r, w] = np.where(np.arange(m) % 2 == 0, 0.72, -0.72) * rng.uniform(0.9, 1.0, m)
        s[5 + 9 * r, w] = rng.uniform(-0.6, 0.6, m)
        s[6 + 9 * r, w] = np.where(np.arange(m) % 2 == 0, 1.0, -1.0)
    be.set_state(s)


def compare_full_step(out, ref, rb, rb_ref, n, what, exact_physics):
    """out/ref: dicts of a backend step and an oracle step from the SAME input state."""
    if exact_physics:
        assert np.array_equal(rb, rb_ref), what + " reset_buf"
        assert np.array_equal(out["timeout"], ref["timeout"]), what + " timeout"
        assert np.array_equal(out["progress_f"], ref["progress_f"]), what + " progress"
        assert_bits_equal(out["rew"], ref["rew"], what + " rew")
        assert_obs_equal(out["term_obs"], ref["term_obs"], what + " term_obs")
        keep = rb == 0
        assert_obs_equal(out["obs"][keep], ref["obs"][keep], what + " obs (kept)")
        assert_obs_equal(out["obs"][~keep], ref["obs"][~keep], what + " obs (reset)", trig_atol=1e-6)
        return 0
    assert np.array_equal(out["progress_f"], ref["progress_f"]), what + " progress"
    tol = lambda a, b: np.abs(a - b) <= PHYS_ATOL + PHYS_RTOL * np.abs(b)
    ok_t = tol(out["term_obs"], ref["term_obs"]).reshape(n, -1).all(1)
    # rewards are differences of positions: same tolerance scaled by the largest weight (10 for goal is
    # integer-valued; grad/move weights 2 and 3)
    ok_r = (np.abs(out["rew"] - ref["rew"]) <= 3 * (PHYS_ATOL + PHYS_RTOL * np.abs(ref["rew"]))).reshape(n, -1).all(1)
    same_done = rb == rb_ref
    good = ok_t & ok_r & same_done & (out["timeout"] == ref["timeout"])
    # fields that agree on the reset decision and were reset must agree on the new state
    both = good & (rb != 0)
    if both.any():
        assert_obs_equal(out["obs"][both], ref["obs"][both], what + " obs after reset", trig_atol=1e-6)
    kept = good & (rb == 0)
    assert np.array_equal(bits(out["obs"][kept]), bits(out["term_obs"][kept])), what + " obs == term_obs when kept"
    return int((~good).sum())


def check_rollout(Backend, n=777, steps=40, seed=7, goff=1 << 33, exact_after=False, **params):
    """Random-action rollout; every step the oracle is re-seeded with the backend's state so
    single-step errors do not compound (chaotic contacts), and outputs are compared."""
    be, p = make_backend_pair(Backend, n, seed, goff, **params)
    rng = np.random.default_rng(seed)
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0  # as if one step had passed: otherwise pre_physics_step zeroes every progress counter
    stage_interesting_state(be, rng)
    flips = dones = timeouts = goals = 0
    for t in range(steps):
        st = oracle_from_backend(be)
        rb_ref = rb.copy()
        actions = rng.uniform(-1.3, 1.3, (n, 2, 3, 2)).astype(np.float32)  # exercises the +-1 clamp
        out = be.step(actions, rb)
        ref = orc.step(p, seed, goff, st, actions, rb_ref)
        flips += compare_full_step(out, ref, rb, rb_ref, n, f"step {t}", exact_physics=False)
        dones += int(rb.sum()); timeouts += int(out["timeout"].sum())
        goals += int((np.abs(out["rew"][:, 0, 0, 0]) > 0).sum())
        # invariants: nothing leaves the field box, unit headings
        got = oracle_from_backend(be)
        assert np.abs(got.ball_pos[:, 0]).max() <= 0.85 + 1e-5 and np.abs(got.ball_pos[:, 1]).max() <= 0.65 + 1e-5
        assert np.abs(got.r_pos[..., 0]).max() <= 0.86 and np.abs(got.r_pos[..., 1]).max() <= 0.66
        np.testing.assert_allclose(np.linalg.norm(got.r_rot, axis=-1), 1.0, atol=2e-6)
    assert dones > 0 and timeouts > 0 and goals > 0, (dones, timeouts, goals)
    assert flips <= max(2, FLIP_FRACTION * n * steps), f"{flips} field-steps outside the physics tolerance"
    return dict(flips=flips, dones=dones, timeouts=timeouts, goals=goals)


def random_post_state(rng, n, ld):
    s = np.zeros((58, ld), np.float32)
    s[0, :n] = rng.uniform(-0.85, 0.85, n); s[1, :n] = rng.uniform(-0.65, 0.65, n)
    s[2:4, :n] = rng.uniform(-1.5, 1.5, (2, n))
    for r in range(6):
        b = 4 + 9 * r
        s[b, :n] = rng.uniform(-0.85, 0.85, n); s[b + 1, :n] = rng.uniform(-0.65, 0.65, n)
        s[b + 2:b + 4, :n] = rng.uniform(-1.2, 1.2, (2, n))
        yaw = rng.uniform(-np.pi, np.pi, n)
        s[b + 4, :n] = np.cos(yaw); s[b + 5, :n] = np.sin(yaw)
        s[b + 6, :n] = rng.uniform(-30, 30, n)
    return s


def check_injected(Backend, n=500, steps=6, seed=99, goff=12345, **params):
    """Identical input states: physics replaced by an injected post-physics state on both sides.
    Everything (rewards, dones, timeouts, progress, obs, terminal obs, masked reset) must agree."""
    be, p = make_backend_pair(Backend, n, seed, goff, **params)
    rng = np.random.default_rng(seed)
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0
    rb[::7] = 1  # some fields still flagged from "the previous step": their progress restarts at 0
    s = be.get_state()
    s.view(np.int32)[58, :n] = rng.integers(380, 400, n)  # many timeouts
    be.set_state(s)
    ndone = 0
    for t in range(steps):
        st = oracle_from_backend(be)
        rb_ref = rb.copy()
        actions = rng.uniform(-1.3, 1.3, (n, 2, 3, 2)).astype(np.float32)
        post = random_post_state(rng, n, be.ld)
        post[0, :8] = [0.8, -0.8, 0.75, 0.76, -0.7500001, 0.0, 0.8, -0.76]   # goal edge cases
        post[1, :8] = [0.05, -0.19, 0.0, 0.2, 0.1, 0.0, 0.3, 0.199]
        out = be.step(actions, rb, post_state=post)
        ref = orc.step(p, seed, goff, st, actions, rb_ref, post_state=post)
        compare_full_step(out, ref, rb, rb_ref, n, f"injected step {t}", exact_physics=True)
        got, want = oracle_from_backend(be), st
        assert np.array_equal(got.progress, want.progress) and np.array_equal(got.episode, want.episode)
        ndone += int(rb.sum())
    assert ndone > 0
    return ndone


def check_views(Backend, view, n=333, steps=12, seed=5, goff=777):
    be, p = make_backend_pair(Backend, n, seed, goff)
    rng = np.random.default_rng(seed + view)
    nv = n * 3 if view == orc.VIEW_DMA else n
    adim = 6 if view == orc.VIEW_CMA else 2
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0
    stage_interesting_state(be, rng)
    abuf = np.zeros((n, 2, 3, 2), np.float32)
    ep_ret, ep_len = np.zeros((nv, 4), np.float32), np.zeros((nv,), np.int32)
    flips = 0
    for t in range(steps):
        st = oracle_from_backend(be)
        rb_ref, abuf_ref, er_ref, el_ref = rb.copy(), abuf.copy(), ep_ret.copy(), ep_len.copy()
        pa = rng.uniform(-1.2, 1.2, (nv, adim)).astype(np.float32)
        step_index = be.step_count
        out = be.step_view(view, pa, abuf, rb, ep_ret, ep_len)
        ref = orc.step_view(p, seed, goff, step_index, st, view, pa, abuf_ref, rb_ref, er_ref, el_ref)
        # the OU noise stream: same Philox counters, libm differences only
        same = rb == rb_ref
        np.testing.assert_allclose(abuf[same], abuf_ref[same], rtol=0, atol=2e-6, err_msg=f"action_buf step {t}")
        assert np.array_equal(out["progress"], ref["progress"])
        per = nv // n
        same_v = np.repeat(same, per)
        tol = lambda a, b: np.abs(a - b) <= 3 * (PHYS_ATOL + PHYS_RTOL * np.abs(b))
        good = same_v & tol(out["term_obs"], ref["term_obs"]).all(1) & tol(out["rews"], ref["rews"]).all(1)
        good &= tol(out["reward"], ref["reward"]) & (out["timeout"] == ref["timeout"]) & (out["done"] == ref["done"])
        good &= tol(out["ret_ret"], ref["ret_ret"]).all(1) & (out["ret_len"] == ref["ret_len"])
        flips += int((~good).sum())
        # internal consistency of the view outputs (exact)
        assert_bits_equal(out["reward"], ((out["rews"][:, 0] + out["rews"][:, 1]) + out["rews"][:, 2]) + out["rews"][:, 3],
                          "reward = rews.sum(-1)")
        assert np.array_equal(out["done"], np.repeat(rb, per))
        assert np.all(abuf[rb != 0] == 0)
        assert np.array_equal(ep_len, np.where(out["done"] != 0, 0, out["ret_len"]))
        # keep the two statistics streams together (tolerance-level differences would otherwise add up)
        ep_ret[...] = er_ref; ep_len[...] = el_ref; abuf[...] = abuf_ref
        if not same.all():
            break  # a flipped done decision: the states have diverged, stop comparing this run
    assert flips <= max(2, FLIP_FRACTION * nv * steps), f"{flips} view rows outside tolerance"
    return flips


def check_golden_injected(Backend, golden_path):
    """The engine's rewards/dones/obs against vectors produced by the REFERENCE's own jit functions
    (tests/golden/jit_functions.npz): prev state -> engine state, cur state -> injected post state."""
    g = np.load(golden_path)
    n = g["ball_pos"].shape[0]
    be, p = make_backend_pair(Backend, n, 1, 0)
    prev = orc.State(n)
    prev.ball_pos[...] = g["prev_ball_pos"]; prev.r_pos[...] = g["prev_r_pos"]
    # progress such that progress+1 equals the golden progress_buf (flagged fields restart from 0)
    prev.progress[...] = g["progress"] - 1
    rb = np.zeros(n, np.int64)
    restart = g["progress"] == 1
    rb[restart] = 1
    prev.progress[restart] = 123
    neg = g["progress"] < 1
    be.set_state(prev.to_soa(be.ld))
    cur = orc.State(n)
    cur.ball_pos[...] = g["ball_pos"]; cur.ball_vel[...] = g["ball_vel"]; cur.r_pos[...] = g["r_pos"]
    cur.r_vel[...] = g["r_vel"]; cur.r_rot[...] = g["r_rot"]; cur.r_w[...] = g["r_w"][..., 0]
    post = cur.to_soa(be.ld)[:58]
    out = be.step(g["acts"], rb, post_state=post)
    ok = ~neg  # progress 0 cannot be produced by a step (it is always >= 1 after the increment)
    assert np.array_equal(rb[ok], g["dones"][ok])
    assert np.array_equal(out["progress_f"][ok], g["progress"][ok].astype(np.float32))
    # rewards with the yaml weights (10, 2, 3, 0): vss.py:225-255
    np.testing.assert_array_equal(out["rew"][..., 0], g["goal_rew"].astype(np.float32) * 10.0)
    np.testing.assert_allclose(out["rew"][..., 1], g["grad_rew"] * 2.0, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out["rew"][..., 2], g["move_rew"] * 3.0, rtol=1e-5, atol=2e-6)
    assert np.all(out["rew"][..., 3] == 0)
    assert_obs_equal(out["term_obs"], g["obs"], "terminal obs vs reference compute_obs", trig_atol=1e-6)
    keep = rb == 0
    assert_obs_equal(out["obs"][keep], g["obs"][keep], "obs vs reference compute_obs", trig_atol=1e-6)
    # energy reward switched on (w_energy > 0 path, vss.py:253-255)
    be2, _ = make_backend_pair(Backend, n, 1, 0)
    be2.set_reward_weights(10.0, 2.0, 3.0, 0.5)
    be2.set_state(prev.to_soa(be2.ld))
    rb2 = np.zeros(n, np.int64)
    out2 = be2.step(g["acts"], rb2, post_state=post)
    np.testing.assert_allclose(out2["rew"][..., 3], g["energy_rew"] * 0.5, rtol=1e-6, atol=1e-7)
    return n


# --------------------------------------------------------------------------- robustness
def chase_actions(obs, noise):
    """All six robots steer at the ball (heavy contact: scrums, pushing along walls)."""
    rel = obs[..., 0:2] - obs[..., 4:6]
    c, s = obs[..., 8], obs[..., 9]
    fwd = rel[..., 0] * c + rel[..., 1] * s
    lat = -rel[..., 0] * s + rel[..., 1] * c
    turn = np.arctan2(lat, fwd)
    return (np.stack([1.0 - 1.5 * turn, 1.0 + 1.5 * turn], -1) + 0.2 * noise).astype(np.float32)


def check_degenerate_contact(Backend, golden_path):
    """Regression: ball centre exactly on a box surface used to give a 0/0 contact normal."""
    g = np.load(golden_path)
    n = 40
    be, p = make_backend_pair(Backend, n, 1, 0)
    s = np.zeros((60, be.ld), np.float32)
    s[:, :n] = g["state"][:, None]
    be.set_state(s)
    rb = np.zeros(n, np.int64)
    st = oracle_from_backend(be)
    rb_ref = rb.copy()
    acts = np.broadcast_to(g["actions"], (n, 2, 3, 2)).copy()
    out = be.step(acts, rb)
    ref = orc.step(p, 1, 0, st, acts, rb_ref)
    for k in ("obs", "term_obs", "rew"):
        assert np.isfinite(out[k]).all() and np.isfinite(ref[k]).all(), k
    assert np.isfinite(be.get_state()[:58, :n]).all()


def check_chase_stress(Backend, n=2048, steps=150, seed=7):
    """Contact-heavy rollout: every state word stays finite and inside the field box."""
    be, p = make_backend_pair(Backend, n, seed, 0)
    rb = np.ones(n, np.int64)
    obs = be.reset_dones(rb)
    rb[:] = 0
    rng = np.random.default_rng(3)
    for t in range(steps):
        out = be.step(chase_actions(obs, rng.uniform(-1.2, 1.2, (n, 2, 3, 2))), rb)
        obs = out["obs"]
        assert np.isfinite(out["obs"]).all() and np.isfinite(out["term_obs"]).all() and np.isfinite(out["rew"]).all(), t
    st = be.get_state()[:58, :n]
    assert np.isfinite(st).all()
    assert np.abs(st[0]).max() <= 0.85 + 1e-5 and np.abs(st[1]).max() <= 0.65 + 1e-5


def check_wall_stress(Backend, n=4096, steps=6, seed=11):
    """Every robot and the ball start within reach of a wall family (side walls, end-wall blocks,
    goal side walls, goal back wall, goal posts), moving into it: exercises the specialised
    point-vs-wall and post-vs-box code against the oracle's generic version, one control step at a
    time from identical states."""
    be, p = make_backend_pair(Backend, n, seed, 0)
    rng = np.random.default_rng(seed)
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0
    s = be.get_state()
    kind = np.arange(n) % 4
    sx = rng.choice([-1.0, 1.0], (7, n)); sy = rng.choice([-1.0, 1.0], (7, n))
    for e in range(7):  # entity 0 = ball, 1..6 = robots
        b = 0 if e == 0 else 4 + 9 * (e - 1)
        reach = 0.03 if e == 0 else 0.06
        # 0: side wall; 1: end wall outside the goal mouth; 2: goal mouth / posts / goal side walls; 3: inside the goal
        x = np.select([kind == 0, kind == 1, kind == 2, kind == 3],
                      [rng.uniform(-0.7, 0.7, n), 0.75 - rng.uniform(0, reach, n), 0.75 + rng.uniform(-reach, 0.02, n),
                       0.85 - rng.uniform(0, reach, n)])
        y = np.select([kind == 0, kind == 1, kind == 2, kind == 3],
                      [0.65 - rng.uniform(0, reach, n), rng.uniform(0.25, 0.6, n), 0.2 + rng.uniform(-reach, reach, n),
                       rng.uniform(0.0, 0.19 - 0.04, n)])
        s[b, :n] = sx[e] * x; s[b + 1, :n] = sy[e] * y
        s[b + 2, :n] = sx[e] * rng.uniform(-0.2, 1.2, n); s[b + 3, :n] = sy[e] * rng.uniform(-0.2, 1.2, n)
        if e:
            yaw = rng.uniform(-np.pi, np.pi, n)
            s[b + 4, :n] = np.cos(yaw); s[b + 5, :n] = np.sin(yaw); s[b + 6, :n] = rng.uniform(-20, 20, n)
    # keep the robots of one field apart (different wall stretches) so that this check is about walls:
    # robots 1..6 of a field get distinct signs / offsets along the wall where possible
    be.set_state(s)
    flips = 0
    for t in range(steps):
        st = oracle_from_backend(be)
        rb_ref = rb.copy()
        actions = rng.uniform(-1.0, 1.0, (n, 2, 3, 2)).astype(np.float32)
        out = be.step(actions, rb)
        ref = orc.step(p, seed, 0, st, actions, rb_ref)
        flips += compare_full_step(out, ref, rb, rb_ref, n, f"wall stress step {t}", exact_physics=False)
        got = oracle_from_backend(be)
        assert np.abs(got.ball_pos[:, 0]).max() <= 0.85 + 1e-5 and np.abs(got.ball_pos[:, 1]).max() <= 0.65 + 1e-5
    assert flips <= max(2, 2 * FLIP_FRACTION * n * steps), f"{flips} field-steps outside the physics tolerance"
    return flips


def check_pair_stress(Backend, n=4096, steps=6, seed=13):
    """Robots start in touching / overlapping pairs and triples (corner-in-box, wheel-in-box, box
    on box at every relative heading) with the ball wedged between some of them: exercises the
    robot-robot feature code and the ball-robot code against the oracle from identical states."""
    be, p = make_backend_pair(Backend, n, seed, 0)
    rng = np.random.default_rng(seed)
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0
    s = be.get_state()
    cx = rng.uniform(-0.5, 0.5, (3, n)); cy = rng.uniform(-0.45, 0.45, (3, n))   # three cluster centres per field
    for r in range(6):
        b = 4 + 9 * r
        g = r // 2 if True else 0
        d = rng.uniform(0.02, 0.06, n) * np.where(np.arange(n) % 3 == 0, 1.0, 0.75)
        ang = rng.uniform(-np.pi, np.pi, n)
        s[b, :n] = cx[g] + (d * np.cos(ang) if r % 2 else 0.0)
        s[b + 1, :n] = cy[g] + (d * np.sin(ang) if r % 2 else 0.0)
        s[b + 2, :n] = rng.uniform(-1.0, 1.0, n); s[b + 3, :n] = rng.uniform(-1.0, 1.0, n)
        yaw = rng.uniform(-np.pi, np.pi, n)
        s[b + 4, :n] = np.cos(yaw); s[b + 5, :n] = np.sin(yaw); s[b + 6, :n] = rng.uniform(-25, 25, n)
    # every fourth field: robots 4 and 5 join cluster 0 (triple contact)
    t3 = np.arange(n) % 4 == 0
    for r in (4, 5):
        b = 4 + 9 * r
        s[b, :n] = np.where(t3, cx[0] + rng.uniform(-0.08, 0.08, n), s[b, :n])
        s[b + 1, :n] = np.where(t3, cy[0] + rng.uniform(-0.08, 0.08, n), s[b + 1, :n])
    # the ball next to cluster 1
    s[0, :n] = cx[1] + rng.uniform(-0.07, 0.07, n); s[1, :n] = cy[1] + rng.uniform(-0.07, 0.07, n)
    s[2, :n] = rng.uniform(-1.5, 1.5, n); s[3, :n] = rng.uniform(-1.5, 1.5, n)
    be.set_state(s)
    flips = 0
    for t in range(steps):
        st = oracle_from_backend(be)
        rb_ref = rb.copy()
        actions = rng.uniform(-1.0, 1.0, (n, 2, 3, 2)).astype(np.float32)
        out = be.step(actions, rb)
        ref = orc.step(p, seed, 0, st, actions, rb_ref)
        flips += compare_full_step(out, ref, rb, rb_ref, n, f"pair stress step {t}", exact_physics=False)
    # deep initial overlaps are resolved through many near-degenerate branches: a looser flip budget
    assert flips <= max(2, 4 * FLIP_FRACTION * n * steps), f"{flips} field-steps outside the physics tolerance"
    return flips


def _place(be, mutate):
    """Reset everything, then overwrite the state through `mutate(State)` (robots parked far apart)."""
    n = be.n
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    st = oracle_from_backend(be)
    st.ball_pos[:] = (0.0, 0.55); st.ball_vel[:] = 0
    park = np.array([[-0.6, -0.5], [-0.3, -0.5], [0.0, -0.5], [0.3, -0.5], [0.6, -0.5], [0.6, 0.5]], np.float32)
    st.r_pos[:] = park.reshape(1, 2, 3, 2); st.r_vel[:] = 0; st.r_w[:] = 0; st.r_act[:] = 0
    st.r_rot[:] = (1.0, 0.0)
    st.progress[:] = 0
    mutate(st)
    be.set_state(st.to_soa(be.ld))
    return np.zeros(n, np.int64)


def check_physics_spec(Backend):
    """The new 2-D model against the numbers the reference's scene description implies (SURVEY App. B;
    vss_robot.urdf, vss.py:427-434): these are properties of the SPEC, not of PhysX's solver."""
    p = orc.default_params()
    r_w, track, wmax = 0.024, 0.0675, 42.0
    v_max, yaw_max = r_w * wmax, r_w * 2 * wmax / track          # 1.008 m/s, 29.87 rad/s
    # 1. straight drive: terminal speed r_w * 42 rad/s, reached within a few control steps, and the
    #    traction limit mu m g caps the acceleration (0.7 * 9.81 * 0.05 = 0.343 m/s per control step)
    def lane(st):
        st.r_pos[:, 0, 0] = (-0.6, 0.0)                          # a free lane along y = 0

    be = Backend(4, seed=1, goff=0, params=p)
    rb = _place(be, lane)
    act = np.zeros((4, 2, 3, 2), np.float32); act[:, 0, 0] = (1.0, 1.0)
    speeds = []
    for t in range(12):
        be.step(act, rb); rb[:] = 0
        speeds.append(float(np.linalg.norm(oracle_from_backend(be).r_vel[0, 0, 0])))
    assert abs(speeds[-1] - v_max) < 0.01 * v_max, speeds
    assert speeds[3] > 0.9 * v_max, speeds                       # "0 -> 1 m/s takes about 3 control steps"
    assert max(np.diff([0.0] + speeds)) <= 0.7 * 9.81 * 0.05 * 1.001, speeds
    st = oracle_from_backend(be)
    assert abs(st.r_vel[0, 0, 0, 1]) < 1e-6 and abs(st.r_w[0, 0, 0]) < 1e-6   # straight: no drift, no spin
    assert np.allclose(st.r_vel[0, 0, 1:], 0, atol=1e-7) and np.allclose(st.r_vel[0, 1], 0, atol=1e-7)
    # 2. spin in place: yaw rate r_w (w_r - w_l) / track; DOF 0 is the left wheel (vss.py / urdf)
    be = Backend(4, seed=1, goff=0, params=p)
    rb = _place(be, lane)
    act = np.zeros((4, 2, 3, 2), np.float32); act[:, 0, 0] = (-1.0, 1.0)
    for t in range(12):
        be.step(act, rb); rb[:] = 0
    st = oracle_from_backend(be)
    assert abs(st.r_w[0, 0, 0] - yaw_max) < 0.01 * yaw_max, st.r_w[0, 0, 0]
    assert np.linalg.norm(st.r_pos[0, 0, 0] - np.array([-0.6, 0.0])) < 1e-3            # turns on the spot
    assert abs(np.linalg.norm(st.r_rot[0, 0, 0]) - 1.0) < 1e-6
    # 3. free rolling ball: exponential decay with the rolling-sphere share (2/7) of PhysX's angular damping 0.5
    be = Backend(4, seed=1, goff=0, params=p)
    def roll(st):
        st.ball_pos[:] = (-0.5, 0.0); st.ball_vel[:] = (0.5, 0.0)
    rb = _place(be, roll)
    act = np.zeros((4, 2, 3, 2), np.float32)
    for t in range(20):                                          # 1 s
        be.step(act, rb); rb[:] = 0
    st = oracle_from_backend(be)
    assert abs(st.ball_vel[0, 0] - 0.5 * np.exp(-0.5 * 2 / 7)) < 1e-4 and abs(st.ball_vel[0, 1]) < 1e-7
    # 4. walls: restitution 0 (the ball keeps no normal velocity), nothing leaves the field, and the ball
    #    only passes the end line through the goal mouth |y| < 0.2 (vss.py:342-345) — which ends the episode
    be = Backend(4, seed=1, goff=0, params=p)
    def shoot(st):
        st.ball_pos[0] = (0.0, 0.6); st.ball_vel[0] = (0.0, 2.0)          # into the side wall
        st.ball_pos[1] = (0.7, 0.4); st.ball_vel[1] = (2.0, 0.0)          # into the end wall beside the goal
        st.ball_pos[2] = (0.7, 0.1); st.ball_vel[2] = (2.0, 0.0)          # into the yellow goal
        st.ball_pos[3] = (-0.7, -0.1); st.ball_vel[3] = (-2.0, 0.0)       # into the blue goal
    rb = _place(be, shoot)
    out = be.step(act, rb)
    st = oracle_from_backend(be)
    assert abs(st.ball_pos[0, 1]) <= 0.65 - 0.02134 + 1e-5 and abs(st.ball_vel[0, 1]) < 1e-6
    assert abs(st.ball_pos[1, 0]) <= 0.75 - 0.02134 + 1e-5 and abs(st.ball_vel[1, 0]) < 1e-6
    assert rb.tolist() == [0, 0, 1, 1]
    assert out["rew"][2, 0, 0, 0] == 10.0 and out["rew"][2, 1, 0, 0] == -10.0     # blue attacks +x (vss.py:589-594)
    assert out["rew"][3, 0, 0, 0] == -10.0 and out["rew"][3, 1, 0, 0] == 10.0
    # 5. a robot driving into the ball pushes it ahead (contact + Coulomb friction 0.5), without tunnelling
    be = Backend(4, seed=1, goff=0, params=p)
    def push(st):
        st.r_pos[:, 0, 0] = (-0.2, 0.0); st.ball_pos[:] = (-0.12, 0.0)
    rb = _place(be, push)
    act = np.zeros((4, 2, 3, 2), np.float32); act[:, 0, 0] = (1.0, 1.0)
    for t in range(10):
        be.step(act, rb); rb[:] = 0
        st = oracle_from_backend(be)
        assert st.ball_pos[0, 0] - st.r_pos[0, 0, 0, 0] >= 0.035 + 0.02134 - 1e-4, t   # ball stays in front of the face
    assert st.ball_vel[0, 0] > 0.8 and abs(st.ball_pos[0, 1]) < 1e-3


def check_nonfinite_guard(Backend, n=96):
    """Safety net: a field whose state is not finite is re-randomised on the spot, reported done with
    zero reward and no timeout; its neighbours are untouched; backend and oracle agree."""
    be, p = make_backend_pair(Backend, n, 3, 0)
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0
    s = be.get_state()
    poisoned = [0, 33, 95]
    s[0, poisoned[0]] = np.nan          # ball x
    s[4 + 9 * 2 + 6, poisoned[1]] = np.inf  # a yaw rate
    s[4 + 9 * 5 + 2, poisoned[2]] = -np.inf  # a velocity
    be.set_state(s)
    st = oracle_from_backend(be)
    rb_ref = rb.copy()
    acts = np.random.default_rng(0).uniform(-1, 1, (n, 2, 3, 2)).astype(np.float32)
    out = be.step(acts, rb)
    ref = orc.step(p, 3, 0, st, acts, rb_ref)
    for k in ("obs", "term_obs", "rew"):
        assert np.isfinite(out[k]).all(), k
    assert all(rb[i] == 1 for i in poisoned) and np.array_equal(rb, rb_ref)
    assert np.all(out["rew"][poisoned] == 0) and np.all(out["timeout"][poisoned] == 0)
    assert np.array_equal(bits(out["obs"][poisoned]), bits(out["term_obs"][poisoned]))   # fresh state in both
    compare_full_step(out, ref, rb, rb_ref, n, "non-finite guard", exact_physics=False)
    assert np.isfinite(be.get_state()[:58, :n]).all()
    if hasattr(be, "sanitised_count"):   # the event is counted (vss_sanitised_count), not silent
        assert be.sanitised_count == len(poisoned)


# --------------------------------------------------------------------------- RNG statistics
def philox4x32_10_np(ctr, key):
    """Vectorised Philox4x32-10 (Salmon et al. 2011): ctr (N,4) uint32, key (2,) -> (N,4) uint32. Written
    from the published algorithm, independent of the product's and the oracle's C versions (checked
    against the oracle's on a few counters by the callers)."""
    c = np.ascontiguousarray(ctr, np.uint32).astype(np.uint64)
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m0, m1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    x, y, z, w = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    for _ in range(10):
        p0, p1 = m0 * x, m1 * z
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        x, y, z, w = hi1 ^ y ^ k0, lo1, hi0 ^ w ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack([x, y, z, w], 1).astype(np.uint32)


def first_attempt_positions(p, seed, goff, n, episode=0):
    """The FIRST placement draw of reset_dones for fields goff .. goff+n-1 (envs/vss.py:283-287: 7 XY pairs
    ~ U(-0.5, 0.5) * field_scale), from the Philox stream DESIGN.md §5 documents, in numpy."""
    gid = np.uint64(goff) + np.arange(n, dtype=np.uint64)
    u = np.empty((n, 16), np.uint32)
    for b in range(4):
        ctr = np.stack([(gid & np.uint64(0xFFFFFFFF)).astype(np.uint32), (gid >> np.uint64(32)).astype(np.uint32),
                        np.full(n, episode, np.uint32), np.full(n, (0 << 28) | b, np.uint32)], 1)
        u[:, 4 * b:4 * b + 4] = philox4x32_10_np(ctr, (seed & 0xFFFFFFFF, seed >> 32))
    u01 = (u >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    px = (u01[:, 0:14:2] - np.float32(0.5)) * np.float32(p.reset_scale_x)
    py = (u01[:, 1:14:2] - np.float32(0.5)) * np.float32(p.reset_scale_y)
    return np.stack([px, py], -1)  # (n, 7, 2): ball, then robots in team-major order


def check_reset_reject_rate(Backend, n=60000, seed=21, goff=3 << 32):
    """envs/vss.py:281-299 redraws all 7 entities while any pair is closer than 0.07 m; SURVEY App. D
    measured P(reject) = 0.1799 per draw on the reference's distribution. Here: the share of fields whose
    final placement is NOT their first draw, and the first draws themselves must be what the rule says."""
    ctr = np.array([1, 2, 3, 4], np.uint32)
    assert np.array_equal(philox4x32_10_np(ctr[None], (7, 9))[0], orc.philox4x32_10(ctr, np.array([7, 9], np.uint32)))
    be, p = make_backend_pair(Backend, n, seed, goff)
    be.reset_dones(np.ones(n, np.int64))
    got = oracle_from_backend(be)
    ent = np.concatenate([got.ball_pos[:, None, :], got.r_pos.reshape(n, 6, 2)], 1)
    first = first_attempt_positions(p, seed, goff, n)
    d = np.linalg.norm(first[:, :, None, :].astype(np.float64) - first[:, None, :, :], axis=-1) + np.eye(7) * 10
    first_ok = d.min((1, 2)) >= 0.07 + 1e-6          # clearly accepted by the rule
    first_bad = d.min((1, 2)) < 0.07 - 1e-6          # clearly rejected
    kept_first = (bits(ent) == bits(first)).reshape(n, -1).all(1)
    assert kept_first[first_ok].all(), "an acceptable first draw was redrawn"
    assert not kept_first[first_bad].any(), "a first draw with a pair closer than 0.07 m was kept"
    rate = 1.0 - kept_first.mean()
    assert abs(rate - 0.18) < 0.01, rate
    return rate


def check_ou_moments(Backend, view=orc.VIEW_SA, n=16384, steps=6, seed=17):
    """The in-kernel opponent noise (envs/wrappers.py:5-19: prev - 0.1 prev + N(0, 0.15), clamp +-1):
    innovation mean 0, std 0.15, Gaussian 4th moment, uncorrelated between steps and between slots."""
    be, p = make_backend_pair(Backend, n, seed, 0)
    nv = n * 3 if view == orc.VIEW_DMA else n
    adim = 6 if view == orc.VIEW_CMA else 2
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0
    abuf = np.zeros((n, 2, 3, 2), np.float32)
    free = np.ones((2, 3, 2), bool)                  # slots the policy does not overwrite
    free[0, 0 if view == orc.VIEW_SA else slice(None)] = False
    innov = []
    for t in range(steps):
        prev = abuf.copy()
        pa = np.zeros((nv, adim), np.float32)
        be.step_view(view, pa, abuf, rb)
        live = rb == 0                               # rows of done fields are zeroed after the step
        x = abuf[live][:, free] - np.float32(0.9) * prev[live][:, free]
        unclamped = np.abs(abuf[live][:, free]) < 1.0
        innov.append(np.where(unclamped, x, np.nan))
        assert np.all(abuf[~live] == 0)
        assert np.all(abuf[live][:, ~free] == 0)     # the policy action (zeros here) landed in its slots
    m = min(len(v) for v in innov)
    z = np.stack([v[:m] for v in innov])             # (steps, fields, free slots)
    flat = z[np.isfinite(z)].astype(np.float64)
    k = flat.size
    assert k > 0.99 * z.size                         # hardly anything reaches the clamp from zero in a few steps
    assert abs(flat.mean()) < 5 * 0.15 / np.sqrt(k), flat.mean()
    assert abs(flat.std() - 0.15) < 5 * 0.15 / np.sqrt(2 * k) + 1e-4, flat.std()
    assert abs(np.mean((flat / 0.15) ** 4) - 3.0) < 5 * np.sqrt(96.0 / k) + 1e-2
    zz = np.nan_to_num(z) / 0.15
    lag = np.mean(zz[1:] * zz[:-1])                  # consecutive steps, same slot
    cross = np.mean(zz[:, :, :-1] * zz[:, :, 1:])    # neighbouring slots, same step
    assert abs(lag) < 5 / np.sqrt(zz[1:].size) and abs(cross) < 5 / np.sqrt(zz[:, :, 1:].size), (lag, cross)
    return dict(mean=flat.mean(), std=flat.std(), samples=k)


# --------------------------------------------------------------------------- packed host rows
def unpack_rows(pk):
    """(obs f32 (n,52) from bf16, reward f32 (n), done u8 (n), timeout u8 (n), pad u16 (n)) of packed rows."""
    pk = np.ascontiguousarray(pk, np.uint8)
    n = pk.shape[0]
    obs16 = pk[:, :104].copy().view(np.uint16).reshape(n, 52)
    obs = (obs16.astype(np.uint32) << 16).view(np.float32)
    reward = pk[:, 104:108].copy().view(np.float32).reshape(n)
    return obs, reward, pk[:, 108], pk[:, 109], pk[:, 110:112].copy().view(np.uint16).reshape(n)


def bf16_rne(x):
    """float32 -> bf16 (round to nearest even) -> float32, in numpy."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    return r.view(np.float32)


def check_packed_rows(Backend, view, n=500, steps=5, seed=23):
    """vss_set_step_packed: the 112-byte rows carry exactly bf16(obs), reward, done, timeout of the same
    launch's ordinary outputs (rows of reset fields hold the post-reset observation, like obs)."""
    be, p = make_backend_pair(Backend, n, seed, 0)
    rng = np.random.default_rng(seed)
    nv = n * 3 if view == orc.VIEW_DMA else n
    adim = 6 if view == orc.VIEW_CMA else 2
    rb = np.ones(n, np.int64)
    be.reset_dones(rb)
    rb[:] = 0
    stage_interesting_state(be, rng)
    abuf = np.zeros((n, 2, 3, 2), np.float32)
    ndone = 0
    for t in range(steps):
        pa = rng.uniform(-1.2, 1.2, (nv, adim)).astype(np.float32)
        out = be.step_view(view, pa, abuf, rb, packed=True)
        obs, reward, done, tmo, pad = unpack_rows(out["packed"])
        assert_bits_equal(obs, bf16_rne(out["obs"]), f"packed obs step {t}")
        assert_bits_equal(reward, out["reward"], f"packed reward step {t}")
        assert np.array_equal(done, out["done"].astype(np.uint8)) and np.array_equal(tmo, out["timeout"])
        assert np.all(pad == 0)
        ndone += int(done.sum())
    assert ndone > 0
    return ndone
