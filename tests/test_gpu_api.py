"""GPU: the reference-facing Python surface (VSS, views, make_env, PPO loop) over libvss_b200.so."""
import types

import numpy as np
import pytest

import parity_checks as pc
from oracle import vss_oracle as orc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _cfg(n):
    from rsoccer_isaac_cleanrl_b200.envs import load_cfg
    cfg = load_cfg()
    cfg["env"]["numEnvs"] = n
    return cfg


def _oracle_state(envs):
    return orc.State.from_soa(envs.engine.get_state().cpu().numpy(), envs.num_fields)


def test_vss_task_contract_and_oracle_parity():
    from rsoccer_isaac_cleanrl_b200.envs import VSS
    n = 300
    envs = VSS(_cfg(n), "cuda:0", "cuda:0", 0, True, False, False, seed=3, global_env_offset=40)
    assert envs.num_envs == n and envs.num_obs == 52 and envs.num_actions == 2
    assert envs.action_space.shape == (2, 3, 2) and envs.observation_space.shape == (2, 3, 52)
    assert envs.reset_buf.dtype == torch.int64 and bool((envs.reset_buf == 1).all())
    assert envs.dof_velocity_buf.shape == (n, 2, 3, 2)
    o = envs.reset()
    assert set(o) == {"obs"} and o["obs"].shape == (n, 2, 3, 52)
    p = orc.default_params()
    rng = np.random.default_rng(0)
    for t in range(6):
        st = _oracle_state(envs)
        rb_ref = envs.reset_buf.cpu().numpy().copy()
        a = torch.from_numpy(rng.uniform(-1.2, 1.2, (n, 2, 3, 2)).astype(np.float32)).cuda()
        obs, rew, reset, extras = envs.step(a)
        assert rew.shape == (n, 2, 3, 4) and reset.dtype == torch.int64 and reset.data_ptr() == envs.reset_buf.data_ptr()
        assert extras["time_outs"].dtype == torch.bool and extras["progress_buffer"].dtype == torch.float32
        assert extras["terminal_observation"].shape == (n, 2, 3, 52)
        ref = orc.step(p, 3, 40, st, a.cpu().numpy(), rb_ref)
        out = dict(obs=obs["obs"].cpu().numpy(), term_obs=extras["terminal_observation"].cpu().numpy(),
                   rew=rew.cpu().numpy(), timeout=extras["time_outs"].cpu().numpy().astype(np.uint8),
                   progress_f=extras["progress_buffer"].cpu().numpy())
        assert pc.compare_full_step(out, ref, reset.cpu().numpy(), rb_ref, n, f"api step {t}", False) <= 1
    # play.py:132-133 pattern: force a full reset through the public buffer
    before = envs.ball_pos.clone()
    envs.reset_buf[:] = 1
    envs.reset_dones()
    assert not torch.equal(before, envs.ball_pos)
    assert bool((envs.dof_velocity_buf == 0).all())
    # reward weights are live attributes (ppo…:389-392)
    envs.w_goal, envs.w_grad, envs.w_move, envs.w_energy = 1.0, 0.0, 0.0, 0.0
    _, rew, _, _ = envs.step(torch.zeros((n, 2, 3, 2), device="cuda"))
    assert bool((rew[..., 1:] == 0).all())


@pytest.mark.parametrize("env_id", ["sa", "cma", "dma"])
def test_views_python_surface(env_id):
    from rsoccer_isaac_cleanrl_b200.envs import RecordEpisodeStatisticsTorch, make_env
    from rsoccer_isaac_cleanrl_b200.ppo import ExtractObsWrapper
    n_agents = 192 * 3 if env_id == "dma" else 192
    args = types.SimpleNamespace(cuda=True, num_envs=n_agents, env_id=env_id, capture_video=False, seed=9)
    raw, view = make_env(args)
    fields = 192
    assert raw.num_fields == fields
    envs = RecordEpisodeStatisticsTorch(ExtractObsWrapper(view), torch.device("cuda:0"))
    adim = 6 if env_id == "cma" else 2
    assert envs.action_space.shape == (adim,) and envs.observation_space.shape == (52,)
    assert envs.num_envs == n_agents
    obs = envs.reset()
    assert obs.shape == (n_agents, 52)
    p = orc.default_params()
    vid = {"sa": orc.VIEW_SA, "cma": orc.VIEW_CMA, "dma": orc.VIEW_DMA}[env_id]
    abuf = np.zeros((fields, 2, 3, 2), np.float32)
    er, el = np.zeros((n_agents, 4), np.float32), np.zeros(n_agents, np.int32)
    rng = np.random.default_rng(1)
    for t in range(5):
        st = _oracle_state(raw)
        rb = raw.reset_buf.cpu().numpy().copy()
        act = rng.uniform(-1, 1, (n_agents, adim)).astype(np.float32)
        step_index = raw.engine.step_count
        obs, reward, done, info = envs.step(torch.from_numpy(act).cuda())
        ref = orc.step_view(p, 9, 0, step_index, st, vid, act, abuf, rb, er, el)
        assert obs.shape == (n_agents, 52) and reward.shape == (n_agents,) and done.shape == (n_agents,)
        assert info["terminal_observation"].shape == (n_agents, 52) and info["rews"].shape == (n_agents, 4)
        assert info["time_outs"].shape == (n_agents,) and info["progress_buffer"].shape == (n_agents,)
        same = np.repeat(raw.reset_buf.cpu().numpy() == rb, n_agents // fields)
        tol = lambda a, b: np.abs(a - b) <= 3 * (pc.PHYS_ATOL + pc.PHYS_RTOL * np.abs(b))
        good = same & tol(info["terminal_observation"].cpu().numpy(), ref["term_obs"]).all(1)
        good &= tol(reward.cpu().numpy(), ref["reward"]) & (done.cpu().numpy() == ref["done"])
        good &= tol(info["r"]["return"].cpu().numpy(), ref["ret_ret"].sum(1)) & (info["l"].cpu().numpy() == ref["ret_len"])
        assert (~good).sum() <= 2, (t, int((~good).sum()))
        np.testing.assert_allclose(view.action_buf.cpu().numpy()[same[::n_agents // fields]],
                                   abuf[same[::n_agents // fields]], atol=2e-6)
        # keep both sides on the engine's stream of states
        abuf[...] = view.action_buf.cpu().numpy()
        er[...] = view.episode_returns.cpu().numpy(); el[...] = view.episode_lengths.cpu().numpy()
    # the raw task's reset() still returns a current observation after fused view steps
    full = raw.reset()["obs"]
    want = obs if env_id != "dma" else obs.view(fields, 3, 52)
    got = full[:, 0, 0, :] if env_id != "dma" else full[:, 0, :, :]
    assert torch.equal(got, want)


def test_ppo_short_training_run(tmp_path):
    from rsoccer_isaac_cleanrl_b200 import ppo
    args = ppo.parse_args(["--env-id", "sa", "--num-envs", "256", "--num-steps", "16", "--total-timesteps",
                           str(256 * 16 * 3), "--update-epochs", "2", "--quiet", "--seed", "2"])
    stats = ppo.train(args)
    assert stats["updates"] == 3 and stats["global_step"] == 256 * 16 * 3 and stats["final_sps"] > 0
    sd = stats["agent"].state_dict()
    assert "critic.0.weight" in sd and "actor_mean.8.bias" in sd and "actor_logstd" in sd
    assert all(torch.isfinite(v).all() for v in sd.values())


def test_play_matches_with_team_policies(tmp_path):
    """play.py surface: teams fill the (N,2,3,2) buffer, the yellow side sees mirrored observations,
    a checkpoint saved by the PPO loop loads through TeamSA / TeamDMA."""
    from rsoccer_isaac_cleanrl_b200 import play, ppo
    from rsoccer_isaac_cleanrl_b200.envs import VSS
    envs = VSS(_cfg(1065), "cuda:0", "cuda:0", 0, True, seed=4)
    envs.w_goal, envs.w_grad, envs.w_move, envs.w_energy = 1.0, 0.0, 0.0, 0.0   # ppo…:389-392
    score, length = play.play_matches(envs, play.get_team("ou"), play.get_team("zero"), 300)
    assert -1.0 <= score <= 1.0 and 1.0 <= length <= 400.0
    # a (barely trained) checkpoint in the reference's state_dict format
    args = ppo.parse_args(["--env-id", "sa", "--num-envs", "128", "--num-steps", "8", "--total-timesteps",
                           str(128 * 8), "--quiet"])
    st = ppo.train(args)
    path = str(tmp_path / "agent.pt")
    torch.save(st["agent"].state_dict(), path)
    blue = play.get_team("ppo-sa", path)
    yellow = play.get_team("ppo-sa-x3", path)
    score, length = play.play_matches(envs, blue, yellow, 200)
    assert -1.0 <= score <= 1.0 and 1.0 <= length <= 400.0
    assert set(play.baseline_teams(str(tmp_path))) == {"zero", "ou"}   # no base_nets checkpoints present


def test_cuda_graph_replay_advances_the_ou_stream():
    """A captured vss_step_view launch must not freeze the OU-noise counter: the step index lives on
    the device and is bumped by a stream-ordered kernel, so every replay draws new noise."""
    from rsoccer_isaac_cleanrl_b200.envs import VSS, SingleAgent
    envs = VSS(_cfg(256), "cuda:0", "cuda:0", 0, True, seed=11)
    view = SingleAgent(envs)
    act = torch.zeros((256, 2), device="cuda")
    view.step(act)                       # eager warm-up
    torch.cuda.synchronize()
    assert envs.engine.step_count == 1
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        view.step(act)
    bufs = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        bufs.append(view.action_buf.clone())
    assert envs.engine.step_count == 4
    assert not torch.equal(bufs[0], bufs[1]) and not torch.equal(bufs[1], bufs[2])
    # and the replayed sequence equals the eager sequence from the same start
    envs2 = VSS(_cfg(256), "cuda:0", "cuda:0", 0, True, seed=11)
    view2 = SingleAgent(envs2)
    for _ in range(4):
        view2.step(act)
    torch.cuda.synchronize()
    assert torch.equal(view2.action_buf, bufs[2])
    assert torch.equal(envs2.engine.get_state(), envs.engine.get_state())


@pytest.mark.parametrize("env_id,n", [("sa", 333), ("cma", 97), ("dma", 1000)])
def test_view_side_outputs_bf16_obs_and_float_flags(env_id, n):
    """vss_set_step_aux: the bf16 / padded copy of the view observation and the float copies of done and
    timeout equal what the conversion launches they replace would produce, for kept and reset fields."""
    from rsoccer_isaac_cleanrl_b200.envs import CMA, DMA, VSS, SingleAgent
    envs = VSS(_cfg(n), "cuda:0", "cuda:0", 0, True, seed=21)
    view = {"sa": SingleAgent, "cma": CMA, "dma": DMA}[env_id](envs)
    nv = view.num_view_envs
    # push many fields to the end of their episodes so that resets happen within the few steps
    st = envs.engine.get_state()
    st[58, :n] = torch.randint(394, 399, (n,), device="cuda", dtype=torch.int32).view(torch.float32)
    envs.engine.set_state(st)
    x16 = torch.full((nv, 64), 7.0, device="cuda", dtype=torch.bfloat16)
    done_f = torch.full((nv,), -1.0, device="cuda")
    tmo_f = torch.full((nv,), -1.0, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    seen_done = 0
    for t in range(8):
        act = torch.rand((nv, view.ACT_DIM), device="cuda", generator=g) * 2 - 1
        obs, reward, done, info = view.step(act, obs_bf16_out=x16, done_f_out=done_f, timeout_f_out=tmo_f)
        torch.cuda.synchronize()
        assert torch.equal(x16[:, :52], obs["obs"].to(torch.bfloat16))
        assert torch.all(x16[:, 52:] == 7.0)          # the padding columns are never written
        assert torch.equal(done_f, done.float()) and torch.equal(tmo_f, info["time_outs"].float())
        seen_done += int(done.sum())
    assert seen_done > 0
    # and switching the side outputs off again leaves the buffers alone
    x16.fill_(3.0)
    view.step(act)
    torch.cuda.synchronize()
    assert torch.all(x16 == 3.0)


@pytest.mark.parametrize("env_id", ["sa", "dma"])
def test_step_host_pipelined_over_field_ranges_is_bit_identical(env_id):
    """`step_host` splits a large batch into field ranges (vss_set_step_range) on two streams; the
    results and the engine state must equal the single-launch step bit for bit (same kernels, RNG
    keyed by global field id and by the whole-engine step index)."""
    from rsoccer_isaac_cleanrl_b200.envs import DMA, VSS, SingleAgent
    n = 5000                                   # ragged: the last range ends at num_envs
    cls = {"sa": SingleAgent, "dma": DMA}[env_id]
    a, b = (cls(VSS(_cfg(n), "cuda:0", "cuda:0", 0, True, seed=31)) for _ in range(2))
    a.HOST_CHUNKS, a.HOST_CHUNK_MIN_FIELDS = 5, 1          # force the pipelined path on the small batch
    b.HOST_CHUNK_MIN_FIELDS = 1 << 30                      # single launch
    for v in (a, b):                                       # some episodes end within the test
        st = v.task.engine.get_state()
        st[58, :n] = (torch.arange(n, device="cuda", dtype=torch.int32) % 7 + 392).view(torch.float32)
        v.task.engine.set_state(st)
    g = torch.Generator(); g.manual_seed(5)
    for t in range(10):
        act = (torch.rand((a.num_view_envs, a.ACT_DIM), generator=g) * 2 - 1).pin_memory()
        oa, ra, da = a.step_host(act)
        ob, rb, db = b.step_host(act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), t
    assert int(da.sum()) >= 0 and torch.equal(a.task.engine.get_state(), b.task.engine.get_state())
    assert torch.equal(a.action_buf, b.action_buf)
    assert a.task.engine.step_count == b.task.engine.step_count == 10
    # a misaligned range is refused
    with pytest.raises(RuntimeError):
        a.task.engine.set_step_range(1, a.task.engine.step_granularity)
        a.step(torch.zeros((a.num_view_envs, a.ACT_DIM), device="cuda"))
    a.task.engine.set_step_range(0, 0)


@pytest.mark.parametrize("n", [100, 3000, 9000, 40000])
def test_outputs_stay_inside_their_buffers(n):
    """Every output of the view step (incl. the side outputs) is a slice of a larger buffer filled with a
    sentinel: the kernel must write every row of the slice and nothing outside it. Sizes cover 8 / 16 /
    32 fields per warp and ragged last tiles."""
    from rsoccer_isaac_cleanrl_b200.envs import DMA, VSS, SingleAgent
    PAD, SENT = 300, -12345.0
    for cls in (SingleAgent, DMA):
        view = cls(VSS(_cfg(n), "cuda:0", "cuda:0", 0, True, seed=n))
        nv = view.num_view_envs

        def guarded(cols, dtype=torch.float32, sent=SENT):
            big = torch.full(((nv + 2 * PAD) * cols,), sent, device="cuda", dtype=dtype)
            return big, big[PAD * cols:(PAD + nv) * cols].view((nv, cols) if cols > 1 else (nv,))

        bufs = {k: guarded(c) for k, c in (("obs", 52), ("term", 52), ("rew", 1), ("done_f", 1), ("tmo_f", 1))}
        bufs["x16"] = guarded(64, torch.bfloat16)
        act = torch.rand((nv, view.ACT_DIM), device="cuda") * 2 - 1
        view.step(act, obs_out=bufs["obs"][1], term_obs_out=bufs["term"][1], reward_out=bufs["rew"][1],
                  obs_bf16_out=bufs["x16"][1], done_f_out=bufs["done_f"][1], timeout_f_out=bufs["tmo_f"][1])
        torch.cuda.synchronize()
        for k, (big, sl) in bufs.items():
            cols = sl.shape[1] if sl.dim() == 2 else 1
            assert torch.all(big[:PAD * cols] == SENT) and torch.all(big[(PAD + nv) * cols:] == SENT), (k, "outside")
            inner = sl[:, :52] if k == "x16" else sl
            assert not torch.any(inner == SENT), (k, "unwritten rows")
        assert torch.all(bufs["x16"][1][:, 52:] == SENT)


def test_terminal_values_come_from_the_next_step_or_the_compacted_critic_pass(tmp_path):
    """ppo…:272 `next_values[step] = agent.get_value(info["terminal_observation"])`: the rollout only runs the critic on
    the rows that ended an episode (and on the observation after the last step) and copies values[t + 1] elsewhere.
    With a zero learning rate (weights unchanged by the update) the result must equal, bit for bit, the critic applied to
    EVERY terminal observation — in the eager first update and in the captured-graph replays. Episodes of 20 steps: every
    env ends at least once per 24-step rollout, some on its last step."""
    import yaml
    from rsoccer_isaac_cleanrl_b200 import ppo
    from rsoccer_isaac_cleanrl_b200.engine import gather_pad_bf16, mlp_forward_fused
    from rsoccer_isaac_cleanrl_b200.envs import load_cfg
    from rsoccer_isaac_cleanrl_b200.tc_mlp import MlpWeights
    T, N = 24, 700
    cfg = load_cfg()
    cfg["env"]["maxEpisodeLength"] = 20
    path = tmp_path / "vss_short.yaml"
    path.write_text(yaml.safe_dump(cfg))
    args = ppo.parse_args(["--env-id", "sa", "--num-envs", str(N), "--num-steps", str(T), "--total-timesteps",
                           str(T * N * 4), "--update-epochs", "1", "--quiet", "--seed", "3", "--learning-rate", "0"])
    args.cfg_path = str(path)
    seen = []

    def hook(update, t):
        mw = MlpWeights(t["agent"].critic)
        x16 = gather_pad_bf16(t["term_obs"].view(T * N, -1), None, 64)
        full, = mlp_forward_fused(x16, [(mw.w16, [b.detach() for b in mw.bs], mw.head_w.detach(), mw.head_b.detach(), None)])
        nv, v, d = t["next_values"], t["values"], t["next_dones"]
        assert torch.equal(nv.view(-1), full.view(-1)), (update, int((nv.view(-1) != full.view(-1)).sum()))
        assert torch.equal(nv[:-1][d[:-1] == 0], v[1:][d[:-1] == 0])   # the shifted copy where no episode ended
        seen.append((int(d.sum()), int(d[-1].sum())))

    stats = ppo.train(args, hook=hook)
    assert stats["updates"] == 4 and len(seen) == 4
    assert all(n_done >= N for n_done, _ in seen) and sum(last for _, last in seen) > 0


def test_compact_nonzero_and_scatter_rows():
    """include/vss_b200.h vss_compact_nonzero / vss_scatter_rows_f32 against torch.nonzero, below and above capacity."""
    from rsoccer_isaac_cleanrl_b200.engine import compact_nonzero, scatter_rows
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 100003
    flags = (torch.rand(n, device="cuda", generator=g) < 0.01).float()
    want = torch.nonzero(flags).view(-1)
    for cap in (4096, 128):
        lst = torch.full((cap,), -1, device="cuda", dtype=torch.int64)
        cnt = torch.full((1,), 99, device="cuda", dtype=torch.int32)
        compact_nonzero(flags, lst, cnt)
        assert int(cnt.item()) == want.numel()
        k = min(cap, want.numel())
        got = lst[:k]
        if cap >= want.numel():
            assert torch.equal(torch.sort(got).values, want) and bool((lst[k:] == -1).all())
        else:
            assert bool(flags[got].bool().all()) and got.unique().numel() == k
        dst = torch.zeros(n, device="cuda")
        src = torch.arange(1, cap + 1, device="cuda", dtype=torch.float32)
        lst_ok = lst.clamp(min=0)
        scatter_rows(dst, lst_ok, src, cnt)
        assert int((dst != 0).sum().item()) == k and torch.equal(dst[got], src[:k])
