"""CPU: `play.py` (teams, get_team, play_matches) and the PPO update schedule (remainder minibatch,
adaptive LR, target-KL stop, LR annealing) against vectors produced by EXECUTING the reference's own
code (tests/golden/make_golden.py::gen_play / gen_ppo_schedule)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import make_golden as mg  # noqa: E402  (fixture builders only; the reference is not read here)

from rsoccer_isaac_cleanrl_b200 import play, ppo  # noqa: E402
from rsoccer_isaac_cleanrl_b200.envs.spaces import Box  # noqa: E402


def _dummy(adim):
    return types.SimpleNamespace(single_observation_space=Box(-np.inf, np.inf, (52,)),
                                 single_action_space=Box(-1.0, 1.0, (adim,)))


def _check_sums(agent, z, prefix):
    for k, v in agent.state_dict().items():
        np.testing.assert_allclose(mg.weight_checksum(v), z[f"{prefix}{k}"], rtol=1e-9,
                                   err_msg=f"{k}: Agent initialisation differs from the reference's")


@pytest.fixture(scope="module")
def play_golden():
    torch.set_num_threads(1)
    return np.load(os.path.join(GOLDEN, "play.npz"))


@pytest.fixture(scope="module")
def checkpoints(play_golden, tmp_path_factory):
    """Checkpoints in the reference's format, rebuilt from the generator's seeds with the product's
    Agent (same layer order, same initialisers -> same weights; verified by checksum)."""
    d = tmp_path_factory.mktemp("ckpt")
    paths = {}
    for adim in (2, 6):
        torch.manual_seed(40 + adim)
        agent = ppo.Agent(_dummy(adim), mlp_backend="torch")
        with torch.no_grad():
            agent.actor_mean[8].weight.mul_(30.0)
            agent.actor_logstd.fill_(-1.0)
        _check_sums(agent, play_golden, f"ckpt{adim}_sum_")
        paths[adim] = str(d / f"agent{adim}.pt")
        torch.save(agent.state_dict(), paths[adim])
    return paths


CASES = {"zero_vs_ou": ("zero", None, "ou", None), "sa_vs_zero": ("ppo-sa", 2, "zero", None),
         "cma_vs_dma": ("ppo-cma", 6, "ppo-dma", 2), "sax3_vs_sa": ("ppo-sa-x3", 2, "ppo-sa", 2)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_play_matches_reproduces_the_reference(name, play_golden, checkpoints):
    z = play_golden
    obs0, steps = mg.play_fixture(np.random.default_rng(21))
    assert abs(obs0.astype(np.float64).sum() - float(z["obs0_sum"])) < 1e-6
    assert abs(steps[-1]["obs"].astype(np.float64).sum() - float(z["last_obs_sum"])) < 1e-6
    ba, bd, ya, yd = CASES[name]
    torch.manual_seed(77)
    blue = play.get_team(ba, checkpoints.get(bd), device="cpu", mlp_backend="torch")
    yellow = play.get_team(ya, checkpoints.get(yd), device="cpu", mlp_backend="torch")
    task = mg.FakeMatchTask(steps, obs0)
    assert obs0.shape[0] > 1065   # fields beyond the first 1065 end an episode every step and must not count
    score, length = play.play_matches(task, blue, yellow, 60)
    assert task.reset_dones_calls == 1 and task.t == int(z[f"{name}_steps"])
    assert abs(score - float(z[f"{name}_score"])) < 1e-12 and abs(length - float(z[f"{name}_length"])) < 1e-9
    for t, a in enumerate(task.seen_actions):
        # same torch generator stream, same op order up to float rounding in random_ou / the sampling
        np.testing.assert_allclose(a, z[f"{name}_act{t}"], rtol=0, atol=2e-6, err_msg=f"{name} action buffer, step {t}")


def test_team_dma_sees_each_robots_own_view(checkpoints):
    """play.py:62-64: the policy is applied to obs (N,3,52) row by row."""
    team = play.get_team("ppo-dma", checkpoints[2], device="cpu", mlp_backend="torch")
    obs = torch.randn(5, 3, 52)
    act = torch.zeros(5, 3, 2)
    torch.manual_seed(1)
    team(act, obs)
    torch.manual_seed(1)
    want = team.agent.get_action_and_value(obs.reshape(15, 52))[0].view(5, 3, 2)
    assert torch.equal(act, want.detach())


def test_get_team_rejects_unknown_algo():
    with pytest.raises(ValueError):
        play.get_team("nope")


# ---------------------------------------------------------------------------------------- schedule
@pytest.fixture(scope="module")
def sched():
    torch.set_num_threads(1)
    return np.load(os.path.join(GOLDEN, "ppo_schedule.npz"))


SCHED_CASES = {"plain": {}, "adaptive": dict(adaptative_lr=True, threshold_kl=0.008),
               "adaptive_up": dict(adaptative_lr=True, threshold_kl=50.0), "target_kl": dict(target_kl=0.002),
               "clipv_nonorm": dict(clip_vloss=True, norm_adv=False)}


@pytest.mark.parametrize("name", sorted(SCHED_CASES))
def test_update_schedule_matches_the_reference(name, sched):
    """ppo…:298-365 executed by the generator vs `update_policy` + `torch_minibatch_grad` + FlatAdam here:
    batch 14, minibatch 3 -> minibatches of 3,3,3,3,2 (the remainder of ppo…:310-312), 3 epochs."""
    z = sched
    fx = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("fx_")}
    torch.manual_seed(9)
    agent = ppo.Agent(_dummy(2), mlp_backend="torch")
    with torch.no_grad():
        agent.actor_mean[8].weight.mul_(20.0)
        agent.actor_logstd.fill_(-0.3)
    _check_sums(agent, z, "init_sum_")
    with torch.no_grad():
        _, lp, _, val = agent.get_action_and_value(fx["obs"], fx["actions"])
    argv = ["--num-envs", "7", "--num-steps", "2", "--update-epochs", "3", "--learning-rate", "0.003"]
    args = ppo.parse_args(argv)
    for k, v in SCHED_CASES[name].items():
        setattr(args, k, v)
    assert args.batch_size == 14 and args.minibatch_size == 3
    batch = dict(b_obs=fx["obs"], b_actions=fx["actions"], b_logprobs=lp + fx["dlogp"], b_advantages=fx["adv"],
                 b_returns=fx["ret"], b_values=val.view(-1) + fx["dval"])
    flat, flat_grad = ppo.flatten_parameters(agent)
    opt = ppo.FlatAdam(flat, flat_grad, lr=args.learning_rate, eps=1e-5)
    stats = {k: torch.zeros(()) for k in ("pg_loss", "v_loss", "entropy", "old_approx_kl", "approx_kl")}
    lr_trace, sizes = [], []

    def run_minibatch(inds):
        opt.sync_lr()
        lr_trace.append(opt.param_groups[0]["lr"])
        sizes.append(int(inds.numel()))
        ppo.torch_minibatch_grad(agent, args, batch, inds, stats)
        gnorm = torch.linalg.vector_norm(flat_grad)
        flat_grad.mul_(torch.clamp(args.max_grad_norm / (gnorm + 1e-6), max=1.0))   # clip_grad_norm_
        opt.step()

    torch.manual_seed(31)
    n_mb, n_ep = ppo.update_policy(args, args.batch_size, lambda n: torch.randperm(n), run_minibatch, opt,
                                   lambda: float(stats["approx_kl"]))
    assert n_mb == int(z[f"{name}_minibatches"]) == len(lr_trace)
    assert sizes[:5] == [3, 3, 3, 3, 2]
    np.testing.assert_allclose(lr_trace, z[f"{name}_lr_at_step"], rtol=1e-12)
    np.testing.assert_allclose(opt.param_groups[0]["lr"], float(z[f"{name}_lr_final"]), rtol=1e-12)
    np.testing.assert_allclose(float(stats["approx_kl"]), float(z[f"{name}_last_kl"]), rtol=2e-3, atol=1e-6)
    # final weights after 5-15 Adam steps. With the LR driven up to 1e-2 (adaptive_up) Adam's early steps are
    # sign-like (|g| / sqrt(v) ~ 1), so last-bit differences between FlatAdam and torch.optim.Adam grow to 1e-4.
    atol = 3e-4 if name == "adaptive_up" else 1e-5
    for k, v in agent.state_dict().items():
        np.testing.assert_allclose(mg.weight_sample(v), z[f"{name}_final_{k}"], rtol=2e-4, atol=atol, err_msg=k)
    if name == "target_kl":
        assert n_ep == 1 and n_mb == 5     # stops after the FIRST epoch's inner loop, not mid-epoch (ppo…:363-365)
    if name == "adaptive":
        assert min(lr_trace) < args.learning_rate          # KL above 2 x threshold: LR / 1.5
    if name == "adaptive_up":
        assert max(lr_trace) > args.learning_rate and max(lr_trace) <= 1e-2   # below threshold / 2: LR x 1.5, capped


def test_anneal_lr_matches_the_reference(sched):
    for u, want in zip(sched["anneal_updates"], sched["anneal_lr"]):
        assert ppo.anneal_lr(int(u), 48, 1e-3) == pytest.approx(float(want), rel=1e-15)
