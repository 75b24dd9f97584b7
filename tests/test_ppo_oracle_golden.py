"""CPU: the numpy restatement of the PPO minibatch arithmetic (oracle/ppo_oracle.py) against vectors
produced by executing the reference's own update code (tests/golden/ppo_update.npz, ppo…:314-354)."""
import os

import numpy as np
import pytest

from oracle import ppo_oracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ppo_update.npz")


@pytest.mark.parametrize("name,clip_vloss", [("a2", False), ("a6v", True)])
def test_loss_statistics_and_output_gradients_match_the_reference(name, clip_vloss):
    g = np.load(GOLD)
    k = lambda s: g[f"{name}_{s}"]
    r = po.ppo_loss(k("mean"), k("value"), k("logstd"), k("b_actions"), k("b_logprobs"), k("b_advantages"),
                    k("b_returns"), k("b_values"), k("mb_inds"), 0.2, 0.005, 4.0, norm_adv=True, clip_vloss=clip_vloss)
    for ours, theirs in (("pg_loss", "pg_loss"), ("v_loss", "v_loss"), ("entropy", "entropy_loss"),
                         ("old_approx_kl", "old_approx_kl"), ("approx_kl", "approx_kl"), ("loss", "loss")):
        assert abs(r[ours] - float(k(theirs))) <= 2e-6 + 2e-6 * abs(float(k(theirs))), (ours, r[ours], float(k(theirs)))
    assert r["clipfrac"] == pytest.approx(float(k("clipfrac")), abs=1e-7)
    np.testing.assert_allclose(r["d_mean"], k("d_mean"), rtol=2e-5, atol=1e-8)
    np.testing.assert_allclose(r["d_value"], k("d_value").reshape(-1), rtol=2e-5, atol=1e-8)
    np.testing.assert_allclose(r["d_logstd"], k("d_logstd"), rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["a2", "a6v"])
def test_clip_and_adam_step_of_logstd_matches_the_reference(name):
    """The one parameter whose full gradient is in the fixture: logstd. The clip coefficient uses the
    norm of ALL gradients (recorded), the Adam step is element-wise."""
    g = np.load(GOLD)
    k = lambda s: g[f"{name}_{s}"]
    norm = float(k("grad_norm"))
    coef = min(1.5 / (norm + 1e-6), 1.0)
    np.testing.assert_allclose(k("d_logstd") * coef, k("clipped_d_logstd"), rtol=5e-6)  # fp32 norm in torch
    # feed the oracle a gradient vector with the recorded total norm: logstd's entries + one filler
    d = k("d_logstd").astype(np.float64)
    filler = np.sqrt(max(norm ** 2 - (d ** 2).sum(), 0.0))
    grad = np.concatenate([d, [filler]])
    p0 = np.concatenate([k("logstd").astype(np.float64), [0.0]])
    p1, gc, _, _, t = po.clip_adam(p0, grad, np.zeros_like(p0), np.zeros_like(p0), 0, 1e-3, 1.5)
    assert t == 1
    np.testing.assert_allclose(gc[:-1], k("clipped_d_logstd"), rtol=1e-5)
    np.testing.assert_allclose(p1[:-1], k("logstd_after"), rtol=0, atol=2e-7)


def test_log_prob_and_entropy_match_the_reference_agent():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "agent.npz"))
    logstd = g["a2_sd_actor_logstd"].reshape(-1)
    logp, ent = po.log_prob_and_entropy(g["a2_mean"], logstd, g["a2_action"])
    np.testing.assert_allclose(logp, g["a2_logp"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(ent, g["a2_ent"], rtol=2e-6)
