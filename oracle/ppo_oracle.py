"""TEST INFRASTRUCTURE — numpy (float64) restatement of the PPO minibatch arithmetic of the reference,
`ppo_continuous_action_isaacgym.py`:
    :155-164  Normal(mean, exp(logstd)).log_prob(action).sum(1), .entropy().sum(1)
    :316-351  ratio, approx-KL, clip fraction, advantage normalisation, clipped surrogate, value loss
              (plain / clipped), entropy bonus, total loss
    :353-354  clip_grad_norm_ + Adam (torch.optim.Adam semantics, no weight decay)
plus the ANALYTIC gradient of the loss w.r.t. the network outputs — what `vss_ppo_loss` computes
instead of running autograd. Pinned by `tests/golden/ppo_update.npz`, which is produced by executing
the reference's own lines (`tests/golden/make_golden.py::gen_ppo_update`).

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline leg may import this module; the
product path (`rsoccer_isaac_cleanrl_b200/`) never does.
"""
import numpy as np

HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)


def log_prob_and_entropy(mean, logstd, action):
    """ppo…:155-164 (torch.distributions.Normal formulas)."""
    mean, logstd, action = (np.asarray(a, np.float64) for a in (mean, logstd, action))
    var = np.exp(2.0 * logstd)
    logp = (-((action - mean) ** 2) / (2.0 * var) - logstd - HALF_LOG_2PI).sum(1)
    entropy = np.broadcast_to(0.5 + HALF_LOG_2PI + logstd, mean.shape).sum(1)
    return logp, entropy


def ppo_loss(mean, value, logstd, b_action, b_logprob, b_adv, b_ret, b_val, inds, clip_coef, ent_coef, vf_coef,
             norm_adv=True, clip_vloss=False):
    """Returns dict(stats..., d_mean, d_value, d_logstd): the statistics of ppo…:316-351 and the
    gradient of `loss` w.r.t. mean (B,A), value (B,) and logstd (A,)."""
    f = lambda a: np.asarray(a, np.float64)
    mean, value, logstd = f(mean), f(value).reshape(-1), f(logstd).reshape(-1)
    j = np.arange(mean.shape[0]) if inds is None else np.asarray(inds)
    act, old_lp, adv, ret = f(b_action)[j], f(b_logprob)[j], f(b_adv)[j], f(b_ret)[j]
    B = mean.shape[0]
    logp, entropy = log_prob_and_entropy(mean, logstd, act)
    logratio = logp - old_lp                                   # :315
    ratio = np.exp(logratio)                                   # :316
    old_kl, kl = (-logratio).mean(), ((ratio - 1.0) - logratio).mean()   # :320-321
    clipfrac = (np.abs(ratio - 1.0) > clip_coef).mean()        # :322
    if norm_adv:                                               # :324-326 (torch .std() is unbiased)
        adv = (adv - adv.mean()) / (adv.std(ddof=1) + 1e-8)
    lo, hi = 1.0 - clip_coef, 1.0 + clip_coef
    t1, t2 = -adv * ratio, -adv * np.clip(ratio, lo, hi)       # :329-330
    pg_loss = np.maximum(t1, t2).mean()                        # :331
    if clip_vloss:                                             # :335-345
        v0 = f(b_val)[j]
        dv = value - v0
        vc = v0 + np.clip(dv, -clip_coef, clip_coef)
        u, c = (value - ret) ** 2, (vc - ret) ** 2
        v_loss = 0.5 * np.maximum(u, c).mean()
        inside_v = (dv >= -clip_coef) & (dv <= clip_coef)
        g_v = np.where(u > c, value - ret, np.where(u < c, np.where(inside_v, vc - ret, 0.0),
                                                    0.5 * (value - ret) + np.where(inside_v, 0.5 * (vc - ret), 0.0)))
    else:                                                      # :347
        v_loss = 0.5 * ((value - ret) ** 2).mean()
        g_v = value - ret
    ent = entropy.mean()                                       # :349
    loss = pg_loss - ent_coef * ent + v_loss * vf_coef         # :350
    # gradient: torch.max splits ties half/half, clamp passes the gradient on [lo, hi]
    inside = (ratio >= lo) & (ratio <= hi)
    g_ratio = np.where(inside | (t1 > t2), -adv, 0.0)
    g_lp = g_ratio * ratio / B
    inv_var = np.exp(-2.0 * logstd)
    d = act - mean
    d_mean = g_lp[:, None] * d * inv_var
    d_logstd = (g_lp[:, None] * (d * d * inv_var - 1.0)).sum(0) - ent_coef
    d_value = vf_coef * g_v / B
    return dict(pg_loss=pg_loss, v_loss=v_loss, entropy=ent, old_approx_kl=old_kl, approx_kl=kl, clipfrac=clipfrac,
                loss=loss, d_mean=d_mean, d_value=d_value, d_logstd=d_logstd)


def clip_adam(p, g, m, v, t, lr, max_norm, grad_scale=1.0, b1=0.9, b2=0.999, eps=1e-5):
    """clip_grad_norm_(max_norm) on g * grad_scale, then one Adam step (ppo…:353-354).
    Returns (p, g_clipped, m, v, t + 1)."""
    p, g, m, v = (np.asarray(a, np.float64) for a in (p, g, m, v))
    g = g * grad_scale
    norm = np.sqrt((g * g).sum())
    g = g * min(max_norm / (norm + 1e-6), 1.0)
    t = t + 1
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    denom = np.sqrt(v) / np.sqrt(1.0 - b2 ** t) + eps
    p = p - (lr / (1.0 - b1 ** t)) * m / denom
    return p, g, m, v, t
