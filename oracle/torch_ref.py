"""TEST / BENCH INFRASTRUCTURE — torch restatements of the reference's own GPU code paths, so that
`bench.py` can time "what the reference runs" on the same B200 beside the fused kernels (BASELINE.md
baselines B2-B4). `/root/reference` does not exist on the GPU box and its sources may not be copied, so
these are restatements of the SAME torch op sequences, pinned in `tests/test_torch_ref_golden.py`
against the golden vectors produced by executing the reference (`tests/golden/make_golden.py`):

  obs_rewards_dones   envs/vss.py:205-265 + jit :530-655  (two compute_obs calls per step, :195 and :203,
                      four reward terms, dones) — cat / repeat_interleave / advanced indexing / norms
  gae_loop            ppo_continuous_action_isaacgym.py:282-296  (python loop over T, ~8 launches per t)
  agent_update        ppo…:314-354 on an fp32 torch `Agent` with torch.optim.Adam (the product's
                      `ppo.torch_minibatch_grad` is that computation; this adds clip_grad_norm_ + Adam.step)

Only tests/ and bench.py import this module; the product never does.
"""
import torch

_MIRROR = (-1.0, -1.0, -1.0, -1.0, -1.0, -1.0, 1.0, 1.0, 1.0)
_PERMS = ((0, 1, 2), (1, 2, 0), (2, 0, 1))


def yaw_of(quats):
    """yaw = atan2(2(wz + xy), w^2 + x^2 - y^2 - z^2) mod 2 pi of xyzw quaternions (torch_utils.get_euler_xyz[2])."""
    x, y, z, w = quats[:, 0], quats[:, 1], quats[:, 2], quats[:, 3]
    return torch.atan2(2.0 * (w * z + x * y), w * w + x * x - y * y - z * z) % (2 * 3.141592653589793)


def compute_obs(b_pos, b_vel, r_pos, r_vel, r_quats, r_w, r_acts):
    """(N,2,3,52): vss.py:530-575 as torch ops (the reference rebuilds the mirror constant on the device
    every call, :533-538, so this does too)."""
    dev = b_pos.device
    mirror = torch.tensor(_MIRROR, dtype=torch.float, device=dev)
    perms = torch.tensor(_PERMS, device=dev)
    n = b_pos.shape[0]
    ball = torch.cat((b_pos, b_vel), -1).repeat_interleave(3, 0).view(n, 1, 3, 4)
    ang = yaw_of(r_quats.reshape(-1, 4)).view(n, 2, 3, 1)
    feats = torch.cat((r_pos, r_vel, torch.cos(ang), torch.sin(ang), r_w, r_acts), -1)           # (N,2,3,9)
    blue = torch.cat((ball, feats[:, 0, perms].view(n, 1, 3, 27),
                      feats[:, 1, :, :7].repeat_interleave(3, 0).view(n, 1, 3, 21)), -1)
    feats = feats * mirror
    yellow = torch.cat((-ball, feats[:, 1, perms].view(n, 1, 3, 27),
                        feats[:, 0, :, :7].repeat_interleave(3, 0).view(n, 1, 3, 21)), -1)
    return torch.cat((blue, yellow), 1)


def goal_rew(reset_buf, ball_pos, half_len=0.75, half_mouth=0.2):
    one = torch.ones_like(reset_buf)
    inside = (ball_pos[:, 0].abs() > half_len) & (ball_pos[:, 1].abs() < half_mouth)
    g = torch.where(inside & (ball_pos[:, 0] > 0), one, torch.zeros_like(one))
    g = torch.where(inside & (ball_pos[:, 0] < 0), -one, g).view(-1, 1, 1).expand(-1, 1, 3)
    return torch.cat((g, -g), 1)


def grad_rew(prev_ball, ball, goal):
    pot = lambda b: torch.norm(b + goal, dim=1) - torch.norm(b - goal, dim=1)
    d = (pot(ball) - pot(prev_ball)).view(-1, 1, 1).expand(-1, 1, 3)
    return torch.cat((d, -d), 1)


def move_rew(prev_robots, robots, prev_ball, ball):
    before = torch.norm(prev_robots.view(-1, 6, 2) - prev_ball.unsqueeze(1), dim=-1)
    after = torch.norm(robots.view(-1, 6, 2) - ball.unsqueeze(1), dim=-1)
    return (before - after).view(-1, 2, 3)


def energy_rew(actions):
    return -actions.abs().mean(-1)


def dones(ball_pos, reset_buf, progress, max_len=400, half_len=0.75, half_mouth=0.2):
    one = torch.ones_like(reset_buf)
    inside = (ball_pos[:, 0].abs() > half_len) & (ball_pos[:, 1].abs() < half_mouth)
    out = torch.where(inside, one, torch.zeros_like(reset_buf))
    return torch.where(progress >= max_len, one, out)


def obs_rewards_dones(s, w=(10.0, 2.0, 3.0, 0.0)):
    """What post_physics_step computes with torch ops per control step (vss.py:189-265): rewards into a
    zeroed (N,2,3,4) buffer, dones, and compute_obs TWICE (terminal observation, then the observation
    after the masked reset). `s`: dict of the state tensors in the reference's layouts."""
    rew = torch.zeros((*s["r_pos"].shape[:3], 4), device=s["ball_pos"].device)
    goal = torch.tensor([0.75, 0.0], device=rew.device)
    if w[0] > 0:
        rew[..., 0] = goal_rew(s["reset_buf"], s["ball_pos"]) * w[0]
    if w[1] > 0:
        rew[..., 1] = grad_rew(s["prev_ball_pos"], s["ball_pos"], goal) * w[1]
    if w[2] > 0:
        rew[..., 2] = move_rew(s["prev_r_pos"], s["r_pos"], s["prev_ball_pos"], s["ball_pos"]) * w[2]
    if w[3] > 0:
        rew[..., 3] = energy_rew(s["acts"]) * w[3]
    reset = dones(s["ball_pos"], s["reset_buf"], s["progress"])
    args = (s["ball_pos"], s["ball_vel"], s["r_pos"], s["r_vel"], s["quats"], s["r_w"], s["acts"])
    term_obs = compute_obs(*args).clone()
    obs = compute_obs(*args)
    return obs, term_obs, rew, reset


def gae_loop(rewards, values, next_values, next_dones, next_timeouts, gamma=0.99, gae_lambda=0.95):
    """ppo…:282-296: bootstrap from V(terminal obs) unless (done and not time-out); the lambda chain is cut by
    done; one python iteration (and ~8 launches) per time step."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = 0
    for t in reversed(range(T)):
        live = 1.0 - next_dones[t] * (1.0 - next_timeouts[t])
        delta = rewards[t] + gamma * next_values[t] * live - values[t]
        adv[t] = last = delta + gamma * gae_lambda * (1.0 - next_dones[t]) * last
    return adv, adv + values


def agent_update(agent, optimizer, args, batch, inds):
    """One minibatch of ppo…:314-354 on an fp32 torch Agent: loss, backward, clip_grad_norm_, Adam."""
    from rsoccer_isaac_cleanrl_b200.ppo import torch_minibatch_grad  # (bench/test only: times the fp32 torch path)
    loss = torch_minibatch_grad(agent, args, batch, inds)
    torch.nn.utils.clip_grad_norm_(agent.parameters(), args.max_grad_norm)
    optimizer.step()
    return loss
