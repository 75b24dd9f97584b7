"""Team policies and the match runner — drop-in for the reference's `play.py`.

Same names and call conventions (reference play.py:26-164): a team is a callable
`team(act, obs)` that fills its slice `act (N,3,2)` of the `(N,2,3,2)` action buffer in place
from its own observations `obs (N,3,52)`; the yellow team receives the 180-degree mirrored
observations, so one policy plays either side. `play_matches` steps the raw `VSS` task (one fused
kernel launch per step).

Difference: the reference builds `BASELINE_TEAMS` from `base_nets/*/agent.pt` at import time;
those checkpoints are missing from the reference tree (`.MISSING_LARGE_BLOBS`), so here the table
is built lazily by `baseline_teams(base_dir)` and only lists the checkpoints that exist (plus
the parameter-free `zero` and `ou` teams).
"""
import os
from abc import ABC, abstractmethod
from collections import namedtuple

import numpy as np
import torch

from .envs.spaces import Box
from .envs.wrappers import random_ou
from .ppo import Agent


class Team(ABC):
    def __init__(self, path=None, env_d=None):
        pass

    @abstractmethod
    def __call__(self, act, obs):
        pass


class TeamZero(Team):
    def __call__(self, act, obs):
        act[:] *= 0


class TeamOU(Team):
    def __call__(self, act, obs):
        act[:] = random_ou(act)


class TeamAgent(Team):
    def __init__(self, path, env_d, device="cuda:0", mlp_backend="tc"):
        self.agent = Agent(env_d, mlp_backend=mlp_backend).to(device)
        self.agent.load_state_dict(torch.load(path, map_location=device))  # reference state_dict keys
        self.agent.eval()


class TeamSA(TeamAgent):
    """A single-agent policy drives robot 0; robots 1-2 follow OU noise (play.py:51-54)."""

    @torch.no_grad()
    def __call__(self, act, obs):
        act[:] = random_ou(act)
        act[:, 0, :] = self.agent.get_action_and_value(obs[:, 0, :].contiguous())[0]


class TeamCMA(TeamAgent):
    """One centralised 6-dim action for the three robots, from robot 0's view (play.py:57-59)."""

    @torch.no_grad()
    def __call__(self, act, obs):
        act[:] = self.agent.get_action_and_value(obs[:, 0, :].contiguous())[0].view(-1, 3, 2)


class TeamDMA(TeamAgent):
    """The same policy applied to each robot's own view (play.py:62-64)."""

    @torch.no_grad()
    def __call__(self, act, obs):
        n = obs.shape[0]
        act[:] = self.agent.get_action_and_value(obs.reshape(n * 3, -1))[0].view(n, 3, 2)


def get_team(algo, path=None, device="cuda:0", mlp_backend="tc"):
    dummy_env = namedtuple("dummy_env", ["single_observation_space", "single_action_space"])
    obs_space = Box(-np.inf, np.inf, (52,))
    if algo == "ppo-sa":
        return TeamSA(path, dummy_env(obs_space, Box(-1.0, 1.0, (2,))), device, mlp_backend)
    if algo in ("ppo-sa-x3", "ppo-dma"):
        return TeamDMA(path, dummy_env(obs_space, Box(-1.0, 1.0, (2,))), device, mlp_backend)
    if algo == "ppo-cma":
        return TeamCMA(path, dummy_env(obs_space, Box(-1.0, 1.0, (6,))), device, mlp_backend)
    if algo == "zero":
        return TeamZero()
    if algo == "ou":
        return TeamOU()
    raise ValueError(f"Unknown algo: {algo}")


def baseline_teams(base_dir="base_nets", device="cuda:0"):
    """The reference's BASELINE_TEAMS table (play.py:105-128), restricted to checkpoints that exist."""
    teams = {"zero": {"00": get_team("zero")}, "ou": {"00": get_team("ou")}}
    for algo, sub in (("ppo-sa", "ppo-sa"), ("ppo-sa-x3", "ppo-sa"), ("ppo-cma", "ppo-cma"), ("ppo-dma", "ppo-dma")):
        for seed in ("10", "20", "30"):
            path = os.path.join(base_dir, f"exp000_{sub}_{seed}", "agent.pt")
            if os.path.exists(path):
                teams.setdefault(algo, {})[seed] = get_team(algo, path, device)
    return teams


def play_matches(envs, blue_team, yellow_team, n_matches, video_path=None, count_envs=1065):
    """Mean blue goal score and mean episode length over `n_matches` finished episodes
    (reference play.py:131-164). Only the first `count_envs` fields are counted, as the
    reference does (hard-coded 1065 there)."""
    if video_path:
        raise NotImplementedError("video capture is out of scope of the B200 engine")
    envs.reset_buf[:] = 1
    envs.reset_dones()
    ep_count = 0
    rew_sum = 0.0
    len_sum = 0.0
    n = envs.cfg["env"]["numEnvs"]
    action_buf = torch.zeros((n,) + envs.action_space.shape, device=envs.device)
    obs = envs.reset()["obs"]
    count_envs = min(count_envs, n)
    while ep_count < n_matches:
        blue_team(action_buf[:, 0], obs[:, 0])
        yellow_team(action_buf[:, 1], obs[:, 1])
        obs, rew, dones, info = envs.step(action_buf)
        obs = obs["obs"]
        env_ids = dones[:count_envs].nonzero(as_tuple=False).squeeze(-1)
        if len(env_ids):
            ep_count += len(env_ids)
            rew_sum += rew[env_ids, 0, 0, 0].sum().item()
            len_sum += info["progress_buffer"][env_ids].sum().item()
    return rew_sum / ep_count, len_sum / ep_count
