"""ctypes/numpy binding of oracle/libvss_oracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module. It never touches a GPU and never imports the product package.

The arrays follow the reference's tensor layouts (envs/vss.py:112-141); see the header of
vss_oracle.c for what is pinned against the reference and what is "parity unpinned".
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvss_oracle.so")

NT, NR, NB, NOBS = 2, 3, 6, 52
STATE_FLOATS, STATE_WORDS = 58, 60


class VssParams(C.Structure):
    """Mirror of `vss_params` in include/vss_b200.h (field order matters)."""

    _fields_ = [
        ("dt", C.c_float), ("substeps", C.c_int32), ("max_episode_length", C.c_int32),
        ("field_half_length", C.c_float), ("field_half_width", C.c_float),
        ("goal_half_width", C.c_float), ("goal_depth", C.c_float),
        ("ball_radius", C.c_float), ("ball_mass", C.c_float), ("ball_drag", C.c_float),
        ("robot_half_size", C.c_float), ("robot_mass", C.c_float), ("robot_inertia", C.c_float),
        ("wheel_radius", C.c_float), ("wheel_half_track", C.c_float), ("wheel_coll_radius", C.c_float),
        ("max_wheel_rad_s", C.c_float), ("drive_damping", C.c_float), ("drive_max_torque", C.c_float),
        ("wheel_inertia", C.c_float), ("mu_traction", C.c_float), ("mu_lateral", C.c_float),
        ("gravity", C.c_float),
        ("restitution", C.c_float), ("mu_ball_robot", C.c_float), ("mu_ball_wall", C.c_float),
        ("mu_robot_wall", C.c_float),
        ("reset_scale_x", C.c_float), ("reset_scale_y", C.c_float), ("min_placement_dist", C.c_float),
        ("ball_reset_speed", C.c_float),
        ("w_goal", C.c_float), ("w_grad", C.c_float), ("w_move", C.c_float), ("w_energy", C.c_float),
        ("ou_theta", C.c_float), ("ou_sigma", C.c_float),
    ]


def default_params() -> VssParams:
    """The reference's constants, restated from envs/vss.py:47-49,342-345,380-434,
    vss_robot.urdf and vss.yaml (NOT read from the product's vss_default_params, so the
    two can be cross-checked in tests)."""
    p = VssParams()
    p.dt, p.substeps, p.max_episode_length = 0.05, 4, 400          # vss.yaml:16,6
    p.field_half_length, p.field_half_width = 1.5 / 2, 1.3 / 2     # vss.py:343
    p.goal_half_width, p.goal_depth = 0.4 / 2, 0.1                 # vss.py:344
    p.ball_radius = 0.02134                                        # vss.py:383
    p.ball_mass = 1130.0 * 4.0 / 3.0 * np.pi * 0.02134 ** 3        # vss.py:382
    p.ball_drag = 0.5 * 2.0 / 7.0                                  # angular damping 0.5 on a rolling sphere
    p.robot_half_size = 0.07 / 2                                   # urdf:16
    p.robot_mass = 0.4 + 2 * 0.02                                  # urdf:6,24,39
    p.robot_inertia = 0.4 * (0.07 ** 2 + 0.07 ** 2) / 12 + 2 * 0.02 * 0.03375 ** 2
    p.wheel_radius, p.wheel_half_track, p.wheel_coll_radius = 0.024, 0.03375, 0.024  # urdf:31,55,33
    p.max_wheel_rad_s = 42.0                                       # vss.py:47
    p.drive_damping, p.drive_max_torque = 0.01, 0.1                # vss.py:430, urdf:59
    p.wheel_inertia = 0.0002 + 0.4 * 0.02 * 0.024 ** 2             # vss.py:431 armature + sphere
    p.mu_traction, p.mu_lateral, p.gravity = 0.7, 0.55, 9.81       # vss.py:373-378,421-424; vss.yaml:17
    p.restitution, p.mu_ball_robot, p.mu_ball_wall, p.mu_robot_wall = 0.0, 0.5, 1.0, 0.5
    p.reset_scale_x, p.reset_scale_y = 1.5 - 0.14, 1.3 - 0.14      # vss.py:142-147
    p.min_placement_dist, p.ball_reset_speed = 0.07, 1.0           # vss.py:49,326
    p.w_goal, p.w_grad, p.w_move, p.w_energy = 10.0, 2.0, 3.0, 0.0 # vss.yaml:8-12
    p.ou_theta, p.ou_sigma = 0.1, 0.15                             # wrappers.py:6-7
    return p


class _OrcState(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("ball_pos", C.c_void_p), ("ball_vel", C.c_void_p), ("r_pos", C.c_void_p), ("r_vel", C.c_void_p),
        ("r_rot", C.c_void_p), ("r_w", C.c_void_p), ("r_act", C.c_void_p),
        ("progress", C.c_void_p), ("episode", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/libvss_oracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "vss_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class State:
    """Field state in the reference's layouts (a stand-in for the PhysX root-state views)."""

    def __init__(self, n: int):
        self.n = n
        self.ball_pos = np.zeros((n, 2), np.float32)
        self.ball_vel = np.zeros((n, 2), np.float32)
        self.r_pos = np.zeros((n, NT, NR, 2), np.float32)
        self.r_vel = np.zeros((n, NT, NR, 2), np.float32)
        self.r_rot = np.zeros((n, NT, NR, 2), np.float32)
        self.r_rot[..., 0] = 1.0
        self.r_w = np.zeros((n, NT, NR), np.float32)
        self.r_act = np.zeros((n, NT, NR, 2), np.float32)
        self.progress = np.zeros((n,), np.int64)
        self.episode = np.zeros((n,), np.uint32)

    def copy(self) -> "State":
        s = State(self.n)
        for k in ("ball_pos", "ball_vel", "r_pos", "r_vel", "r_rot", "r_w", "r_act", "progress", "episode"):
            getattr(s, k)[...] = getattr(self, k)
        return s

    def _c(self) -> _OrcState:
        return _OrcState(self.n, _p(self.ball_pos), _p(self.ball_vel), _p(self.r_pos), _p(self.r_vel),
                         _p(self.r_rot), _p(self.r_w), _p(self.r_act), _p(self.progress), _p(self.episode))

    # ---- conversion to / from the engine's SoA words (include/vss_b200.h) ----
    def to_soa(self, ld: int = None) -> np.ndarray:
        ld = ld or ((self.n + 31) // 32 * 32)
        w = np.zeros((STATE_WORDS, ld), np.float32)
        n = self.n
        w[0:2, :n] = self.ball_pos.T
        w[2:4, :n] = self.ball_vel.T
        for r in range(NB):
            t, j = divmod(r, NR)
            b = 4 + 9 * r
            w[b + 0:b + 2, :n] = self.r_pos[:, t, j].T
            w[b + 2:b + 4, :n] = self.r_vel[:, t, j].T
            w[b + 4:b + 6, :n] = self.r_rot[:, t, j].T
            w[b + 6, :n] = self.r_w[:, t, j]
            w[b + 7:b + 9, :n] = self.r_act[:, t, j].T
        w.view(np.int32)[58, :n] = self.progress.astype(np.int32)
        w.view(np.uint32)[59, :n] = self.episode
        return w

    @staticmethod
    def from_soa(w: np.ndarray, n: int) -> "State":
        s = State(n)
        w = np.ascontiguousarray(w, np.float32)
        s.ball_pos[...] = w[0:2, :n].T
        s.ball_vel[...] = w[2:4, :n].T
        for r in range(NB):
            t, j = divmod(r, NR)
            b = 4 + 9 * r
            s.r_pos[:, t, j] = w[b + 0:b + 2, :n].T
            s.r_vel[:, t, j] = w[b + 2:b + 4, :n].T
            s.r_rot[:, t, j] = w[b + 4:b + 6, :n].T
            s.r_w[:, t, j] = w[b + 6, :n]
            s.r_act[:, t, j] = w[b + 7:b + 9, :n].T
        s.progress[...] = w.view(np.int32)[58, :n]
        s.episode[...] = w.view(np.uint32)[59, :n]
        return s


# ------------------------------------------------------------------ jit-function restatements
def compute_obs(b_pos, b_vel, r_pos, r_vel, r_rot, r_w, r_acts):
    """envs/vss.py:530-575 with (cos, sin) of the yaw given directly."""
    b_pos, b_vel, r_pos, r_vel, r_rot, r_w, r_acts = map(_f32, (b_pos, b_vel, r_pos, r_vel, r_rot, r_w, r_acts))
    n = b_pos.shape[0]
    obs = np.empty((n, NT, NR, NOBS), np.float32)
    lib().orc_compute_obs(C.c_int64(n), _p(b_pos), _p(b_vel), _p(r_pos), _p(r_vel), _p(r_rot), _p(r_w),
                          _p(r_acts), _p(obs))
    return obs


def compute_goal_rew(ball_pos, field_width=1.5, goal_height=0.4):
    ball_pos = _f32(ball_pos)
    n = ball_pos.shape[0]
    out = np.empty((n, NT, NR), np.int64)
    lib().orc_compute_goal_rew(C.c_int64(n), _p(ball_pos), C.c_float(field_width), C.c_float(goal_height), _p(out))
    return out


def compute_grad_rew(prev_ball_pos, ball_pos, yellow_goal=(0.75, 0.0)):
    prev_ball_pos, ball_pos, yg = _f32(prev_ball_pos), _f32(ball_pos), _f32(yellow_goal)
    n = ball_pos.shape[0]
    out = np.empty((n, NT, NR), np.float32)
    lib().orc_compute_grad_rew(C.c_int64(n), _p(prev_ball_pos), _p(ball_pos), _p(yg), _p(out))
    return out


def compute_move_rew(p_robots, robots, p_ball, ball):
    p_robots, robots, p_ball, ball = map(_f32, (p_robots, robots, p_ball, ball))
    n = ball.shape[0]
    out = np.empty((n, NT, NR), np.float32)
    lib().orc_compute_move_rew(C.c_int64(n), _p(p_robots), _p(robots), _p(p_ball), _p(ball), _p(out))
    return out


def compute_energy_rew(actions):
    actions = _f32(actions)
    n = actions.shape[0]
    out = np.empty((n, NT, NR), np.float32)
    lib().orc_compute_energy_rew(C.c_int64(n), _p(actions), _p(out))
    return out


def compute_dones(ball_pos, progress, max_episode_length=400, field_width=1.5, goal_height=0.4):
    ball_pos = _f32(ball_pos)
    progress = np.ascontiguousarray(progress, np.int64)
    n = ball_pos.shape[0]
    out = np.empty((n,), np.int64)
    lib().orc_compute_dones(C.c_int64(n), _p(ball_pos), _p(progress), C.c_float(max_episode_length),
                            C.c_float(field_width), C.c_float(goal_height), _p(out))
    return out


# ------------------------------------------------------------------ sequencing
def reset_dones(params, seed, global_offset, state: State, reset_buf):
    reset_buf = np.ascontiguousarray(reset_buf, np.int64)
    cs = state._c()
    lib().orc_reset_dones(C.byref(params), C.c_uint64(seed), C.c_int64(global_offset), C.byref(cs), _p(reset_buf))


def physics(params, state: State):
    cs = state._c()
    lib().orc_physics(C.byref(params), C.byref(cs))


def step(params, seed, global_offset, state: State, actions, reset_buf, post_state=None, out=None):
    """VSS.step (envs/vss.py:180-203 + VecTask.step). reset_buf (N) int64 is updated in
    place. Returns dict(obs, term_obs, rew, timeout, progress_f); `out` = a dict returned by an
    earlier call, to write into the same (persistent) buffers as the reference does."""
    n = state.n
    actions = _f32(actions).reshape(n, NT, NR, 2)
    assert reset_buf.dtype == np.int64 and reset_buf.flags.c_contiguous
    out = out if out is not None else dict(
        obs=np.empty((n, NT, NR, NOBS), np.float32), term_obs=np.empty((n, NT, NR, NOBS), np.float32),
        rew=np.empty((n, NT, NR, 4), np.float32), timeout=np.empty((n,), np.uint8),
        progress_f=np.empty((n,), np.float32))
    ld = 0
    if post_state is not None:
        post_state = _f32(post_state)
        ld = post_state.shape[1]
    cs = state._c()
    lib().orc_step(C.byref(params), C.c_uint64(seed), C.c_int64(global_offset), C.byref(cs), _p(actions),
                   _p(post_state), C.c_int64(ld), _p(reset_buf), _p(out["obs"]), _p(out["term_obs"]),
                   _p(out["rew"]), _p(out["timeout"]), _p(out["progress_f"]))
    return out


VIEW_SA, VIEW_CMA, VIEW_DMA = 0, 1, 2


def step_view(params, seed, global_offset, step_index, state: State, view, policy_action, action_buf,
              reset_buf, ep_ret=None, ep_len=None, out=None):
    """SingleAgent/CMA/DMA.step + RecordEpisodeStatisticsTorch.step (envs/wrappers.py). `out` = a dict
    returned by an earlier call (same view, same statistics arguments): reuse its buffers."""
    n = state.n
    nv = n * 3 if view == VIEW_DMA else n
    policy_action = _f32(policy_action)
    assert action_buf.dtype == np.float32 and action_buf.flags.c_contiguous
    out = out if out is not None else dict(
        obs=np.empty((nv, NOBS), np.float32), term_obs=np.empty((nv, NOBS), np.float32),
        rews=np.empty((nv, 4), np.float32), reward=np.empty((nv,), np.float32),
        done=np.empty((nv,), np.int64), timeout=np.empty((nv,), np.uint8),
        progress=np.empty((nv,), np.float32))
    ret_ret = ret_len = None
    if ep_ret is not None:
        if "ret_ret" not in out:
            out["ret_ret"], out["ret_len"] = np.empty((nv, 4), np.float32), np.empty((nv,), np.int32)
        ret_ret, ret_len = out["ret_ret"], out["ret_len"]
    cs = state._c()
    lib().orc_step_view(C.byref(params), C.c_uint64(seed), C.c_int64(global_offset), C.c_uint32(step_index),
                        C.byref(cs), C.c_int(view), _p(policy_action), _p(action_buf), _p(reset_buf),
                        _p(out["obs"]), _p(out["term_obs"]), _p(out["rews"]), _p(out["reward"]),
                        _p(out["done"]), _p(out["timeout"]), _p(out["progress"]),
                        _p(ep_ret), _p(ep_len), _p(ret_ret), _p(ret_len))
    return out


def gae(rewards, values, next_values, next_dones, next_timeouts, gamma=0.99, gae_lambda=0.95):
    """ppo_continuous_action_isaacgym.py:282-296. Arrays (T,N) f32."""
    rewards, values, next_values, next_dones, next_timeouts = map(
        _f32, (rewards, values, next_values, next_dones, next_timeouts))
    T, N = rewards.shape
    adv, ret = np.empty((T, N), np.float32), np.empty((T, N), np.float32)
    lib().orc_gae(_p(rewards), _p(values), _p(next_values), _p(next_dones), _p(next_timeouts), _p(adv),
                  _p(ret), C.c_int32(T), C.c_int64(N), C.c_double(gamma), C.c_double(gae_lambda))
    return adv, ret


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.empty(4, np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())


def view_outputs(view, obs, term_obs, rew, reset_buf, timeout, progress_f, action_buf, ep_ret=None, ep_len=None):
    """Slicing/aggregation part of SingleAgent/CMA/DMA.step (+ episode statistics) applied to raw
    VSS.step outputs (envs/wrappers.py:104-115,136-148,166-180,68-81). action_buf rows of done
    fields are zeroed in place."""
    obs, term_obs, rew, progress_f = map(_f32, (obs, term_obs, rew, progress_f))
    reset_buf = np.ascontiguousarray(reset_buf, np.int64)
    timeout = np.ascontiguousarray(timeout, np.uint8)
    n = obs.shape[0]
    nv = n * 3 if view == VIEW_DMA else n
    assert action_buf.dtype == np.float32 and action_buf.flags.c_contiguous
    out = dict(
        obs=np.empty((nv, NOBS), np.float32), term_obs=np.empty((nv, NOBS), np.float32),
        rews=np.empty((nv, 4), np.float32), reward=np.empty((nv,), np.float32),
        done=np.empty((nv,), np.int64), timeout=np.empty((nv,), np.uint8),
        progress=np.empty((nv,), np.float32))
    ret_ret = ret_len = None
    if ep_ret is not None:
        ret_ret, ret_len = np.empty((nv, 4), np.float32), np.empty((nv,), np.int32)
        out["ret_ret"], out["ret_len"] = ret_ret, ret_len
    lib().orc_view_outputs(C.c_int64(n), C.c_int(view), _p(obs), _p(term_obs), _p(rew), _p(reset_buf),
                           _p(timeout), _p(progress_f), _p(action_buf), _p(out["obs"]), _p(out["term_obs"]),
                           _p(out["rews"]), _p(out["reward"]), _p(out["done"]), _p(out["timeout"]),
                           _p(out["progress"]), _p(ep_ret), _p(ep_len), _p(ret_ret), _p(ret_len))
    return out


def set_num_threads(n: int):
    lib().orc_set_num_threads(C.c_int(int(n)))
