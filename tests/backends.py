"""Two interchangeable backends for the parity tests.

EmuBackend  host build of the product's per-lane code (tests/emu) — runs in the GPU-less
            container, checks the logic the CUDA kernel compiles.
GpuBackend  the real thing: libvss_b200.so through the C-ABI on cuda:0 (tests marked gpu).

Both expose numpy in / numpy out so the comparison code is shared.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NT, NR, NOBS = 2, 3, 52
PACKED_ROW_BYTES = 112
VIEW_FULL, VIEW_SA, VIEW_CMA, VIEW_DMA = -1, 0, 1, 2


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _nv(n, view):
    return n * 3 if view == VIEW_DMA else n


class EmuBackend:
    name = "emu"
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            d = os.path.join(ROOT, "tests", "emu")
            subprocess.run(["make", "-C", d], check=True, capture_output=True)
            cls._lib = C.CDLL(os.path.join(d, "libvss_emu.so"))
        return cls._lib

    wpt = 1   # warps per tile: 1 = the k_step structure, 2..8 = the k_step_cta structure
    fpt = 32  # fields per tile of the k_step_cta structure (32, 16 or 8)

    def __init__(self, n, seed=0, goff=0, params=None):
        from oracle import vss_oracle as orc
        self.n, self.seed, self.goff = n, seed, goff
        self.ld = (n + 31) // 32 * 32
        self.params = params if params is not None else orc.default_params()
        self.state = np.zeros((60, self.ld), np.float32)
        self.step_count = 0

    def set_reward_weights(self, goal, grad, move, energy):
        self.params.w_goal, self.params.w_grad, self.params.w_move, self.params.w_energy = goal, grad, move, energy

    def get_state(self):
        return self.state.copy()

    def set_state(self, s):
        self.state[...] = s

    def reset_dones(self, reset_buf):
        obs = np.zeros((self.n, NT, NR, NOBS), np.float32)
        rb = np.ascontiguousarray(reset_buf, np.int64)
        rc = self.lib().emu_reset_dones(C.byref(self.params), _p(self.state), C.c_longlong(self.n),
                                        C.c_longlong(self.ld), C.c_ulonglong(self.goff), C.c_ulonglong(self.seed),
                                        _p(rb), _p(obs))
        assert rc == 0
        return obs

    def _call(self, view, actions, inject, reset_buf, obs, term_obs, rew, timeout, progress_f, policy_action=None,
              action_buf=None, reward_v=None, done_v=None, ep_ret=None, ep_len=None, ret_ret=None, ret_len=None,
              packed=None):
        rc = self.lib().emu_step(C.byref(self.params), C.c_int(view), _p(self.state), C.c_longlong(self.n),
                                 C.c_longlong(self.ld), C.c_ulonglong(self.goff), C.c_ulonglong(self.seed),
                                 C.c_uint(self.step_count & 0xFFFFFFFF), _p(actions), _p(inject), _p(reset_buf),
                                 _p(obs), _p(term_obs), _p(rew), _p(timeout), _p(progress_f), _p(policy_action),
                                 _p(action_buf), _p(reward_v), _p(done_v), _p(ep_ret), _p(ep_len), _p(ret_ret),
                                 _p(ret_len), _p(packed), C.c_int(self.wpt), C.c_int(self.fpt))
        assert rc == 0
        self.step_count += 1

    def step(self, actions, reset_buf, post_state=None):
        n = self.n
        actions = np.ascontiguousarray(actions, np.float32).reshape(n, NT, NR, 2)
        assert reset_buf.dtype == np.int64
        out = dict(obs=np.zeros((n, NT, NR, NOBS), np.float32), term_obs=np.zeros((n, NT, NR, NOBS), np.float32),
                   rew=np.zeros((n, NT, NR, 4), np.float32), timeout=np.zeros((n,), np.uint8),
                   progress_f=np.zeros((n,), np.float32))
        inj = None if post_state is None else np.ascontiguousarray(post_state, np.float32)
        self._call(VIEW_FULL, actions, inj, reset_buf, out["obs"], out["term_obs"], out["rew"], out["timeout"],
                   out["progress_f"])
        return out

    def step_view(self, view, policy_action, action_buf, reset_buf, ep_ret=None, ep_len=None, packed=False):
        n, nv = self.n, _nv(self.n, view)
        policy_action = np.ascontiguousarray(policy_action, np.float32)
        pk = np.zeros((nv, PACKED_ROW_BYTES), np.uint8) if packed else None
        out = dict(obs=np.zeros((nv, NOBS), np.float32), term_obs=np.zeros((nv, NOBS), np.float32),
                   rews=np.zeros((nv, 4), np.float32), reward=np.zeros((nv,), np.float32),
                   done=np.zeros((nv,), np.int64), timeout=np.zeros((nv,), np.uint8),
                   progress=np.zeros((nv,), np.float32))
        ret_ret = ret_len = None
        if ep_ret is not None:
            ret_ret, ret_len = np.zeros((nv, 4), np.float32), np.zeros((nv,), np.int32)
            out["ret_ret"], out["ret_len"] = ret_ret, ret_len
        self._call(view, None, None, reset_buf, out["obs"], out["term_obs"], out["rews"], out["timeout"],
                   out["progress"], policy_action, action_buf, out["reward"], out["done"], ep_ret, ep_len, ret_ret,
                   ret_len, packed=pk)
        if packed:
            out["packed"] = pk
        return out


class GpuBackend:
    """libvss_b200.so through the product's Engine wrapper (C-ABI) on cuda:0."""

    name = "gpu"
    wpt = None   # None = the library's automatic launch shape; 1..8 forces warps per tile
    fpt = None   # None = automatic; 8 / 16 / 32 forces the fields per tile

    def __init__(self, n, seed=0, goff=0, params=None):
        import torch
        import rsoccer_isaac_cleanrl_b200 as R
        self.torch = torch
        self.n = n
        p = R.default_params()
        if params is not None:  # copy field by field from the oracle's struct
            for k, _ in p._fields_:
                setattr(p, k, getattr(params, k))
        self.params = p
        self.eng = R.Engine(n, "cuda:0", seed=seed, global_env_offset=goff, params=p)
        if self.wpt is not None:
            self.eng.warps_per_tile = self.wpt
        if self.fpt is not None:
            self.eng.fields_per_tile = self.fpt
        self.ld = self.eng.ld
        self.dev = torch.device("cuda:0")

    @property
    def step_count(self):
        return self.eng.step_count

    @step_count.setter
    def step_count(self, v):
        self.eng.step_count = v

    def _t(self, a):
        return None if a is None else self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    @property
    def sanitised_count(self):
        return self.eng.sanitised_count

    def set_reward_weights(self, goal, grad, move, energy):
        self.eng.set_reward_weights(goal, grad, move, energy)

    def get_state(self):
        return self.eng.get_state().cpu().numpy()

    def set_state(self, s):
        self.eng.set_state(self._t(np.ascontiguousarray(s, np.float32)))
        self.torch.cuda.synchronize()

    def reset_dones(self, reset_buf):
        t = self.torch
        obs = t.zeros((self.n, NT, NR, NOBS), dtype=t.float32, device=self.dev)
        self.eng.reset_dones(self._t(np.ascontiguousarray(reset_buf, np.int64)), obs)
        return obs.cpu().numpy()

    def step(self, actions, reset_buf, post_state=None):
        t, n = self.torch, self.n
        z = lambda shape, dt: t.zeros(shape, dtype=dt, device=self.dev)
        rb = self._t(reset_buf)
        obs, tobs = z((n, NT, NR, NOBS), t.float32), z((n, NT, NR, NOBS), t.float32)
        rew, tout, prog = z((n, NT, NR, 4), t.float32), z((n,), t.uint8), z((n,), t.float32)
        acts = self._t(np.ascontiguousarray(actions, np.float32).reshape(n, NT, NR, 2))
        self.eng.step(acts, rb, obs, tobs, rew, tout, prog,
                      post_state=None if post_state is None else self._t(np.ascontiguousarray(post_state, np.float32)))
        t.cuda.synchronize()
        reset_buf[...] = rb.cpu().numpy()
        return dict(obs=obs.cpu().numpy(), term_obs=tobs.cpu().numpy(), rew=rew.cpu().numpy(),
                    timeout=tout.cpu().numpy(), progress_f=prog.cpu().numpy())

    def step_view(self, view, policy_action, action_buf, reset_buf, ep_ret=None, ep_len=None, packed=False):
        t, n = self.torch, self.n
        nv = _nv(n, view)
        z = lambda shape, dt: t.zeros(shape, dtype=dt, device=self.dev)
        rb, ab = self._t(reset_buf), self._t(action_buf)
        obs, tobs, rews, reward = z((nv, NOBS), t.float32), z((nv, NOBS), t.float32), z((nv, 4), t.float32), z((nv,), t.float32)
        done, tout, prog = z((nv,), t.int64), z((nv,), t.uint8), z((nv,), t.float32)
        er = el = rr = rl = None
        if ep_ret is not None:
            er, el = self._t(ep_ret), self._t(ep_len)
            rr, rl = z((nv, 4), t.float32), z((nv,), t.int32)
        pk = z((nv, PACKED_ROW_BYTES), t.uint8) if packed else None
        self.eng.step_view(view, self._t(np.ascontiguousarray(policy_action, np.float32)), ab, rb, obs, tobs, rews,
                           reward, done, tout, prog, er, el, rr, rl, packed=pk)
        t.cuda.synchronize()
        reset_buf[...] = rb.cpu().numpy()
        action_buf[...] = ab.cpu().numpy()
        out = dict(obs=obs.cpu().numpy(), term_obs=tobs.cpu().numpy(), rews=rews.cpu().numpy(),
                   reward=reward.cpu().numpy(), done=done.cpu().numpy(), timeout=tout.cpu().numpy(),
                   progress=prog.cpu().numpy())
        if packed:
            out["packed"] = pk.cpu().numpy()
        if ep_ret is not None:
            ep_ret[...] = er.cpu().numpy()
            ep_len[...] = el.cpu().numpy()
            out["ret_ret"], out["ret_len"] = rr.cpu().numpy(), rl.cpu().numpy()
        return out


def with_wpt(Backend, wpt, fpt=None):
    """The same backend with a forced launch shape (warps per tile, and optionally fields per tile)."""
    attrs = {"wpt": wpt}
    if fpt is not None:
        attrs["fpt"] = fpt
    return type(f"{Backend.__name__}W{wpt}F{fpt}", (Backend,), attrs)
