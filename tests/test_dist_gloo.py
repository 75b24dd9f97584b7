"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (flat-gradient all-reduce,
field sharding by global id)."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rsoccer_isaac_cleanrl_b200 import ppo
    envs = types.SimpleNamespace(single_observation_space=types.SimpleNamespace(shape=(52,)),
                                 single_action_space=types.SimpleNamespace(shape=(2,)))
    torch.manual_seed(0)
    agent = ppo.Agent(envs)
    flat, flat_grad = ppo.flatten_parameters(agent)
    dist.broadcast(flat, 0)
    g = torch.Generator().manual_seed(5)
    x, a, adv = torch.randn(64, 52, generator=g), torch.randn(64, 2, generator=g), torch.randn(64, generator=g)
    lo, hi = rank * 32, (rank + 1) * 32            # each rank: its half of the global minibatch
    _, logp, ent, v = agent.get_action_and_value(x[lo:hi], a[lo:hi])
    loss = (-(adv[lo:hi] * logp.exp())).mean() - 0.005 * ent.mean() + 2.0 * (v.view(-1) ** 2).mean()
    flat_grad.zero_()
    loss.backward()
    dist.all_reduce(flat_grad)
    flat_grad.div_(world)
    if rank == 0:
        ret["grad"] = flat_grad.clone().numpy()
        ret["flat"] = flat.clone().numpy()
    # sharding: rank r owns global field ids [r*n, (r+1)*n)
    ret[f"offset{rank}"] = rank * 96
    dist.barrier()
    dist.destroy_process_group()


def test_allreduced_gradient_equals_single_process_gradient():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    from rsoccer_isaac_cleanrl_b200 import ppo
    envs = types.SimpleNamespace(single_observation_space=types.SimpleNamespace(shape=(52,)),
                                 single_action_space=types.SimpleNamespace(shape=(2,)))
    torch.manual_seed(0)
    agent = ppo.Agent(envs)
    flat, flat_grad = ppo.flatten_parameters(agent)
    np.testing.assert_array_equal(flat.numpy(), ret["flat"])
    g = torch.Generator().manual_seed(5)
    x, a, adv = torch.randn(64, 52, generator=g), torch.randn(64, 2, generator=g), torch.randn(64, generator=g)
    _, logp, ent, v = agent.get_action_and_value(x, a)
    loss = (-(adv * logp.exp())).mean() - 0.005 * ent.mean() + 2.0 * (v.view(-1) ** 2).mean()
    flat_grad.zero_()
    loss.backward()
    np.testing.assert_allclose(ret["grad"], flat_grad.numpy(), rtol=1e-4, atol=1e-6)
    assert ret["offset0"] == 0 and ret["offset1"] == 96


def test_field_sharding_is_invariant_on_the_host_build():
    """Two shards keyed by global field id reproduce one engine over all fields (emu backend)."""
    import parity_checks as pc
    from backends import EmuBackend
    n = 64
    whole, lo, hi = EmuBackend(2 * n, seed=4, goff=0), EmuBackend(n, seed=4, goff=0), EmuBackend(n, seed=4, goff=n)
    ow = whole.reset_dones(np.ones(2 * n, np.int64))
    ol, oh = lo.reset_dones(np.ones(n, np.int64)), hi.reset_dones(np.ones(n, np.int64))
    assert np.array_equal(pc.bits(ow[:n]), pc.bits(ol)) and np.array_equal(pc.bits(ow[n:]), pc.bits(oh))
    rng = np.random.default_rng(0)
    rbw, rbl, rbh = np.zeros(2 * n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
    for t in range(10):
        a = rng.uniform(-1, 1, (2 * n, 2, 3, 2)).astype(np.float32)
        w, l, h = whole.step(a, rbw), lo.step(a[:n], rbl), hi.step(a[n:], rbh)
        for k in ("obs", "term_obs", "rew", "timeout", "progress_f"):
            assert np.array_equal(w[k][:n], l[k]) and np.array_equal(w[k][n:], h[k]), (t, k)
