"""GPU (two or more B200s on one node): the hand-written gradient all-reduce over NVLink peer memory
(csrc/peer_reduce.cu) against NCCL's all-reduce of the same buffers. Skipped on a one-GPU box."""
import os

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, out_q):
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from rsoccer_isaac_cleanrl_b200.peer import PeerGradients
    dev = torch.device("cuda", rank)
    pg = PeerGradients(n, dev, rank, world)
    assert pg.buffer.numel() >= n and pg.buffer.numel() % 4 == 0 and bool((pg.buffer == 0).all())
    out = torch.zeros_like(pg.buffer)
    worst = 0.0
    for it in range(6):   # consecutive calls: the step counter, both flag phases and buffer reuse
        g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
        pg.buffer.copy_(torch.randn(pg.buffer.numel(), device=dev, generator=g) * (1.0 + it))
        want = pg.buffer.clone()
        dist.all_reduce(want)
        pg.allreduce(out)
        pg.buffer.zero_()          # what the next backward pass does: legal as soon as the kernel has run
        torch.cuda.synchronize()
        worst = max(worst, float((out - want).abs().max() / want.abs().max()))
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], x) for x in gathered), "ranks disagree on the sum"
    # inside a CUDA graph, replayed
    pg.buffer.fill_(float(rank + 1))
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        pg.allreduce(out)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert bool((out == world * (world + 1) / 2).all())
    dist.barrier()
    if rank == 0:
        out_q.put(worst)
    pg.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1_080_077, 4100])
def test_peer_allreduce_matches_nccl(n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (n % 50)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) < 1e-6   # fp32 sums in a different order than NCCL's
