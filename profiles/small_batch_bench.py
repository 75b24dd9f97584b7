import sys, os, torch
sys.path.insert(0, os.getcwd())
from rsoccer_isaac_cleanrl_b200.envs import VSS, SingleAgent, DMA, load_cfg
out=[]
def timed_graph(fn, k=20, reps=20):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(k): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (k * reps) * 1e3
for n in (1024, 4096, 8192, 16384, 21845, 32768):
    cfg = load_cfg(); cfg["env"]["numEnvs"] = n
    envs = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=0)
    envs.reset_buf.zero_()
    st = envs.engine.get_state(); st[58, :n] = torch.randint(0, 400, (n,), device="cuda", dtype=torch.int32).view(torch.float32); envs.engine.set_state(st); del st
    acts = [torch.rand((n, 2, 3, 2), device="cuda") * 2 - 1 for _ in range(4)]
    full = timed_graph(lambda i: envs.step(acts[i & 3]))
    view = SingleAgent(envs); pa = torch.rand((n, 2), device="cuda")
    sa = timed_graph(lambda i: view.step(pa))
    out.append("%d: full %.1f sa %.1f" % (n, full, sa))
    del envs, acts, view
print("fpw", os.environ.get("VSS_FPW","auto"), " | ".join(out), flush=True)
