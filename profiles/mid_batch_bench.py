import sys, os, torch
sys.path.insert(0, os.getcwd())
from rsoccer_isaac_cleanrl_b200.envs import VSS, load_cfg
out=[]
for n in (16384, 32768, 65536, 131072):
    cfg = load_cfg(); cfg["env"]["numEnvs"] = n
    envs = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=0)
    envs.reset_buf.zero_()
    st = envs.engine.get_state(); st[58, :n] = torch.randint(0, 400, (n,), device="cuda", dtype=torch.int32).view(torch.float32); envs.engine.set_state(st); del st
    acts = [torch.rand((n, 2, 3, 2), device="cuda") * 2 - 1 for _ in range(4)]
    for i in range(30): envs.step(acts[i & 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(300): envs.step(acts[i & 3])
    e1.record(); torch.cuda.synchronize()
    out.append("%d:%.1f" % (n, e0.elapsed_time(e1) / 300 * 1e3))
    del envs, acts
print(os.environ.get("VSS_WPB","-"), os.environ.get("VSS_SYNC","-"), " ".join(out), flush=True)
