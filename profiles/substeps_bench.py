import sys, os, torch
sys.path.insert(0, os.getcwd())
from rsoccer_isaac_cleanrl_b200.envs import VSS, load_cfg
n = 1 << 20
for sub in (1, 2, 4, 8):
    cfg = load_cfg(); cfg["env"]["numEnvs"] = n; cfg["sim"]["substeps"] = sub
    envs = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=0)
    envs.reset_buf.zero_()
    st = envs.engine.get_state(); st[58, :n] = torch.randint(0, 400, (n,), device="cuda", dtype=torch.int32).view(torch.float32); envs.engine.set_state(st); del st
    acts = [torch.rand((n, 2, 3, 2), device="cuda") * 2 - 1 for _ in range(4)]
    for i in range(20): envs.step(acts[i & 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): envs.step(acts[i & 3])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print(f"substeps {sub}: {ms*1e3:.1f} us/step, {n/ms*1e3:.3e} env-steps/s, frac {3141*n/(ms*1e-3)/1e9/6543.1:.3f}", flush=True)
    del envs, acts
