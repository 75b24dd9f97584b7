import torch,sys,os
sys.path.insert(0,os.getcwd())
from rsoccer_isaac_cleanrl_b200.engine import gae
for N in (4096,65536,196608):
    T=128
    a=[torch.randn((T,N),device="cuda") for _ in range(3)]
    d=(torch.rand((T,N),device="cuda")<0.01).float(); to=d*(torch.rand((T,N),device="cuda")<0.5).float()
    adv,ret=torch.empty_like(d),torch.empty_like(d)
    for i in range(5): gae(*a,d,to,0.99,0.95,adv,ret)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): gae(*a,d,to,0.99,0.95,adv,ret)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/50
    print(64, N, "%.1f us"%(ms*1e3), "%.0f GB/s"%(28*T*N/ms/1e6), "%.3f"%(28*T*N/ms/1e6/6543.1))
