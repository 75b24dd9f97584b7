import os, sys, torch
sys.path.insert(0, "/root/repo")
from rsoccer_isaac_cleanrl_b200.engine import head_backward, head_forward
M=131072
for no in (1,2,6):
    h=torch.tanh(torch.randn(M,256,device="cuda")).to(torch.bfloat16); W=torch.randn(no,256,device="cuda")*0.1; b=torch.zeros(no,device="cuda")
    dout=torch.randn(M,no,device="cuda"); dW=torch.zeros(no,256,device="cuda"); db=torch.zeros(no,device="cuda"); cs=torch.zeros(256,device="cuda")
    for fn,name in ((lambda: head_backward(dout,h,W,dW=dW,db=db,dz_colsum=cs),"bwd"),(lambda: head_forward(h,W,b),"fwd")):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize(); print(no,name,f"{e0.elapsed_time(e1)/20*1e3:.1f} us")
