"""The rollout's per-step policy + value forward: one fused launch (vss_mlp_forward_fused) against the two
layer-by-layer chains side by side on two streams (2 x (4 GEMMs + head)), both replayed from CUDA graphs of 20 steps.
   python profiles/mlp_fused_bench.py [rows,rows,...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from rsoccer_isaac_cleanrl_b200 import _lib  # noqa: E402
from rsoccer_isaac_cleanrl_b200.engine import gather_pad_bf16, mlp_forward_fused  # noqa: E402

if os.environ.get("VSS_AB_LIB"):  # A/B runs of experimental builds on the same box (this script only)
    _lib.LIB_PATH = os.path.abspath(os.environ["VSS_AB_LIB"])
from rsoccer_isaac_cleanrl_b200.tc_mlp import forward_explicit  # noqa: E402
from test_gpu_gemm import _mlp_pair  # noqa: E402

REPS = 20


def graph_time_us(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (10 * REPS)


def main():
    sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1024,4096,16384,65535,131072".split(","))]
    for n_act in ((2,) if os.environ.get("VSS_AB_LIB") else (2, 6)):
        actor, critic = _mlp_pair(n_act, 0)
        for M in sizes:
            x16 = gather_pad_bf16(torch.randn(M, 52, device="cuda"), None, 64)
            oa = torch.empty((M, n_act), device="cuda"); oc = torch.empty((M, 1), device="cuda")
            nets = [(mw.w16, [b.detach() for b in mw.bs], mw.head_w.detach(), mw.head_b.detach(), o)
                    for mw, o in ((actor, oa), (critic, oc))]
            side = torch.cuda.Stream()

            def chains():
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    forward_explicit(critic, x16, out=oc)
                forward_explicit(actor, x16, out=oa)
                torch.cuda.current_stream().wait_stream(side)

            t_chain = graph_time_us(chains)
            t4 = graph_time_us(lambda: mlp_forward_fused(x16, nets, epilogue_warps=4))
            t8 = graph_time_us(lambda: mlp_forward_fused(x16, nets, epilogue_warps=8))
            if M == sizes[0] or M == 4096:
                st = torch.zeros(11, device="cuda", dtype=torch.int64)
                for ew in (4, 8):
                    for _ in range(3):
                        mlp_forward_fused(x16, nets, epilogue_warps=ew, stamps=st)
                    torch.cuda.synchronize()
                    t = (st - st[0]).tolist()
                    print(f"   first CTA, {ew} epilogue warps, ns since entry: setup {t[1]} | " +
                          " | ".join(f"L{l} acc {t[2 + 2 * l]} epi {t[3 + 2 * l]}" for l in range(4)) + f" | exit {t[10]}")
            flop = 2 * 2 * M * (64 * 256 + 256 * 512 + 512 * 512 + 512 * 256)
            print(f"rows {M:7d} head {n_act}+1: two chains {t_chain:7.1f} us | fused 4 epilogue warps {t4:7.1f} us, 8: {t8:7.1f} us "
                  f"({flop / t8 * 1e-6:6.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
