"""Tuning run: step-kernel time per launch shape (warps per tile) and batch size, full contract and SA view.
   python profiles/step_shape_bench.py > gpurun_out/step_shapes.txt
Each point: 300 launches after 30 warm-up launches, CUDA events, inputs resident in HBM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200 import _lib  # noqa: E402
from rsoccer_isaac_cleanrl_b200.envs import VSS, SingleAgent, load_cfg  # noqa: E402

if os.environ.get("VSS_AB_LIB"):  # A/B runs of experimental builds on the SAME box (box-to-box spread is ~10 %): this
    _lib.LIB_PATH = os.path.abspath(os.environ["VSS_AB_LIB"])  # script only - the product never reads the environment


def time_us(fn, reps=300, warm=30):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else
                              "1024,4096,8192,16384,21845,32768,49152,65536,98304,131072,262144".split(","))]
    # a shape is "warps per tile" or "warps per tile x fields per tile" (8x16); fields default to the automatic choice
    shapes = [tuple(int(y) for y in x.split("x")) if "x" in x else (int(x), 0)
              for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,2,3,4,7,8".split(","))]
    print("fields  contract  " + "  ".join(f"{w}x{f:<3d}" for w, f in shapes) + "   (us per launch; * = automatic shape)")
    for n in sizes:
        cfg = load_cfg()
        cfg["env"]["numEnvs"] = n
        task = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=1)
        task.reset_buf.zero_()
        st = task.engine.get_state()
        st[58, :n] = torch.randint(0, 400, (n,), device="cuda", dtype=torch.int32).view(torch.float32)
        task.engine.set_state(st)
        auto = (task.engine.warps_per_tile, task.engine.fields_per_tile)
        acts = [torch.rand((n, 2, 3, 2), device="cuda") * 2 - 1 for _ in range(4)]
        view = SingleAgent(task)
        pas = [torch.rand((n, 2), device="cuda") * 2 - 1 for _ in range(4)]
        rows = {"full": [], "sa": []}
        for w, f in shapes:
            if w > 1 and n > 400000:
                rows["full"].append(float("nan")); rows["sa"].append(float("nan"))
                continue
            task.engine.warps_per_tile = w
            task.engine.fields_per_tile = f
            rows["full"].append(time_us(lambda i: task.step(acts[i & 3])))
            rows["sa"].append(time_us(lambda i: view.step(pas[i & 3])))
        for k, v in rows.items():
            print(f"{n:7d}  {k:8s}  " + "  ".join(f"{x:6.1f}{'*' if (w, f) == auto or (f == 0 and w == auto[0]) else ' '}" for x, (w, f) in zip(v, shapes)))
        sys.stdout.flush()
        del task, view, acts, pas, st


if __name__ == "__main__":
    main()
