"""Where a small-batch step spends its time: CUDA-graph-timed launches (no host overhead) as a function of
the substep count (0 = no physics at all) and the launch shape.   python profiles/step_decompose.py [fields]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200.envs import VSS, SingleAgent, load_cfg  # noqa: E402


def graph_time_us(fn, per_graph=50, reps=6):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(per_graph):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * per_graph)


def main():
    sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1024", "4096", "16384", "65536"])]
    shapes = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "4", "7"])]
    for n in sizes:
        for sub in (0, 1, 2, 4):
            cfg = load_cfg()
            cfg["env"]["numEnvs"] = n
            cfg.setdefault("sim", {})["substeps"] = sub
            task = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=1)
            task.reset_buf.zero_()
            st = task.engine.get_state()
            st[58, :n] = torch.randint(0, 400, (n,), device="cuda", dtype=torch.int32).view(torch.float32)
            task.engine.set_state(st)
            acts = [torch.rand((n, 2, 3, 2), device="cuda") * 2 - 1 for _ in range(4)]
            view = SingleAgent(task)
            pas = [torch.rand((n, 2), device="cuda") * 2 - 1 for _ in range(4)]
            out = []
            for w in shapes:
                task.engine.warps_per_tile = w
                eng, v = task.engine, view
                full = graph_time_us(lambda i: eng.step(acts[i & 3], task.reset_buf, task.obs_buf, task.terminal_obs_buf,
                                                        task.rew_buf, task._timeout_u8, task._progress_f))
                sa = graph_time_us(lambda i: v.step(pas[i & 3]))
                out.append(f"wpt={w}: full {full:6.1f} sa {sa:6.1f}")
            print(f"{n:7d} fields, {sub} substeps | " + " | ".join(out))
            sys.stdout.flush()
            del task, view


if __name__ == "__main__":
    main()
