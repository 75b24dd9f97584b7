"""Where the time of a small-batch step launch goes: per-CTA clock totals of the phases of k_step_cta.
Needs the profiling build of the library (NOT the product build):
   make -C rsoccer_isaac_cleanrl_b200/csrc -B EXTRA=-DVSS_PHASE_PROFILE && python profiles/step_phase_profile.py [fields] ; make -C ... -B
Prints, over the CTAs of one launch in steady state: mean / p90 / max cycles per phase, and the phase split of the slowest CTAs."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsoccer_isaac_cleanrl_b200 import _lib  # noqa: E402
from rsoccer_isaac_cleanrl_b200.envs import VSS, SingleAgent, load_cfg  # noqa: E402

NAMES = ["load+actions", "integrate", "broadphase", "contacts", "walls", "outputs+obs+store", "total"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    if os.environ.get("VSS_AB_LIB"):  # the profiling build as a second file (this script only)
        _lib.LIB_PATH = os.path.abspath(os.environ["VSS_AB_LIB"])
    lib = _lib.load_library()
    assert hasattr(lib, "vss_prof_read"), "not the profiling build"
    cfg = load_cfg()
    cfg["env"]["numEnvs"] = n
    task = VSS(cfg, "cuda:0", "cuda:0", 0, True, seed=1)
    if len(sys.argv) > 2:
        task.engine.fields_per_tile = int(sys.argv[2])
    if len(sys.argv) > 3:
        task.engine.warps_per_tile = int(sys.argv[3])
    view = SingleAgent(task)
    view.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for i in range(600):  # steady state: robots spread to the walls, episodes at all ages
        view.step(torch.rand((n, 2), device="cuda", generator=g) * 2 - 1)
    torch.cuda.synchronize()
    ctas = (n + task.engine.fields_per_tile - 1) // task.engine.fields_per_tile
    acc = np.zeros((ctas, 8), np.float64)
    reps = 20
    for i in range(reps):
        view.step(torch.rand((n, 2), device="cuda", generator=g) * 2 - 1)
        torch.cuda.synchronize()
        buf = np.zeros((ctas, 8), np.uint64)
        assert lib.vss_prof_read(buf.ctypes.data_as(C.c_void_p), ctas) == 0
        acc += buf
        if i == reps - 1:
            last = buf.astype(np.float64)
    acc /= reps
    print(f"{n} fields, {ctas} CTAs of {task.engine.warps_per_tile} warps, {task.engine.fields_per_tile} fields per tile; cycles (SM clock)")
    print(f"{'phase':20s} {'mean':>9s} {'p90':>9s} {'max':>9s}   (one launch: last of {reps})")
    for k, name in enumerate(NAMES):
        c = last[:, k]
        print(f"{name:20s} {c.mean():9.0f} {np.percentile(c, 90):9.0f} {c.max():9.0f}")
    order = np.argsort(-last[:, 6])[:5]
    print("slowest CTAs of that launch: " + "; ".join(
        f"#{i}: " + " ".join(f"{int(last[i, k])}" for k in range(7)) for i in order))
    if hasattr(lib, "vss_prof_read_integrate"):
        ib = np.zeros((ctas, 16), np.uint64)
        assert lib.vss_prof_read_integrate(ib.ctypes.data_as(C.c_void_p), ctas) == 0
        ib = ib.astype(np.float64)
        print("integrate, per warp (cycles per step, mean over CTAs): compute " + " ".join(f"{x:.0f}" for x in ib[:, :8].mean(0)) +
              " | wait at the barrier " + " ".join(f"{x:.0f}" for x in ib[:, 8:].mean(0)))
    print("mean over launches of the per-launch MAX total:", end=" ")
    print("(per-phase mean over CTAs and launches) " + " ".join(f"{NAMES[k]}={acc[:, k].mean():.0f}" for k in range(7)))


if __name__ == "__main__":
    main()
