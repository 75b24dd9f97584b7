#!/usr/bin/env python
"""bench.py — VSS env-steps/s (and PPO SPS) on N B200s, with roofline, CPU baseline and e2e.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E_PER_GPU] [--impl native|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL only for timing)

A "step" is one pass of the hot path over one batch: one fused `vss_step` launch advancing E
fields per GPU by one control step (full VSS.step contract: obs + terminal obs + rewards + dones
+ masked reset). Fields shard across GPUs with no data-path collective (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic HBM bytes of one field-step of the full VSS.step contract:
# state 240 in + 240 out, actions 48, reset flags 8 in + 8 out, obs 1248, terminal obs 1248,
# rewards 96, timeout 1, progress 4
BYTES_PER_FIELD_STEP = 240 + 240 + 48 + 8 + 8 + 1248 + 1248 + 96 + 1 + 4
# sa view: state 480, action_buf 48 in + 48 out, policy action 8, reset 16, obs 208, term obs 208,
# rews 16, reward 4, done 8, timeout 1, progress 4
BYTES_PER_FIELD_STEP_SA = 480 + 96 + 8 + 16 + 208 + 208 + 16 + 4 + 8 + 1 + 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--envs", type=int, default=1 << 20, help="fields per GPU")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-envs", type=int, default=65536, help="fields per step of the CPU reference arm")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-ppo", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the N in {1K..1M} sweep (BASELINE configs[4])")
    ap.add_argument("--ppo-env-id", default="sa")
    ap.add_argument("--ppo-envs", type=int, default=4096, help="agents per GPU of the PPO leg (configs[1])")
    ap.add_argument("--ppo-updates", type=int, default=6)
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm (oracle)
def cpu_oracle_run(n_envs, seconds=None, steps=None, warmup=1):
    """Times the oracle's VSS.step (CPU restatement, all host threads) on a bounded sample."""
    import numpy as np
    from oracle import vss_oracle as orc
    p = orc.default_params()
    st = orc.State(n_envs)
    rb = np.ones(n_envs, np.int64)
    orc.reset_dones(p, 0, 0, st, rb)
    rb[:] = 0
    rng = np.random.default_rng(0)
    st.progress[:] = rng.integers(0, 400, n_envs)
    acts = [rng.uniform(-1, 1, (n_envs, 2, 3, 2)).astype(np.float32) for _ in range(4)]
    # use as many host threads as actually help (containers often expose more cpus than their quota)
    best = (None, 1)
    for nt in sorted({1, max(1, (os.cpu_count() or 1) // 2), os.cpu_count() or 1}):
        orc.set_num_threads(nt)
        orc.step(p, 0, 0, st, acts[0], rb)
        t0 = time.perf_counter()
        orc.step(p, 0, 0, st, acts[1], rb)
        dt = time.perf_counter() - t0
        if best[0] is None or dt < best[0]:
            best = (dt, nt)
    orc.set_num_threads(best[1])
    for i in range(warmup):
        orc.step(p, 0, 0, st, acts[i % 4], rb)
    t0 = time.perf_counter()
    k = 0
    while True:
        orc.step(p, 0, 0, st, acts[k % 4], rb)
        k += 1
        el = time.perf_counter() - t0
        if (steps is not None and k >= steps) or (steps is None and el >= seconds and k >= 3):
            break
    return {"value": n_envs * k / el, "unit": "env-steps/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"{k} oracle VSS.step calls over {n_envs} fields ({el:.1f} s, OpenMP over fields, "
                      f"{os.cpu_count()} host cpus); CPU restatement incl. the new 2-D physics, NOT PhysX"}, el / k


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, sec_per_step = cpu_oracle_run(args.ref_envs, steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {
        "impl": "reference", "metric": "vss_env_steps_per_s", "value": cb["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"random-action VSS 3v3 env step, full VSS.step contract, {args.ref_envs} fields "
                               "per step on the host cores (bounded sample of the native arm's workload)",
                   "note": "the reference's own CPU PhysX pipeline cannot run (IsaacGym absent); this is the "
                           "oracle port: reference obs/reward/reset semantics + the new 2-D physics"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- native arm
def time_steps(torch, dist, world, fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def make_task(torch, n, rank, local, seed=0):
    """The public task object (drop-in VSS) with progress counters staggered to steady state."""
    from rsoccer_isaac_cleanrl_b200.envs import VSS, load_cfg
    cfg = load_cfg()
    cfg["env"]["numEnvs"] = n
    dev = f"cuda:{local}"
    envs = VSS(cfg, dev, dev, 0, True, seed=seed, global_env_offset=rank * n)
    envs.reset_buf.zero_()
    st = envs.engine.get_state()
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    st[58, :n] = torch.randint(0, 400, (n,), device=dev, generator=g, dtype=torch.int32).view(torch.float32)
    envs.engine.set_state(st)
    del st
    acts = [torch.rand((n, 2, 3, 2), device=dev, generator=g) * 2 - 1 for _ in range(4)]
    return envs, acts


def run_native(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.envs
    envs, acts = make_task(torch, n, rank, local)

    def step_full(i):
        envs.step(acts[i & 3])  # the drop-in VSS.step: one fused vss_step launch

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = time_steps(torch, dist, world, step_full, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = world * n * args.steps / (ms * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = BYTES_PER_FIELD_STEP * n / (ms * 1e-3 / args.steps) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, scaled per field
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_step_traffic.json")))
        traffic = tr["dram_bytes_per_field_step"] * n
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_step<full>", "algorithmic_bytes_per_launch": BYTES_PER_FIELD_STEP * n,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"}

    launches = args.steps  # timed region: K x k_step (its last CTA advances the device-resident step index)

    # ---- GAE reverse scan (BASELINE roofline row: 28 B per (t, env)), T = 128, N = 65536
    gae_res = None
    if world == 1:
        from rsoccer_isaac_cleanrl_b200.engine import gae as gae_kernel
        # two sizes: T=128 x 65536 columns (configs[3] read as 65 536 agents: 235 MB per call, latency and
        # launch ramp still visible) and T=128 x 196608 (65 536 dma fields = 196 608 agents: 705 MB per call)
        gae_res = []
        for Tn, Nn in ((128, 65536), (128, 196608)):
            gin = [torch.randn((Tn, Nn), device="cuda") for _ in range(3)]
            gd = (torch.rand((Tn, Nn), device="cuda") < 0.01).float()
            gto = gd * (torch.rand((Tn, Nn), device="cuda") < 0.5).float()
            adv, ret = torch.empty_like(gd), torch.empty_like(gd)
            # >= 235 MB per call: larger than L2, so consecutive calls do not hit in cache
            ms_g = time_steps(torch, dist, 1, lambda i: gae_kernel(*gin, gd, gto, 0.99, 0.95, adv, ret), 50, 5)
            gbs = 28.0 * Tn * Nn / (ms_g * 1e-3 / 50) / 1e9
            gae_res.append({"elements_per_s": Tn * Nn * 50 / (ms_g * 1e-3), "us_per_call": ms_g * 1e3 / 50,
                            "achieved_gbs": gbs, "frac": gbs / peak,
                            "workload": f"vss_gae T={Tn} N={Nn}, 28 B per (t, env)"})
        del gin, gd, gto, adv, ret
    # ---- e2e: the user-facing call with HOST buffers (SingleAgent view): pinned policy action ->
    #      device, fused view step, view obs/reward/done -> pinned host, every step.
    e2e = None
    if not args.skip_e2e:
        from rsoccer_isaac_cleanrl_b200.envs import SingleAgent
        view = SingleAgent(envs)
        pa = [torch.rand((n, 2)).mul_(2).sub_(1).pin_memory() for _ in range(2)]
        k_e2e = max(10, min(args.steps, 100))
        chunks = view.HOST_CHUNKS if n >= view.HOST_CHUNK_MIN_FIELDS else 1
        variants = {}
        for name, kw in (("bf16_direct", dict(obs_dtype=torch.bfloat16, host_write="direct")),
                         ("bf16_staged", dict(obs_dtype=torch.bfloat16, host_write="staged")),
                         ("f32_copies", dict(obs_dtype=torch.float32))):
            ms_v = time_steps(torch, dist, world, lambda i: view.step_host(pa[i & 1], **kw), k_e2e, 3)
            variants[name] = {"value": world * n * k_e2e / (ms_v * 1e-3), "ms_per_step": ms_v / k_e2e,
                              "d2h_bytes_per_step": view.d2h_bytes(kw["obs_dtype"])}
            launches += (k_e2e + 3) * chunks
        # the ceiling: a plain cudaMemcpyAsync of the same bytes, device -> pinned host, all ranks at once
        from rsoccer_isaac_cleanrl_b200.hostmem import pinned_empty
        nbytes = view.d2h_bytes(torch.bfloat16)
        src, dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), pinned_empty((nbytes,), torch.uint8, f"cuda:{local}")
        ms_c = time_steps(torch, dist, world, lambda i: dst.copy_(src, non_blocking=True), 20, 3)
        d2h_gbs = nbytes * 20 / (ms_c * 1e-3) / 1e9
        del src, dst
        best = variants["bf16_direct" if view.HOST_WRITE == "direct" else "bf16_staged"]
        e2e = {"value": best["value"], "unit": "env-steps/s",
               "h2d_bytes_per_step": view.h2d_bytes_per_step, "d2h_bytes_per_step": best["d2h_bytes_per_step"],
               "steps": k_e2e, "ms_per_step": best["ms_per_step"],
               "api": "SingleAgent.step_host(pinned policy action (N,2)) -> pinned packed rows (N,112 B): obs (N,52) "
                      "bf16, reward (N) f32, done (N) u8, time-out (N) u8 (what envs/wrappers.py:108-115 returns to "
                      f"the policy); the step runs as {chunks} field ranges (vss_set_step_range) on two streams; "
                      f"host_write={view.HOST_WRITE}; host sync every step",
               "variants": variants,
               "d2h_memcpy_gbs_per_gpu": d2h_gbs,
               "d2h_frac_of_memcpy": best["d2h_bytes_per_step"] / (best["ms_per_step"] * 1e-3) / 1e9 / d2h_gbs}

    sweep = None
    if not args.no_sweep:
        # BASELINE configs[4]: random-action step throughput from 1K to 1M fields per GPU (same full
        # VSS.step contract; sizes below ~64K fields fit in L2 and are launch/latency bound)
        sweep = []
        del envs, acts
        for ne in (1024, 4096, 16384, 65536, 262144, 1048576):
            e2, a2 = make_task(torch, ne, rank, local)
            f = lambda i: e2.step(a2[i & 3])
            m = time_steps(torch, dist, world, f, 200, 20)
            sweep.append({"envs_per_gpu": ne, "env_steps_per_s": world * ne * 200 / (m * 1e-3),
                          "us_per_step": m * 1e3 / 200,
                          "hbm_frac": BYTES_PER_FIELD_STEP * ne / (m * 1e-3 / 200) / 1e9 / peak})
            del e2, a2

    # ---- PPO SPS (BASELINE configs[1]: ppo-sa, 4096 envs per GPU, OU-noise opponents as in training)
    ppo_res = None
    if not args.skip_ppo:
        from rsoccer_isaac_cleanrl_b200 import ppo as ppo_mod
        try:
            del envs, acts
        except NameError:
            pass
        torch.cuda.empty_cache()
        pa_ = ppo_mod.parse_args(["--env-id", args.ppo_env_id, "--num-envs", str(args.ppo_envs), "--quiet",
                                  "--total-timesteps", str(args.ppo_envs * 128 * world * (args.ppo_updates + 1))])
        st = ppo_mod.train(pa_)
        sps_list = st["sps"]
        # SPS exactly as ppo…:376 (global_step / wall since start), plus the steady-state rate of the
        # updates after the first (which pays one-off allocation and cuBLAS/NCCL initialisation)
        # steady state: updates after the eager first one and the graph-capturing second one
        tail_r, tail_u = st["rollout_wall"][2:] or st["rollout_wall"], st["update_wall"][2:] or st["update_wall"]
        per_update = sorted(a + b for a, b in zip(tail_r, tail_u))[len(tail_r) // 2]
        ppo_res = {"sps": st["final_sps"], "unit": "samples/s (global_step / wall, ppo…:257,376), all GPUs",
                   "sps_steady": args.ppo_envs * 128 * world / per_update,
                   "steady_note": "median over updates >= 3 (update 1 runs eagerly, update 2 records the CUDA graphs)",
                   "rollout_s_steady": sorted(tail_r)[len(tail_r) // 2], "update_s_steady": sorted(tail_u)[len(tail_u) // 2],
                   "workload": f"ppo-{args.ppo_env_id}, {args.ppo_envs} agents per GPU x 128 steps, 4 minibatches x "
                               f"8 epochs, {st['updates']} updates, defaults of ppo…:71-108",
                   "rollout_s": st["rollout_s"], "update_s": st["update_s"], "wall_s": st["wall"],
                   "mlp_backend": st.get("mlp_backend", "torch-fp32"), "n_gpus": world}
        del st

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cpu_baseline, _ = cpu_oracle_run(32768, seconds=args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": "vss_env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"random-action VSS 3v3 env step (BASELINE configs[4] sweep top), full VSS.step "
                                   f"contract, {n} fields per GPU, U(-1,1) actions resident in HBM, progress "
                                   "counters staggered to the steady-state reset rate",
                       "envs_per_gpu": n, "l2": f"per-step working set {BYTES_PER_FIELD_STEP * n / 1e6:.0f} MB "
                                               "(> 126 MB L2 when envs_per_gpu >= 65536)",
                       "parallelism": f"fields sharded over {world} GPU(s), no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "ppo": ppo_res, "gae": gae_res,
        }
        if sweep:
            line["sweep"] = sweep
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The ONE JSON line goes to the real stdout; everything libraries print (e.g. NCCL's version
    banner) was diverted to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
