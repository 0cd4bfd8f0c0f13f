#!/usr/bin/env python
"""bench.py — converged patches/s of the patch-refinement hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm on the host cores

A "step" = one pass of the hot path over one batch: `--patches` synthetic TYPE_EXPAND candidates per GPU go through
Patch::refine() + removeInvisibleCamera() (seam 2 of include/pmvs_b200.h). Workload at N=1 = BASELINE.json configs[1]:
5 views 1600x1200, patchRadius 15, 3 pyramid levels, adaptive distance+difference weights, README sample swarm
(15 particles x 30 iterations). `value` times the device-resident call (inputs already in HBM, CUDA events on the
launching stream, L2 flushed between steps); `e2e` times the host-buffer C-ABI call including H2D/D2H copies.
N>1: every rank refines its own shard of the candidates (weak scaling) and the ranks all-gather the converged patch
records over NCCL after each pass (the exchange step between expansion rounds).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pais-mvs_b200", "python"))

import numpy as np  # noqa: E402

from pmvs_b200 import abi, scene, shard  # noqa: E402

METRIC = "converged patches/sec (r=15, 5 views 1600x1200)"
UNIT = "patches/s"
WORKLOAD = "configs[1]: 5 views 1600x1200, patchRadius=15, 3 pyramid levels, adaptive distance+difference on"


def bench_config():
    cfg = abi.readme_config()        # README.md:110-207 sample config.txt over TMVS.cpp:26-52 defaults
    cfg.patchRadius = 15
    cfg.patchSize = 31
    cfg.distWeighting = 5.0
    cfg.maxLOD = 2                   # 3 pyramid levels
    cfg.adaptiveDistanceEnable, cfg.adaptiveDifferenceEnable, cfg.adaptiveGradientEnable = 1, 1, 0
    return cfg


def make_scene(cfg, views=5, width=1600, height=1200):
    return scene.SynthScene(cfg, nviews=views, width=width, height=height, seed=1234)


def alg_bytes_per_eval(cfg, V):
    """the loads the reference algorithm issues per getFitness call (patch.cpp:986, :1014-1017,
    :1031, :1037): 4 u8 taps per sample and view, the mask byte, the f64 distance weight, the f64 edge value."""
    S = cfg.patchSize * cfg.patchSize
    return 4 * S * V + S + 8 * S * int(cfg.adaptiveDistanceEnable) + 8 * S * int(cfg.adaptiveGradientEnable)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_per_launch(args, n):
    """dram read+write bytes of one refine_kernel launch from the committed ncu capture (profiles/r1_traffic.json),
    valid only for the launch size it was captured with; --traffic overrides."""
    if args.traffic is not None:
        return args.traffic
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        if t.get("patches") == n:
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    return None


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, sc, n_patches, threads, seed=42, first_id=0, patch_seed=5678):
    """The reference's CPU algorithm on host cores: f64 restatement of patch.cpp driven by the UNMODIFIED reference
    solver when oracle/_ref is present (else the restated solver), patches spread over `threads` OpenMP threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    use_ref = orc.ref_lib() is not None
    o = orc.Oracle(cfg, sc.records, seed=seed, use_ref_pso=use_ref)
    ps = sc.patches(n_patches, seed=patch_seed, first_id=first_id)
    t0 = time.perf_counter()
    out = o.refine_batch(ps, flags=abi.F_POST_REMOVE_INVISIBLE, patch_threads=threads)
    dt = time.perf_counter() - t0
    kept = sum(1 for q in out if not q.drop)
    return dt, kept, ("unmodified reference PSO (oracle/_ref) + f64 restatement of patch.cpp" if use_ref else
                      "f64 restatement (oracle/) of patch.cpp + psosolver.cpp")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = bench_config()
    sc = make_scene(cfg, args.views, args.width, args.height)
    cores = os.cpu_count() or 1
    n = args.cpu_patches if args.cpu_patches > 0 else max(cores * 128, 64)
    times = []
    for s in range(args.warmup + args.steps):
        dt, kept, how = cpu_reference_run(cfg, sc, n, cores, patch_seed=5678 + s)
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "patches_per_step": n, "particles": cfg.particleNum, "iterations": cfg.maxIteration},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d candidate patches per step, OpenMP over patches on %d host threads; %s" % (n, cores, how)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pmvs_b200 import lib as pmvs_lib
    from pmvs_b200.api import PatchRefiner

    if not os.path.exists(pmvs_lib.LIB_PATH) and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        pmvs_lib.build()            # fresh checkout: compile the CUDA library (there is still no CPU path)
    for _ in range(600):            # other ranks wait for rank 0's build
        if os.path.exists(pmvs_lib.LIB_PATH):
            break
        time.sleep(0.5)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = bench_config()
    if args.particles > 0:
        cfg.particleNum = args.particles
    if args.iterations > 0:
        cfg.maxIteration = args.iterations
    sc = make_scene(cfg, args.views, args.width, args.height)
    V = len(sc.cams)
    n = args.patches
    pr = PatchRefiner(cfg, sc.records, device=local, seed=42)

    in_bytes, out_bytes = C.sizeof(abi.PmvsPatchIn) * n, C.sizeof(abi.PmvsPatchOut) * n
    total_steps = args.warmup + args.steps
    # a different candidate set per step; resident in HBM before the timed region
    host_in = [sc.patches(n, seed=5678 + 1000 * rank + s, first_id=(rank * total_steps + s) * n) for s in range(total_steps)]
    d_in = [torch.frombuffer(bytearray(bytes(h)), dtype=torch.uint8).to(dev) for h in host_in]
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)          # a real (non-NULL) stream: kernel, events and NCCL all on it
    torch.cuda.set_stream(stream)
    flags = abi.F_POST_REMOVE_INVISIBLE

    fit_off, drop_off = abi.PmvsPatchOut.fitness.offset, abi.PmvsPatchOut.drop.offset

    def exchange():
        """Exchange step between expansion rounds: every rank's converged records (centre, normal, fitness, drop)."""
        rec = d_out.view(n, C.sizeof(abi.PmvsPatchOut))
        geo = rec[:, :48].contiguous().view(torch.float64)
        fit = rec[:, fit_off:fit_off + 8].contiguous().view(torch.float64)
        drp = rec[:, drop_off:drop_off + 4].contiguous().view(torch.int32).double()
        return shard.allgather_records(torch.cat([geo, fit, drp], dim=1), world * n, rank, world)

    def one_pass(s):
        pr.refine_device(n, d_in[s].data_ptr(), d_out.data_ptr(), flags, stream=stream.cuda_stream)
        if world > 1:
            exchange()

    for s in range(args.warmup):
        one_pass(s)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    step_ms, kernel_ms, evals, wevals, kept = [], [], 0, 0, 0
    launches0 = pr.launch_count()
    for s in range(args.warmup, total_steps):
        flush.zero_()                                                      # L2 flush between timed iterations
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        pr.refine_device(n, d_in[s].data_ptr(), d_out.data_ptr(), flags, stream=stream.cuda_stream)
        e1.record(stream)
        if world > 1:
            exchange()
        e2.record(stream)
        e2.synchronize()
        step_ms.append(e0.elapsed_time(e2))
        kernel_ms.append(e0.elapsed_time(e1))
        o = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=scene.PATCH_OUT_DTYPE)
        evals += int(o["evaluations"].sum())
        wevals += int(o["windowEvaluations"].sum())
        kept += int((o["drop"] == 0).sum())
    torch.cuda.synchronize()
    launches = pr.launch_count() - launches0          # kernels of this library launched inside the timed region
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    counts = torch.tensor([evals, wevals, kept], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_s = float(total_ms.item()) / 1e3
    value = world * n * args.steps / total_s

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed call)
    pin_in = torch.empty(in_bytes, dtype=torch.uint8).pin_memory()
    pin_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    e2e_t = []
    for s in range(0 if args.no_e2e else total_steps):
        pin_in.numpy()[:] = np.frombuffer(bytes(host_in[s]), dtype=np.uint8)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = pr.L.pmvs_refine_batch(pr.h, n, C.cast(pin_in.data_ptr(), C.POINTER(abi.PmvsPatchIn)),
                                    C.cast(pin_out.data_ptr(), C.POINTER(abi.PmvsPatchOut)), flags)
        dt = time.perf_counter() - t0
        pr._check(rc)
        if s >= args.warmup:
            e2e_t.append(dt)
    e2e_total = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(e2e_total.item()) if e2e_t else None

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bpe = alg_bytes_per_eval(cfg, V)
        k_ms = sum(kernel_ms) / len(kernel_ms)
        wev_per_launch = counts[1].item() / (world * args.steps)
        achieved = bpe * wev_per_launch / (k_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "patches_per_step_per_gpu": n, "views": V, "particles": cfg.particleNum,
                           "iterations": cfg.maxIteration, "l2": "flushed between timed steps (256 MiB memset)",
                           "converged_kept_fraction": counts[2].item() / (world * n * args.steps),
                           "evaluations_per_patch": counts[0].item() / (world * n * args.steps),
                           "parallelism": "patches sharded by index over %d GPU(s); all-gather of converged records per pass" % world},
                "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "refine_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic_per_launch(args, n),
                             "peak_source": peak_src,
                             "alg_bytes_per_eval": bpe, "window_evals_per_launch": wev_per_launch, "kernel_ms": k_ms,
                             "note": "algorithmic tap-bytes model of SURVEY.md 8(d); not HBM-bound: issue slots 49 %, FP64 pipe 32 %, DRAM 1 % of peak — see DESIGN.md section 2"}}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ncpu = args.cpu_patches if args.cpu_patches > 0 else max(cores * 512, 256)     # ~10-20 s of CPU work
            dt, ckept, how = cpu_reference_run(cfg, sc, ncpu, cores)
            line["cpu_baseline"] = {"value": ncpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d candidate patches of the same workload, OpenMP over patches on %d host threads, %.1f s; %s"
                                              % (ncpu, cores, dt, how)}
        print(json.dumps(line))
    pr.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patches", type=int, default=16384, help="candidate patches per GPU per step")
    ap.add_argument("--views", type=int, default=5)
    ap.add_argument("--particles", type=int, default=0, help="override particleNum (0 = README config, 15)")
    ap.add_argument("--iterations", type=int, default=0, help="override maxIteration (0 = README config, 30)")
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--cpu-patches", type=int, default=0, help="CPU sample size (0 = 128 per host core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs: skip the host-buffer end-to-end leg")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from an ncu --set full capture")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: note: fewer than 3 warm-up steps", file=sys.stderr)
    return run_reference(args) if args.impl == "reference" else run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
