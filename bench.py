#!/usr/bin/env python
"""bench.py — VOLDOR EM hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cpu-port|abi-dropin]

One "step" = one VO window (one pass of the hot path over one batch of synthetic input): BASELINE.json
configs[1] — a single 640x480 frame with 8 flows, 30 EM iterations, monocular (bootstrap pose/depth
injected), 8192 pose hypotheses per camera, window truncation disabled.  metric = EM
iterations per second, whole job (all ranks).  With N > 1 each rank runs its own independent window
(window-per-GPU) and contributes its poses to one NCCL all_gather per step; "scaling": "weak".

  value : windows whose inputs already live in HBM when the timed region starts (device pointers are passed
          to the same C-ABI call; every bulk copy in the library is cudaMemcpyDefault)
  e2e   : the same call with pinned HOST buffers: H2D of the flows and D2H of poses/covariances/depth/
          confidence are inside the timed region
  roofline : dominant kernel = fused cost + random depth search, algorithmic bytes W*H*(12N+60) per launch
             (flows 8N, rigidness weights 4N, depth r/w 8, cost w 4, XORWOW state r/w 48) over
             its CUDA-event duration measured live in the library (vb_profile_*), against MEASURED_PEAKS.json
  cpu_baseline : the CPU port (oracle/cpu_kernels.cpp, OpenMP) under the restated reference orchestration on
             a bounded sample (same window, 2 EM iterations)

--impl reference runs the REFERENCE arm: the reference's own CUDA kernels (oracle/_ref, its unmodified .cu
files rebuilt for sm_100a) driven by the reference's host orchestration restated OpenCV-free over the ABI
(oracle/host_voldor.cpp) — i.e. voldor.cpp's geometry loop on the host cores calling its gpu-kernels
library, which is the only implementation of this path the reference has (there is no CPU E/M
step in the reference).  --impl cpu-port times the CPU port instead; --impl abi-dropin times the same reference
host orchestration over THIS library's gpu_kernels.h entry points (what a maintainer gets by only re-linking).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, NFLOWS, EM_ITERS, NPOSES = 640, 480, 8, 30, 8192
CONFIG = f"--silent --max_iters {EM_ITERS} --no_trunc_iters 1000 --n_poses_to_sample {NPOSES}"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, n in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(rank):
    import synth

    win = synth.make_window(W, H, NFLOWS, seed=100 + rank)
    boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05, seed=7 + rank))
    return win, boot


def run_cpu_sample(win, boot, iters=2):
    """bounded CPU sample: same window, `iters` EM iterations, OpenMP over all host cores"""
    import oracle_host

    cfg = f"--silent --max_iters {iters} --no_trunc_iters 1000 --n_poses_to_sample {NPOSES}"
    t0 = time.time()
    r = oracle_host.run_window("cpu", win["flows"], win["fx"], win["fy"], win["cx"], win["cy"], config=cfg, boot=boot)
    dt = time.time() - t0
    return r["iters"] / dt, dt, r


def bench_reference(args, rank, world):
    """reference arm: rank 0 only"""
    if rank != 0:
        return
    import ffi
    import oracle_host

    win, boot = make_inputs(0)
    cores = os.cpu_count()
    if args.impl == "cpu-port" or not os.path.exists(ffi.REF):
        times, iters = [], 0
        for i in range(args.warmup + args.steps):
            v, dt, r = run_cpu_sample(win, boot, 1)
            if i >= args.warmup:
                times.append(dt)
                iters += r["iters"]
        total = sum(times)
        value = iters / total
        kind, sample = "port", f"one {W}x{H}x{NFLOWS} window, 1 EM iteration per step, CPU port with OpenMP on {cores} cores"
        ms_per_step = 1e3 * total / max(1, args.steps)
    else:
        import torch

        torch.cuda.set_device(0)
        clocks = ClockSampler(0)
        times, iters = [], 0
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                clocks.start()
            ffi.libc_srand(1000 + i)
            torch.cuda.synchronize()
            t0 = time.time()
            backend = "ours_abi" if args.impl == "abi-dropin" else "ref"
            r = oracle_host.run_window(backend, win["flows"], win["fx"], win["fy"], win["cx"], win["cy"],
                                       config=CONFIG, boot=boot)
            torch.cuda.synchronize()
            if i >= args.warmup:
                times.append(time.time() - t0)
                iters += r["iters"]
        total = sum(times)
        value = iters / total
        kind = "reference"
        if args.impl == "abi-dropin":
            kind = "abi-dropin"
        sample = (f"full workload: {args.steps} windows of {W}x{H}x{NFLOWS}, {EM_ITERS} EM iterations each; reference "
                  "CUDA kernels (its own .cu files rebuilt for sm_100a, oracle/_ref) on 1 B200 driven by the reference "
                  "host orchestration (voldor.cpp/geometry.cpp restated OpenCV-free) on 1 host thread")
        if args.impl == "abi-dropin":
            sample = sample.replace("reference CUDA kernels (its own .cu files rebuilt for sm_100a, oracle/_ref)",
                                    "voldor_b200's gpu_kernels.h entry points (library-level drop-in, host buffers)")
        ms_per_step = 1e3 * total / max(1, args.steps)
        clk = clocks.stop()
    line = {
        "impl": "reference" if args.impl != "abi-dropin" else "abi-dropin", "metric": "EM-iters/sec", "value": value, "unit": "EM-iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: single {W}x{H} frame, {NFLOWS} flows, {EM_ITERS} EM iters, "
                               f"monocular, {NPOSES} hypotheses/camera", "parallelism": "single window (rank 0)"},
        "cpu_baseline": {"value": value, "unit": "EM-iterations/s", "cores": 1 if kind != "port" else cores,
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "EM-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_sec": value * NFLOWS / EM_ITERS,
    }
    if kind != "port":
        line["clocks"] = clk
    print(json.dumps(line))


def bench_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import voldor_b200
    from voldor_b200 import pyvoldor_vo as pv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = voldor_b200.load_library()
    lib.vb_profile_enable.argtypes = [C.c_int]
    lib.vb_profile_get.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    win, boot = make_inputs(rank)
    voldor_b200.set_bootstrap_override(*boot)
    fx, fy, cx, cy = float(win["fx"]), float(win["fy"]), float(win["cx"]), float(win["cy"])

    flows_pinned = torch.from_numpy(win["flows"]).pin_memory()
    flows_dev = flows_pinned.to(dev)
    out_h = {k: torch.zeros(s, dtype=torch.float32).pin_memory() for k, s in
             (("poses", (NFLOWS, 6)), ("covar", (NFLOWS, 36)), ("depth", (H, W)), ("conf", (H, W)))}
    out_d = {k: torch.zeros_like(v, device=dev) for k, v in out_h.items() if k in ("depth", "conf")}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    FP = C.POINTER(C.c_float)

    def p(t):
        return C.cast(t.data_ptr(), FP)

    def window(resident):
        n, it = C.c_int(0), C.c_int(0)
        stats = np.zeros(4, np.float32)
        fl = flows_dev if resident else flows_pinned
        depth = out_d["depth"] if resident else out_h["depth"]
        conf = out_d["conf"] if resident else out_h["conf"]
        lib.vb_py_voldor_wrapper_ex(p(fl), None, None, None, None, None, fx, fy, cx, cy, 0.0, NFLOWS, 0, W, H,
                                    CONFIG.encode(), C.byref(n), p(out_h["poses"]), p(out_h["covar"]), p(depth),
                                    p(conf), C.byref(it), stats.ctypes.data_as(FP))
        return n.value, it.value, stats

    gathered = [torch.zeros(NFLOWS * 6, device=dev) for _ in range(world)] if world > 1 else None

    def timed(resident, steps, warmup, clocks=None):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        iters = frames = 0
        wall = 0.0
        for i in range(warmup + steps):
            if i == warmup:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                if clocks:
                    clocks.start()
            flush.fill_(i & 255)  # evict L2 between steps (outside the timed region)
            torch.cuda.synchronize()
            k = i - warmup
            if k >= 0:
                ev[k][0].record()
            t0 = time.time()
            n, it, _ = window(resident)
            if world > 1:  # the single collective of the path: all ranks learn all poses
                dist.all_gather(gathered, out_h["poses"].to(dev, non_blocking=True).reshape(-1))
            if k >= 0:
                ev[k][1].record()
                torch.cuda.synchronize()
                wall += time.time() - t0
                iters += it
                frames += n
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        ms = max(dev_ms, 1e3 * wall)  # the call is synchronous: both clocks see the same region
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        cnt = torch.tensor([iters, frames], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        return float(t.item()), float(cnt[0].item()), float(cnt[1].item())

    clocks = ClockSampler(local_rank) if rank == 0 else None
    ms_res, iters_res, frames_res = timed(True, args.steps, args.warmup, clocks)
    clk = clocks.stop() if clocks else None
    ms_e2e, iters_e2e, _ = timed(False, args.steps, max(1, args.warmup // 2))

    # roofline: per-launch duration of the dominant kernel, measured in the library with CUDA events
    roof = None
    launches_per_window = None
    if rank == 0:
        lib.vb_profile_enable(1)
        window(True)
        sm, sl = C.c_double(0), C.c_longlong(0)
        lib.vb_profile_get(C.byref(sm), C.byref(sl))
        lib.vb_profile_enable(0)
        alg_bytes = W * H * (12 * NFLOWS + 60)
        avg_ms = sm.value / max(1, sl.value)
        peak, peak_src = _peaks()
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        # DRAM traffic per launch of the same kernel from the committed `ncu --set full` capture (profiles/)
        traffic, issue_pct = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_ncu_search_kernel.json")) as f:
                cap = json.load(f)
            traffic, issue_pct = cap.get("dram_bytes_per_launch"), cap.get("issue_active_pct")
        except (OSError, ValueError):
            pass
        roof = {"bound": "hbm", "kernel": "k_cost_and_random_search", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms, "launches_timed": sl.value,
                "issue_slots_busy_pct_ncu": issue_pct,
                # compulsory bytes of one whole EM iteration W*H*(76N+56+16*N_dp) + 128*N*P over the
                # measured time per iteration of the resident arm
                "whole_iteration": {
                    "algorithmic_bytes_per_iter": W * H * (76 * NFLOWS + 56) + 128 * NFLOWS * NPOSES,
                    "achieved_per_gpu": (W * H * (76 * NFLOWS + 56) + 128 * NFLOWS * NPOSES) * iters_res / world /
                    (ms_res * 1e-3) / 1e9,
                    "frac": (W * H * (76 * NFLOWS + 56) + 128 * NFLOWS * NPOSES) * iters_res / world /
                    (ms_res * 1e-3) / 1e9 / peak,
                },
                "note": "instruction-issue-bound kernel (6 powf + expf + logf + ~10 IEEE divisions per likelihood "
                        "term, all pinned by bit-parity): HBM fraction reported as BASELINE.json asks; the window "
                        "state is L2 resident; see DESIGN.md §6 and profiles/r01_ncu_search_kernel.md"}
        # kernel launches per window (counted once, outside the timed region, with the CUPTI-based profiler)
        try:
            from torch.profiler import ProfilerActivity, profile

            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                window(True)
                torch.cuda.synchronize()
            launches_per_window = sum(1 for e in prof.events() if e.device_type.name == "CUDA" and
                                      ("k_" in e.name or "vb" in e.name) and "Memcpy" not in e.name)
        except Exception:
            launches_per_window = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, _ = run_cpu_sample(win, boot, 2)
        cpu = {"value": v, "unit": "EM-iterations/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"the same {W}x{H}x{NFLOWS} window, 2 EM iterations ({dt:.1f} s), CPU port with OpenMP on all "
                         "host cores under the restated reference orchestration"}

    voldor_b200.set_bootstrap_override()
    if rank == 0:
        value = iters_res / (ms_res * 1e-3)
        e2e = iters_e2e / (ms_e2e * 1e-3)
        h2d = int(flows_pinned.numel() * 4)
        d2h = int(sum(v.numel() for v in out_h.values()) * 4)
        line = {
            "metric": "EM-iters/sec", "value": value, "unit": "EM-iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: single {W}x{H} frame, {NFLOWS} flows, {EM_ITERS} EM iters, "
                                   f"monocular (injected bootstrap), {NPOSES} hypotheses/camera, truncation off",
                       "parallelism": f"window-per-GPU x{world}" + (", 1 NCCL all_gather of poses per step" if world > 1 else ""),
                       "l2": "256 MB device write between timed steps (flush), per-step CUDA events"},
            "frames_per_sec": frames_res / (ms_res * 1e-3),
            "e2e": {"value": e2e, "unit": "EM-iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": (launches_per_window or 0) * args.steps,
            "gpu_launches_per_step": launches_per_window,
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-port", "abi-dropin"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl != "ours":
        bench_reference(args, rank, world)
        return
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; voldor_b200 has no CPU fallback")
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        bench_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
