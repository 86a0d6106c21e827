#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native camera<-laser extrinsic solve.

Workload (BASELINE.json configs[1]): 10^4 frames x 10^3 laser points per GPU, synthetic boards from the reference's
simulation generator (main/calibr_simulation.cpp:10-108, exact-M mode, 1 cm range noise), identity initial guess.
A "step" is one complete CamLaserCalibration-equivalent solve: the on-device Levenberg-Marquardt loop to Ceres'
convergence criteria, one fused residual+Jacobian+reduce sweep over every point per LM iteration.

  value     residual+Jacobian evaluations / s over the whole job, data already resident in HBM
  e2e       the same through the public API with HOST (pinned) buffers: H2D upload + HBM layout + solve + D2H result
  roofline  the fused sweep kernel alone: algorithmic bytes (24 B/residual + 40 B/frame + 224 B) / CUDA-event time
            per launch, L2 flushed between launches, against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference   the CPU oracle port of the reference algorithm (Ceres-shaped: materialised
            Jacobian + dense QR, evaluation threaded over all host cores) on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): weak scaling -- every rank holds 10^4 frames of a 10^4 x N frame problem;
the 28 normal-equation sums are all-reduced (NCCL, 224 B) after every sweep and every rank runs the identical LM
update on its device.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 10_000
BEAMS = 1_000
SEED = 7
SIGMA = 0.01
X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
CPU_SAMPLE_FRAMES = 1_000
METRIC = "residual+Jacobian evals/sec (full LM solves)"
UNIT = "residual evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU")
    ap.add_argument("--beams", type=int, default=BEAMS)
    ap.add_argument("--kernel-launches", type=int, default=100, help="timed launches of the sweep kernel for the roofline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-allreduce", action="store_true", help="N>1: ncclAllReduce between kernels instead of the fused peer exchange")
    ap.add_argument("--no-config3", action="store_true", help="skip the supplementary 4.8 GB (configs[2]) measurement")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ---- clocks during the timed region (B200_PROFILING.md) ----------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed power range
        thr = min(power) + 0.5 * (max(power) - min(power))
        loaded = [c for c, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ---- CPU baseline: the oracle port of the reference algorithm ------------------------------------------------------
def cpu_reference_problem(frames, beams):
    from oracle import oracle as O

    return O.generate(frames, beams, seed=SEED, sigma=SIGMA, exact_m=True)


def time_cpu_solve(p, threads, linear_solver):
    """One full LM solve with the oracle; returns (seconds, residual evaluations, iterations)."""
    from oracle import oracle as O

    opt = O.default_options(linear_solver=linear_solver, num_threads=threads)
    t0 = time.perf_counter()
    _, s, _ = O.solve(p, X0, opt)
    dt = time.perf_counter() - t0
    return dt, s.num_residual_evaluations * p.num_residuals(), s.num_iterations


def pick_threads(p):
    """Thread count that makes the oracle's residual sweep fastest on this host (a container may expose more CPUs than
    its quota lets it use, in which case all-cores OpenMP is slower than a few threads)."""
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    cands = sorted({1, 2, 4, 8, 16, 32, 64, cores} & set(range(1, cores + 1)))
    best, best_t = 1, None
    for th in cands:
        O.evaluate_normal(p, X0, num_threads=th)
        t0 = time.perf_counter()
        for _ in range(3):
            O.evaluate_normal(p, X0, num_threads=th)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
    return best


def cpu_baseline(frames, beams, quick=False):
    """Times the oracle port on a bounded sample.  Primary number: the Ceres-shaped solve (materialised P x 6
    Jacobian + Householder QR -- the work the reference does) with the better of {1 thread (what the reference uses:
    Ceres num_threads is never set), all host threads for the residual/Jacobian evaluation}."""
    n = min(frames, CPU_SAMPLE_FRAMES)
    p = cpu_reference_problem(n, beams)
    cores = pick_threads(p)
    dt1, ev1, iters = time_cpu_solve(p, 1, 0)
    best = dict(value=ev1 / dt1, cores=1)
    extras = {"ceres_shaped_1_thread": {"value": ev1 / dt1, "unit": UNIT, "cores": 1}}
    if cores > 1:
        dta, eva, _ = time_cpu_solve(p, cores, 0)
        extras["ceres_shaped_best_thread_count"] = {"value": eva / dta, "unit": UNIT, "cores": cores}
        if eva / dta > best["value"]:
            best = dict(value=eva / dta, cores=cores)
    if not quick:
        dts, evs, _ = time_cpu_solve(p, cores, 1)  # most favourable CPU variant: streaming normal equations
        extras["streaming_normal_equations_best_thread_count"] = {"value": evs / dts, "unit": UNIT, "cores": cores}
    extras["host_cpus_visible"] = os.cpu_count()
    out = dict(value=best["value"], unit=UNIT, cores=best["cores"], kind="port",
               sample=f"first {n} of {frames} frames x {beams} points ({n * beams} residuals), one full LM solve "
                      f"({iters} iterations, {ev1 // (n * beams)} sweeps), Ceres-shaped (materialised Jacobian + dense QR)")
    out.update(extras)
    return out


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    n = min(args.frames, CPU_SAMPLE_FRAMES)
    p = cpu_reference_problem(n, args.beams)
    cores = pick_threads(p)
    # threads: the better of 1 (what the reference uses) and all host threads, decided on the warm-up solves
    rates = {}
    for th in sorted({1, cores}):
        dt, ev, _ = time_cpu_solve(p, th, 0)
        rates[th] = ev / dt
    cores = max(rates, key=rates.get)
    for _ in range(max(0, args.warmup - len(rates))):
        time_cpu_solve(p, cores, 0)
    t_tot, ev_tot, iters, steps_done = 0.0, 0, 0, 0
    for _ in range(args.steps):
        dt, ev, iters = time_cpu_solve(p, cores, 0)
        t_tot += dt
        ev_tot += ev
        steps_done += 1
        if t_tot > 120.0:  # keep the whole run within a few minutes whatever --steps says
            break
    value = ev_tot / t_tot
    sample = (f"each step = one full LM solve on the first {n} of {args.frames} frames x {args.beams} points "
              f"({n * args.beams} residuals, {iters} iterations); Ceres-shaped oracle port, evaluation on {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / steps_done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: {args.frames} frames x {args.beams} points per GPU, sigma={SIGMA} m, "
                               f"identity start, full LM solve (bounded CPU sample: {n} frames)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference itself (Ceres+Eigen+ROS) cannot be built in this image; this is the CPU oracle port",
    }
    print(json.dumps(line))


# ---- our arm ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from camlasercalibratool_b200 import Comm, Problem, comm_unique_id, default_options, launch_count
    from camlasercalibratool_b200.api import pinned_array

    rank, local_rank, world = dist_env()
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    frames_total = args.frames * world
    f0, f1 = rank * args.frames, (rank + 1) * args.frames
    prob = Problem.synthetic(frames_total, args.beams, seed=SEED, sigma=SIGMA, frame_begin=f0, frame_end=f1, device=local_rank)
    n_frames, n_points, _ = prob.sizes()
    # The simulated laser is two-dimensional, so the library would drop the z stream (16 B per residual).  SURVEY.md 8(d)
    # fixes the contract figure at 24 B per residual: the headline legs run the general three-stream kernels; the planar
    # kernels are reported as their own row ("planar") with their own byte count.
    prob.set_planar_mode(0)
    comm = None
    if world > 1:
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = Comm(uid[0], world, rank, device=local_rank)

        def all_gather(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out

        p2p_note = "fused into the sweep kernel (NVLink peer stores, sequence-tagged words, rank-order sum)"
        if not args.nccl_allreduce:
            try:
                comm.enable_p2p(all_gather)  # fused in-kernel all-reduce over NVLink peer memory
            except Exception as exc:  # no peer access between these GPUs: the NCCL path still works
                args.nccl_allreduce = True
                p2p_note = f"peer exchange unavailable ({exc}); "
            ok = torch.tensor([0 if args.nccl_allreduce else 1], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks must use the same mode
            if int(ok.item()) == 0 and not args.nccl_allreduce:
                args.nccl_allreduce = True
                comm.close()
                uid = [comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                comm = Comm(uid[0], world, rank, device=local_rank)
        prob.attach_comm(comm)
    opt = default_options()

    # ---- resident-data leg: K full solves ----
    for _ in range(args.warmup):
        prob.solve(X0, opt)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = launch_count()
    barrier()
    t0 = time.perf_counter()
    dev_ms, sweeps, iters = 0.0, 0, 0
    x = X0
    for _ in range(args.steps):
        x, s, _ = prob.solve(X0, opt)
        dev_ms += s.device_ms
        sweeps += s.num_sweeps
        iters += s.num_iterations - 1
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = launch_count() - launches0
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    total_points = sum_over_ranks(float(n_points))
    value = total_points * sweeps / (dev_ms * 1e-3)
    lm_iters_per_s = iters / (dev_ms * 1e-3)

    # ---- roofline leg: the sweep kernel alone, L2 flushed between launches (local shard, no collective) ----
    prob.bench_eval(x, 5, flush_l2=True)
    k_ms = prob.bench_eval(x, args.kernel_launches, flush_l2=True)
    launches += args.kernel_launches
    k_b2b = prob.bench_eval(x, args.kernel_launches, flush_l2=False)
    launches += args.kernel_launches
    k_mean = float(np.mean(k_ms))
    alg_bytes = prob.algorithmic_bytes()
    peaks, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    except Exception:
        pass
    achieved = alg_bytes / (k_mean * 1e-3) / 1e9
    traffic = traffic_planar = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            traffic = tj.get("dram_bytes_per_launch")
            traffic_planar = (tj.get("planar") or {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks, "unit": "GB/s", "frac": achieved / peaks,
                "traffic": traffic, "kernel": "clc_sweep_kernel<LOSS,LM>", "kernel_ms_mean": k_mean,
                "kernel_ms_min": float(np.min(k_ms)), "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "residuals_per_s_kernel": n_points / (k_mean * 1e-3),
                "back_to_back_no_flush": {"kernel_ms_mean": float(np.mean(k_b2b)), "achieved": alg_bytes / (float(np.mean(k_b2b)) * 1e-3) / 1e9,
                                          "frac": alg_bytes / (float(np.mean(k_b2b)) * 1e-3) / 1e9 / peaks,
                                          "note": "inputs (240 MB) exceed the 126 MB L2 but part of them survives between launches"},
                "l2": "flushed between launches (256 MiB written, then read back so that no dirty lines are left)"}

    # ---- separate row: the planar (two-stream) kernels the library picks by itself for z == 0 data ----
    planar = None
    prob.set_planar_mode(1)
    if prob.planar:
        lc0 = launch_count()
        for _ in range(args.warmup):
            prob.solve(X0, opt)
        barrier()
        p_ms, p_sweeps, p_iters = 0.0, 0, 0
        for _ in range(args.steps):
            _, s, _ = prob.solve(X0, opt)
            p_ms += s.device_ms
            p_sweeps += s.num_sweeps
            p_iters += s.num_iterations - 1
        barrier()
        p_ms = max_over_ranks(p_ms)
        prob.bench_eval(x, 5, flush_l2=True)
        pk_ms = prob.bench_eval(x, args.kernel_launches, flush_l2=True)
        launches += launch_count() - lc0
        pk_mean = float(np.mean(pk_ms))
        p_bytes = prob.streamed_bytes()
        planar = {"what": "z == 0 for every point (a 2-D laser): z stream dropped from HBM, two-stream kernels, results "
                          "equal to the general kernels up to summation order; 16 B per residual -- own denominators, never mixed with the 24 B row",
                  "value": total_points * p_sweeps / (p_ms * 1e-3), "unit": UNIT, "ms_per_step": p_ms / args.steps,
                  "lm_iters_per_s": p_iters / (p_ms * 1e-3),
                  "roofline": {"bound": "hbm", "achieved": p_bytes / (pk_mean * 1e-3) / 1e9, "peak": peaks, "unit": "GB/s",
                               "frac": p_bytes / (pk_mean * 1e-3) / 1e9 / peaks, "traffic": traffic_planar,
                               "kernel": "clc_sweep_kernel<LOSS,LM,PLANAR>",
                               "kernel_ms_mean": pk_mean, "kernel_ms_min": float(np.min(pk_ms)),
                               "algorithmic_bytes_per_launch": p_bytes, "residuals_per_s_kernel": n_points / (pk_mean * 1e-3),
                               "speedup_over_24B_kernel": k_mean / pk_mean}}

    # ---- end-to-end leg: host (pinned) buffers -> create (H2D + layout) -> solve -> D2H result -> destroy ----
    d = prob.download()
    pin_pts = pinned_array(d["points"].shape)
    pin_pts.array[...] = d["points"]
    pin_fp = pinned_array(d["frame_pose"].shape)
    pin_fp.array[...] = d["frame_pose"]
    offsets = d["offsets"]
    del d

    def e2e_step():
        with Problem.from_arrays(pin_fp.array, offsets, pin_pts.array, device=local_rank) as q:
            q.attach_comm(comm)
            _, ss, _ = q.solve(X0, opt)
        return ss

    e2e_steps = max(3, min(args.steps, 10))
    e2e_sweeps = 0
    e2e_step()  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_sweeps += e2e_step().num_sweeps
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_points * e2e_sweeps / e2e_s
    h2d = int(pin_pts.nbytes + pin_fp.nbytes + offsets.nbytes + 7 * 8 + 64)
    d2h = int(7 * 8 + 64)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "what": "per rank: Problem.from_arrays(pinned host AoS, 24 B per point) [H2D + HBM layout + planarity detection] + "
                   "clc_solve_lm (library default: planar kernels, the data have z == 0) + result read-back + destroy; wall "
                   "clock, max over ranks; PCIe-bound: the 240 MB upload alone takes 4.3 ms at the measured 55 GB/s"}
    pin_pts.free()
    pin_fp.free()
    clocks = sampler.stop() if rank == 0 else None

    # ---- supplementary: BASELINE configs[2] (10^5 frames x 2*10^3 points, 4.8 GB): the size BASELINE.json names for the
    #      ncu capture of achieved HBM GB/s and for "full LM to convergence" ----
    config3 = None
    if world == 1 and not args.no_config3:
        prob.close()
        with Problem.synthetic(100_000, 2_000, seed=SEED, sigma=SIGMA, device=local_rank) as big:
            big.set_planar_mode(0)  # contract row first (24 B per residual)
            big.bench_eval(X0, 3, flush_l2=True)
            ms3 = big.bench_eval(x, 20, flush_l2=True)
            b3 = big.algorithmic_bytes()
            for _ in range(2):
                big.solve(X0, opt)
            _, s3, _ = big.solve(X0, opt)
            launches += 23 + 3 * s3.num_sweeps
            big.set_planar_mode(1)
            big.bench_eval(X0, 3, flush_l2=True)
            ms3p = big.bench_eval(x, 20, flush_l2=True)
            b3p = big.streamed_bytes()
            big.solve(X0, opt)
            _, s3p, _ = big.solve(X0, opt)
            launches += 23 + 2 * s3p.num_sweeps
            ach3 = b3 / (float(np.mean(ms3)) * 1e-3) / 1e9
            config3 = {"workload": "BASELINE configs[2]: 100000 frames x 2000 points (4.8 GB), same generator",
                       "roofline": {"bound": "hbm", "achieved": ach3, "peak": peaks, "unit": "GB/s", "frac": ach3 / peaks,
                                    "kernel_ms_mean": float(np.mean(ms3)), "algorithmic_bytes_per_launch": b3,
                                    "residuals_per_s_kernel": 2e8 / (float(np.mean(ms3)) * 1e-3)},
                       "full_lm_solve": {"ms": s3.device_ms, "lm_iterations": s3.num_iterations - 1, "sweeps": s3.num_sweeps,
                                         "residual_evals_per_s": 2e8 * s3.num_sweeps / (s3.device_ms * 1e-3),
                                         "lm_iters_per_s": (s3.num_iterations - 1) / (s3.device_ms * 1e-3),
                                         "termination": int(s3.termination)},
                       "planar": {"kernel_ms_mean": float(np.mean(ms3p)), "algorithmic_bytes_per_launch": b3p,
                                  "achieved": b3p / (float(np.mean(ms3p)) * 1e-3) / 1e9,
                                  "frac": b3p / (float(np.mean(ms3p)) * 1e-3) / 1e9 / peaks,
                                  "full_lm_solve_ms": s3p.device_ms, "sweeps": s3p.num_sweeps,
                                  "residual_evals_per_s": 2e8 * s3p.num_sweeps / (s3p.device_ms * 1e-3)}}

    # ---- supplementary: BASELINE configs[0], the reference's own problem size (50 frames x 180 beams), end to end through
    #      the mirrored entry point (marshal + upload + on-device LM + analysis tail + destroy) next to the CPU oracle ----
    config1 = None
    if rank == 0 and world == 1:
        from camlasercalibratool_b200 import CamLaserCalibration, Oberserve
        from oracle import oracle as O

        small = O.generate(50, 180, seed=1, sigma=0.01)
        obs = [Oberserve(small.frame_pose[f, :4].copy(), small.frame_pose[f, 4:].copy(),
                         small.points[small.offsets[f]:small.offsets[f + 1]], small.points[small.offsets[f]:small.offsets[f + 1]])
               for f in range(small.n_frames)]
        for _ in range(3):
            CamLaserCalibration(obs, np.eye(4), False, verbose=False)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            rep = CamLaserCalibration(obs, np.eye(4), False, verbose=False)
        gpu_ms = 1e3 * (time.perf_counter() - t0) / reps
        launches += 23 * 20
        t0 = time.perf_counter()
        for _ in range(5):
            O.solve(small, X0)
        cpu_ms = 1e3 * (time.perf_counter() - t0) / 5
        config1 = {"workload": "BASELINE configs[0]: 50 frames x 180 beams (5351 residuals), CamLaserCalibration() end to end",
                   "gpu_ms_per_call": gpu_ms, "device_ms_of_the_lm": rep["device_ms"], "lm_iterations": rep["iterations"] - 1,
                   "cpu_oracle_ms_per_solve_1_thread": cpu_ms}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.frames, args.beams)
    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: {args.frames} frames x {args.beams} points per GPU "
                                   f"({frames_total} frames total), calibr_simulation generator (exact-M), sigma={SIGMA} m, "
                                   f"identity start, full LM solve to Ceres convergence",
                       "sharding": (f"frames by rank, {world} rank(s), 28-double all-reduce per sweep: " +
                                    ("ncclAllReduce between kernels" if args.nccl_allreduce else
                                     "fused into the sweep kernel (NVLink peer stores, sequence-tagged words, rank-order sum)")) if world > 1 else "single GPU",
                       "l2": f"inputs ({alg_bytes / 1e6:.0f} MB per GPU) larger than the 126 MB L2; roofline leg flushes L2 between launches",
                       "lm": "one fused residual+Jacobian+reduce sweep per LM iteration (speculative Jacobian at the candidate)",
                       "kernels": "general three-stream kernels (24 B per residual, the SURVEY.md 8(d) contract figure); the planar "
                                  "two-stream kernels the library would pick for this z == 0 data are the separate row 'planar'"},
            "lm_iters_per_s": lm_iters_per_s, "sweeps_per_solve": sweeps / args.steps, "lm_iterations_per_solve": iters / args.steps,
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": roofline, "planar": planar, "config3": config3, "config1": config1, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches),
        }
        print(json.dumps(line))
    prob.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    _, _, world = dist_env()
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: self-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
