#!/usr/bin/env python
"""bench.py -- particle-steps/s of the explicit MPH / total-Lagrangian FSI step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mphx|reference] [--particles P]

Workload (config.workload): the configuration the metric is quoted on -- "3D dam break on elastic
plate, synthetic 10M particles" (BASELINE.json configs[3]); it fits one B200, so it is the N=1
workload too.  A "step" is one pass of the hot path (bucket rebuild, two stencil sweeps, solid
sub-steps, integration) over all particles.

  value      whole-job particle-steps/s with the state resident in HBM, device-timed with CUDA
             events on the library's own stream (mphx_timed_steps), max over ranks.
  e2e        the same metric through the C-ABI with HOST buffers: every step copies Position and
             Velocity host->device from page-locked memory (mphx_upload_state), steps once, and
             reads Position and Velocity back (mphx_download); copies are inside the timed region.
  roofline   dominant kernel (normally pass 2 over the candidate list): algorithmic HBM bytes per
             launch (108 B/fluid, 60 B wall/solid ... see DESIGN.md) / average launch
             duration measured live with CUDA events on the library's stream, against
             MEASURED_PEAKS.json hbm_gbs.  The sweeps are not HBM bound; the ncu facts that say what
             binds them (FP64 pipe, issue slots, L1 data pipe) ride along under roofline.ncu.
  cpu_baseline  the reference's own CPU build (oracle/_ref, all host threads) on a bounded sample.

--impl reference times the reference's CPU implementation (oracle/_ref when built, else the oracle
port) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
WORKLOAD = "3D dam break on elastic plate, synthetic 10M particles (BASELINE.json configs[3])"
WORKLOAD_100M = "3D FSI scaling case, synthetic 100M particles (BASELINE.json configs[4])"


def peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fn):
        try:
            return json.load(open(fn)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(n_sample: float, steps: int, warmup: int, threads: int):
    """the reference's CPU implementation on a bounded sample of the workload -> (value, info)"""
    from particlemethod_fsi_b200 import cases
    case = cases.fsi3d_for_count(n_sample)
    from oracle import refharness
    variant = "3d_dam_nb128"
    if refharness.available(variant):
        d = tempfile.mkdtemp(prefix="mphx_ref_")
        cases.write_grid_file(os.path.join(d, "c.grid"), case)
        cases.write_data_file(os.path.join(d, "c.data"), case.params, case.rc)
        devnull = os.open(os.devnull, os.O_WRONLY)
        so, se = os.dup(1), os.dup(2)
        os.dup2(devnull, 1), os.dup2(devnull, 2)   # the reference prints a line per invalid .data row
        try:
            h = refharness.RefHarness(variant, os.path.join(d, "c.data"), os.path.join(d, "c.grid"), nthreads=threads)
            h.init()
            if warmup:
                h.step(warmup)
            t0 = time.perf_counter()
            h.step(steps)
            dt = time.perf_counter() - t0
        finally:
            os.dup2(so, 1), os.dup2(se, 2)
        kind = "reference"
        what = f"oracle/_ref/libref_{variant}.so (untouched src/main.cpp, g++ -O3 -fopenmp, MAX_NEIGHBOR_COUNT 128)"
    else:
        from oracle.oracle import Oracle
        os.environ.setdefault("OMP_NUM_THREADS", str(threads))
        o = Oracle.from_case(case, max_neighbor_count=128)
        o.init()
        if warmup:
            o.step(warmup)
        t0 = time.perf_counter()
        o.step(steps)
        dt = time.perf_counter() - t0
        kind = "port"
        what = "oracle/liboracle.so (C restatement, OpenMP)"
    value = case.n * steps / dt
    info = {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{case.n} particles ({case.name}: same geometry at coarser spacing), {steps} steps after "
                      f"{warmup} warm-up, {what}", "seconds": dt}
    return value, info, case


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_cores()
    steps = max(1, min(args.steps, 3))
    value, info, case = cpu_reference_run(args.ref_particles, steps, min(args.warmup, 1), threads)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * info["seconds"] / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "sample_particles": case.n, "dim": 3,
                       "note": "reference CPU build on host cores; bounded sample of the workload"},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_mphx(args):
    import torch
    import torch.distributed as dist
    import particlemethod_fsi_b200 as pm
    from particlemethod_fsi_b200 import cases

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from particlemethod_fsi_b200 import slab
        return slab.bench_main(args, METRIC, UNIT, WORKLOAD, peaks, ClockSampler, cpu_reference_run, host_cores)
    if not torch.cuda.is_available() or pm.solver.device_count() == 0:
        raise SystemExit("bench.py: no B200 visible -- the product path has no CPU fallback")
    torch.cuda.set_device(local)

    case = cases.fsi3d_for_count(args.particles)
    n = case.n
    nf, ns, nw = case.counts()
    s = pm.Solver.from_case(case, device=local, list_reuse=None if args.list_reuse < 0 else bool(args.list_reuse))
    K, W = args.steps, args.warmup

    # ---- resident-state throughput ("value") -------------------------------------------------------
    s.step(W, sync=True)
    l0 = s.launch_count
    s.set_timing(True)
    sampler = ClockSampler(local)
    sampler.start()
    ms = s.timed_steps(K)
    clocks = sampler.stop()
    phase = s.timers_ms()
    s.set_timing(False)
    launches = s.launch_count - l0
    value = n * K / (ms * 1e-3)
    # per-kernel durations for the roofline: a few more steps with the solid sub-steps serialised on the
    # main stream (in the timed region above they overlap pass 2 on a second stream, so an event pair around
    # pass 2 would time both)
    KK = max(3, min(K, 5))
    s.set_overlap(False)
    s.step(1, sync=True)
    s.set_timing(True)
    s.step(KK, sync=True)
    kms = [t * K / KK for t in s.kernel_timers_ms()]   # [rebuild, k_filter, pass 1, pass 2, solid]; scaled to K steps
    s.set_timing(False)
    s.set_overlap(True)

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    pk, pk_kind = peaks()
    # algorithmic HBM bytes per launch:
    #   pass 2: R x 24, v 24, type 4, P 8; W x 24, v 24 = 108 B per fluid particle; wall/solid particles
    #           read the same 60 B and write nothing the model counts;
    #   pass 1: R x 24, v 24, type 4; W PressureP 8 = 60 B per particle.  The model has no separate filter
    #           kernel: k_filter reads the positions it tests (24 B + type 4 per particle).
    cand = [("k_filter<3>", kms[1] / K, 28.0 * n),
            ("k_pass1_v3<3,false,true>", kms[2] / K, 60.0 * n),
            ("k_pass2_v3<3,false,true>", kms[3] / K, 108.0 * nf + 60.0 * (nw + ns))]
    dom_name, dom_ms, dom_bytes = max(cand, key=lambda c: c[1])
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # the second roof: FP64.  Peak = a pure DFMA kernel timed on this device now; work = SURVEY 8(d)'s algorithmic count,
    # 15 flop per candidate examined + 45 per in-radius pair and sweep, with both counts taken from the live lists.
    fp64_peak = pm.solver.measure_fp64_peak(local)
    cand_n, inr_n = s.count_pairs()
    sweep_flop = 15.0 * cand_n + 45.0 * inr_n
    fp64 = {"peak_measured_tflops": fp64_peak, "peak_source": "k_fp64_peak (8 DFMA chains per thread) timed with CUDA events in this run",
            "candidates_per_particle": cand_n / n, "in_radius_pairs_per_particle": inr_n / n,
            "algorithmic_flop_per_sweep": sweep_flop, "model": "15 flop per candidate + 45 per in-radius pair (SURVEY 8(d))"}
    for nm, t in (("pass1", kms[2] / K), ("pass2", kms[3] / K)):
        tfl = sweep_flop / (t * 1e-3) / 1e12
        fp64[nm] = {"achieved_tflops": tfl, "frac": tfl / fp64_peak}
    # the gather level (L1/L2, not HBM; SURVEY 8(d): "report HBM-level and L2-level throughput separately"): what the list
    # kernels pull per candidate -- a 4-byte list entry and two 32-byte records -- counted from the live lists
    gather_bytes = 68.0 * cand_n
    gather = {"model": "per candidate: list entry 4 B + records 2 x 32 B, served by L1/L2 (not HBM traffic)", "bytes_per_sweep": gather_bytes,
              "pass1_gbs": gather_bytes / (kms[2] / K * 1e-3) / 1e9, "pass2_gbs": gather_bytes / (kms[3] / K * 1e-3) / 1e9}
    # ncu counters of the dominant kernel: a committed capture (DRAM traffic per launch cannot be measured without the
    # profiler); the entry says which commit / round it was taken at
    traffic, ncu_facts = None, None
    tf = os.path.join(ROOT, "profiles", "sweep_ncu_facts.json")
    if os.path.exists(tf):
        try:
            facts = json.load(open(tf))
            ncu_facts = facts.get(dom_name)
            if ncu_facts is not None:
                ncu_facts = dict(ncu_facts, captured_at=facts.get("_captured_at"))
            traffic = ncu_facts.get("dram_bytes_per_launch") if ncu_facts else None
        except Exception:
            traffic = None
    step_bytes = s.algorithmic_bytes_per_step
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                "peak_source": pk_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                "kernel_ms_per_step": {"bucket_rebuild": kms[0] / K, "k_filter": kms[1] / K, "pass1_list": kms[2] / K,
                                       "pass2_list": kms[3] / K, "solid_substeps": kms[4] / K},
                "phase_ms_per_step": {"bucket_rebuild": phase[0] / K, "pass1": phase[1] / K, "pass2": phase[2] / K,
                                      "solid_substeps": phase[3] / K},
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes * K / (ms * 1e-3) / 1e9,
                               "frac": step_bytes * K / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
                "fp64": fp64,
                "gather_level": gather,
                "ncu": ncu_facts,
                "note": "frac is the HBM fraction the metric asks for; it is small by construction (370 algorithmic bytes but ~2e4 "
                        "fp64 flop per particle-step).  What binds the sweeps is under fp64 (live) and ncu (committed capture): FP64 "
                        "pipe, instruction issue and the L1 data pipe.  kernel_ms_per_step: isolated kernels (solid sub-steps "
                        "serialised); phase_ms_per_step: the timed region, where the sub-steps overlap pass 2"}

    # ---- end to end through the C-ABI with host buffers ------------------------------------------------
    hx = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    hv = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    s.download_ptr(position=hx.data_ptr(), velocity=hv.data_ptr())
    ke = max(1, min(K, args.e2e_steps))
    for _ in range(1):  # warm the path
        s.upload_state_ptr(hx.data_ptr(), hv.data_ptr())
        s.step(1)
        s.download_ptr(position=hx.data_ptr(), velocity=hv.data_ptr())
    t0 = time.perf_counter()
    for _ in range(ke):
        s.upload_state_ptr(hx.data_ptr(), hv.data_ptr())
        s.step(1)
        s.download_ptr(position=hx.data_ptr(), velocity=hv.data_ptr())   # synchronises
    te = time.perf_counter() - t0
    e2e = {"value": n * ke / te, "unit": UNIT, "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 48 * n,
           "steps": ke, "ms_per_step": 1e3 * te / ke,
           "path": "mphx_upload_state + mphx_step + mphx_download (pinned host buffers, wall clock)"}
    assert bool(torch.isfinite(hx).all())
    status = s.status()
    assert status["err"] == 0, status
    s.close()

    # ---- companion number on a developed (disordered, moving) state ------------------------------------------
    # the headline state is the generator's lattice shortly after release: one particle per bucket, lists that never
    # expire.  Here every fluid particle is displaced by a uniform +-0.2 l0 per axis (seed 12345: 0..3 particles per
    # bucket) and the water moves at 0.5 m/s, so candidate lists carry real skin candidates and expire every few steps.
    developed = None
    if args.state in ("both", "developed"):
        rng = np.random.default_rng(12345)
        fl = case.property < 2
        dcase = cases.Case(case.name + "_developed", case.params.copy(), case.rc, case.property,
                           case.position.copy(), case.initial_position, case.velocity.copy())
        dcase.position[fl] += rng.uniform(-0.2, 0.2, size=(int(fl.sum()), 3)) * case.params.particle_spacing
        dcase.velocity[fl, 0] = 0.5
        d = pm.Solver.from_case(dcase, device=local, list_reuse=None if args.list_reuse < 0 else bool(args.list_reuse))
        d.step(W, sync=True)
        st0 = d.status()
        dms = d.timed_steps(K)
        st1 = d.status()
        dc, di = d.count_pairs()
        g = d.download("position")
        developed = {"value": n * K / (dms * 1e-3), "unit": UNIT, "ms_per_step": dms / K, "steps": K,
                     "state": "fluid displaced by uniform +-0.2 l0 per axis (seed 12345), fluid velocity 0.5 m/s in +x",
                     "lists_built": st1["builds"] - st0["builds"], "steps_reusing_a_list": st1["reuses"] - st0["reuses"],
                     "candidates_per_particle": dc / n, "in_radius_pairs_per_particle": di / n,
                     "all_finite": bool(np.isfinite(g["position"]).all()), "error_flags": st1["err"]}
        d.close()

    # ---- CPU baseline on the host cores (bounded sample) ---------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        try:
            _v, cpu, _c = cpu_reference_run(args.ref_particles, 2, 1, host_cores())
        except Exception as e:  # never let the baseline leg kill the measurement
            cpu = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "unavailable", "sample": repr(e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "particles": n, "fluid": nf, "solid": ns, "wall": nw, "dim": 3,
                       "particle_spacing": case.params.particle_spacing, "dt": case.params.dt,
                       "solid_substeps": int(case.params.dt / case.params.elastic_dt + 0.5),
                       "cache": "inputs larger than L2 (state ~%.1f GB vs 126 MB L2)" % (n * 240 / 1e9),
                       "parallelism": "1 GPU", "state": "generator lattice, first steps after release (the headline)",
                       "developed_state": developed,
                       "candidate_list": {"reuse": bool(status["skin_on"]), "lists_built": status["builds"],
                                          "steps_reusing_a_list": status["reuses"],
                                          "note": "all steps of the run incl. warm-up and the e2e leg (every e2e step uploads a new "
                                                  "state and therefore rebuilds)"}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mphx", choices=["mphx", "reference"])
    ap.add_argument("--particles", type=float, default=1.0e7)
    ap.add_argument("--ref-particles", type=float, default=1.0e6, help="sample size of the CPU reference arm (10-30 s of host work)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--state", default="both", choices=["both", "lattice", "developed"],
                    help="lattice: the headline only; both: also the developed-state companion number (config.developed_state)")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the ring-vs-single-context check of the exchange")
    ap.add_argument("--verify-particles", type=float, default=2.0e5)
    ap.add_argument("--trace", default="", help="N > 1: after the timed region, write a device-side timeline of 3 steps per rank to <prefix>_rank<r>.json")
    ap.add_argument("--list-reuse", type=int, default=-1, help="1/0: candidate-list reuse on/off (default: the library's default, on)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "mphx" else args.warmup
    global WORKLOAD
    if args.particles >= 5.0e7:
        WORKLOAD = WORKLOAD_100M
    elif abs(args.particles - 1.0e7) > 1.0e6:
        WORKLOAD = "3D dam break on elastic plate, synthetic %.3gM particles (BASELINE.json configs[3] geometry)" % (args.particles / 1e6)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mphx(args)


if __name__ == "__main__":
    main()
