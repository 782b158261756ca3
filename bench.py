#!/usr/bin/env python
"""bench.py — the ssa_sdpd hot path on B200, measured as BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload box|tank|cylinder|ens_*]

Workload (default `box`): the unit the metric's target is quoted on — BASELINE configs[4], the synthetic 3-D SDPD + sSSA box, 200^3 =
8 M moving particles per GPU (spatialpy_b200/configs.py:box_sdpd_rdme).  One bench "step" = SPS engine
timesteps, each = cell list / predictor / neighbour search (when due) / pairwise force sweep / corrector / BVF sweep / sSSA windows.
  N = 1   one 8 M box on one GPU.  `value` = particle-steps/s with the model resident in HBM, CUDA events on the engine's stream;
          `e2e` = the same through the C-ABI call a user's Solver.run makes (ssb_run: upload of the initial state from host memory,
          stepping, output snapshots copied back to pinned host memory), host clock.  The line also carries `same_n` (our engine on
          the very instance the reference arm runs: the like-for-like ratio) and `sub_records` (cylinder = configs[1], tank = configs[2]).
  N > 1   ONE box of N x 8 M particles split into N slabs along x (weak scaling), halo exchange through peer-mapped windows over
          NVLink (ssb_slab_*, include/ssb.h): `value` = owned particles of all ranks x steps / max-over-ranks device time.
`--impl reference` times the UNMODIFIED reference engine (oracle/_ref/bench_*/fast/ssa_sdpd.exe, built from /root/reference by
oracle/oracle_build.py) on the host cores on a bounded instance of the same workload (24^3 = 13 824 particles: the reference
cannot be compiled at 8 M, one C++ source line per particle), as BASELINE.md 3.2 says: wall of the run minus a zero-step run.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # see spatialpy_b200/engine.py (ensemble lanes)

METRIC = "particle-steps/s (SDPD+sSSA)"
UNIT = "particle-steps/s"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------
def make_workload(name, scale):
    from spatialpy_b200 import configs
    if name == "cylinder":
        delta = 0.03155 if scale >= 1.0 else 0.03155 / scale ** (1.0 / 3.0)
        fm = configs.cylinder_rdme(delta=delta, nt=1000, output_every=100, dt=1e-3)
        desc = f"BASELINE configs[1]: 3D_Cylinder_Demo static RDME+PDE, A+B->0 with end-cap sources, jittered lattice delta={delta:.5f}"
    elif name == "tank":
        n = max(12, int(round(120 * scale ** (1.0 / 3.0))))
        fm = configs.tank_sdpd(n=n, nt=1000, output_every=100, dt=1e-5)
        desc = f"BASELINE configs[2] stand-in: 3-D SDPD tank {n}^3 lattice under gravity with one advected reacting species"
    elif name == "box":
        n = max(16, int(round(200 * scale ** (1.0 / 3.0))))
        fm = configs.box_sdpd_rdme(n, n, n, nt=100, output_every=100, dt=1e-5)
        desc = f"BASELINE configs[4] per-GPU unit: synthetic 3-D SDPD+sSSA box {n}^3"
    else:
        raise SystemExit(f"unknown workload {name}")
    return fm, desc


def algorithmic_bytes(fm, moving):
    """SURVEY.md §8(d) contract figures, per particle-step and per RDME event."""
    Sc, Sd, R = fm.num_chem_species, fm.num_stoch_species, fm.num_stoch_rxns
    per_step = (698 + 64 * Sc) + (68 + 12 * Sd + 8 * R if Sd else 0) if moving else (48 + 40 * Sc)
    return per_step


# ---------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------
REF_NAME = {"cylinder": "bench_cylinder", "tank": "bench_tank", "box": "bench_box"}


def _time_reference(exe, threads, seed=1000):
    d = tempfile.mkdtemp(prefix="ssb_ref_")
    t0 = time.perf_counter()
    subprocess.run([exe, "-t", str(threads), "-s", str(seed)], cwd=d, stdout=subprocess.DEVNULL, check=True)
    dt = time.perf_counter() - t0
    subprocess.run(["rm", "-rf", d])
    return dt


def _reference_instance(workload):
    base = os.path.join(ROOT, "oracle", "_ref", REF_NAME[workload])
    exe = os.path.join(base, "fast", "ssa_sdpd.exe")
    if not os.path.exists(exe):
        return None
    meta = json.load(open(os.path.join(base, "meta.json")))
    zero = os.path.join(ROOT, "oracle", "_ref", REF_NAME[workload] + "_zero", "fast", "ssa_sdpd.exe")
    return base, exe, (zero if os.path.exists(zero) else None), meta


def run_reference(args, rank, world):
    """The reference's own CPU implementation on the box's host cores (rank 0 only).  Stepping-loop time = wall of the run minus
    the wall of a zero-step run of the same model, so that process start, particle construction and the first
    output do not count against the reference."""
    if rank != 0:
        return
    if args.workload not in REF_NAME:
        print(json.dumps({"impl": "reference", "unavailable": f"no reference instance for workload {args.workload}"}))
        return
    inst = _reference_instance(args.workload)
    if inst is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (oracle/oracle_build.py needs /root/reference)"}))
        return
    base, exe, zero, meta = inst
    threads = os.cpu_count() or 1
    t_zero = min(_time_reference(zero, threads, 900 + k) for k in range(2)) if zero else 0.0
    times = []
    for it in range(args.warmup + args.steps):
        dt = _time_reference(exe, threads, 1000 + it)
        if it >= args.warmup:
            times.append(max(dt - t_zero, 1e-9))
    ms = 1e3 * sum(times) / len(times)
    value = meta["N"] * meta["nt"] / (ms / 1e3)
    sample = (f"unmodified reference engine (g++ -O3), {meta['builder']}{meta['kwargs']}: N={meta['N']} particles x {meta['nt']} steps per run, "
              f"-t {threads}; wall of the executable minus a zero-step run of the same model ({t_zero:.3f} s)" if zero else
              f"unmodified reference engine (g++ -O3), {meta['builder']}{meta['kwargs']}: N={meta['N']} particles x {meta['nt']} steps per run, "
              f"-t {threads}; wall clock of the whole executable (includes particle construction and VTK output)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} (bounded CPU instance of the GPU workload)", "particles": meta["N"], "engine_steps_per_step": meta["nt"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(workload):
    """Bounded sample of the reference on the host cores (rank 0, N=1 only).  `value` is the strongest CPU arm (g++ -O3, all
    cores, stepping loop = wall minus a zero-step run where that executable exists); `variants` adds what SURVEY.md 8(d) asks to see
    beside it: one thread, the reference's own default thread cap (min(8, cores), E/propensity_file_template.cpp:129-132) and the
    build the reference actually ships (no -O, E/build/SConstruct:23)."""
    inst = _reference_instance(workload)
    if inst is None:
        return None
    base, exe, zero, meta = inst
    cores = os.cpu_count() or 1
    work = meta["N"] * meta["nt"]
    t_zero = _time_reference(zero, cores, 900) if zero else 0.0
    dt = max(_time_reference(exe, cores) - t_zero, 1e-9)
    variants = {f"O3_t{cores}": work / dt}
    try:
        # the extra arms stay inside the bench's time budget: skipped on a host where the strongest arm already takes long
        if dt < 12.0:
            t1 = max(_time_reference(exe, 1) - t_zero, 1e-9)
            variants["O3_t1"] = work / t1
            if cores > 8:
                variants["O3_t8"] = work / max(_time_reference(exe, 8) - t_zero, 1e-9)
            shipped = os.path.join(base, "shipped", "ssa_sdpd.exe")
            if os.path.exists(shipped) and t1 < 60.0:
                variants[f"shipped_noopt_t{cores}"] = work / max(_time_reference(shipped, cores) - t_zero, 1e-9)
    except Exception:   # noqa: BLE001 - the extra arms are informative only
        pass
    return {"value": work / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"unmodified reference engine (g++ -O3) on {meta['builder']}{meta['kwargs']}: N={meta['N']} x {meta['nt']} steps, -t {cores}, "
                      f"{dt:.2f} s stepping loop (wall minus {t_zero:.2f} s zero-step run)",
            "variants": {k: round(v, 1) for k, v in variants.items()}}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def _peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per particle per launch of `kernel`, from this round's `ncu --set full` capture
    (profiles/ncu_traffic.json, written by profiles/extract_traffic.py from the committed raw-page CSVs); None if not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(path))[workload][kernel]["bytes_per_particle"])
    except (OSError, KeyError, ValueError, TypeError):
        return None


def kernel_bytes_table(fm, moving):
    """Algorithmic bytes per particle per launch (DESIGN.md section 4 = SURVEY.md 8d per-phase figures)."""
    Sc, Sd, R = fm.num_chem_species, fm.num_stoch_species, fm.num_stoch_rxns
    return {
        "force": (104 + 8 * Sc + 56 + 8 * Sc) if moving else (48 + 16 * Sc),
        "rdme_window": 8 + 4 * Sd,
        "finish": (80 + 16 * Sc + 56 + 8 * Sc) if moving else (24 * Sc + 8),
        "predictor": (116 + 16 * Sc + 88 + 8 * Sc) if moving else (24 * Sc + 8),
        "cells": 96, "search": 24 + 24 + 4, "corrector": 70 + 32, "diff_init": 68 + 12 * Sd, "rdme_init": 4 * Sd + 8 * R + 24,
        "output": 0,
    }


def roofline_of(workload, fm, moving, prof, N, nnz, value_per_gpu):
    """Roofline record of the dominant kernel from the live per-category CUDA-event timers (ssb_profile)."""
    peak, peak_src = _peak()
    kb = kernel_bytes_table(fm, moving)
    dom = max(prof, key=lambda k: prof[k]["ms"])
    mean_nbr = nnz / max(N, 1)
    stream_bytes = dict(kb)
    # bytes the kernel must stream given the stored index-only candidate lists: 4 B index per pair (+ 8 B cached coefficient
    # per pair on the static fast path) on top of the SURVEY figure
    stream_bytes["force"] = kb["force"] + (12.0 if not moving else 4.0) * mean_nbr
    dn = max(prof[dom]["launches"], 1)
    dom_ms = prof[dom]["ms"] / dn
    achieved = kb[dom] * N / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    traffic = _ncu_traffic(workload, dom)
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (traffic * N if traffic else None),
            "traffic_note": "dram bytes per launch of the dominant kernel: profiles/ncu_traffic.json (this round's ncu --set full capture), scaled to this N",
            "stream_achieved": (stream_bytes[dom] * N / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0),
            "stream_frac": (stream_bytes[dom] * N / (dom_ms / 1e3) / 1e9 / peak if dom_ms > 0 else 0.0),
            "stream_note": "algorithmic bytes + the stored neighbour-list stream (4 B/pair index, +8 B/pair cached coefficient when static)",
            "peak_source": peak_src, "avg_launch_ms": dom_ms, "launches": prof[dom]["launches"],
            "share_of_step": prof[dom]["ms"] / max(sum(p["ms"] for p in prof.values()), 1e-30),
            "algorithmic_bytes_per_particle": kb[dom],
            "whole_step_frac": (algorithmic_bytes(fm, moving) * value_per_gpu) / (peak * 1e9),
            "kernels_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}}, mean_nbr, dom_ms


def measure_single(args, workload, fm, desc, rank, local_rank, world, steps, warmup, SPS, comm, with_fp64=False, with_cpu=False):
    """One trajectory of `fm` per GPU: device-resident value, roofline of the dominant kernel, end-to-end through ssb_run."""
    from spatialpy_b200.engine import Engine, FLAG_NO_VTK, FLAG_SKIP_STATIC_FORCES
    import numpy as np
    moving = not fm.static_domain
    N = fm.num_particles
    eng = Engine(fm, device=local_rank, flags=FLAG_SKIP_STATIC_FORCES | FLAG_NO_VTK | args.extra_flags)
    # ---- device-resident run: every bench step is the first SPS engine steps of a fresh trajectory (the same segment the
    # end-to-end arm runs); the state upload (ssb_reset) happens BEFORE the timed region of each step -------------------
    for w in range(warmup):
        eng.reset(1000 + rank + 31 * w)
        eng.step_timed(SPS)
    ev_total, win_total, launches = 0, 0, 0
    eng.profile(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    comm.barrier()
    dev_ms = 0.0
    for k in range(steps):
        eng.reset(5000 + rank + 17 * k)            # untimed: inputs are resident in HBM when the timed region starts
        l0 = eng.launch_count()
        dev_ms += eng.step_timed(SPS)              # CUDA events on the engine stream, synchronised on both sides
        c = eng.counters()
        ev_total += c["reactions"] + c["diffusions"]
        win_total += c["windows"]
        launches += eng.launch_count() - l0
    comm.barrier()
    clk = clocks.stop()
    prof = eng.profile_read()
    eng.profile(False)
    skin = eng.skin_stats() if moving else None
    dev_ms = comm.allmax(dev_ms)
    events = comm.allsum(float(ev_total))
    value = world * N * SPS * steps / (dev_ms / 1e3)
    cap, nnz = eng.nbr_stats()
    roofline, mean_nbr, dom_ms = roofline_of(workload, fm, moving, prof, N, nnz, value / world)
    eng.close()

    # ---- end to end: the C-ABI call behind Solver.run, host buffers in, host buffers out -------------------------------
    fm.nt = SPS
    fm.output_steps = np.array([0, SPS], dtype="uint32")
    t_create0 = time.perf_counter()
    eng2 = Engine(fm, device=local_rank, flags=FLAG_SKIP_STATIC_FORCES | FLAG_NO_VTK | args.extra_flags)
    create_s = time.perf_counter() - t_create0
    eng2.run_no_files(2000 + rank, 1)           # warm-up trajectory
    comm.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        eng2.run_no_files(3000 + rank + 17 * k, 1)   # reset (H2D of the initial state) + SPS steps + output snapshots (D2H)
    comm.barrier()
    e2e_s = comm.allmax(time.perf_counter() - t0)
    h2d, d2h = eng2.io_bytes()
    e2e_value = world * N * SPS * steps / e2e_s
    eng2.close()

    if with_fp64 and rank == 0:
        # fp64 companion of the HBM roofline: the moving-domain sweeps are fp64-issue / gather bound, so their flop
        # rate is reported against the MEASURED fp64 FMA peak of this device (child process: it cannot take the line with it)
        pk = fp64_peak_sample(local_rank)
        roofline["fp64"] = pk
        if "fp64_tflops" in pk and moving and prof["force"]["launches"]:
            sweep_ms = prof["force"]["ms"] / prof["force"]["launches"]
            if sweep_ms > 0:
                tf = 230.0 * mean_nbr * N / (sweep_ms / 1e3) / 1e12
                pk.update(force_sweep_tflops=tf, force_sweep_frac=tf / pk["fp64_tflops"], force_sweep_ms=sweep_ms,
                          note="SURVEY 8(d) pair figure: 230 flop per stored candidate-list entry of the force sweep")
    cpu = cpu_baseline_sample(workload) if (with_cpu and rank == 0 and world == 1 and not args.no_cpu) else None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "particles_per_gpu": N, "engine_steps_per_step": SPS, "static_domain": not moving,
                   "species": fm.num_species, "reactions": fm.num_reactions, "mean_neighbours": nnz / N,
                   "sssa_windows_per_engine_step": win_total / max(SPS * steps, 1),
                   "trajectory_segment": f"each bench step = engine steps 0..{SPS} of a fresh trajectory (incl. the step-0 list build on the first one)",
                   "parallelism": f"one trajectory per GPU x {world}" + ("" if world == 1 else " (independent replicas, no collective)"),
                   "verlet_skin": skin,
                   "l2": "working set (candidate lists + gather records + state) exceeds the 126 MB L2; no explicit flush"},
        "rdme_events_per_s": events / (dev_ms / 1e3),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "includes": "ssb_run (the C-ABI call Solver.run makes): H2D of the initial state from host memory, stepping, output snapshots D2H "
                            "into pinned host memory; no VTK text", "engine_create_s": create_s},
        "gpu_launches": int(launches),
        "clocks": clk,
    }


class Comm:
    """barrier / max / sum over the ranks of the torchrun launch (NCCL), or no-ops at N = 1."""

    def __init__(self, local_rank, world):
        import torch
        self.torch, self.world = torch, world
        torch.cuda.set_device(local_rank)
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _red(self, v, op):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(self, v):
        return self._red(v, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def allsum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM if self.world > 1 else None)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def same_n_record(args, workload, rank, local_rank, comm):
    """Our engine on EXACTLY the instance the reference arm runs (oracle/_ref/<name>/meta.json): the like-for-like companion of
    the headline ratio.  Small models are launch-latency bound on a B200 — this is the honest same-size number, not the headline."""
    inst = _reference_instance(workload)
    if inst is None or rank != 0:
        return None
    from spatialpy_b200 import configs
    meta = inst[3]
    fm = getattr(configs, meta["builder"])(**meta["kwargs"])
    nt = int(meta["nt"])
    line = measure_single(args, workload, fm, f"{meta['builder']}{meta['kwargs']}", rank, local_rank, 1, 5, 2, nt, comm)
    return {"particles": fm.num_particles, "engine_steps_per_step": nt, "value": line["value"], "e2e": line["e2e"]["value"], "unit": UNIT,
            "ms_per_step": line["ms_per_step"], "note": "same model, same size, same number of steps as `--impl reference` runs on the host cores"}


def run_ours(args, rank, local_rank, world):
    comm = Comm(local_rank, world)
    fm, desc = make_workload(args.workload, args.scale)
    moving = not fm.static_domain
    SPS = args.sps if args.sps is not None else (50 if moving else 200)
    line = measure_single(args, args.workload, fm, desc, rank, local_rank, world, args.steps, args.warmup, SPS, comm,
                          with_fp64=True, with_cpu=True)
    if world == 1 and not args.no_sub:
        try:
            line["same_n"] = same_n_record(args, args.workload, rank, local_rank, comm)
        except Exception as err:      # noqa: BLE001 - never lose the headline line to a companion measurement
            line["same_n"] = {"error": f"{type(err).__name__}: {err}"[:300]}
        subs = {}
        for wl in ("tank", "cylinder"):
            if wl == args.workload:
                continue
            try:
                f2, d2 = make_workload(wl, args.scale)
                sps2 = 50 if not f2.static_domain else 200
                sub = measure_single(args, wl, f2, d2, rank, local_rank, 1, 3, 3, sps2, comm)
                subs[wl] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "rdme_events_per_s", "roofline", "e2e", "gpu_launches", "clocks")}
            except Exception as err:  # noqa: BLE001
                subs[wl] = {"error": f"{type(err).__name__}: {err}"[:300]}
        for key, name, ntraj, batch in (("ens_birth_death", "birth_death", 256, 256), ("ens_cdc42_full", "cdc42_full", 128, 64)):
            try:
                subs[key] = ensemble_sub_record(name, ntraj, batch, local_rank, not args.no_cpu) if rank == 0 else None
            except Exception as err:  # noqa: BLE001
                subs[key] = {"error": f"{type(err).__name__}: {err}"[:300]}
        line["sub_records"] = subs
    if rank == 0:
        print(json.dumps(line))
    comm.close()


# ---------------------------------------------------------------------------------------------------------
# slab-decomposed arm: ONE domain split over the ranks (BASELINE configs[4]: weak scaling, fixed particles per GPU)
# ---------------------------------------------------------------------------------------------------------
def run_slab(args, rank, local_rank, world):
    comm = Comm(local_rank, world)
    line = slab_measure(args, rank, local_rank, world, args.steps, args.warmup, args.sps or 50, comm)
    if rank == 0:
        print(json.dumps(line))
    comm.close()


def fp64_peak_sample(device):
    """Measured fp64 FMA peak (TFLOP/s) from `python -m spatialpy_b200.peaks` (libssb_peaks.so, include/ssb_peaks.h)."""
    import subprocess
    try:
        out = subprocess.run([sys.executable, "-m", "spatialpy_b200.peaks", str(device)], cwd=ROOT, capture_output=True,
                             text=True, timeout=120)
        if out.returncode != 0:
            tail = (out.stderr.strip().splitlines() or ["?"])[-1]
            return {"error": tail[:200]}
        res = json.loads(out.stdout.strip().splitlines()[-1])
        res["how"] = "issue-bound DFMA loop, 8 chains x 2048 threads per SM, best of 4 launches (CUDA events)"
        return res
    except Exception as err:   # noqa: BLE001 - a helper measurement must never cost the bench line
        return {"error": f"{type(err).__name__}: {err}"[:200]}


def slab_measure(args, rank, local_rank, world, steps, warmup, SPS, comm):
    """ONE box of world x n^3 particles (n = 200 at scale 1: 8 M per GPU), slab-decomposed along x over the ranks; the bench line."""
    import numpy as np
    from spatialpy_b200 import configs
    from spatialpy_b200.slab import SlabEngine
    n = max(16, int(round(200 * args.scale ** (1.0 / 3.0))))
    part = configs.box_slab(rank, world, nx_per_rank=n, ny=n, nz=n)
    se = SlabEngine(part, rank, world, device=local_rank, transport=args.transport)
    fm = part.local
    # the same protocol as the N = 1 line: every bench step is the first SPS engine steps of a fresh trajectory (state upload
    # untimed, the step-0 list build and the initial A + B annihilation transient inside the timed region)
    for w in range(warmup):
        se.reset(1000 + 31 * w)
        se.step(SPS)
    se.eng.profile(True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    comm.barrier()
    dev_ms, wall_s, launches, ev = 0.0, 0.0, 0, 0
    for k in range(steps):
        se.reset(5000 + 17 * k)                 # untimed: the rank's initial state is resident in HBM when the timed region starts
        comm.barrier()
        l0 = se.eng.launch_count()
        t0 = time.perf_counter()
        se.eng.mark(0)
        se.step(SPS)
        se.eng.mark(1)
        dev_ms += se.eng.mark_elapsed_ms()      # CUDA events on the engine stream (includes the waits for the halo messages)
        wall_s += time.perf_counter() - t0
        c1 = se.eng.counters()
        ev += c1["reactions"] + c1["diffusions"]
        launches += int(se.eng.launch_count() - l0)
    comm.barrier()
    clk = clocks.stop()
    prof = se.eng.profile_read()
    se.eng.profile(False)
    dev_ms = comm.allmax(dev_ms)
    wall_s = comm.allmax(wall_s)
    owned_total = comm.allsum(float(part.n_owned))
    events = comm.allsum(float(ev))
    nsteps = SPS * steps
    value = owned_total * nsteps / (dev_ms / 1e3)
    cap, nnz = se.eng.nbr_stats()
    roofline, mean_nbr, dom_ms = roofline_of("box", fm, True, prof, part.n_owned, nnz, value / world)
    ghosts = fm.num_particles - part.n_owned
    Sc = fm.num_chem_species
    halo_bytes = sum(len(v) for v in part.send_ids.values()) * 8 * (7 + Sc + 1 + 4)
    # ---- end to end: host buffers in (ssb_reset uploads the rank's initial state), the same steps, owned state read back to the host
    comm.barrier()
    t1 = time.perf_counter()
    se.reset(2000)
    se.step(SPS)
    d2h = 0
    for f in ("x", "v", "rho", "C", "xx"):
        d2h += se.eng.get(f).nbytes
    comm.barrier()
    e2e_s = comm.allmax(time.perf_counter() - t1)
    h2d = se.eng.io_bytes()[0]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"BASELINE configs[4]: synthetic 3-D SDPD+sSSA box of {world} x {n}^3 particles, slab-decomposed along x",
                   "particles_per_gpu": part.n_owned, "ghosts_per_gpu": ghosts, "engine_steps_per_step": SPS,
                   "trajectory_segment": f"each bench step = engine steps 0..{SPS} of a fresh trajectory (incl. the step-0 list build), as at N = 1",
                   "mean_neighbours": nnz / max(part.n_owned, 1),
                   "parallelism": (f"spatial slabs x {world}; transport = {se.transport}: " +
                                   ("pack kernels store into the neighbour's receive window over NVLink (CUDA IPC peer memory), sequence flags, "
                                    "stream-ordered; 3 field groups + sSSA inbox entries + 3 scalar boards per step" if se.transport == "native"
                                    else "host-orchestrated NCCL send/recv")),
                   "halo_bytes_sent_per_engine_step_per_rank": halo_bytes, "repartitions": se.repartitions,
                   "l2": "working set exceeds the 126 MB L2; no explicit flush"},
        "rdme_events_per_s": events / (dev_ms / 1e3),
        "roofline": roofline,
        "cpu_baseline": None,
        "e2e": {"value": owned_total * SPS / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "includes": "per rank: ssb_reset (H2D of the initial state from host memory) + the same engine steps with halo exchanges + "
                            "owned x, v, rho, C, xx read back to host memory; host clock, max over ranks",
                "stepping_loop_wall_value": owned_total * nsteps / wall_s},
        "gpu_launches": launches, "clocks": clk}
    se.close()
    return line


# ---------------------------------------------------------------------------------------------------------
# ensemble arm: many trajectories of a SMALL model (BASELINE configs[0] birth-death, configs[3] Cdc42), k mod G over the GPUs
# and several concurrent engine handles per GPU
# ---------------------------------------------------------------------------------------------------------
def _ensemble_cpu_arm(name, fm):
    """Reference arm of an ensemble: trajectories are independent processes (solver.py:547-605), so the host runs
    one per core side by side; trajectories/s = processes / wall.  None when oracle/_ref/bench_<name>/ was not built."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"bench_{name}", "fast", "ssa_sdpd.exe")
    if not os.path.exists(exe):
        return None
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    procs, dirs = [], []
    for k in range(cores):
        d = tempfile.mkdtemp(prefix="ssb_ref_ens_")
        dirs.append(d)
        procs.append(subprocess.Popen([exe, "-t", "1", "-s", str(1000 + k)], cwd=d, stdout=subprocess.DEVNULL))
    for p in procs:
        p.wait()
    wall = time.perf_counter() - t0
    for d in dirs:
        subprocess.run(["rm", "-rf", d])
    return {"value": fm.num_particles * fm.nt * cores / wall, "unit": UNIT, "cores": cores, "kind": "reference",
            "trajectories_per_s": cores / wall,
            "sample": f"unmodified reference engine (g++ -O3), {cores} trajectories of the same model as {cores} concurrent processes "
                      f"(-t 1 each), {wall:.2f} s wall incl. process start and VTK output"}


def ensemble_sub_record(name, ntraj, batch, device, with_cpu):
    """A bounded single-GPU sample of an ensemble config (BASELINE configs[0] birth-death, configs[3] Cdc42 at its named size) for the
    default line's `sub_records`: `ntraj` trajectories, `batch` at a time as disjoint copies in one engine handle (what `Solver.run`
    picks from 16 trajectories on), host wall clock around the whole call (state upload, stepping, read-back of every trajectory's populations)."""
    import torch
    from spatialpy_b200 import FlatModel
    from spatialpy_b200.ensemble import run_ensemble_batched
    fm = FlatModel.load(os.path.join(ROOT, "tests", "golden", f"{name}.model.npz"))
    run_ensemble_batched(fm, min(batch, 8), 1, device=device, batch=min(batch, 8))      # warm-up: unit build, module load
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    res = run_ensemble_batched(fm, ntraj, 1000, device=device, batch=batch)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    ev = float(res["counters"]["reactions"] + res["counters"]["diffusions"])
    return {"value": fm.num_particles * fm.nt * ntraj / dt, "unit": UNIT, "trajectories": ntraj, "trajectories_per_s": ntraj / dt,
            "rdme_events_per_s": ev / dt, "ms_per_step": dt * 1e3,
            "config": {"workload": f"ensemble of {ntraj} trajectories of the {name} fixture model ({fm.num_particles} particles, {fm.nt} steps, "
                                   f"{fm.num_species} species, {fm.num_reactions} reactions), one engine handle holding {batch} copies at a time, on one GPU"},
            "cpu_baseline": _ensemble_cpu_arm(name, fm) if with_cpu else None}


def run_ensemble_bench(args, rank, local_rank, world):
    import torch
    from spatialpy_b200 import FlatModel
    from spatialpy_b200.ensemble import default_lanes, run_ensemble
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # ens_cdc42_full = BASELINE configs[3] at its named size (create_cdc42_model(DX=50): 2 500 particles) on a bounded horizon
    # (t = 0.001, ~6e6 events per trajectory; the notebook's t = 100 is ~6e11 events per trajectory for any engine)
    name = {"ens_birth_death": "birth_death", "ens_cdc42": "cdc42", "ens_cdc42_full": "cdc42_full"}[args.workload]
    fm = FlatModel.load(os.path.join(ROOT, "tests", "golden", f"{name}.model.npz"))
    lanes = args.lanes or default_lanes(fm.num_particles)
    per_gpu = args.trajectories
    total = per_gpu * world

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()

    if args.batch:
        # opt-in: `--batch n` trajectories per engine handle as disjoint copies of the model (ensemble.replicate_model); rank r runs
        # its per_gpu trajectories with seeds 1000 + r * per_gpu ...
        from spatialpy_b200.ensemble import run_ensemble_batched
        run_ensemble_batched(fm, min(args.batch, per_gpu), 1, device=local_rank, batch=args.batch)              # warm-up
        barrier()
        t0 = time.perf_counter()
        resb = run_ensemble_batched(fm, per_gpu, 1000 + rank * per_gpu, device=local_rank, batch=args.batch)
        barrier()
        dt = time.perf_counter() - t0
        ev = float(resb["counters"]["reactions"] + resb["counters"]["diffusions"])
        lanes = f"1 handle x {min(args.batch, per_gpu)} copies"
    else:
        run_ensemble(fm, lanes * world, 1, devices=[local_rank], lanes=lanes, rank=rank, world_size=world)      # warm-up (JIT, contexts)
        barrier()
        t0 = time.perf_counter()
        res = run_ensemble(fm, total, 1000, devices=[local_rank], lanes=lanes, rank=rank, world_size=world)
        barrier()
        dt = time.perf_counter() - t0
        ev = float(sum(c["reactions"] + c["diffusions"] for c in res.values()))
    if use_dist:
        t = torch.tensor([dt, ev], dtype=torch.float64, device="cuda")
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, ev = float(tm[0].item()), float(t[1].item())
    cpu = _ensemble_cpu_arm(name, fm) if rank == 0 and not args.no_cpu else None
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": fm.num_particles * fm.nt * total / dt, "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"ensemble of {total} trajectories of the {name} fixture model ({fm.num_particles} particles, {fm.nt} steps, "
                                   f"{fm.num_species} species, {fm.num_reactions} reactions)", "trajectories_per_gpu": per_gpu,
                       "concurrent_engine_handles_per_gpu": lanes, "parallelism": f"ensemble: trajectory k -> GPU k mod {world}"},
            "trajectories_per_s": total / dt, "rdme_events_per_s": ev / dt, "cpu_baseline": cpu,
            "e2e": {"value": fm.num_particles * fm.nt * total / dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "includes": "host wall clock of ssb_run per trajectory incl. state upload and output staging (no VTK text)"}}))
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="box", choices=["box", "tank", "cylinder", "ens_birth_death", "ens_cdc42", "ens_cdc42_full"])
    ap.add_argument("--trajectories", type=int, default=128, help="ensemble workloads: trajectories per GPU")
    ap.add_argument("--lanes", type=int, default=0, help="ensemble workloads: concurrent engine handles per GPU (0 = auto)")
    ap.add_argument("--batch", type=int, default=0, help="ensemble workloads: trajectories per engine handle as disjoint copies of the model (0 = off)")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full particle count (1.0 = BASELINE size)")
    ap.add_argument("--sps", type=int, default=None, help="engine timesteps per bench step (default: 200 static, 50 moving, 20 slab)")
    ap.add_argument("--extra-flags", type=int, default=0, help="extra SSB_FLAG_* bits for the engine (diagnostics)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--no-sub", action="store_true", help="skip the same-N companion and the cylinder / tank sub-records")
    ap.add_argument("--decomp", default="auto", choices=["auto", "ensemble", "slab"],
                    help="N>1: ONE box domain split into slabs (auto: the box workload) or independent trajectories per GPU")
    ap.add_argument("--transport", default=None, choices=["native", "host"], help="slab runs: halo transport (default native)")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload.startswith("ens_"):
        run_ensemble_bench(args, rank, local_rank, world)
        return
    if args.decomp == "slab" or (args.decomp == "auto" and args.workload == "box" and world > 1):
        run_slab(args, rank, local_rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
