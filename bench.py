#!/usr/bin/env python
"""Benchmark of the TMM hot path (contract: see the task statement; summary in DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]

One "step" = one pass of the fused hot path (table prep + chain kernel) over the whole synthetic grid of BASELINE
configs[1] (DBR microcavity, 8192 wavelengths x 1024 angles x s/p, 83 layers) with all inputs resident in HBM.
At N > 1 (torchrun, one rank per GPU) every rank evaluates its own 8192-wavelength slab of an N-times finer
wavelength grid: weak scaling, no data-path collective.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "TMM spectral points/s (fp64, lambda x theta x pol)"
UNIT = "points/s"
FLOP_PER_LAYER_STEP = 56.0  # first-column mode, transcendentals not counted


def load_measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (recipe in B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ workloads
def make_workload(name: str, rank: int, world: int):
    """The rank's own slab (weak scaling: per-GPU work fixed)."""
    from pistachio_b200 import workloads as wk

    if name == "c2":
        w = wk.c2_dbr_cavity(n_wl=8192 * world, n_theta=1024)
        w.wl = w.wl[rank * 8192:(rank + 1) * 8192]
        desc = "configs[1] DBR microcavity: 8192 wl x 1024 theta x s/p per GPU, 83 layers (81 interior)"
    elif name == "c1":
        w = wk.c1_fabry_perot(2000)
        desc = "configs[0] Fabry-Perot: 2000 wl x 1 theta x s/p, 5 layers"
    elif name == "c3":
        w = wk.c3_polariton(n_d=512 * world, n_wl=4096, n_theta=64)
        w.sweep_d = w.sweep_d[rank * 512:(rank + 1) * 512]
        desc = "configs[2] polariton sweep chunk: 512 d x 4096 wl x 64 theta x s/p per GPU, 5 layers"
    elif name == "c4":
        w = wk.c4_chirped(10000, 16384 * world)
        w.wl = w.wl[rank * 16384:(rank + 1) * 16384]
        desc = "configs[3] chirped mirror: 10000 layers x 16384 wl per GPU, s-wave, scan kernel"
    elif name == "c5":
        w = wk.c5_box_sweep(n_wl=8192, n_theta=256, n_d=32 * world)
        w.sweep_d = w.sweep_d[rank * 32:(rank + 1) * 32]
        desc = "configs[4] box sweep slab: 32 d x 8192 wl x 256 theta x s/p per GPU, 5 layers"
    else:
        raise SystemExit(f"unknown workload {name}")
    return w, desc


def workload_pols(name):
    return ("s-wave",) if name == "c4" else ("s-wave", "p-wave")


# ------------------------------------------------------------------------------------------------ flop model
def algorithmic_flops(w, prep, n_pol, cse=True, power=True):
    """Algorithmic FP64 flops per spectral point of the path the kernel takes for this workload.
    Counted: the multiply/add/FMA(=2) of the phase, the layer-matrix entries and the 2-vector layer step, and the
    epilogue.  NOT counted: sincos / exp evaluations, rescaling, address arithmetic.
      real (lossless-material) layer:   build 4 + 2/pol,  step 12/pol
      complex layer:                    build 20 + 4/pol, step 28/pol   (28 + 28 per pol)
    With layer CSE a matrix is built once per distinct (material, thickness).  K3 per thickness: the phase (2 or 4),
    cos/sin of the complex phase from sincos and cosh/sinh (4, absorbing layer only) and, per polarisation, the two
    2-term sums M00, M10 (2 x 14 complex, 2 x 6 real), |.|^2 of both (6) and R, T, A (5).  K2 carries both columns of
    a segment product; for an all-lossless stack they are four reals (12 flop per layer and polarisation)."""
    lossy = (prep["mat_k"] != 0).any(dim=1).cpu().numpy()
    mats = np.asarray(prep["layer_mat"])
    ds = np.asarray(prep["layer_d"])
    interior = list(range(1, len(mats) - 1))
    build = lambda m: (20 + 4 * n_pol) if lossy[m] else (4 + 2 * n_pol)
    step = lambda m: (28 if lossy[m] else 12) * n_pol
    epilogue = 30 * n_pol
    if w.sweep_d is not None and len(w.sweep_d) >= 4:
        m = mats[w.sweep_layer]
        per_thread = (8 + 39 * n_pol) if lossy[m] else (2 + 23 * n_pol)  # per (wavelength, theta, thickness)
        kind = "K3 sweep (prefix/suffix amortised over the chunk, not counted)"
    else:
        n_types = len({(int(mats[j]), float(ds[j])) for j in interior})
        typed = cse and 0 < n_types <= 8 and len(interior) >= 2 * n_types and w.interior_layers < 128 * 16
        cols = 1
        if w.interior_layers >= 128 and len(w.wl) * len(w.theta) < 148 * 512:
            # K2 segments carry both columns; for an all-lossless stack they are four reals = one column's worth
            cols, typed = (1 if not lossy[mats[interior]].any() else 2), False
        builds = {(int(mats[j]), float(ds[j])): build(mats[j]) for j in interior} if typed else None
        steps = cols * sum(step(mats[j]) for j in interior)
        kind = "K1 typed (layer CSE)" if typed else ("K2 segmented" if (w.interior_layers >= 128 and len(w.wl) * len(w.theta) < 148 * 512) else "K1 untyped")
        if typed and power:
            # run-length ops as the host builds them; a lossless pair repeated k >= 3 times costs per polarisation:
            # period matrix 12 + (k - 2) recurrence FMAs * 2 + closed form 6 + one real-structured step 12
            seq = [(int(mats[j]), float(ds[j])) for j in reversed(interior)]

            def repeats(at):
                k = 1
                while at + 2 * k + 1 < len(seq) and seq[at + 2 * k] == seq[at] and seq[at + 2 * k + 1] == seq[at + 1]:
                    k += 1
                return k

            steps, i, prev_power, shape = 0, 0, None, []
            while i < len(seq):
                if i + 1 < len(seq):
                    k = repeats(i)
                    if k == 1 and i + 2 < len(seq) and repeats(i + 1) >= 2:  # lone layer in front of a periodic run
                        shape.append("single")
                        steps += step(seq[i][0])
                        i += 1
                        continue
                    if k >= 3 and not lossy[seq[i][0]] and not lossy[seq[i + 1][0]]:
                        key = (frozenset((seq[i], seq[i + 1])), k)
                        shape.append("power")
                        if key == prev_power:  # same or mirrored period, same k: the recurrence is shared
                            steps += n_pol * (12 + 6 + 12)
                            i += 2 * k
                            continue
                        prev_power = key
                        steps += n_pol * (12 + 2 * (k - 2) + 6 + 12)
                    else:
                        shape.append("pairs")
                        steps += k * (step(seq[i][0]) + step(seq[i + 1][0]))
                    i += 2 * k
                else:
                    shape.append("single")
                    steps += step(seq[i][0])
                    i += 1
            if shape == ["power", "single", "power"] and prev_power is not None and steps > 0:
                # mirror / spacer / mirror with equal pair counts (the straight-line program): the second mirror reuses
                # the first one's period matrix (exchanged diagonal) -- only its 2-vector step remains
                first_k = prev_power[1]
                n_runs_same = sum(1 for x in shape if x == "power")
                if n_runs_same == 2 and len(seq) == 4 * first_k + 1:
                    steps -= n_pol * (12 + 6)
            kind = "K1 typed (layer CSE + closed-form power of repeated lossless pairs)"
        per_thread = (sum(builds.values()) if typed else sum(build(mats[j]) for j in interior)) + steps + epilogue
    return {"flop_per_point": per_thread / n_pol, "kernel_path": kind}


# ------------------------------------------------------------------------------------------------ CPU baseline
def _cpu_task(args):
    """One (theta, pol) task of the reference call sequence on the oracle port."""
    wl, n_list, d_list, theta, pol = args
    from oracle import tmm_oracle as orc

    T, R = orc.spectrum(orc.Stack(wl, n_list, d_list), theta, pol)
    return float(T.sum() + R.sum())


def cpu_reference_run(name: str, steps: int, warmup: int):
    """Times the oracle port (loop-faithful numpy restatement of the reference, pinned to it in the build container)
    on the host cores with the reference's own fan-out: a multiprocessing pool over (theta, pol) tasks
    (transfer_matrix.py:614-623).  Each step is a bounded sub-sample of the workload."""
    import multiprocessing as mp

    from oracle import tmm_oracle as orc
    from pistachio_b200 import workloads as wk

    w, desc = make_workload(name, 0, 1)
    cores = os.cpu_count() or 1
    n_l = min(len(w.wl), 256 if name != "c4" else max(4, cores))
    n_t = min(len(w.theta), max(8, cores // 2))
    sel_l = np.linspace(0, len(w.wl) - 1, n_l).astype(int)
    sel_t = np.linspace(0, len(w.theta) - 1, n_t).astype(int)
    wl = w.wl[sel_l]
    n_list = []
    for l in w.layers:
        n, k = orc.resample_table(l.tab_wl, l.tab_n, l.tab_k, wl)
        n_list.append(n + 1j * k)
    d_list = [l.thickness for l in w.layers]
    pols = workload_pols(name)
    # tasks = (theta, pol) pairs as in the reference; when there are fewer pairs than cores (single-angle workloads)
    # the wavelength axis is cut as well so that every core has work
    n_cut = max(1, min(n_l, -(-2 * cores // (n_t * len(pols)))))
    cuts = np.array_split(np.arange(n_l), n_cut)
    tasks = [(wl[c], [n[c] for n in n_list], d_list, float(w.theta[i]), p) for i in sel_t for p in pols for c in cuts]
    points = n_l * n_t * len(pols)
    procs = min(cores, len(tasks))
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_task, [(wl[:1], [n[:1] for n in n_list], d_list, 0.0, pols[0])] * procs, chunksize=1)  # spin-up
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_task, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    sample = (f"{n_l} wl x {n_t} theta x {len(pols)} pol = {points} points of {desc}; "
              f"multiprocessing pool of {procs} over {len(tasks)} (theta, pol[, wl-chunk]) tasks")
    return {"value": points / sec, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
            "layer_steps_per_s": points * w.interior_layers / sec, "seconds_per_step": sec, "host_cpus": cores}, w, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, w, desc = cpu_reference_run(args.workload, max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "note": "reference CPU path = oracle port (numpy loops identical to the reference's; "
                   "the pure-Python reference cannot travel to the GPU box), bounded sample per step"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from pistachio_b200 import dist as pdist
    from pistachio_b200.engine import get_engine

    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = None
    if world_env == 1 and not args.no_cpu_baseline:
        # timed first, before this process owns a CUDA context: the pool forks
        cpu, _, _ = cpu_reference_run(args.workload, 1, 0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "layer_steps_per_s")}
    rank, world, local = pdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: pistachio_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    eng = get_engine(local)
    w, desc = make_workload(args.workload, rank, world)
    pols = workload_pols(args.workload)
    prep = eng.prepare_workload(w)
    n_sw = 1 if w.sweep_d is None else len(w.sweep_d)
    shape = (n_sw, len(pols), len(w.wl), len(w.theta))
    points_rank = int(np.prod(shape))
    out = {k: torch.empty(shape, dtype=torch.float64, device=eng.device) for k in ("R", "T", "A")}
    stream = torch.cuda.current_stream()

    flags = (16 if args.no_layer_cse else 0) | (128 if args.no_periodic_power else 0)  # PST_FLAG_NO_LAYER_CSE / _NO_PERIODIC_POWER
    flags |= 512 if args.fp32 else 0  # PST_FLAG_FP32 (optional mode; the default and every reported headline number is fp64)

    def step():
        eng.eval_prepared(prep, pols=pols, out=out, flags=flags)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps (outputs are also
    # 403 MB per step, > the 126 MB L2)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = eng.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        eng.flush_l2()
        a.record(stream)
        step()
        b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = torch.tensor([sum(ms) / 1e3], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    sec_total = float(t_dev.item())
    ms_per_step = sec_total / args.steps * 1e3
    value = points_rank * world * args.steps / sec_total

    # ---- ablations of the algorithmic shortcuts (same grid, same outputs; results agree to rounding / bit for bit)
    ablation = None
    if args.workload == "c2" and flags == 0:
        ablation = {}
        for name, fl in (("layer_by_layer_pairs", 128), ("no_layer_cse", 128 | 16), ("pol_degenerate_opt_in", 1024),
                         ("op_interpreter", 2048)):
            for _ in range(3):
                eng.eval_prepared(prep, pols=pols, out=out, flags=fl)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            barrier()
            for a_, b_ in evs:
                eng.flush_l2()
                a_.record(stream)
                eng.eval_prepared(prep, pols=pols, out=out, flags=fl)
                b_.record(stream)
            barrier()
            t_ab = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in evs) / 1e3], dtype=torch.float64, device=eng.device)
            if world > 1:
                dist.all_reduce(t_ab, op=dist.ReduceOp.MAX)
            ablation[name] = {"value": points_rank * world * 20 / float(t_ab.item()), "ms_per_step": float(t_ab.item()) / 20 * 1e3}
        eng.eval_prepared(prep, pols=pols, out=out, flags=flags)  # leave the default path's results in `out`
        ablation["note"] = ("layer_by_layer_pairs: PST_FLAG_NO_PERIODIC_POWER (every layer applied one by one, cached "
                            "matrices); no_layer_cse: additionally rebuild every layer matrix (sincos per layer); "
                            "op_interpreter: PST_FLAG_NO_SHAPE_FASTPATH (the mirror/spacer/mirror program through the general op loop "
                            "instead of straight-line code; bit-identical); "
                            "pol_degenerate_opt_in: PST_FLAG_POL_DEGENERATE -- NOT the default and not part of `value`: "
                            "the reference's model makes s and p the same function (similarity transform), so one chain "
                            "is computed and written to both planes")

    # ---- e2e: the plugin-style call with HOST buffers (pinned), H2D of the inputs and D2H of R, T, A in the timed region
    mat_n_h = prep["mat_n"].cpu().pin_memory()
    mat_k_h = prep["mat_k"].cpu().pin_memory()
    wl_h = torch.from_numpy(np.ascontiguousarray(w.wl)).pin_memory()
    th_h = torch.from_numpy(np.ascontiguousarray(w.theta)).pin_memory()
    sw_h = torch.from_numpy(np.ascontiguousarray(w.sweep_d)).pin_memory() if w.sweep_d is not None else None
    out_h = {k: torch.empty(shape, dtype=torch.float64).pin_memory() for k in ("R", "T", "A")}

    def e2e_step():
        eng.eval_grid_host(wl_h, th_h, mat_n_h, mat_k_h, prep["layer_mat"], prep["layer_d"], pols,
                           sweep_layer=w.sweep_layer, sweep_d=sw_h, out=out_h, flags=flags)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = points_rank * world * e2e_steps / float(t_e2e.item())
    h2d = 8 * (wl_h.numel() + th_h.numel() + mat_n_h.numel() + mat_k_h.numel() + (sw_h.numel() if sw_h is not None else 0))
    d2h = 8 * 3 * points_rank
    same = all(bool((out_h[k] == out[k].cpu()).all()) for k in ("R", "T", "A"))

    # ---- sweep workloads: the same end-to-end call with the fused spectral reduction --
    # transmittance spectra stay on the device, only peak lists (the "Cavity Length Determiner" analysis) come back
    fused = None
    if w.sweep_d is not None:
        nu_d = 1.0e4 / prep["wl"]
        max_peaks = 32
        pk_h = {"idx": torch.empty(shape[:2] + (shape[3], max_peaks), dtype=torch.int32).pin_memory(),
                "cnt": torch.empty(shape[:2] + (shape[3],), dtype=torch.int32).pin_memory(),
                "spc": torch.empty(shape[:2] + (shape[3],), dtype=torch.float64).pin_memory()}

        def fused_step():
            dev = {k: v.to(eng.device, non_blocking=True) for k, v in (("wl", wl_h), ("th", th_h), ("n", mat_n_h),
                                                                       ("k", mat_k_h), ("sw", sw_h))}
            res = eng.eval_grid(dev["wl"], dev["th"], dev["n"], dev["k"], prep["layer_mat"], prep["layer_d"], pols,
                                sweep_layer=w.sweep_layer, sweep_d=dev["sw"], want=("T",), out={"T": out["T"]}, flags=flags)
            idx, _, cnt, spc = eng.find_peaks(res.T, height=0.01, distance=20, max_peaks=max_peaks, axis=nu_d, reverse=True)
            pk_h["idx"].copy_(idx, non_blocking=True)
            pk_h["cnt"].copy_(cnt, non_blocking=True)
            pk_h["spc"].copy_(spc, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            fused_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fused_step()
        barrier()
        t_f = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(t_f, op=dist.ReduceOp.MAX)
        fused = {"value": points_rank * world * e2e_steps / float(t_f.item()), "unit": UNIT,
                 "d2h_bytes_per_step": sum(v.numel() * v.element_size() for v in pk_h.values()),
                 "what": "eval_grid (T only) + pst_find_peaks(height=1 %, distance=20) on the device; peak indices, counts and "
                         "mean peak spacing are the only results copied to the host",
                 "spectra": int(np.prod(shape[:2]) * shape[3]), "overflowed_spectra": int((pk_h["cnt"] < 0).sum())}

    # ---- optional assembly on every device: one NCCL all-gather per output (not part of `value`)
    gather = None
    if world > 1:
        axis = "sweep" if n_sw > 1 else "wl"
        sizes = [shape[0] if axis == "sweep" else shape[2]] * world
        # NCCL connection set-up and the allocator's first cudaMalloc of the three assembled arrays are not part of the number
        warm = [pdist.allgather_result(out[k], axis, sizes) for k in ("R", "T", "A")]
        del warm
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        full = [pdist.allgather_result(out[k], axis, sizes) for k in ("R", "T", "A")]
        g1.record(stream)
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=eng.device)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        recv = 3 * 8 * points_rank * (world - 1)
        gather = {"ms": float(gms.item()), "recv_GB_per_rank": recv / 1e9, "GBps_per_rank": recv / 1e9 / (float(gms.item()) / 1e3)}
        del full
        # the same assembly fused into the compute kernel: every result is stored straight into all GPUs' copies
        # (symmetric memory, one peer store per GPU over NVLink -- dist.AssembledResult); layout [rank slab][...slab dims]
        try:
            asm = pdist.AssembledResult(world, shape, rank, world, eng.device, mode="unicast")

            def fused_step():
                asm.eval_into(eng, prep, pols, flags=flags)
                asm.finish()

            fused_step()
            best = None
            for _ in range(5):
                barrier()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record(stream)
                fused_step()
                f1.record(stream)
                barrier()
                fms = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=eng.device)
                dist.all_reduce(fms, op=dist.ReduceOp.MAX)
                best = float(fms.item()) if best is None else min(best, float(fms.item()))
            same_f = all(bool(torch.equal(asm.full[k][rank], out[k])) for k in ("R", "T", "A"))
            gather["fused_compute_plus_assembly_ms"] = best
            gather["nccl_compute_plus_allgather_ms"] = ms_per_step + float(gms.item())
            gather["fused_identical_to_local_result"] = same_f
            gather["fused_how"] = "unicast peer stores from the chain kernel into symmetric memory (pst_peer_outputs), then a device barrier"
            del asm
        except Exception as e:
            gather["fused_unavailable"] = repr(e)[:200]

    if rank != 0:
        return
    peaks, peaks_src = load_measured_peaks()
    fp64_peak = eng.fp64_peak_tflops(8192)
    layer_steps = points_rank * w.interior_layers
    flop_model = algorithmic_flops(w, prep, len(pols), cse=not (args.no_layer_cse or args.fp32), power=not args.no_periodic_power)
    achieved_tf = flop_model["flop_per_point"] * points_rank / (ms_per_step / 1e3) / 1e12
    survey_tf = layer_steps * FLOP_PER_LAYER_STEP / (ms_per_step / 1e3) / 1e12
    hbm_gbs = 24.0 * points_rank / (ms_per_step / 1e3) / 1e9
    fp64_frac = achieved_tf / fp64_peak if fp64_peak else 0.0
    hbm_frac = hbm_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else 0.0
    if hbm_frac > fp64_frac:
        roof = {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": hbm_frac,
                "traffic": None, "bytes_per_point": 24, "peak_source": f"MEASURED_PEAKS.json ({peaks_src})"}
    else:
        roof = {"bound": "fp64_pipe", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_frac,
                "traffic": None,
                "peak_source": "independent-DFMA probe run in this process (pst_fp64_peak_probe); MEASURED_PEAKS.json has no FP64 entry"}
    roof["fp64"] = {"achieved_tflops": achieved_tf, "peak_tflops": fp64_peak, "frac": fp64_frac, **flop_model,
                    "survey_56flop_per_layer_step_tflops": survey_tf,
                    "note": "algorithmic flops of the formulas the kernel evaluates (FMA = 2), transcendentals, rescaling and "
                            "address arithmetic not counted; see DESIGN.md section 4 and the ncu FP64-pipe utilisation in profiles/"}
    roof["hbm_store"] = {"achieved_gbs": hbm_gbs, "peak_gbs": peaks.get("hbm_gbs"), "frac": hbm_frac}
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this workload
        import glob

        cands = []
        for path in glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")):
            with open(path) as fh:
                t = json.load(fh)
            if t.get("workload") == args.workload and flags == t.get("flags", 0):
                import re

                cands.append((tuple(int(x) for x in re.findall(r"\d+", os.path.basename(path))), os.path.basename(path), t))
        tr = max(cands, key=lambda c: c[:2])[2] if cands else {}  # newest capture of this workload (files are named rNN_vMM_traffic.json)
        if tr:
            roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roof["traffic_source"] = tr["source"]
            roof["algorithmic_bytes"] = 24 * points_rank
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 chain / f64 phase (optional mode)" if args.fp32 else "f64",
        "data": "synthetic",
        "config": {"workload": desc, "layer_cse": not args.no_layer_cse, "periodic_power": not (args.no_periodic_power or args.no_layer_cse), "points_per_gpu": points_rank, "interior_layers": w.interior_layers,
                   "layer_steps_per_s": value * w.interior_layers, "l2": "flushed between steps (256 MiB memset); outputs 403 MB/step > L2",
                   "timing": "CUDA events per step on the launch stream, max over ranks", "wall_s_incl_flush": t_wall},
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "identical_to_device_path": same,
                "d2h_gbs_per_gpu": d2h * e2e_steps / float(t_e2e.item()) / 1e9,
                "bound": "PCIe device->host copy of R, T, A (24 B/point); kernels overlap the copies",
                "api": "pst_eval_grid_host via TMMEngine.eval_grid_host, pinned host buffers"},
        "ablation": ablation,
        "e2e_fused_reduction": fused,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if gather:
        line["allgather"] = gather
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layer-cse", action="store_true", help="ablation: rebuild every layer matrix (PST_FLAG_NO_LAYER_CSE)")
    ap.add_argument("--fp32", action="store_true", help="optional fp32 mode (PST_FLAG_FP32): fp64 phase, fp32 chain, <= 1e-5")
    ap.add_argument("--no-periodic-power", action="store_true",
                    help="ablation: apply repeated lossless pairs layer by layer (PST_FLAG_NO_PERIODIC_POWER)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
