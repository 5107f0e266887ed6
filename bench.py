#!/usr/bin/env python
"""Benchmark of the TMM hot path (contract: see the task statement; summary in DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]

One "step" = one pass of the fused hot path (table prep + chain kernel) over the whole synthetic grid of BASELINE
configs[1] (DBR microcavity, 8192 wavelengths x 1024 angles x s/p, 83 layers) with all inputs resident in HBM.
At N > 1 (torchrun, one rank per GPU) every rank evaluates its own 8192-wavelength slab of an N-times finer
wavelength grid: weak scaling, no data-path collective.  Rank 0 prints ONE JSON line.  Besides the headline keys it
carries `configs` (every other BASELINE config, device-timed with its own roofline), `sustained` (a >= 2 s
back-to-back segment with clocks sampled), `e2e` (host buffers through pst_eval_grid_host, with the probed pinned
copy rate as its roofline) and, at N > 1, `assembly` (configs[4] at full size: strong scaling, with and without
gathering the results on one / on every device).
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
    """SM clock and throttle reasons sampled DURING a timed region.  NVML from a thread (about 1 ms per sample) so that
    even a 40 ms region collects samples; falls back to an nvidia-smi loop (50 ms period) when NVML is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.002):
        self.index, self.period = index, period_s
        self.sm, self.power, self.reasons = [], [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self.proc = None
        self.how = None

    def _nvml_loop(self, nv, h):
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.REASONS:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def _smi_pump(self):
        names = [n for n, _ in self.REASONS]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                self.sm.append(float(parts[0]))
                self.sm_max = max(self.sm_max or 0.0, float(parts[1]))
                self.power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.replace(",", "").isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml"
            self._thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self._thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi -lms 50"
            threading.Thread(target=self._smi_pump, daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable"]}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": self.sm_max, "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.sm), "reasons": sorted(self.reasons), "how": self.how}


# ------------------------------------------------------------------------------------------------ workloads
DESCS = {
    "c2": "configs[1] DBR microcavity: 8192 wl x 1024 theta x s/p per GPU, 83 layers (81 interior)",
    "c1": "configs[0] Fabry-Perot: 2000 wl x 1 theta x s/p, 5 layers",
    "c3": "configs[2] polariton sweep chunk: 512 d x 4096 wl x 64 theta x s/p per GPU, 5 layers",
    "c4": "configs[3] chirped mirror: 10000 layers x 16384 wl per GPU, s-wave, scan kernel",
    "c5": "configs[4] box sweep slab: 32 d x 8192 wl x 256 theta x s/p per GPU, 5 layers",
}


def make_workload(name: str, rank: int, world: int):
    """The rank's own slab (weak scaling: per-GPU work fixed)."""
    from pistachio_b200 import workloads as wk

    if name == "c2":
        w = wk.c2_dbr_cavity(n_wl=8192 * world, n_theta=1024)
        w.wl = w.wl[rank * 8192:(rank + 1) * 8192]
    elif name == "c1":
        w = wk.c1_fabry_perot(2000)
    elif name == "c3":
        w = wk.c3_polariton(n_d=512 * world, n_wl=4096, n_theta=64)
        w.sweep_d = w.sweep_d[rank * 512:(rank + 1) * 512]
    elif name == "c4":
        w = wk.c4_chirped(10000, 16384 * world)
        w.wl = w.wl[rank * 16384:(rank + 1) * 16384]
    elif name == "c5":
        w = wk.c5_box_sweep(n_wl=8192, n_theta=256, n_d=32 * world)
        w.sweep_d = w.sweep_d[rank * 32:(rank + 1) * 32]
    else:
        raise SystemExit(f"unknown workload {name}")
    return w, DESCS[name]


def workload_pols(name):
    return ("s-wave",) if name == "c4" else ("s-wave", "p-wave")


# ------------------------------------------------------------------------------------------------ flop model
def algorithmic_flops(w, prep, n_pol, cse=True, power=True):
    """Algorithmic FP64 flops per spectral point of the path the kernel takes for this workload.
    Counted: the multiply/add/FMA(=2) of the phase, the layer-matrix entries and the 2-vector layer step, and the
    epilogue.  NOT counted: sincos / exp evaluations, rescaling, address arithmetic.
      real (lossless-material) layer:   build 4 + 2/pol,  step 12/pol
      complex layer:                    build 20 + 4/pol, step 28/pol   (28 + 28 per pol)
    With layer CSE a matrix is built once per distinct (material, thickness).  A lossless pair repeated k >= 3 times is
    one power of its period: the period is formed once per point from the polarisation-independent parts of the two
    layers (phase 2 + 2 each, period 18), powered by binary exponentiation (6 per squaring, 5 per set bit below the
    leading one), turned into the matrix (6 + 2/pol) and applied as one real-structured step (12/pol); a second mirror
    with the same pair count reuses the matrix.  K3 per thickness: the phase (2 or 4), cos/sin of the complex phase
    from sincos and cosh/sinh (4, absorbing layer only) and, per polarisation, the two 2-term sums M00, M10 (2 x 14
    complex, 2 x 6 real), |.|^2 of both (6) and R, T, A (5).  K2 carries both columns of a segment product; for an
    all-lossless stack they are four reals (12 flop per layer and polarisation)."""
    # evaluations of sin+cos (35 flop: magic-number rint 3, two-term Cody-Waite 4, r^2 1, two degree-6 polynomials 20,
    # reconstruction 7 -- pst_math.cuh sincos_core) and of the cosh/sinh pair of an absorbing layer (43 flop: exp_split 35
    # + cosh_sinh_scaled 8) are reported next to the conservative count, not inside it
    SINCOS_FLOP, EXP_FLOP = 35, 43
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
        n_sincos, n_exp = 1, (1 if lossy[m] else 0)
        kind = "K3 sweep (prefix/suffix amortised over the chunk, not counted)"
    else:
        n_types = len({(int(mats[j]), float(ds[j])) for j in interior})
        typed = cse and 0 < n_types <= 8 and len(interior) >= 2 * n_types and w.interior_layers < 128 * 16
        cols = 1
        deep = w.interior_layers >= 128 and len(w.wl) * len(w.theta) < 148 * 512
        if deep:
            # K2 segments carry both columns; for an all-lossless stack they are four reals = one column's worth
            cols, typed = (1 if not lossy[mats[interior]].any() else 2), False
        steps = cols * sum(step(mats[j]) for j in interior)
        builds = sum(build(mats[j]) for j in interior)
        n_sincos, n_exp = len(interior), int(sum(1 for j in interior if lossy[mats[j]]))
        kind = "K2 segmented" if deep else "K1 untyped"
        if typed:
            kind = "K1 typed (layer CSE)"
            seq = [(int(mats[j]), float(ds[j])) for j in reversed(interior)]

            def repeats(at):
                k = 1
                while at + 2 * k + 1 < len(seq) and seq[at + 2 * k] == seq[at] and seq[at + 2 * k + 1] == seq[at + 1]:
                    k += 1
                return k

            built, steps, builds, i, powers = set(), 0, 0, 0, {}
            n_sincos = n_exp = 0
            while i < len(seq):
                k = repeats(i) if i + 1 < len(seq) else 1
                if i + 1 < len(seq) and not (k == 1 and i + 2 < len(seq) and repeats(i + 1) >= 2):
                    a, b = seq[i], seq[i + 1]
                    if power and k >= 3 and not lossy[a[0]] and not lossy[b[0]]:
                        key = (frozenset((a, b)), k)
                        if key not in powers:  # phases + bare layers + period + binary powering + matrix
                            hb, ones = k.bit_length() - 1, bin(k).count("1") - 1
                            powers[key] = True
                            n_sincos += 2
                            builds += 2 * (2 + 2) + 18 + 6 * hb + 5 * ones + 6 + 2 * n_pol
                        steps += 12 * n_pol
                        kind = "K1 cavity/typed (layer CSE + closed-form power of repeated lossless pairs, polarisation-shared period)"
                    else:
                        for t in (a, b):
                            if t not in built:
                                built.add(t)
                                builds += build(t[0])
                                n_sincos += 1
                                n_exp += int(lossy[t[0]])
                        steps += k * (step(a[0]) + step(b[0]))
                    i += 2 * k
                else:
                    if seq[i] not in built:
                        built.add(seq[i])
                        builds += build(seq[i][0])
                        n_sincos += 1
                        n_exp += int(lossy[seq[i][0]])
                    steps += step(seq[i][0])
                    i += 1
        per_thread = builds + steps + epilogue
    trans = SINCOS_FLOP * n_sincos + EXP_FLOP * n_exp
    return {"flop_per_point": per_thread / n_pol, "kernel_path": kind,
            "sincos_per_thread": n_sincos, "exp_pairs_per_thread": n_exp,
            "flop_per_point_with_transcendentals": (per_thread + trans) / n_pol}


# ------------------------------------------------------------------------------------------------ CPU arm
_REF = {}


def _load_installed_reference():
    """The UNMODIFIED reference from baseline/_ref (pip install --target of /root/reference, made in the build container
    and shipped with the snapshot), imported through the ruamel.yaml stand-in of oracle/ref_loader.py."""
    path = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(path, "pistachio", "transfer_matrix.py")):
        return None
    try:
        from oracle import ref_loader

        ref_loader._install_ruamel_shim()
        if path not in sys.path:
            sys.path.insert(0, path)
        import importlib
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return importlib.import_module("pistachio.transfer_matrix")
    except Exception:
        return None


def _cpu_task(args):
    """One (theta, pol[, wavelength chunk]) task of the reference call sequence."""
    chunk, theta, pol = args
    kind, obj = _REF["structs"][chunk]
    if kind == "reference":
        s = obj
        s.calculate_layer_matrices(theta, pol)                       # tm.py:516
        M = s.calculate_all_transfer_matrices(theta, pol)            # tm.py:536
        T, R = s.calculate_all_t_r(M)                                # tm.py:490
    else:
        from oracle import tmm_oracle as orc

        T, R = orc.spectrum(obj, theta, pol)
    return float(np.sum(T) + np.sum(R))


def cpu_reference_run(name: str, steps: int, warmup: int):
    """Times the reference's CPU path on the host cores with the reference's own fan-out: a multiprocessing pool over
    (theta, pol) tasks (transfer_matrix.py:614-623).  The real reference (baseline/_ref) when it is there, else the
    oracle port (loop-faithful numpy restatement, pinned bit for bit to the reference in the build container).  Each
    step is a bounded sub-sample of the workload."""
    import multiprocessing as mp

    from oracle import tmm_oracle as orc

    w, desc = make_workload(name, 0, 1)
    cores = os.cpu_count() or 1
    n_l = min(len(w.wl), 256 if name != "c4" else max(4, cores))
    n_t = min(len(w.theta), max(8, cores // 2))
    sel_l = np.linspace(0, len(w.wl) - 1, n_l).astype(int)
    sel_t = np.linspace(0, len(w.theta) - 1, n_t).astype(int)
    wl = w.wl[sel_l]
    pols = workload_pols(name)
    # tasks = (theta, pol) pairs as in the reference; when there are fewer pairs than cores (single-angle workloads)
    # the wavelength axis is cut as well so that every core has work
    n_cut = max(1, min(n_l, -(-2 * cores // (n_t * len(pols)))))
    cuts = np.array_split(np.arange(n_l), n_cut)
    tm = _load_installed_reference()
    structs = []
    for c in cuts:
        if tm is not None:
            s = tm.Structure()
            for l in w.layers:
                layer = tm.Layer(material=l.material, thickness=l.thickness)
                layer.wavelengths = np.array(l.tab_wl, dtype=np.float64)
                layer.refractive_index = np.array(l.tab_n, dtype=np.float64)
                layer.extinction_coeff = np.array(l.tab_k, dtype=np.float64)
                layer.make_datapoints(wl[c])                          # tm.py:117 (scipy interp1d)
                s.add_layer(layer)
            s.wavelengths = wl[c]
            structs.append(("reference", s))
        else:
            n_list = []
            for l in w.layers:
                n, k = orc.resample_table(l.tab_wl, l.tab_n, l.tab_k, wl[c])
                n_list.append(n + 1j * k)
            structs.append(("port", orc.Stack(wl[c], n_list, [l.thickness for l in w.layers])))
    _REF["structs"] = structs  # inherited by the forked workers
    kind = "reference" if tm is not None else "port"
    tasks = [(ic, float(w.theta[i]), p) for i in sel_t for p in pols for ic in range(len(cuts))]
    points = n_l * n_t * len(pols)
    procs = min(cores, len(tasks))
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_task, [(0, 0.0, pols[0])] * procs, chunksize=1)  # spin-up
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_task, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    sample = (f"{n_l} wl x {n_t} theta x {len(pols)} pol = {points} points of {desc}; "
              f"multiprocessing pool of {procs} over {len(tasks)} (theta, pol[, wl-chunk]) tasks; "
              + ("the unmodified reference from baseline/_ref: calculate_layer_matrices -> calculate_all_transfer_matrices -> "
                 "calculate_all_t_r per task" if kind == "reference" else "oracle port (baseline/_ref not present)"))
    return {"value": points / sec, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample,
            "layer_steps_per_s": points * w.interior_layers / sec, "seconds_per_step": sec, "host_cpus": cores}, w, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, w, desc = cpu_reference_run(args.workload, max(1, min(args.steps, 5)), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "note": "the reference's own CPU path on the host cores, bounded sample per step "
                   "(at most 5 timed samples whatever --steps says: one sample is ~0.3 s of 16 cores)"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class Ctx:
    """What every measurement below shares."""

    def __init__(self, eng, rank, world, stream):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.eng, self.rank, self.world, self.stream = eng, rank, world, stream
        self.peaks, self.peaks_src = load_measured_peaks()
        self.fp64_peak = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.eng.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, fn, steps: int, flush: bool = True) -> float:
        """Seconds for `steps` calls of fn: one CUDA-event pair per step on the launch stream (L2 flushed before each),
        barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in ev:
            if flush:
                self.eng.flush_l2()
            a.record(self.stream)
            fn()
            b.record(self.stream)
        self.barrier()
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) / 1e3)


def traffic_from_profiles(workload: str, flags: int):
    """DRAM bytes per launch (and the FP64-pipe utilisation) of the dominant kernel from the newest committed
    `ncu --set full` capture of this workload (profiles/rNN_vMM_traffic.json, written by scripts/summarise_ncu.py)."""
    import glob
    import re

    cands = []
    for path in glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")):
        try:
            with open(path) as fh:
                t = json.load(fh)
        except Exception:
            continue
        if t.get("workload") == workload and flags == t.get("flags", 0):
            cands.append((tuple(int(x) for x in re.findall(r"\d+", os.path.basename(path))), os.path.basename(path), t))
    return max(cands, key=lambda c: c[:2])[2] if cands else None


def roofline_of(cx: Ctx, name, w, prep, pols, ms_per_step, points_rank, flags, cse=True, power=True):
    """The `roofline` object of one workload: whichever of the two candidate bounds (HBM result stores, FP64 pipe) the
    kernel sits closer to, with both reported."""
    flop_model = algorithmic_flops(w, prep, len(pols), cse=cse, power=power)
    sec = ms_per_step / 1e3
    achieved_tf = flop_model["flop_per_point"] * points_rank / sec / 1e12
    survey_tf = points_rank * w.interior_layers * FLOP_PER_LAYER_STEP / sec / 1e12
    hbm_gbs = 24.0 * points_rank / sec / 1e9
    fp64_frac = achieved_tf / cx.fp64_peak if cx.fp64_peak else 0.0
    hbm_frac = hbm_gbs / cx.peaks["hbm_gbs"] if cx.peaks.get("hbm_gbs") else 0.0
    if hbm_frac >= 0.5:  # the result stream owns the kernel (K3 sweeps); otherwise the FP64 pipe does
        roof = {"bound": "hbm", "achieved": hbm_gbs, "peak": cx.peaks.get("hbm_gbs"), "unit": "GB/s", "frac": hbm_frac,
                "traffic": None, "bytes_per_point": 24,
                "peak_source": f"MEASURED_PEAKS.json ({cx.peaks_src}): burst COPY bandwidth (read + write); a write-only result "
                               "stream can exceed it"}
    else:
        roof = {"bound": "fp64_pipe", "achieved": achieved_tf, "peak": cx.fp64_peak, "unit": "TFLOP/s", "frac": fp64_frac,
                "traffic": None,
                "peak_source": "independent-DFMA probe run in this process (pst_fp64_peak_probe); MEASURED_PEAKS.json has no FP64 entry"}
    with_tr_tf = flop_model["flop_per_point_with_transcendentals"] * points_rank / sec / 1e12
    roof["fp64"] = {"achieved_tflops": achieved_tf, "peak_tflops": cx.fp64_peak, "frac": fp64_frac, **flop_model,
                    "with_transcendentals": {"achieved_tflops": with_tr_tf, "frac": with_tr_tf / cx.fp64_peak if cx.fp64_peak else 0.0,
                                             "note": "the same count plus the polynomial evaluations of sin/cos (35 flop each) and of "
                                                     "an absorbing layer's cosh/sinh pair (43) that the formulas call for; "
                                                     "`roofline.frac` stays the conservative figure without them"},
                    "survey_56flop_per_layer_step_tflops": survey_tf,
                    "note": "algorithmic flops of the formulas the kernel evaluates (FMA = 2), transcendentals, rescaling and "
                            "address arithmetic not counted; see DESIGN.md section 4 and the ncu FP64-pipe utilisation in profiles/"}
    roof["hbm_store"] = {"achieved_gbs": hbm_gbs, "peak_gbs": cx.peaks.get("hbm_gbs"), "frac": hbm_frac}
    tr = traffic_from_profiles(name, flags)
    if tr:
        roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        roof["traffic_source"] = tr["source"]
        roof["algorithmic_bytes"] = 24 * points_rank
        if "fp64_pipe_pct" in tr:
            roof["fp64"]["pipe_util_ncu_pct"] = tr["fp64_pipe_pct"]
    return roof


def bench_workload(cx: Ctx, name: str, steps: int, warmup: int, flags: int = 0):
    """Device-timed throughput of one workload (each rank its own slab), results left on the device."""
    torch = cx.torch
    w, desc = make_workload(name, cx.rank, cx.world)
    pols = workload_pols(name)
    prep = cx.eng.prepare_workload(w)
    n_sw = 1 if w.sweep_d is None else len(w.sweep_d)
    shape = (n_sw, len(pols), len(w.wl), len(w.theta))
    points_rank = int(np.prod(shape))
    out = {k: torch.empty(shape, dtype=torch.float64, device=cx.eng.device) for k in ("R", "T", "A")}

    def step():
        cx.eng.eval_prepared(prep, pols=pols, out=out, flags=flags)

    for _ in range(max(warmup, 3)):
        step()
    launches0 = cx.eng.launch_count()
    sec_total = cx.time_steps(step, steps)
    launches = cx.eng.launch_count() - launches0
    return {"w": w, "desc": desc, "pols": pols, "prep": prep, "shape": shape, "points_rank": points_rank, "out": out,
            "step": step, "sec_total": sec_total, "ms_per_step": sec_total / steps * 1e3,
            "value": points_rank * cx.world * steps / sec_total, "launches": launches}


def pinned_copy_probe(cx: Ctx, nbytes: int = 256 << 20):
    """Pinned device->host and host->device copy rate of this process's GPU, all ranks copying at the same time (the
    host side is shared): the roofline of `e2e`."""
    torch = cx.torch
    n = nbytes // 8
    dev = torch.empty(n, dtype=torch.float64, device=cx.eng.device)
    host = torch.empty(n, dtype=torch.float64).pin_memory()
    rates = {}
    for kind in ("d2h", "h2d"):
        fn = (lambda: host.copy_(dev, non_blocking=True)) if kind == "d2h" else (lambda: dev.copy_(host, non_blocking=True))
        fn()
        best = None
        for _ in range(3):
            cx.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(cx.stream)
            fn()
            b.record(cx.stream)
            cx.barrier()
            ms = cx.max_over_ranks(a.elapsed_time(b))
            best = ms if best is None else min(best, ms)
        rates[kind + "_gbs_per_gpu"] = nbytes / 1e9 / (best / 1e3)
    rates["bytes"] = nbytes
    rates["how"] = "torch copy_ between a pinned host tensor and HBM, all ranks at once, best of 3, CUDA events, max over ranks"
    return rates


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
    stream = torch.cuda.current_stream()
    cx = Ctx(eng, rank, world, stream)
    cx.fp64_peak = eng.fp64_peak_tflops(8192)

    flags = (16 if args.no_layer_cse else 0) | (128 if args.no_periodic_power else 0)  # PST_FLAG_NO_LAYER_CSE / _NO_PERIODIC_POWER
    flags |= 2048 if args.no_shape_fastpath else 0  # PST_FLAG_NO_SHAPE_FASTPATH
    flags |= 512 if args.fp32 else 0  # PST_FLAG_FP32 (optional mode; the default and every reported headline number is fp64)

    # ---- headline: K timed steps, CUDA events on the launching stream, L2 flushed between steps (outputs are also
    # 403 MB per step, > the 126 MB L2), clocks sampled during the region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    main = bench_workload(cx, args.workload, args.steps, args.warmup, flags)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    w, desc, pols, prep, shape = main["w"], main["desc"], main["pols"], main["prep"], main["shape"]
    points_rank, out, step = main["points_rank"], main["out"], main["step"]
    ms_per_step, value, launches = main["ms_per_step"], main["value"], main["launches"]

    # ---- sustained segment: the same step back to back for >= 2 s (no L2 flush, no gaps), clocks sampled every few ms.
    # The headline region is a burst of tens of milliseconds; second-long sweeps run at the sustained clock.
    sustained = None
    if not args.no_sustained:
        n_sus = max(200, int(2.2 / max(ms_per_step / 1e3, 1e-6)))
        sus_sampler = ClockSampler(local, period_s=0.02)
        if rank == 0:
            sus_sampler.start()
        cx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n_sus):
            step()
        b.record(stream)
        cx.barrier()
        sec = cx.max_over_ranks(a.elapsed_time(b) / 1e3)
        sus_clocks = sus_sampler.stop() if rank == 0 else None
        sustained = {"value": points_rank * world * n_sus / sec, "unit": UNIT, "steps": n_sus, "seconds": sec,
                     "ms_per_step": sec / n_sus * 1e3, "clocks": sus_clocks,
                     "note": "back-to-back steps on one stream, no L2 flush between them (outputs 403 MB/step > L2 anyway)"}

    # ---- ablations of the algorithmic shortcuts (same grid, same outputs; results agree to rounding / bit for bit)
    ablation = None
    if args.workload == "c2" and flags == 0 and not args.no_ablation:
        ablation = {}
        for name, fl in (("layer_by_layer_pairs", 128), ("no_layer_cse", 128 | 16), ("pol_degenerate_opt_in", 1024),
                         ("op_interpreter", 2048)):
            fn = lambda fl=fl: eng.eval_prepared(prep, pols=pols, out=out, flags=fl)
            for _ in range(3):
                fn()
            sec = cx.time_steps(fn, 10)
            ablation[name] = {"value": points_rank * world * 10 / sec, "ms_per_step": sec / 10 * 1e3}
        step()  # leave the default path's results in `out`
        ablation["note"] = ("layer_by_layer_pairs: PST_FLAG_NO_PERIODIC_POWER (every layer applied one by one, cached "
                            "matrices, typed interpreter); no_layer_cse: additionally rebuild every layer matrix (sincos per layer: "
                            "what a stack without repetition costs); op_interpreter: PST_FLAG_NO_SHAPE_FASTPATH (the mirror/spacer/"
                            "mirror program through the typed kernel's op loop instead of the cavity kernel; bit-identical); "
                            "pol_degenerate_opt_in: PST_FLAG_POL_DEGENERATE -- NOT the default and not part of `value`: "
                            "the reference's model makes s and p the same function (similarity transform), so one chain "
                            "is computed and written to both planes")

    # ---- e2e: the plugin-style call with HOST buffers (pinned), H2D of the inputs and D2H of the results inside the timed
    # region.  The call returns what the reference returns -- (T, R) per point (tm.py:490-514); A = (1 - R) - T is one
    # subtraction on the host for whoever wants it and is NOT shipped over PCIe (16 B/point instead of 24).
    mat_n_h = prep["mat_n"].cpu().pin_memory()
    mat_k_h = prep["mat_k"].cpu().pin_memory()
    wl_h = torch.from_numpy(np.ascontiguousarray(w.wl)).pin_memory()
    th_h = torch.from_numpy(np.ascontiguousarray(w.theta)).pin_memory()
    sw_h = torch.from_numpy(np.ascontiguousarray(w.sweep_d)).pin_memory() if w.sweep_d is not None else None
    e2e_keys = ("T", "R")
    out_h = {k: torch.empty(shape, dtype=torch.float64).pin_memory() for k in e2e_keys}

    def e2e_step():
        eng.eval_grid_host(wl_h, th_h, mat_n_h, mat_k_h, prep["layer_mat"], prep["layer_d"], pols,
                           sweep_layer=w.sweep_layer, sweep_d=sw_h, want=e2e_keys, out=out_h, flags=flags)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    cx.barrier()
    t_e2e = cx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = points_rank * world * e2e_steps / t_e2e
    h2d = 8 * (wl_h.numel() + th_h.numel() + mat_n_h.numel() + mat_k_h.numel() + (sw_h.numel() if sw_h is not None else 0))
    d2h = 8 * len(e2e_keys) * points_rank
    same = all(bool((out_h[k] == out[k].cpu()).all()) for k in e2e_keys)
    probe = pinned_copy_probe(cx)
    d2h_rate = d2h * e2e_steps / t_e2e / 1e9
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "identical_to_device_path": same, "returns": "T, R (what tm.py:490-514 returns); A on request",
           "d2h_gbs_per_gpu": d2h_rate,
           "roofline": {"bound": "pcie_d2h", "achieved": d2h_rate, "peak": probe["d2h_gbs_per_gpu"], "unit": "GB/s",
                        "frac": d2h_rate / probe["d2h_gbs_per_gpu"] if probe["d2h_gbs_per_gpu"] else None,
                        "bytes_per_point": 8 * len(e2e_keys), "probe": probe},
           "bound": "PCIe device->host copy of T, R (16 B/point); kernels overlap the copies",
           "api": "pst_eval_grid_host via TMMEngine.eval_grid_host, pinned host buffers"}
    del out_h

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
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fused_step()
        cx.barrier()
        t_f = cx.max_over_ranks(time.perf_counter() - t0)
        fused = {"value": points_rank * world * e2e_steps / t_f, "unit": UNIT,
                 "d2h_bytes_per_step": sum(v.numel() * v.element_size() for v in pk_h.values()),
                 "what": "eval_grid (T only) + pst_find_peaks(height=1 %, distance=20) on the device; peak indices, counts and "
                         "mean peak spacing are the only results copied to the host",
                 "spectra": int(np.prod(shape[:2]) * shape[3]), "overflowed_spectra": int((pk_h["cnt"] < 0).sum())}

    # ---- optional assembly of the bench workload on every device: one NCCL all-gather per output, or the gather fused
    # into the compute kernel (not part of `value`)
    gather = None
    if world > 1:
        gather = measure_assembly_of_step(cx, pdist, prep, pols, out, shape, points_rank, ms_per_step, flags)

    # ---- every other BASELINE config, device-timed (same timing rules, fewer steps)
    roof = roofline_of(cx, args.workload, w, prep, pols, ms_per_step, points_rank, flags,
                       cse=not (args.no_layer_cse or args.fp32), power=not args.no_periodic_power) if rank == 0 else None
    del out, main, step
    torch.cuda.empty_cache()
    configs = None
    if not args.no_configs and flags == 0:
        configs = {}
        for name in ("c1", "c3", "c4", "c5"):
            if name == args.workload:
                continue
            try:
                r = bench_workload(cx, name, args.config_steps, 3)
                entry = {"workload": r["desc"], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                         "steps": args.config_steps, "points_per_gpu": r["points_rank"],
                         "layer_steps_per_s": r["value"] * r["w"].interior_layers, "run": "driver-visible (default bench run)"}
                if rank == 0:
                    entry["roofline"] = roofline_of(cx, name, r["w"], r["prep"], r["pols"], r["ms_per_step"], r["points_rank"], 0)
                if name == "c2":
                    pass
                configs[name] = entry
                del r
            except Exception as e:  # a config that does not fit must not cost the headline line
                configs[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        # configs[4] at FULL size (2^30 points): strong scaling over the ranks, with and without assembly
        try:
            configs["c5_full"] = c5_full_size(cx, pdist)
        except Exception as e:
            configs["c5_full"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
        # configs[2] at FULL size (5.24e10 points, chunked): strong scaling, R/T/A to HBM and T + fused peak search
        try:
            configs["c3_full"] = c3_full_size(cx, pdist)
        except Exception as e:
            configs["c3_full"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    if rank != 0:
        return
    layer_steps = points_rank * w.interior_layers
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 chain / f64 phase (optional mode)" if args.fp32 else "f64",
        "data": "synthetic",
        "config": {"workload": desc, "layer_cse": not args.no_layer_cse, "periodic_power": not (args.no_periodic_power or args.no_layer_cse), "points_per_gpu": points_rank, "interior_layers": w.interior_layers,
                   "layer_steps_per_s": value * w.interior_layers, "l2": "flushed between steps (256 MiB memset); outputs 403 MB/step > L2",
                   "timing": "CUDA events per step on the launch stream, max over ranks", "wall_s_incl_flush_and_warmup": t_wall},
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "sustained": sustained,
        "configs": configs,
        "ablation": ablation,
        "e2e_fused_reduction": fused,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if gather:
        line["allgather"] = gather
    print(json.dumps(line), flush=True)


def measure_assembly_of_step(cx: Ctx, pdist, prep, pols, out, shape, points_rank, ms_per_step, flags):
    """The bench workload's results assembled across the ranks (outside `value`): NCCL all-gather of T, R after the
    kernel, the same gather fused into the compute kernel (peer stores into symmetric memory), and the rooted variant
    (every rank stores into rank 0's copy only -- north_star: 'when results must be assembled on one device')."""
    torch, dist, eng, world, rank, stream = cx.torch, cx.dist, cx.eng, cx.world, cx.rank, cx.stream
    n_sw = shape[0]
    axis = "sweep" if n_sw > 1 else "wl"
    sizes = [shape[0] if axis == "sweep" else shape[2]] * world
    keys = ("T", "R")  # A = (1 - R) - T is recomputed where it is wanted instead of crossing NVLink
    warm = [pdist.allgather_result(out[k], axis, sizes) for k in keys]  # connection set-up and first allocation
    del warm
    cx.barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    full = [pdist.allgather_result(out[k], axis, sizes) for k in keys]
    g1.record(stream)
    cx.barrier()
    gms = cx.max_over_ranks(g0.elapsed_time(g1))
    recv = len(keys) * 8 * points_rank * (world - 1)
    res = {"arrays": "T, R (16 B/point on the wire; A recomputed on the receiver)", "ms": gms, "recv_GB_per_rank": recv / 1e9,
           "GBps_per_rank": recv / 1e9 / (gms / 1e3), "nccl_compute_plus_allgather_ms": ms_per_step + gms,
           "link_peak_GBps": 900.0,
           "byte_arithmetic": f"{world - 1} slabs x {points_rank} points x 16 B = {recv / 1e9:.3f} GB into every GPU; at the 900 GB/s "
                              f"NVLink 5 receive rate that is {recv / 900e9 * 1e3:.2f} ms, against {ms_per_step:.3f} ms to compute a slab: "
                              "assembling everything everywhere is bound by the receive side of one link, whatever issues the stores"}
    del full
    for mode, root in (("unicast", None), ("unicast", 0)):
        tag = "fused_allgather" if root is None else "fused_gather_to_rank0"
        try:
            asm = pdist.AssembledResult(world, shape, rank, world, eng.device, mode=mode, root=root, keys=keys)

            def fused_step():
                asm.eval_into(eng, prep, pols, flags=flags)
                asm.finish()

            fused_step()
            best = None
            for _ in range(5):
                cx.barrier()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record(stream)
                fused_step()
                f1.record(stream)
                cx.barrier()
                ms = cx.max_over_ranks(f0.elapsed_time(f1))
                best = ms if best is None else min(best, ms)
            if root is None or rank == root:
                ok = all(bool(torch.equal(asm.full[k][rank], out[k])) for k in keys)
            else:
                ok = True
            okt = torch.tensor([int(ok)], device=eng.device)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            res[tag] = {"compute_plus_assembly_ms": best, "recv_GBps_on_receiver": recv / 1e9 / (best / 1e3),
                        "assembled_points_per_s": points_rank * world / (best / 1e3), "identical_to_local_result": bool(okt.item())}
            del asm
            torch.cuda.empty_cache()
        except Exception as e:
            res[tag] = {"unavailable": repr(e)[:200]}
    res["fused_how"] = ("the chain kernel stores T and R straight into the receiving GPUs' symmetric-memory copies (pst_peer_outputs: "
                        "all GPUs, or rank 0 only), then a device barrier; no separate gather pass")
    return res


def c5_full_size(cx: Ctx, pdist):
    """configs[4] at full size: 256 thicknesses x 8192 wl x 256 theta x s/p = 2^30 points, thickness axis sharded over the
    ranks (STRONG scaling).  Compute alone, compute + NCCL all-gather of T, R, and the fused peer-store assembly."""
    torch, dist, eng, world, rank = cx.torch, cx.dist, cx.eng, cx.world, cx.rank
    from pistachio_b200 import workloads as wk

    n_d = 256
    pols = ("s-wave", "p-wave")
    w = wk.c5_box_sweep(n_wl=8192, n_theta=256, n_d=n_d)
    per = n_d // world
    lo = rank * per
    prep = eng.prepare_workload(w)
    p = dict(prep)
    p["sweep_d"] = prep["sweep_d"][lo:lo + per].contiguous()
    pts = n_d * 2 * 8192 * 256
    keys = ("T", "R")
    plane = (2, 8192, 256)
    if world == 1:
        out = {k: torch.empty((n_d,) + plane, dtype=torch.float64, device=eng.device) for k in ("R", "T", "A")}
        fn = lambda: eng.eval_prepared(p, pols=pols, out=out)
        fn()
        sec = min(cx.time_steps(fn, 1) for _ in range(3))
        viol = int(((out["A"] != (1.0 - out["R"]) - out["T"]) | (out["R"] < 0) | (out["R"] > 1 + 1e-12)).sum().item())
        return {"workload": "configs[4] box sweep at full size: 256 d x 8192 wl x 256 theta x s/p = 2^30 points", "scaling": "strong",
                "n_gpus": 1, "compute_ms": sec * 1e3, "points_per_s_compute": pts / sec, "property_violations": viol,
                "hbm_store_GBps": 24.0 * pts / sec / 1e9}
    asm = pdist.AssembledResult(n_d, plane, rank, world, eng.device, mode="nccl", keys=keys)
    local_A = torch.empty((per,) + plane, dtype=torch.float64, device=eng.device)
    mine = {k: asm.local_slab(k) for k in keys}
    mine["A"] = local_A

    def compute():
        eng.eval_prepared(p, pols=pols, out=mine)

    def both():
        compute()
        asm.finish()

    both()
    ms_c = min(cx.time_steps(compute, 1) for _ in range(3)) * 1e3
    ms_both = min(cx.time_steps(both, 1, flush=False) for _ in range(3)) * 1e3
    ref = {k: asm.full[k].clone() for k in keys}
    recv = 8.0 * len(keys) * pts * (world - 1) / world
    res = {"workload": "configs[4] box sweep at full size: 256 d x 8192 wl x 256 theta x s/p = 2^30 points", "scaling": "strong",
           "n_gpus": world, "compute_ms": ms_c, "points_per_s_compute": pts / (ms_c / 1e3),
           "arrays_on_the_wire": "T, R (A recomputed by whoever needs it)",
           "nccl": {"compute_plus_allgather_ms": ms_both, "points_per_s_assembled": pts / (ms_both / 1e3),
                    "recv_GB_per_rank": recv / 1e9, "recv_GBps_per_rank": recv / 1e9 / (max(ms_both - ms_c, 1e-6) / 1e3)}}
    del asm, mine
    torch.cuda.empty_cache()
    for tag, root in (("fused_allgather", None), ("fused_gather_to_rank0", 0)):
        try:
            asm = pdist.AssembledResult(n_d, plane, rank, world, eng.device, mode="unicast", root=root, keys=keys)

            def fused():
                asm.eval_into(eng, p, pols)
                asm.finish()

            fused()
            ms = min(cx.time_steps(fused, 1, flush=False) for _ in range(3)) * 1e3
            ok = all(bool(torch.equal(asm.full[k], ref[k])) for k in keys) if (root is None or rank == root) else True
            okt = torch.tensor([int(ok)], device=eng.device)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            res[tag] = {"ms": ms, "points_per_s_assembled": pts / (ms / 1e3), "recv_GBps_on_receiver": recv / 1e9 / (ms / 1e3),
                        "bit_identical_to_nccl_assembly": bool(okt.item())}
            del asm
            torch.cuda.empty_cache()
        except Exception as e:
            res[tag] = {"unavailable": repr(e)[:200]}
    return res


def c3_full_size(cx: Ctx, pdist, chunk: int = 2048):
    """BASELINE configs[2] at FULL size: 1e5 cavity thicknesses x 4096 wavelengths x 64 angles x s/p = 5.24e10 spectral
    points, strong scaling (the thickness axis is cut over the ranks, no communication).  The result would be 1.26 TB of
    R, T, A, so the sweep runs in chunks of `chunk` thicknesses whose buffers are reused (2048 thicknesses = 8.6 GB per
    array: 262 144 spectra per peak-search launch keep that kernel at the HBM rate; 512 left it at half): once writing R, T, A to HBM
    (with the path's properties checked on every value in the warm-up pass), once in the form the reference's
    notebooks consume it -- T only, fused with the device peak search (pst_find_peaks, "Cavity Length Determiner")."""
    torch, dist, eng, rank, world = cx.torch, cx.dist, cx.eng, cx.rank, cx.world
    from pistachio_b200 import workloads as wk

    n_d, pols = 100000, ("s-wave", "p-wave")
    w = wk.c3_polariton(n_d=n_d, n_wl=4096, n_theta=64)
    lo, hi = pdist.shard_bounds(n_d, world, rank)
    prep = eng.prepare_workload(w)
    d_all = prep["sweep_d"]
    out = {k: torch.empty((chunk, 2, 4096, 64), dtype=torch.float64, device=eng.device) for k in ("R", "T", "A")}
    nu = 1.0e4 / prep["wl"]
    bad = torch.zeros((), dtype=torch.int64, device=eng.device)
    peaks = torch.zeros((), dtype=torch.int64, device=eng.device)

    def sub(i0, i1):
        p = dict(prep)
        p["sweep_d"] = d_all[i0:i1]
        return p, {k: v[:i1 - i0] for k, v in out.items()}

    def pass_rta(check=False):
        for i0 in range(lo, hi, chunk):
            p, o = sub(i0, min(i0 + chunk, hi))
            eng.eval_prepared(p, pols=pols, out=o)
            if check:
                bad.add_(((o["A"] != (1.0 - o["R"]) - o["T"]) | (o["R"] < 0) | (o["R"] > 1) | (o["T"] < 0)).sum())

    def pass_fused():
        for i0 in range(lo, hi, chunk):
            p, o = sub(i0, min(i0 + chunk, hi))
            res = eng.eval_prepared(p, pols=pols, want=("T",), out={"T": o["T"]})
            _, _, cnt, _ = eng.find_peaks(res.T, height=0.01, distance=20, max_peaks=32, axis=nu, reverse=True)
            peaks.add_(cnt.clamp_min(0).sum())

    pass_rta(check=True)  # warm-up + properties
    ms_rta = cx.time_steps(pass_rta, 1, flush=False) * 1e3
    pass_fused()
    peaks.zero_()
    ms_fused = cx.time_steps(pass_fused, 1, flush=False) * 1e3
    if world > 1:
        dist.all_reduce(bad)
        dist.all_reduce(peaks)
    pts = n_d * 2 * 4096 * 64
    return {"workload": f"configs[2] polariton cavity-length sweep at full size: {n_d} d x 4096 wl x 64 theta x s/p = {pts:.3e} points",
            "scaling": "strong", "n_gpus": world, "chunk_thicknesses": chunk,
            "R_T_A_to_hbm": {"ms": ms_rta, "points_per_s": pts / (ms_rta / 1e3),
                             "hbm_store_GBps_per_gpu": 24.0 * pts / world / (ms_rta / 1e3) / 1e9},
            "T_plus_fused_peak_search": {"ms": ms_fused, "points_per_s": pts / (ms_fused / 1e3), "peaks_found": int(peaks.item())},
            "property_violations": int(bad.item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--config-steps", type=int, default=10, help="timed steps of each entry of `configs`")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` object (the other BASELINE configs)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained segment")
    ap.add_argument("--no-ablation", action="store_true")
    ap.add_argument("--no-layer-cse", action="store_true", help="ablation: rebuild every layer matrix (PST_FLAG_NO_LAYER_CSE)")
    ap.add_argument("--no-shape-fastpath", action="store_true",
                    help="ablation: mirror/spacer/mirror programs through the typed kernel's op interpreter (PST_FLAG_NO_SHAPE_FASTPATH)")
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
