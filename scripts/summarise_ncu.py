"""Turn a gpurun_out ncu capture into the tracked evidence files under profiles/.

    python scripts/summarise_ncu.py TAG [--rep gpurun_out/prof.ncu-rep] [--launches gpurun_out/launches.csv]
                                        [--workload c2]

writes  profiles/TAG_chain_kernel_ncu_raw.txt   selected raw metrics of every captured kernel instance
        profiles/TAG_launches.csv               the launch list (copied)
        profiles/TAG_launch_share.txt           per-kernel count / mean / total / share of the launch list
        profiles/TAG_traffic.json               DRAM bytes of the first captured instance (bench.py reads the newest)
        profiles/TAG_opcodes.txt                executed warp instructions per opcode (source page of the capture)
Needs `ncu` (reads reports without a GPU).
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
)


def ncu_csv(rep: str, page: str) -> list[list[str]]:
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep", default=os.path.join(ROOT, "gpurun_out", "prof.ncu-rep"))
    ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "launches.csv"))
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--name", default="chain_kernel")
    a = ap.parse_args()
    prof = os.path.join(ROOT, "profiles")
    os.makedirs(prof, exist_ok=True)

    rows = ncu_csv(a.rep, "raw")
    hdr, units, inst = rows[0], rows[1], rows[2:]
    iname = hdr.index("Kernel Name")
    raw_path = os.path.join(prof, f"{a.tag}_{a.name}_ncu_raw.txt")
    with open(raw_path, "w") as fh:
        fh.write("# kernels: " + " | ".join(r[iname] for r in inst) + "\n")
        for k, (h, u) in enumerate(zip(hdr, units)):
            if any(h.startswith(key) for key in KEYS):
                fh.write(f"{h} | {u} | {[r[k] for r in inst]}\n")
    print("wrote", raw_path)

    def val(name: str, r: list[str]) -> float:
        k = hdr.index(name)
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(units[k], 1.0)
        return float(r[k]) * scale

    tr = {"workload": a.workload, "kernel": inst[0][iname], "source": f"ncu --set full, profiles/{os.path.basename(raw_path)}",
          "dram_bytes_read": int(val("dram__bytes_read.sum", inst[0])), "dram_bytes_write": int(val("dram__bytes_write.sum", inst[0]))}
    for key, name in (("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                      ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active")):
        if name in hdr:
            tr[key] = float(inst[0][hdr.index(name)])  # bench.py copies fp64_pipe_pct into roofline.fp64.pipe_util_ncu_pct
    with open(os.path.join(prof, f"{a.tag}_traffic.json"), "w") as fh:
        json.dump(tr, fh, indent=1)
        fh.write("\n")

    # opcode histogram from the source page of the first kernel
    src = ncu_csv(a.rep, "source")
    h2 = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    sh = src[h2]
    i_src, i_inst, i_smp = sh.index("Source"), sh.index("Instructions Executed"), sh.index("# Samples")
    ops: collections.Counter = collections.Counter()
    smp: collections.Counter = collections.Counter()
    for r in src[h2 + 1:]:
        if len(r) <= i_inst or not r[i_inst].isdigit():
            break  # next kernel instance
        tok = r[i_src].split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        ops[op] += int(r[i_inst])
        smp[op] += int(r[i_smp])
    tot, ts = sum(ops.values()), max(sum(smp.values()), 1)
    grid = val("launch__grid_size", inst[0]) * val("launch__block_size", inst[0]) / 32.0
    with open(os.path.join(prof, f"{a.tag}_opcodes.txt"), "w") as fh:
        fh.write(f"# {inst[0][iname]}\n# executed warp instructions {tot} = {tot / grid:.1f} per warp ({grid:.0f} warps)\n")
        fh.write("# opcode | per warp | share of instructions | share of stall samples\n")
        for op, c in ops.most_common(40):
            fh.write(f"{op:10s} {c / grid:9.1f} {100 * c / tot:6.1f}% {100 * smp[op] / ts:6.1f}%\n")
    print("opcodes:", tot / grid, "instructions per warp")

    if os.path.exists(a.launches):
        shutil.copy(a.launches, os.path.join(prof, f"{a.tag}_launches.csv"))
        with open(a.launches) as fh:
            lines = [l for l in fh if not l.startswith("==")]
        lr = list(csv.reader(io.StringIO("".join(lines))))
        lh = lr[0]
        kn, mv, mu = lh.index("Kernel Name"), lh.index("Metric Value"), lh.index("Metric Unit")
        agg: dict[str, list[float]] = collections.defaultdict(list)
        for r in lr[1:]:
            if len(r) <= mv:
                continue
            t = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
            agg[r[kn]].append(t)
        total = sum(sum(v) for v in agg.values())
        with open(os.path.join(prof, f"{a.tag}_launch_share.txt"), "w") as fh:
            fh.write("# kernel | launches | mean us | total us | share of listed GPU time\n")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                fh.write(f"{k[:110]} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {100 * sum(v) / total:.1f}%\n")
        print("launch share written")


if __name__ == "__main__":
    sys.exit(main())
