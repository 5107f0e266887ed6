#!/usr/bin/env python
"""Executed warp instructions and stall samples of one profiled kernel, grouped by SOURCE LINE.

    python scripts/ncu_by_line.py gpurun_out/prof.ncu-rep <mangled kernel name> [--lib pistachio_b200/libpistachio_b200.so]
                                  [--min 3] [--warps N]

Joins the SASS page of an `ncu --set full --import-source on` capture (per-instruction executed counts) with the line
table of the same kernel in the library's cubin (nvdisasm --print-line-info; the library is built with -lineinfo).
The library must be the build that was profiled.  Inlined code is attributed to the innermost source line.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(lib, kernel):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
        cubin = next(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin"))
        txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    table, cur, on = {}, None, False
    for l in txt.split("\n"):
        if l.lstrip().startswith(".section"):
            on = (".text." + kernel + ",") in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    argv = sys.argv[1:]
    opts = {"--lib": os.path.join(ROOT, "pistachio_b200", "libpistachio_b200.so"), "--min": "3", "--warps": "0"}
    for k in list(opts):
        if k in argv:
            i = argv.index(k)
            opts[k] = argv[i + 1]
            del argv[i:i + 2]
    rep, kernel = argv[0], argv[1]
    table = line_table(opts["--lib"], kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h2 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    sh = rows[h2]
    ia, i_src, i_inst, i_smp = sh.index("Address"), sh.index("Source"), sh.index("Instructions Executed"), sh.index("# Samples")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    warps = float(opts["--warps"])
    if not warps:
        g, b = rr[0].index("launch__grid_size"), rr[0].index("launch__block_size")
        warps = float(rr[2][g].replace(",", "")) * float(rr[2][b].replace(",", "")) / 32.0
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0, collections.Counter()])
    base = None
    for r in rows[h2 + 1:]:
        if len(r) <= i_inst or not r[i_inst].isdigit():
            break
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        base = a if base is None else base
        key = table.get(a - base, ("?", 0))
        tok = r[i_src].split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        n = int(r[i_inst]) / warps
        v = agg[key]
        v[0] += n
        v[1] += n if op in ("DFMA", "DMUL", "DADD", "DSETP") else 0.0
        v[2] += int(r[i_smp])
        v[3][op] += n
    tot, tots = sum(v[0] for v in agg.values()), max(sum(v[2] for v in agg.values()), 1)
    print(f"# {kernel}: {tot:.1f} warp instructions per warp, {sum(v[1] for v in agg.values()):.1f} FP64; {warps:.0f} warps")
    src = {}
    for key in sorted(agg):
        v = agg[key]
        if v[0] < float(opts["--min"]):
            continue
        if key[0] not in src:
            p = os.path.join(ROOT, "pistachio_b200", "csrc", key[0])
            src[key[0]] = open(p).read().split("\n") if os.path.exists(p) else []
        text = src[key[0]][key[1] - 1].strip()[:70] if 0 < key[1] <= len(src[key[0]]) else ""
        other = ", ".join(f"{o}:{c:.0f}" for o, c in sorted(v[3].items(), key=lambda x: -x[1]) if o not in ("DFMA", "DMUL", "DADD"))
        print(f"{key[0]:16s}{key[1]:5d} inst {v[0]:6.1f} fp64 {v[1]:6.1f} stall {100 * v[2] / tots:5.1f}% | {text} | {other[:80]}")


if __name__ == "__main__":
    main()
