#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of libpistachio_b200.so (cuobjdump -sass), for profiles/*_opcodes.txt.

    python scripts/sass_hist.py <substring of the mangled kernel name> [path/to/lib.so] [--dump out.sass]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_sass(lib, key):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            on = key in m.group(1)
            if on:
                out.append(line)
            continue
        if on:
            out.append(line)
    return out


def main():
    argv = sys.argv[1:]
    dump = None
    if "--dump" in argv:
        i = argv.index("--dump")
        dump = argv[i + 1]
        del argv[i:i + 2]
    args = argv
    key = args[0]
    lib = args[1] if len(args) > 1 else os.path.join(ROOT, "pistachio_b200", "libpistachio_b200.so")
    lines = kernel_sass(lib, key)
    if dump:
        with open(dump, "w") as fh:
            fh.write("\n".join(lines))
    hist = collections.Counter()
    for line in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            hist[m.group(1)] += 1
    total = sum(hist.values())
    fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"{key}: {total} instructions, {fp64} FP64 ({100.0 * fp64 / max(total, 1):.1f} %)")
    for k, v in hist.most_common(40):
        print(f"{v:6d} {k}")


if __name__ == "__main__":
    main()
