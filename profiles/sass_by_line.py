#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel in an ncu report.

    python profiles/sass_by_line.py <report.ncu-rep> <kernel regex> <cubin-name e.g. peaks> [top N]

ncu's CSV source page lists SASS instructions with counters but no line numbers; nvdisasm -g
lists the same SASS with //## File ... line markers.  Both come from the same libmfpa.so
(built with -lineinfo), so the two listings are joined by instruction address.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "musicfpaugment_b200", "libmfpa.so")


def sass_lines(cubin_name, kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_name + ".")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    addr2line, cur_fn, line, inlined = {}, None, None, None
    for l in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur_fn and re.search(kernel_re, cur_fn):
            addr2line[(cur_fn, int(m.group(1), 16))] = (line, m.group(2).strip())
    return addr2line


def main():
    rep, kre, cub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre],
                         capture_output=True, text=True).stdout
    # the CSV holds one block per profiled launch; keep the first
    blocks = re.split(r'(?m)^"Kernel Name",', out)
    body = blocks[1]
    name = body.splitlines()[0]
    rows = list(csv.DictReader(io.StringIO("\n".join(body.splitlines()[1:]))))
    a2l = sass_lines(cub, kre)
    fns = {k[0] for k in a2l}
    # pick the function whose instruction count matches
    best = max(fns, key=lambda f: sum(1 for k in a2l if k[0] == f) == len(rows))
    base = None
    per_line = defaultdict(lambda: [0, 0, 0])
    per_op = defaultdict(int)
    total = 0
    for r in rows:
        addr = int(r["Address"], 16) if r["Address"].startswith("0x") else int(r["Address"])
        if base is None:
            base = addr
        key = (best, addr - base)
        line = a2l.get(key, (None, ""))[0]
        n = int(float(r["Instructions Executed"] or 0))
        s = int(float(r["# Samples"] or 0))
        per_line[line][0] += n
        per_line[line][1] += s
        per_line[line][2] += 1
        per_op[r["Source"].split()[0] if not r["Source"].startswith("@") else r["Source"].split()[1]] += n
        total += n
    print(f"# {name[:120]}\n# function {best}\n# total warp instructions {total}")
    print("# line  warp-instr  share  samples  sass-count")
    for line, (n, s, k) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{str(line):28s} {n:12d} {100.0 * n / max(total, 1):6.2f}% {s:8d} {k:5d}")
    print("# opcode totals")
    for op, n in sorted(per_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{op:24s} {n:12d} {100.0 * n / max(total, 1):6.2f}%")


if __name__ == "__main__":
    main()
