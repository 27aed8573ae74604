#!/usr/bin/env python
"""Attribute an ncu SASS-level source page to CUDA source lines and functions.

usage: ncu_by_line.py <ncu --page source --csv> <nvdisasm -g output> <kernel mangled name> [top]

nvdisasm -g prints `//## File "...", line N` markers (from -lineinfo) ahead of the SASS they
cover; the ncu CSV lists the same instructions in order with executed counts and stall samples.
"""
import csv
import os
import re
import sys
from collections import defaultdict


def sass_lines(path, kernel):
    loc = ("?", 0)
    out = []
    inside = False
    for ln in open(path):
        if ln.startswith(".text." + kernel + ":"):
            inside = True
            continue
        if not inside:
            continue
        if ln.startswith("//-----") or ln.startswith(".text.") or ln.lstrip().startswith(".section"):
            if out:
                break
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), loc, m.group(2).strip()))
    return out


def functions_of(path):
    """line -> enclosing function name (crude: NB_D/NB_HD/__global__/template definitions)"""
    names = []
    pat = re.compile(r"^(?:NB_H?D|__global__|__device__|inline|template).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    try:
        src = open(path).read().split("\n")
    except OSError:
        return lambda n: "?"
    for i, l in enumerate(src, 1):
        m = pat.match(l)
        if m and not l.rstrip().endswith(";"):
            names.append((i, m.group(1)))

    def f(n):
        cur = "?"
        for i, nm in names:
            if i <= n:
                cur = nm
            else:
                break
        return cur
    return f


def main():
    csvp, sassp, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(csvp)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    data = rows[hi + 1:]
    sass = sass_lines(sassp, kernel)
    if len(sass) != len(data):
        print("warning: %d SASS lines vs %d ncu rows" % (len(sass), len(data)))
    n = min(len(sass), len(data))
    by_line = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for k in range(n):
        r = data[k]
        ex = int(float(r[col["Instructions Executed"]] or 0))
        th = int(float(r[col["Thread Instructions Executed"]] or 0))
        sm = int(float(r[col["# Samples"]] or 0))
        key = sass[k][1]
        for j, v in enumerate((ex, th, sm)):
            by_line[key][j] += v
            tot[j] += v
    srcdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nest_b200", "csrc")
    fn_cache = {}
    by_fn = defaultdict(lambda: [0, 0, 0])
    for (f, l), v in by_line.items():
        if f not in fn_cache:
            fn_cache[f] = functions_of(os.path.join(srcdir, f))
        name = "%s:%s" % (f, fn_cache[f](l))
        for j in range(3):
            by_fn[name][j] += v[j]
    print("kernel %s: %d SASS instructions, %d warp-instr executed, %d thread-instr, %d stall samples" %
          (kernel, n, tot[0], tot[1], tot[2]))
    print("\n== by function (innermost inlined frame) ==")
    print("%-44s %8s %8s %8s %6s" % ("function", "warp-ins%", "thr-ins%", "samples%", "thr/w"))
    for name, v in sorted(by_fn.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%-44s %8.2f %8.2f %8.2f %6.1f" % (name, 100. * v[0] / tot[0], 100. * v[1] / tot[1],
                                              100. * v[2] / max(tot[2], 1), v[1] / max(v[0], 1)))
    print("\n== by source line ==")
    for (f, l), v in sorted(by_line.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%-28s %8.2f %8.2f %8.2f %6.1f" % ("%s:%d" % (f, l), 100. * v[0] / tot[0], 100. * v[1] / tot[1],
                                              100. * v[2] / max(tot[2], 1), v[1] / max(v[0], 1)))


if __name__ == "__main__":
    main()
