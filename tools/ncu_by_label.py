#!/usr/bin/env python
"""Executed warp-instructions of one kernel by SASS subroutine label and by source line inside each label.

usage: ncu_by_label.py <file.ncu-rep> <kernel mangled name> [top]
Rebuilds the SASS listing of the CURRENT libnest_b200.so (cuobjdump + nvdisasm -g) -- the report must come from
the same build.  Unlike ncu_by_line.py this keeps out-of-line helpers (accurate_pow, div slow paths, philox_block,
binomial, box_muller) apart: they carry no line info of their own and would otherwise be booked on whatever source
line precedes them in the listing."""
import csv
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    sassd = os.path.join(ROOT, "gpurun_out", "sass")
    os.makedirs(sassd, exist_ok=True)
    for fn in os.listdir(sassd):
        if fn.endswith(".cubin"):
            os.remove(os.path.join(sassd, fn))
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("NB200_SASS_LIB", os.path.join(ROOT, "nest_b200", "csrc", "libnest_b200.so"))], cwd=sassd,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cub = [fn for fn in os.listdir(sassd) if fn.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", cub], cwd=sassd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                                          text=True).stdout.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(rows[hi])}
    data = rows[hi + 1:]
    lab, loc, info, inside = "main", ("?", 0), [], False
    for ln in sass.splitlines():
        if ln.startswith(".text." + kern + ":"):
            inside = True
            continue
        if not inside:
            continue
        if ln.startswith("//-----") and info:
            break
        m = re.match(r"^(\$[^:]+|_ZN[^:]+):", ln)
        if m:
            lab = m.group(1).split("$")[-1]
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*;", ln):
            info.append((lab, loc))
    if len(info) != len(data):
        sys.exit("SASS listing (%d) and report (%d) differ: not the same build" % (len(info), len(data)))
    tot = 0.0
    bylab, byline = defaultdict(lambda: [0.0, 0.0]), defaultdict(lambda: [0.0, 0.0])
    for (lab, loc), r in zip(info, data):
        ex = float(r[col["Instructions Executed"]] or 0)
        th = float(r[col["Thread Instructions Executed"]] or 0)
        tot += ex
        bylab[lab][0] += ex
        bylab[lab][1] += th
        byline[(lab, loc)][0] += ex
        byline[(lab, loc)][1] += th
    print("kernel %s: %.4g warp instructions executed" % (kern, tot))
    print("\n== by SASS subroutine ==\n%-58s %8s %6s" % ("label", "warp-ins%", "lanes"))
    for lab, (ex, th) in sorted(bylab.items(), key=lambda kv: -kv[1][0])[:14]:
        print("%-58s %8.2f %6.1f" % (lab[-58:], 100 * ex / tot, th / max(ex, 1)))
    src = {}
    print("\n== by source line within its subroutine ==")
    for (lab, loc), (ex, th) in sorted(byline.items(), key=lambda kv: -kv[1][0])[:top]:
        path = os.path.join(ROOT, "nest_b200", "csrc", loc[0])
        if path not in src:
            src[path] = open(path).read().split("\n") if os.path.exists(path) else []
        text = src[path][loc[1] - 1].strip()[:70] if 0 < loc[1] <= len(src[path]) else ""
        print("%-22s %-18s %6.2f%% lanes %4.1f | %s" % (lab[-22:], "%s:%d" % loc, 100 * ex / tot, th / max(ex, 1), text))


if __name__ == "__main__":
    main()
