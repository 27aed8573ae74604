#!/bin/bash
# time the three chain kernels for each launch-configuration variant built under csrc/variants/
for v in "$@"; do
  NB200_LIB=$PWD/nest_b200/csrc/variants/lib_$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/sweep_$v.csv python tools/profile_run.py --events 4e6 --launches 2 > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, sys
v = sys.argv[1]
t = {}
for r in csv.reader(open("gpurun_out/sweep_%s.csv" % v)):
    if len(r) > 14 and r[0].isdigit() and int(r[0]) >= 3:
        t[r[4].split("(")[0]] = float(r[14]) / 1e6
print("variant %s: " % v + "  ".join("%s %.3f ms" % kv for kv in t.items()) + "  total %.3f ms" % sum(t.values()))
PY
done
