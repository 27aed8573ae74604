#!/bin/bash
# k_lar per species: timing (bench.py's LAr leg only would need the whole bench; this is the bare kernel) and an ncu capture
# usage: tools/gpu_lar.sh <tag> [species to capture ...]
tag=$1; shift
mkdir -p gpurun_out
python tools/time_lar.py > gpurun_out/${tag}_lar_times.txt 2>&1 || { tail -20 gpurun_out/${tag}_lar_times.txt; exit 1; }
cat gpurun_out/${tag}_lar_times.txt
for sp in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_lar -s 1 -c 1 -f -o gpurun_out/${tag}_k_lar${sp} \
     python tools/profile_lar.py ${sp} > gpurun_out/${tag}_k_lar${sp}.log 2>&1
  tail -2 gpurun_out/${tag}_k_lar${sp}.log
done
