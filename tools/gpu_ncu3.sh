#!/bin/bash
# ncu --set full capture of one launch of each chain kernel (profile_run.py, config via $2), reports into gpurun_out/
# usage: tools/gpu_ncu3.sh <tag> [config] [kernels...]
tag=$1; cfg=${2:-NR}; shift; shift
ks=${@:-k_pre k_hits k_post}
for k in $ks; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 1 -c 1 -f -o gpurun_out/${tag}_$k \
     python tools/profile_run.py --events 4e6 --launches 2 --config $cfg > gpurun_out/${tag}_$k.log 2>&1
  tail -2 gpurun_out/${tag}_$k.log
done
ls -la gpurun_out/${tag}_*.ncu-rep
