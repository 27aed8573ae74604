#!/bin/bash
# everything profiles/ is made from: the bench line of both arms, the ncu launch list of the bench command, and one
# ncu --set full capture per kernel (each after its own command ran without ncu).  Two parts, because one gpurun call
# brings back at most 64 MiB: `gpu_profiles.sh <tag> a` (bench lines, launch list, k_pre, k_hits) and
# `gpu_profiles.sh <tag> b` (k_hits_finish, k_post, k_lar)
tag=$1; part=${2:-a}
if [ "$part" = a ]; then
  python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err || exit 1
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
     python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
  bash tools/gpu_ncu3.sh ${tag} NR k_pre k_hits
else
  bash tools/gpu_ncu3.sh ${tag} NR k_hits_finish k_post
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_lar -s 1 -c 1 -f -o gpurun_out/${tag}_k_lar \
     python tools/profile_lar.py 1 > gpurun_out/${tag}_k_lar.log 2>&1
  tail -2 gpurun_out/${tag}_k_lar.log
fi
ls -la gpurun_out/${tag}_*
