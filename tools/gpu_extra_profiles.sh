#!/bin/bash
# ncu --set full captures of the two kernels gpu_profiles.sh does not cover: k_band_fit (Gaussian fits of a 250-point NR
# sweep's bands) and k_lar for NR (gpu_profiles.sh takes ER)
tag=$1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_band_fit -s 1 -c 1 -f -o gpurun_out/${tag}_k_band_fit \
   python tools/time_band_fit.py > gpurun_out/${tag}_k_band_fit.log 2>&1
tail -3 gpurun_out/${tag}_k_band_fit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_lar -s 1 -c 1 -f -o gpurun_out/${tag}_k_lar_NR \
   python tools/profile_lar.py 0 > gpurun_out/${tag}_k_lar_NR.log 2>&1
tail -2 gpurun_out/${tag}_k_lar_NR.log
