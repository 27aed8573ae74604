mkdir -p gpurun_out
echo "--- minb 3"; python tools/time_lar.py 2>&1 | tee gpurun_out/r5_lar3.txt
echo "--- minb 4"; NB200_LIB=$PWD/nest_b200/csrc/variants/lib_lar4.so python tools/time_lar.py 2>&1 | tee gpurun_out/r5_lar4.txt
