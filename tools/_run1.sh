mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_lar.py -x -q 2>&1 | tail -4
timeout 200 python -m pytest tests/test_gpu_sweep.py -x -q 2>&1 | tail -4
