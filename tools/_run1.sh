mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_variants.py tests/test_gpu_chain.py tests/test_gpu_sweep.py tests/test_gpu_zz_stage_api.py -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
python - <<PY
import json
b=json.load(open("gpurun_out/r6_bench.json"))
print("value %.4g e2e %.4g ms/step %.2f stages %s" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["stage_ms_per_step"]))
print(b.get("config4_sweep"))
PY
