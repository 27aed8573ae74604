#!/bin/bash
# one gpurun call: parity tests of the touched samplers, per-kernel times of library variants, a short bench
# usage: tools/gpu_session.sh <tag> [variant ...]
tag=$1; shift
mkdir -p gpurun_out
{
  timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_hitmath.py tests/test_gpu_variants.py -x -q 2>&1 | tail -15
  timeout 1200 python -m pytest ${TESTS:-tests/test_gpu_chain.py} -x -q 2>&1 | tail -25
} > gpurun_out/${tag}_tests.log 2>&1
timeout 600 bash tools/sweep_variants.sh "$@" > gpurun_out/${tag}_sweep.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -5 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_sweep.log; python - <<PY
import json
try:
    b=json.load(open("gpurun_out/${tag}_bench.json"))
    print("value %.4g e2e %.4g ms/step %.2f stages %s" % (b["value"], b["e2e"]["value"], b["ms_per_step"], b["roofline"]["stage_ms_per_step"]))
except Exception as e:
    print("bench parse failed", e)
PY
