#!/bin/bash
# full default bench.py run with its wall time, and a digest of the line
tag=$1; shift
start=$(date +%s.%N)
python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
rc=$?
end=$(date +%s.%N)
echo "bench rc=$rc wall $(echo "$end - $start" | bc -l 2>/dev/null || python -c "print($end - $start)") s"
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
b=json.load(open("gpurun_out/${tag}_bench.json"))
for k in ("value","ms_per_step","gpu_launches","other_workloads_events_per_s","lar_hbm_gbs_per_gpu","lar_hbm_frac","config4_sweep","band_checksum_fixed_ids","band_chi2","cpu_baseline"):
    print(k, b.get(k))
print("e2e", {k:v for k,v in b["e2e"].items() if k in ("value","packed_f32","host_ingest_gbs","record_stream_gbs","record_stream_gbs_packed")})
print("roofline frac", b["roofline"]["frac"], b["roofline"]["stage_ms_per_step"], b["clocks"])
PY
