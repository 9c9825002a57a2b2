#!/bin/bash
# two-lane upload: forced-path test, then the e2e leg of the default bench with and without it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "upload" > gpurun_out/r2_t_upload.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_t_upload.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/r2_bench_c3_up2.json 2> gpurun_out/r2_bench_c3_up2.err; echo "bench rc=$?"; grep "e2e" gpurun_out/r2_bench_c3_up2.err | tail -3
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c3_up2.json"))
print("two-lane e2e", d["e2e"]["ms_per_step"], d["e2e"]["call_ms"], "step", d["ms_per_step"])
PY
CSC_UPLOAD_THREADS=8 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/r2_bench_c3_up2_t8.json 2> gpurun_out/r2_bench_c3_up2_t8.err; echo "bench t8 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c3_up2_t8.json"))
print("two-lane 8 threads e2e", d["e2e"]["ms_per_step"], d["e2e"]["call_ms"])
PY
