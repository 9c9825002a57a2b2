#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_configs.py -m gpu -q -x -k "hvg or c1 or object or loess" > gpurun_out/r2_t_hvg3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t_hvg3.log
CSC_LOG_LEVEL=2 timeout 120 python tools/hvg_trace.py 33000 > gpurun_out/r2_hvg_trace2.log 2>&1; echo "trace rc=$?"; tail -12 gpurun_out/r2_hvg_trace2.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_hvg3.json 2> gpurun_out/r2_bench_c3_hvg3.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_hvg3.err | tail -1
