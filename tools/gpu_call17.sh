#!/bin/bash
# host lane with 4 sub-streams, new argsort / 16 fit threads in the HVG host part
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_configs.py -m gpu -q -x -k "upload or hvg or c1" > gpurun_out/r2_t_hvg2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t_hvg2.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/r2_bench_c3_up3.json 2> gpurun_out/r2_bench_c3_up3.err; echo "bench rc=$?"; grep "stage ms\|e2e pass" gpurun_out/r2_bench_c3_up3.err | tail -3
