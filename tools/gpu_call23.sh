#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_configs.py -m gpu -q -x -k "normalize or moments or c1" > gpurun_out/r2_t_norm4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_t_norm4.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_n4.json 2> gpurun_out/r2_bench_c3_n4.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_n4.err | tail -1
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c2_n4.json 2> gpurun_out/r2_bench_c2_n4.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c2_n4.err | tail -1
