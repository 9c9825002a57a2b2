#!/bin/bash
# device-side find_variable_genes finish: the whole GPU test suite, then C3 / C1 stage times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t_gpu_hvgdev.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_t_gpu_hvgdev.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_h1.json 2> gpurun_out/r2_bench_c3_h1.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_h1.err | tail -1
timeout 300 python bench.py --workload C1 --steps 10 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c1_h1.json 2> gpurun_out/r2_bench_c1_h1.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c1_h1.err | tail -1
