#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "knn" -s 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/r2_bench_c3_d.json 2> gpurun_out/r2_bench_c3_d.err; echo "bench c3 rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_d.err
