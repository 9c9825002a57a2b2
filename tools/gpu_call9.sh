#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "split_transform or unproven" -s 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/r2_bench_c3_e.json 2> gpurun_out/r2_bench_c3_e.err; echo "bench c3 rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_e.err
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/r2_bench_c2_e.json 2> gpurun_out/r2_bench_c2_e.err; echo "bench c2 rc=$?"; grep "stage ms" gpurun_out/r2_bench_c2_e.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
