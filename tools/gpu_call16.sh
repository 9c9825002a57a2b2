#!/bin/bash
# sparse front end at C3: capped L2 prefetch, BIG normalize variant, partial E table, predicated value loads in select_fill
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_configs.py -m gpu -q -x -k "normalize or moments or select or hvg or scale or c1" > gpurun_out/r2_t_sparse2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_t_sparse2.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_sp2.json 2> gpurun_out/r2_bench_c3_sp2.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_sp2.err | tail -1
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c2_sp2.json 2> gpurun_out/r2_bench_c2_sp2.err; echo "bench c2 rc=$?"; grep "stage ms" gpurun_out/r2_bench_c2_sp2.err | tail -1
