#!/bin/bash
# candidate sets in shared memory: parity tests, scan diagnostics, C3 / C2 stage times, k = 30
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "knn" > gpurun_out/r2_t_knn3.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t_knn3.log
timeout 200 python tools/knn_bench.py 1000000 50 20 2 > gpurun_out/r2_knn_diag3.log 2>&1; echo "rc=$?"
CSC_KNN_CNT=1 timeout 200 python tools/knn_bench.py 1000000 50 20 1 >> gpurun_out/r2_knn_diag3.log 2>&1; echo "rc=$?"
timeout 200 python tools/knn_bench.py 1000000 50 30 2 >> gpurun_out/r2_knn_diag3.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_knn_diag3.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_k3.json 2> gpurun_out/r2_bench_c3_k3.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_k3.err | tail -1
