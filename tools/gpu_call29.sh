#!/bin/bash
mkdir -p gpurun_out
CSC_KNN_L=48 timeout 200 python tools/knn_bench.py 1000000 50 20 2 > gpurun_out/r2_knn_diag4.log 2>&1; echo "rc=$?"
CSC_KNN_L=48 CSC_KNN_CNT=1 timeout 200 python tools/knn_bench.py 1000000 50 20 1 >> gpurun_out/r2_knn_diag4.log 2>&1; echo "rc=$?"
CSC_KNN_L=48 CSC_KNN_DBG=1 timeout 200 python tools/knn_bench.py 1000000 50 20 1 >> gpurun_out/r2_knn_diag4.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_knn_diag4.log
CSC_KNN_L=48 timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_k4.json 2> gpurun_out/r2_bench_c3_k4.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_k4.err | tail -1
