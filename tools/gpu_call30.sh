#!/bin/bash
# insert-warp diagnostics (smem candidate sets): loop trips, flagged rows, mask-building cycles; on the clustered micro data and on C3
mkdir -p gpurun_out
CSC_KNN_CNT=1 timeout 200 python tools/knn_bench.py 1000000 50 20 1 > gpurun_out/r2_knn_diag5.log 2>&1; echo "rc=$?"
CSC_KNN_CNT=1 timeout 400 python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_k5.json 2> gpurun_out/r2_bench_c3_k5.err; echo "bench rc=$?"
grep "knn f16" gpurun_out/r2_bench_c3_k5.err | tail -1 >> gpurun_out/r2_knn_diag5.log
cat gpurun_out/r2_knn_diag5.log
