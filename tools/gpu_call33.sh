#!/bin/bash
# tensor-core k-means label pass of the kNN key ordering: parity tests, A/B against the scalar kernel, C3 stage times
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "knn" > gpurun_out/r2_t_knn8.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t_knn8.log
CSC_KNN_KM_MMA=0 timeout 200 python tools/knn_bench.py 1000000 50 20 2 > gpurun_out/r2_knn_diag8.log 2>&1; echo "rc=$?"
timeout 200 python tools/knn_bench.py 1000000 50 20 2 >> gpurun_out/r2_knn_diag8.log 2>&1; echo "rc=$?"
cat gpurun_out/r2_knn_diag8.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_k8.json 2> gpurun_out/r2_bench_c3_k8.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_k8.err | tail -1
