#!/bin/bash
# cluster size of the kNN key ordering (keys per k-means cluster): 2048 (shipped), 1024, 512, 256 -- scan time and overhead at C3
mkdir -p gpurun_out
for kd in 2048 1024 512 256; do
  echo "== CSC_KNN_KDIV=$kd"
  CSC_KNN_KDIV=$kd timeout 300 python bench.py --steps 3 --warmup 2 --cpu-sample 0 --no-e2e > gpurun_out/r2_bench_c3_kd$kd.json 2> gpurun_out/r2_bench_c3_kd$kd.err; echo "bench rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_kd$kd.err | tail -1 | sed 's/.*stage ms.step//'
done
timeout 200 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "graph" > gpurun_out/r2_t_graph9.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t_graph9.log
