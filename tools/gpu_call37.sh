#!/bin/bash
# eigensolver profile (CSC_PCA_PROF=1) on the C3 Gram (2000 x 2000): where the 3.9 ms go
mkdir -p gpurun_out
CSC_PCA_PROF=1 timeout 300 python bench.py --workload C1 --steps 3 --warmup 2 --cpu-sample 0 --no-e2e > gpurun_out/r2_pcaprof_c1.json 2> gpurun_out/r2_pcaprof_c1.err; echo "rc=$?"
grep "pca prof" gpurun_out/r2_pcaprof_c1.err | tail -2
grep "stage ms" gpurun_out/r2_pcaprof_c1.err | tail -1
