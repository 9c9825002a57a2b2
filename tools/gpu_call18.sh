#!/bin/bash
# where the host part of find_variable_genes spends its time on the GPU box; ncu --set full of the sparse front end at C3 shape
mkdir -p gpurun_out
CSC_LOG_LEVEL=2 timeout 120 python tools/hvg_trace.py 33000 > gpurun_out/r2_hvg_trace.log 2>&1; echo "trace rc=$?"; tail -24 gpurun_out/r2_hvg_trace.log
timeout 300 python tools/sparse_bench.py 33000 1000000 0.05 2000 2 > gpurun_out/r2_sparse_c3.log 2>&1; echo "sparse rc=$?"; tail -3 gpurun_out/r2_sparse_c3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_normalize|k_select_fill|k_select_count|k_scale_stats_fx" -c 4 -o gpurun_out/r2_sparse_c3_full -f python tools/sparse_bench.py 33000 1000000 0.05 2000 1 > gpurun_out/r2_sparse_c3_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
