#!/bin/bash
# last call of the round: the whole GPU suite on the closing tree
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q > gpurun_out/r2_t_gpu_final4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t_gpu_final4.log
