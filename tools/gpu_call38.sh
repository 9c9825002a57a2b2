#!/bin/bash
# new tests of the device-side HVG finish (tied means, spans 0.05 / 1.0, device vs host, zero-variance rejection)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "hvg" > gpurun_out/r2_t_hvgdev2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_t_hvgdev2.log
