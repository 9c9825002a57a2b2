#!/bin/bash
# round-2 call 1: probe the box, GPU tests, C3/C2 bench lines with the round-1 kernels + the round-2 correctness fixes
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi -L; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; lscpu | head -20; } > gpurun_out/r2_probe.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t_gpu1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_gpu1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_a.json 2> gpurun_out/r2_bench_c3_a.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_a.json 2> gpurun_out/r2_bench_c2_a.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err; echo "ref rc=$?"
tail -3 gpurun_out/r2_t_gpu1.log
