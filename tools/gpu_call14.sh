#!/bin/bash
# round-2 final 1-GPU call: the whole GPU suite, the default bench line (C3) with e2e / cpu_baseline, C2 for continuity,
# the reference arm, Leiden on the C3 graph
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t_gpu_final.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r2_bench_c3_final.err; echo "bench c3 rc=$?"; grep "stage ms\|e2e pass" gpurun_out/r2_bench_c3_final.err | tail -2
timeout 400 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_final.json 2> gpurun_out/r2_bench_c2_final.err; echo "bench c2 rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err; echo "ref rc=$?"
timeout 300 python tools/leiden_bench.py C3 1e-4 > gpurun_out/r2_leiden_c3.json 2> gpurun_out/r2_leiden_c3.err; echo "leiden rc=$?"; cat gpurun_out/r2_leiden_c3.json
python __graft_entry__.py smoke 2>&1 | tail -1
