#!/bin/bash
# round-2 closing 1-GPU call: whole GPU suite, default bench line (C3) with e2e / cpu_baseline, the other workloads,
# the reference arm, smoke, and the ncu launch list of the default bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t_gpu_final3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t_gpu_final3.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_final3.json 2> gpurun_out/r2_bench_c3_final3.err; echo "bench c3 rc=$?"; grep "stage ms\|e2e pass" gpurun_out/r2_bench_c3_final3.err | tail -2
timeout 400 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_final3.json 2> gpurun_out/r2_bench_c2_final3.err; echo "bench c2 rc=$?"
timeout 400 python bench.py --workload C1 --steps 10 --warmup 3 > gpurun_out/r2_bench_c1_final3.json 2> gpurun_out/r2_bench_c1_final3.err; echo "bench c1 rc=$?"
timeout 600 python bench.py --workload C4 --steps 5 --warmup 3 --cpu-sample 0 > gpurun_out/r2_bench_c4_final3.json 2> gpurun_out/r2_bench_c4_final3.err; echo "bench c4 rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final3.json 2> gpurun_out/r2_bench_ref_final3.err; echo "ref rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_c3_v3.csv python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample 0 > gpurun_out/r2_ncu_l3.log 2>&1; echo "ncu rc=$?"
