#!/bin/bash
# round-2 call 6 (1 GPU): tests, the default bench line, ncu launch list + full captures of the C3 top kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t_gpu6.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed" gpurun_out/r2_t_gpu6.log | head
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_c.json 2> gpurun_out/r2_bench_c3_c.err; echo "bench c3 rc=$?"
timeout 600 python tools/knn_sweep.py 4000000 30 2 > gpurun_out/r2_knn_c5_g1.json 2> gpurun_out/r2_knn_c5_g1.err; echo "c5 knn rc=$?"; cat gpurun_out/r2_knn_c5_g1.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample 0 > gpurun_out/r2_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_c3.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample 0 > gpurun_out/r2_ncu_l.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --cpu-sample 0 > gpurun_out/r2_plain2.log 2>&1 && \
timeout 2400 ncu --set full --clock-control none --import-source on \
    -k regex:"k_knn_scan_f16|k_normalize|k_scale_split|k_transform_sparse|k_gram_tc|k_scale_stats_fx|k_select_fill|k_select_count|k_knn_refine" \
    -s 9 -c 9 -o gpurun_out/r2_c3_top python bench.py --steps 1 --warmup 1 --no-e2e --cpu-sample 0 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
