#!/bin/bash
# 8 GPUs, closing code: the default bench line (C3 sharded over 8 ranks, NCCL inside the library), no e2e leg
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) "$@"; }
timeout 280 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --cpu-sample 0" > gpurun_out/r2_bench_c3_g8_d.json 2> gpurun_out/r2_bench_c3_g8_d.err; echo "bench g8 rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_g8_d.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/r2_bench_c3_g8_d.json')); print(d['ms_per_step'], d['value'], d['diag']['one_vs_n'])"
