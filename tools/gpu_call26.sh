#!/bin/bash
# 8 GPUs: the default bench line (C3) with e2e and the 1-vs-N check
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) "$@"; }
timeout 300 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 10 --warmup 3" > gpurun_out/r2_bench_c3_g8_c.json 2> gpurun_out/r2_bench_c3_g8_c.err; echo "bench c3 g8 rc=$?"
grep "stage ms" gpurun_out/r2_bench_c3_g8_c.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/r2_bench_c3_g8_c.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['diag']['one_vs_n'])"
nproc
