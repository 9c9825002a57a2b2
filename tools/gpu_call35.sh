#!/bin/bash
# 2 GPUs after the label-pass / graph-key changes: world-size-2 NCCL parity tool + the default bench line without e2e
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) "$@"; }
timeout 150 bash -c "$(declare -f run); run 2 tools/multigpu_check.py 33000 20000" > gpurun_out/r2_mg2d.log 2>&1; echo "mgcheck rc=$?"; tail -2 gpurun_out/r2_mg2d.log
timeout 300 bash -c "$(declare -f run); run 2 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --cpu-sample 0" > gpurun_out/r2_bench_c3_g2_d.json 2> gpurun_out/r2_bench_c3_g2_d.err; echo "bench g2 rc=$?"; grep "stage ms" gpurun_out/r2_bench_c3_g2_d.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/r2_bench_c3_g2_d.json')); print(d['ms_per_step'], d['diag']['one_vs_n'])"
