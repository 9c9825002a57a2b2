#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) "$@"; }
for n in 8 4; do
  timeout 240 bash -c "$(declare -f run); run $n bench.py --gpus $n --steps 10 --warmup 3" > gpurun_out/r2_bench_c3_g${n}_b.json 2> gpurun_out/r2_bench_c3_g${n}_b.err; echo "bench c3 g$n rc=$?"
  grep "stage ms" gpurun_out/r2_bench_c3_g${n}_b.err | tail -1
done
timeout 200 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 10 --warmup 3 --workload C2 --no-e2e" > gpurun_out/r2_bench_c2_g8_b.json 2> gpurun_out/r2_bench_c2_g8_b.err; echo "bench c2 g8 rc=$?"
grep "stage ms" gpurun_out/r2_bench_c2_g8_b.err | tail -1
