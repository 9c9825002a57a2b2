#!/bin/bash
# round-2 scaling call (8 GPUs): C3 at N = 8, 4, 2 through torchrun (NCCL inside the library), 1-vs-N check in diag,
# C5 kNN sweep, C2 at N = 8
mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
for n in 8 4 2; do
  timeout 600 bash -c "$(declare -f run); run $n bench.py --gpus $n --steps 10 --warmup 3" > gpurun_out/r2_bench_c3_g$n.json 2> gpurun_out/r2_bench_c3_g$n.err; echo "bench c3 g$n rc=$?"
  grep "stage ms" gpurun_out/r2_bench_c3_g$n.err | tail -1
done
for n in 8 4 2; do
  timeout 600 bash -c "$(declare -f run); run $n tools/knn_sweep.py 4000000 30 2" > gpurun_out/r2_knn_c5_g$n.json 2> gpurun_out/r2_knn_c5_g$n.err; echo "c5 g$n rc=$?"; cat gpurun_out/r2_knn_c5_g$n.json
done
timeout 600 bash -c "$(declare -f run); run 8 bench.py --gpus 8 --steps 10 --warmup 3 --workload C2 --no-e2e" > gpurun_out/r2_bench_c2_g8.json 2> gpurun_out/r2_bench_c2_g8.err; echo "bench c2 g8 rc=$?"
grep "stage ms" gpurun_out/r2_bench_c2_g8.err | tail -1
timeout 600 bash -c "$(declare -f run); run 8 tools/multigpu_check.py 3000 80000" > gpurun_out/r2_mg8.log 2>&1; echo "mgcheck rc=$?"; tail -2 gpurun_out/r2_mg8.log
