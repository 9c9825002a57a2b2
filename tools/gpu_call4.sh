#!/bin/bash
# round-2 call 4 (2 GPUs): full GPU tests, NCCL-in-library equality check, C3 bench at N=2 and N=1, k=30 kNN timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t_gpu4.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_t_gpu4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py 3000 40000 > gpurun_out/r2_mg2.log 2>&1; echo "mgcheck rc=$?"; tail -3 gpurun_out/r2_mg2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_g2_a.json 2> gpurun_out/r2_bench_c3_g2_a.err; echo "bench g2 rc=$?"
CSC_BENCH_COMM=torch timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-verify > gpurun_out/r2_bench_c3_g2_torchcomm.json 2> gpurun_out/r2_bench_c3_g2_torchcomm.err; echo "bench g2 torch rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/r2_bench_c3_b.json 2> gpurun_out/r2_bench_c3_b.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --cpu-sample 0 --k 30 > gpurun_out/r2_bench_c3_k30.json 2> gpurun_out/r2_bench_c3_k30.err; echo "bench c3 k30 rc=$?"
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/r2_bench_c2_b.json 2> gpurun_out/r2_bench_c2_b.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-e2e --cpu-sample 0 --k 30 > gpurun_out/r2_bench_c2_k30.json 2> gpurun_out/r2_bench_c2_k30.err; echo "bench c2 k30 rc=$?"
grep -h "stage ms" gpurun_out/r2_bench_c3_g2_a.err gpurun_out/r2_bench_c3_g2_torchcomm.err gpurun_out/r2_bench_c3_b.err gpurun_out/r2_bench_c3_k30.err gpurun_out/r2_bench_c2_b.err gpurun_out/r2_bench_c2_k30.err
