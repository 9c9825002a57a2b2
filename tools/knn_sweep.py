"""C5 of BASELINE.json: kNN (k = 30) on a 4M-cell x 50-PC embedding, cells sharded over the ranks.
    torchrun --nproc-per-node N tools/knn_sweep.py [n_cells=4000000] [k=30] [reps=3]
The embedding comes from the same latent model as the count generator (cell-type means + cell noise in 50 dimensions
with a decaying spectrum), generated per rank; every rank all-gathers the keys through the library's NCCL communicator
and answers the queries of its own cells.  Prints one JSON line (rank 0): ms per pass (max over ranks, CUDA events on the
library's stream), candidates/s, tensor-roofline fraction of the scan kernel."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import csb200 as cs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
D = 50
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = cs.Context(local)
comm = cs.NcclComm.from_torch(ctx) if world > 1 else cs.LocalComm()
offsets = cs.partition_cells(N, world)
lo, hi = int(offsets[rank]), int(offsets[rank + 1])
# embedding: 64 cell types, type means + noise, component scale decaying like a PCA spectrum (same for every rank)
g = torch.Generator(device="cuda").manual_seed(1234)
scale = torch.tensor(np.sqrt(16.0 * 0.955 ** np.arange(D) + 1.0), device="cuda", dtype=torch.float32)
means = torch.randn(64, D, generator=g, device="cuda") * 0.75
gl = torch.Generator(device="cuda").manual_seed(99 + rank)
typ = torch.randint(0, 64, (hi - lo,), generator=gl, device="cuda")
emb_t = (means[typ] + 0.65 * torch.randn(hi - lo, D, generator=gl, device="cuda")) * scale
emb = ctx.dense_upload(emb_t.cpu().numpy())
del emb_t
times, scans = [], []
for it in range(REPS + 1):
    ctx.flush_l2()
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    ctx.stage_times(reset=True)
    ctx.timer_start()
    keys = comm.allgather_rows(ctx, emb, offsets)
    idx, dd = ctx.knn(emb, keys, K, query_offset=lo)
    ms = ctx.timer_stop()
    st, _ = ctx.stage_times()
    fb = ctx.knn_last_fallback_rows()
    if it > 0:
        times.append(ms)
        scans.append(st["knn_scan"])
    del keys, idx, dd
t = torch.tensor([float(np.mean(times)), float(np.mean(scans)), float(fb)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms, scan_ms, fbm = (float(x) for x in t.tolist())
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    flops = 2.0 * (hi - lo) * N * D
    print(json.dumps({"workload": f"C5 kNN: {N} cells x {D} PCs, k={K}, cosine", "n_gpus": world, "ms_per_pass": ms,
                      "cells_per_s": N / (ms * 1e-3), "knn_scan_ms": scan_ms,
                      "scan_TFLOP/s_per_gpu": flops / (scan_ms * 1e-3) / 1e12,
                      "scan_frac_of_measured_bf16_sustained": flops / (scan_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                      "max_fallback_rows_per_rank": fbm}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
