"""Multi-GPU path on real GPUs: `torchrun --nproc-per-node N tools/multigpu_check.py [n_genes] [n_cells]`.
Every rank generates its cell shard on its device and runs the sharded path (NCCL collectives); rank 0 also runs
the whole matrix alone and compares: HVG set, explained variance, embedding (sign-aligned), neighbour table."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import csb200 as cs

G = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = cs.Context(local)
comm = cs.TorchComm(device=torch.device("cuda", local)) if os.environ.get("CSC_BENCH_COMM") == "torch" else cs.NcclComm.from_torch(ctx)
offsets = cs.partition_cells(N, world)
lo, hi = int(offsets[rank]), int(offsets[rank + 1])
p = cs.synth.make_params(G, N, density=0.06, seed=11)
params = cs.PathParams(n_features=800, n_pcs=30, k=15, graph_mode="sum")
raw = ctx.generate_counts(p, lo, hi)
out = cs.run_path(ctx, comm, raw, N, lo, params, offsets)
emb = out["embedding"].download(np.float64)
idx = out["knn_idx"].download().reshape(hi - lo, params.k)
pv = out["pca"].info()["principalvars"]
feats = out["hvg"]["features"]
gcp, grv, gv = out["graph"].download(0)
embs = [None] * world
idxs = [None] * world
dist.all_gather_object(embs, emb)
dist.all_gather_object(idxs, idx)
nnzs = [None] * world
dist.all_gather_object(nnzs, int(gv.size))
if rank == 0:
    emb_all, idx_all = np.concatenate(embs), np.concatenate(idxs)
    raw1 = ctx.generate_counts(p, 0, N)
    one = cs.run_path(ctx, cs.LocalComm(), raw1, N, 0, params)
    e1 = one["embedding"].download(np.float64)
    i1 = one["knn_idx"].download().reshape(N, params.k)
    pv1 = one["pca"].info()["principalvars"]
    g1 = one["graph"].download(0)[2].size
    ok_f = np.array_equal(feats, one["hvg"]["features"])
    sign = np.sign(np.sum(emb_all * e1, axis=0))
    err = np.abs(emb_all * sign - e1).max() / np.abs(e1).max()
    same = (idx_all == i1).all(axis=1).mean()
    print(f"world={world}: HVG identical {ok_f}; principalvars rel diff {np.abs(pv - pv1).max() / pv1[0]:.2e}; "
          f"embedding max diff {err:.2e}; rows with identical neighbour lists {same:.5f}; "
          f"graph nnz {sum(nnzs)} vs {g1}", flush=True)
    assert ok_f and err < 1e-4 and same > 0.995 and abs(sum(nnzs) - g1) <= 0.001 * g1
    print("MULTIGPU OK", flush=True)
dist.barrier()
dist.destroy_process_group()
