"""The N > 1 host path on CPU: world_size 2 over gloo (127.0.0.1), with the numpy test double standing in
for the per-rank CUDA kernels.  Checks that sharded partials + collectives + finish equal the
single-rank result and the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir, balance_nnz):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import csb200 as cs
    from _numpy_backend import NpSparse, NumpyBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = sp.load_npz(os.path.join(tmpdir, "counts.npz")).tocsc()
    n = counts.shape[1]
    offsets = cs.partition_cells(n, world, counts.indptr if balance_nnz else None)
    lo, hi = int(offsets[rank]), int(offsets[rank + 1])
    be = NumpyBackend()
    params = cs.PathParams(n_features=120, n_pcs=8, k=6, graph_mode="sum", keep_norm=True)
    out = cs.run_path(be, cs.TorchComm(device=None), NpSparse(counts[:, lo:hi]), n, lo, params, offsets)
    np.savez(os.path.join(tmpdir, f"rank{rank}.npz"), lo=lo, hi=hi, features=out["hvg"]["features"],
             emb=out["embedding"].a, idx=out["knn_idx"].a, pv=out["pca"].info()["principalvars"],
             adj_data=out["graph"].m.data, adj_indices=out["graph"].m.indices, adj_indptr=out["graph"].m.indptr,
             norm=out["norm"].m.data)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("balance_nnz", [False, True])
def test_two_ranks_equal_one_rank(tmp_path, cs, oracle, balance_nnz):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _numpy_backend import NpSparse, NumpyBackend
    from conftest import make_counts
    counts = make_counts(cs, oracle, 500, 401, density=0.1, seed=3)
    sp.save_npz(str(tmp_path / "counts.npz"), counts)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), balance_nnz), nprocs=2, join=True)
    n = counts.shape[1]
    params = cs.PathParams(n_features=120, n_pcs=8, k=6, graph_mode="sum", keep_norm=True)
    single = cs.run_path(NumpyBackend(), cs.LocalComm(), NpSparse(counts), n, 0, params)
    want = oracle.run_pipeline(counts, nFeatures=120, n_pcs=8, k=6, pca_method="gram")
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert parts[0]["lo"] == 0 and parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == n
    assert 0 < parts[0]["hi"] < n
    for p in parts:
        assert np.array_equal(p["features"], want["hvg"]["features"])          # HVG set identical on every rank
        assert np.allclose(p["pv"], want["pca"]["principalvars"], rtol=1e-9)
    emb = np.concatenate([p["emb"] for p in parts])
    sign = np.sign((emb * want["pca"]["embedding"]).sum(axis=0))
    assert np.allclose(emb * sign, want["pca"]["embedding"], atol=1e-7)
    assert np.allclose(emb, single["embedding"].a, atol=1e-9) or np.allclose(np.abs(emb), np.abs(single["embedding"].a), atol=1e-9)
    idx = np.concatenate([p["idx"].reshape(-1, 6) for p in parts])
    assert np.array_equal(idx, single["knn_idx"].a.reshape(-1, 6))
    full = oracle.adjacency_atlas(idx)
    for p in parts:
        sub = full[:, int(p["lo"]):int(p["hi"])]
        assert np.array_equal(p["adj_indptr"], sub.indptr) and np.array_equal(p["adj_indices"], sub.indices)
        assert np.array_equal(p["adj_data"], sub.data)
    assert np.allclose(np.concatenate([p["norm"] for p in parts]), want["norm"].data)
