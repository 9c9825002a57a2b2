"""f1: the community detection that run_clustering hands the graph to (src/scrna/processing.jl:233-234, 279-280:
Leiden.leiden(adj_mat, resolution = res)) on the device.  The reference's Leiden is sequential and seeded from Julia's
RNG, so partitions are compared by what the algorithm guarantees and optimises: CPM quality against the oracle's
sequential Leiden, connected communities, recovery of planted clusters, and the fields run_clustering fills."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def knn_graph(cs, oracle, ctx, n, d, n_types, k, spread, seed, mode="sum"):
    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(n_types, d)) * spread
    truth = rng.integers(0, n_types, n)
    x = (cent[truth] + rng.normal(size=(n, d))).astype(np.float32)
    e = ctx.dense_upload(x)
    idx, _ = ctx.knn(e, e, k)
    g = ctx.graph_build(idx, n, k, 0, n, {"sum": cs._capi.GRAPH_SUM, "binary": cs._capi.GRAPH_BINARY, "snn": cs._capi.GRAPH_SNN_JACCARD}[mode])
    return g, truth


@pytest.mark.parametrize("n,k,gamma,mode", [(3000, 15, 0.02, "sum"), (3000, 15, 0.05, "binary"), (2000, 20, 0.01, "snn"),
                                            (6000, 30, 0.06, "sum")])
def test_leiden_quality_and_guarantees_vs_oracle(cs, oracle, ctx, n, k, gamma, mode):
    g, truth = knn_graph(cs, oracle, ctx, n, 12, 9, k, 3.0, seed=n + k, mode=mode)
    adj = g.to_scipy(np.float64)
    got = ctx.leiden(g, gamma)
    lab = got["labels"]
    assert lab.shape == (n,) and lab.min() == 0 and lab.max() == got["n_communities"] - 1
    assert np.array_equal(np.unique(lab), np.arange(got["n_communities"]))
    # the reported quality is the CPM quality of the reported partition
    assert np.isclose(got["quality"], oracle.cpm_quality(adj, lab, gamma), rtol=1e-9)
    # Leiden's guarantee: every community is connected
    assert oracle.communities_connected(adj, lab)
    # as good as the sequential Leiden of the oracle (random, seeded): within 3 % of its quality, far above trivial partitions
    wlab, wq = oracle.leiden_cpm(adj, gamma, seed=1)
    assert got["quality"] >= 0.97 * wq, (got["quality"], wq)
    assert got["quality"] > max(oracle.cpm_quality(adj, np.zeros(n, int), gamma), 0.0)
    # deterministic
    again = ctx.leiden(g, gamma)
    assert np.array_equal(again["labels"], lab)


def test_leiden_recovers_planted_clusters(cs, oracle, ctx):
    n, types = 8000, 6
    g, truth = knn_graph(cs, oracle, ctx, n, 10, types, 15, 6.0, seed=4, mode="sum")
    got = ctx.leiden(g, 0.002)
    lab = got["labels"]
    # every found community is (almost) pure, and the big ones map one-to-one onto the planted clusters
    big = [c for c in range(got["n_communities"]) if (lab == c).sum() >= 200]
    assert len(big) == types
    for c in big:
        counts = np.bincount(truth[lab == c], minlength=types)
        assert counts.max() >= 0.99 * counts.sum()
    assert sum((lab == c).sum() for c in big) >= 0.99 * n


def test_leiden_edge_cases(cs, oracle, ctx):
    # resolution 0: every connected component becomes one community; a huge resolution: every node alone
    n = 1500
    g, _ = knn_graph(cs, oracle, ctx, n, 8, 5, 10, 8.0, seed=2, mode="binary")
    adj = g.to_scipy(np.float64)
    import scipy.sparse.csgraph as csg
    n_comp, comp = csg.connected_components(adj, directed=False)
    zero = ctx.leiden(g, 0.0)
    assert zero["n_communities"] == n_comp
    assert all(np.unique(zero["labels"][comp == c]).size == 1 for c in range(n_comp))
    lone = ctx.leiden(g, 1e6)
    assert lone["n_communities"] == n and np.isclose(lone["quality"], 0.0)


def test_run_clustering_fills_the_reference_fields(cs, oracle, ctx, small_counts):
    """clustData.clustering (cluster, cell_id) ordered like colnames(obj); metaData.cluster; ledein_res.partition."""
    genes = [f"g{i}" for i in range(small_counts.shape[0])]
    cells = [f"c{i}" for i in range(small_counts.shape[1])]
    obj = cs.scRNAObject(cs.RawCountObject(small_counts, cells, genes))
    obj = cs.normalize_object(obj, ctx=ctx)
    obj = cs.find_variable_genes(obj, nFeatures=300, ctx=ctx)
    obj = cs.scale_object(obj, features=obj.varGene.var_gene, ctx=ctx)
    obj = cs.run_pca(obj, maxoutdim=10, ctx=ctx)
    obj = cs.run_clustering(obj, n_neighbors=15, res=0.03, ctx=ctx)
    cl = obj.clustData
    n = len(obj.rawCount.cell_name)
    assert list(cl.clustering.columns) == ["cluster", "cell_id"] and list(cl.clustering["cell_id"]) == obj.rawCount.cell_name
    assert list(obj.metaData["cluster"]) == list(cl.clustering["cluster"])
    part = cl.ledein_res["partition"]
    assert sum(len(p) for p in part) == n and len(part[0]) >= len(part[-1])
    assert set(cl.clustering["cluster"]) == {str(i + 1) for i in range(len(part))}
    lab = np.empty(n, dtype=np.int64)
    for i, p in enumerate(part):
        lab[p] = i
    assert np.isclose(cl.ledein_res["quality"], oracle.cpm_quality(cl.adj_mat, lab, 0.03), rtol=1e-9)
    # find_all_markers runs straight on the result
    markers = cs.find_all_markers(obj, ctx=ctx)
    assert "cluster" in markers.columns
