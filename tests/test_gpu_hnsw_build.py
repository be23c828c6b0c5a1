"""GPU-side HNSW build (hnsw_build.cu):
  * inserted one vector at a time it must produce EXACTLY the graph of the oracle's deterministic restatement of
    index_embedding / create_node_edges / add_neighbor (same level RNG, same root vector);
  * with concurrent batches (like the reference's rayon build) the graph differs, so: structural invariants,
    CUDA search == oracle search on the exported graph (bit-exact), and recall against brute force."""
import numpy as np
import pytest

import cosdata_b200 as cdb
import oracle as orc
from oracle import pyhnsw

pytestmark = pytest.mark.gpu
ST, MK = cdb.StorageType, cdb.DistanceMetricKind
EMPTY = 0xFFFFFFFF


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def clustered(n, dim, seed):
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(32, dim)).astype(np.float32)
    v = (centres[rng.integers(0, 32, n)] + 0.35 * rng.normal(size=(n, dim))).astype(np.float32)
    return (v / (np.abs(v).max() * 1.01)).astype(np.float32)


def root_vector(seed, dim, lo=-1.0, hi=1.0):
    s = orc.synth(seed ^ 0x526F6F74, 0, dim)
    return (np.float32(lo) + (s + np.float32(1.0)) * np.float32(0.5) * np.float32(hi - lo)).astype(np.float32)


def build_gpu(vecs, st, metric, levels, nb, nb0, efc, max_batch, seed):
    n, dim = vecs.shape
    ix = cdb.DenseIndex(dim=dim, storage_type=st, metric=metric, capacity=n + 1, keep_raw_f32=True)
    ix.append(vecs)
    ix.build_graph(levels, nb, nb0, efc, 64, max_batch, seed)
    assert len(ix) == n + 1
    return ix, ix.read_graph()


def flat_from(ix, g, st, metric, dim, n):
    codes, mags = ix.read_codes(0, n + 1)
    return pyhnsw.FlatGraph(int(metric), int(st), dim, codes, mags, n, g["num_levels"], g["neighbors_count"],
                            g["level0_neighbors_count"], g["entry"], g["node_row"], g["adj"], g["child"])


@pytest.mark.parametrize("st,metric", [(ST.HalfPrecisionFP, MK.Cosine), (ST.UnsignedByte, MK.Cosine), (ST.SubByte2, MK.DotProduct)])
def test_sequential_gpu_build_equals_oracle_builder(st, metric):
    n, dim, seed = 1200, 32, 5
    vecs = clustered(n, dim, 11)
    ix, g = build_gpu(vecs, st, metric, levels=4, nb=8, nb0=16, efc=32, max_batch=1, seed=seed)
    fg = pyhnsw.build(int(metric), int(st), vecs, root_vector(seed, dim), num_levels=4, neighbors_count=8,
                      level0_neighbors_count=16, ef_construction=32, shortlist_size=64, seed=seed)
    codes, mags = ix.read_codes(0, n + 1)
    assert np.array_equal(codes, fg.codes) and np.array_equal(bits(mags), bits(fg.mags))       # same root vector too
    assert g["entry"] == fg.entry
    for lv in range(5):
        assert np.array_equal(g["node_row"][lv], fg.node_row[lv]), lv
        assert np.array_equal(g["adj"][lv], fg.adj[lv]), lv
        if lv > 0:
            assert np.array_equal(g["child"][lv], fg.child[lv]), lv
    ix.close()


def test_batched_gpu_build_invariants_search_parity_and_recall():
    n, dim, k = 6000, 48, 10
    vecs = clustered(n, dim, 21)
    st, metric = ST.HalfPrecisionFP, MK.Cosine
    ix, g = build_gpu(vecs, st, metric, levels=5, nb=16, nb0=32, efc=64, max_batch=256, seed=9)
    # structure
    assert np.array_equal(g["node_row"][0], np.arange(n + 1, dtype=np.uint32)) and g["root_row"] == n
    for lv in range(6):
        nbc = g["level0_neighbors_count"] if lv == 0 else g["neighbors_count"]
        adj = g["adj"][lv].reshape(-1, nbc)
        valid = adj[adj != EMPTY]
        assert valid.size == 0 or valid.max() < adj.shape[0]
        assert not np.any(adj == np.arange(adj.shape[0], dtype=np.uint32)[:, None])
        for i in range(0, adj.shape[0], 7):
            row = adj[i][adj[i] != EMPTY]
            assert row.size == np.unique(row).size
        if lv >= 1:
            assert g["node_row"][lv][0] == n
            assert np.array_equal(g["node_row"][lv - 1][g["child"][lv]], g["node_row"][lv])
    assert 0.15 * n < g["node_row"][1].size - 1 < 0.35 * n
    deg0 = (g["adj"][0].reshape(-1, 32) != EMPTY).sum(1)
    assert (deg0[:n] > 0).mean() > 0.99
    # CUDA search == oracle search on the exported graph
    fg = flat_from(ix, g, st, metric, dim, n)
    rng = np.random.default_rng(2)
    queries = (vecs[rng.integers(0, n, 48)] + 0.05 * rng.normal(size=(48, dim))).astype(np.float32)
    ids, scores, counts, err = ix.batch_search(queries, k, cdb.SearchMode.HNSW, ef_search=64, shortlist_size=64)
    w = pyhnsw.search_batch(fg, vecs, queries, k, ef_search=64)
    assert np.array_equal(ids, w[0]) and np.array_equal(bits(scores), bits(w[1])) and np.array_equal(counts, w[2])
    # recall against brute force, and no worse than a graph from the sequential oracle builder
    gt, _ = orc.brute_topk_f32(vecs, queries, k)
    recall = np.mean([len(set(ids[i]) & set(gt[i])) / k for i in range(len(queries))])
    fo = pyhnsw.build(int(metric), int(st), vecs, root_vector(9, dim), num_levels=5, neighbors_count=16,
                      level0_neighbors_count=32, ef_construction=64, seed=9)
    wo = pyhnsw.search_batch(fo, vecs, queries, k, ef_search=64)
    recall_o = np.mean([len(set(wo[0][i]) & set(gt[i])) / k for i in range(len(queries))])
    assert recall >= 0.85 and recall >= recall_o - 0.05, (recall, recall_o)
    ix.close()


# ----------------------------------------------------------------------------- collections with a metadata schema
def build_replicas_gpu(pop, st, metric, levels, nb, nb0, efc, max_batch, seed):
    from tests import mdgraph  # noqa: F401
    vecs = pop["vecs"]
    n, dim = vecs.shape
    ix = cdb.DenseIndex(dim=dim, storage_type=st, metric=metric, capacity=n + 2, keep_raw_f32=True)
    ix.append(vecs)
    failed = ix.build_graph_replicas(pop["row"], pop["node_id"], pop["base_id"], pop["md_row"], pop["max_level"], pop["md_bits"],
                                     pop["md_mags"], pop["main_root_md"], pop["pseudo_root_md"], levels, nb, nb0, efc, 64, max_batch, seed)
    assert len(ix) == n + 2
    return ix, failed


def oracle_replicas(ix, pop, st, metric, levels, nb, nb0, efc):
    from oracle import pymeta
    n, dim = pop["vecs"].shape
    codes, mags = ix.read_codes(0, n + 2)                               # rows n / n+1: the roots the library appended
    rows = np.where(pop["row"] == EMPTY, n + 1, pop["row"]).astype(np.uint32)
    rl = pymeta.ReplicaList(rows, pop["node_id"], pop["base_id"], pop["md_row"], pop["max_level"], n, pop["main_root_md"], n + 1,
                            pop["pseudo_root_md"])
    return pymeta.build_md(int(metric), int(st), dim, codes, mags, pop["md_bits"], pop["md_mags"], rl, num_levels=levels,
                           neighbors_count=nb, level0_neighbors_count=nb0, ef_construction=efc)


@pytest.mark.parametrize("st,metric", [(ST.HalfPrecisionFP, MK.Cosine), (ST.UnsignedByte, MK.Cosine), (ST.SubByte2, MK.DotProduct),
                                        (ST.FullPrecisionFP, MK.Cosine)])
def test_sequential_replica_build_equals_oracle_builder(st, metric):
    # Pseudo <-> Metadata edges only on cs == 1.0, Metadata <-> Metadata not on -1.0 (vector_store.rs:1014-1040);
    # inserted one node at a time the GPU build must equal the oracle's restatement array for array
    from tests import mdgraph
    levels, nb, nb0, efc = 4, 8, 16, 32
    pop = mdgraph.replica_population(n=500, dim=32, md_dims=6, levels=levels, seed=13)
    ix, failed = build_replicas_gpu(pop, st, metric, levels, nb, nb0, efc, max_batch=1, seed=5)
    mg, ofailed = oracle_replicas(ix, pop, st, metric, levels, nb, nb0, efc)
    assert np.array_equal(failed, ofailed) and failed.sum() == 0
    g = ix.read_graph()
    ids, mds = ix.read_graph_metadata()
    assert g["entry"] == mg.fg.entry == 0 and mg.pseudo_entry == 1
    kinds_seen = 0
    for lv in range(levels + 1):
        assert np.array_equal(g["node_row"][lv], mg.fg.node_row[lv]), lv
        assert np.array_equal(ids[lv], mg.node_id[lv]) and np.array_equal(mds[lv], mg.node_md[lv]), lv
        assert np.array_equal(g["adj"][lv], mg.fg.adj[lv]), lv
        if lv > 0:
            assert np.array_equal(g["child"][lv], mg.fg.child[lv]), lv
        kinds_seen += int((g["adj"][lv] != EMPTY).sum())
    assert kinds_seen > 4 * len(pop["row"])
    ix.close()


def test_batched_replica_build_filtered_search_parity_and_recall():
    from oracle import pymeta
    from tests import mdgraph
    levels, nb, nb0, efc, k = 5, 16, 32, 64, 10
    st, metric = ST.HalfPrecisionFP, MK.Cosine
    pop = mdgraph.replica_population(n=4000, dim=48, md_dims=6, levels=levels, n_patterns=7, seed=17)
    vecs = pop["vecs"]
    n, dim = vecs.shape
    ix, failed = build_replicas_gpu(pop, st, metric, levels, nb, nb0, efc, max_batch=256, seed=9)
    assert failed.sum() == 0
    g = ix.read_graph()
    ids, mds = ix.read_graph_metadata()
    codes, mags = ix.read_codes(0, n + 2)
    fg = pyhnsw.FlatGraph(int(metric), int(st), dim, codes, mags, n, levels, nb, nb0, g["entry"], g["node_row"], g["adj"], g["child"])
    mg = pymeta.MdGraph(fg, pop["md_bits"], pop["md_mags"], ids, mds, 1)
    # edge rules hold under concurrency too
    for lv in range(levels + 1):
        nbc = nb0 if lv == 0 else nb
        adj = g["adj"][lv].reshape(-1, nbc)
        has = (mds[lv] != EMPTY) & (pop["md_mags"][np.minimum(mds[lv], pop["md_mags"].size - 1)] != 0)
        a_idx, slot = np.nonzero(adj != EMPTY)
        b_idx = adj[a_idx, slot]
        assert np.array_equal(has[a_idx], has[b_idx]), lv                           # Base nodes never meet the pseudo component
        both = has[a_idx]
        pa, pb = pop["md_bits"][mds[lv][a_idx[both]]], pop["md_bits"][mds[lv][b_idx[both]]]
        is_ps = lambda x: (ids[lv][x] >= 0xFFFFFFFF - 257) & (ids[lv][x] <= 0xFFFFFFFF - 2)
        mixed = ~(is_ps(a_idx[both]) & is_ps(b_idx[both]))                           # at least one Metadata end: identical dims
        assert np.array_equal(pa[mixed], pb[mixed]), lv
    # filtered + unfiltered queries: CUDA search == oracle search on the exported graph
    rng = np.random.default_rng(4)
    nq = 60
    src = rng.integers(0, n, nq)
    q = (vecs[src] + 0.03 * rng.normal(size=(nq, dim))).astype(np.float32).clip(-0.999, 0.999)
    filters = [None if i % 3 == 0 else [pop["md_bits"][2 + int(rng.integers(0, 7))].astype(np.int8)] for i in range(nq)]
    r_ids, r_scores, r_counts, r_err = ix.batch_search_filtered(q, filters, k, ef_search=64, shortlist_size=64)
    w = pymeta.search_batch_md(mg, np.concatenate([vecs, np.zeros((2, dim), np.float32)]), q, filters, k, ef_search=64)
    assert np.array_equal(r_err, w[3]) and np.array_equal(r_counts, w[2])
    assert np.array_equal(r_ids, w[0]) and np.array_equal(bits(r_scores), bits(w[1]))
    # a filtered query only returns replicas that carry the filter's pattern; unfiltered ones reach base replicas
    id_to_md = dict(zip(pop["node_id"].tolist(), pop["md_row"].tolist()))
    hits = 0
    for i in range(nq):
        got = r_ids[i, : r_counts[i]]
        if filters[i] is None:
            assert all(id_to_md[int(x)] == 0 for x in got) and r_counts[i] == k
            continue
        want_md = int(np.flatnonzero((pop["md_bits"] == filters[i][0].astype(np.int32)).all(1))[0])
        assert all(id_to_md[int(x)] == want_md for x in got), i
        hits += int(r_counts[i])
    assert hits > 0.5 * k * sum(f is not None for f in filters)
    ix.close()
