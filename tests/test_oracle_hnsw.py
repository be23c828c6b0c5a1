"""Pins oracle/hnsw_oracle.c: an independent pure-Python restatement of traverse_find_nearest /
ann_search / remove_duplicates_and_filter (src/vector_store.rs:256-402, 1112-1204,
src/models/common.rs:381-412, src/models/fixedset.rs) must agree with the C oracle on graphs
produced by the deterministic builder, plus structural and recall properties."""
import heapq

import numpy as np
import pytest

import oracle as orc
from oracle import pyhnsw

ROOT_ID, QUERY_ID, EMPTY = 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFF


def make_graph(n=700, dim=32, st=orc.ST_F16, metric=orc.METRIC_COSINE, levels=4, nb=8, nb0=16, efc=32, seed=3):
    vecs = orc.synth_matrix(9000 + n + dim, n, dim)
    root = orc.synth_matrix(9999, 1, dim)[0]
    fg = pyhnsw.build(metric, st, vecs, root, num_levels=levels, neighbors_count=nb, level0_neighbors_count=nb0,
                      ef_construction=efc, shortlist_size=64, seed=seed)
    return fg, vecs


class PyFixedSet:
    def __init__(self, length):
        self.b = [0] * length
        self.len = length

    def _bk(self, v):
        return (v >> 6) & ((self.len - 1) & 0xFFFFFFFF)

    def insert(self, v):
        self.b[self._bk(v)] |= 1 << (v & 0x3F)

    def member(self, v):
        return (self.b[self._bk(v)] >> (v & 0x3F)) & 1


def py_traverse(fg, level, entry, qcode, qmag, ef, shortlist, final_len, fs):
    nb = fg.nbrs(level)
    rows, adj = fg.node_row[level], fg.adj[level].reshape(-1, nb)
    cb = orc.code_bytes(fg.storage_type, fg.dim)
    codes = fg.codes.reshape(-1, cb)

    def nid(row):
        return int(row) if row < fg.n else ROOT_ID

    def dist(row):
        rc, v = orc.distance(fg.metric, fg.storage_type, fg.dim, qcode, qmag, codes[row], fg.mags[row])
        assert rc == 0
        return v

    def key(v, i):
        return (orc.order_key(fg.metric, v) << 32) | ((~i) & 0xFFFFFFFF)

    d = dist(rows[entry])
    fs.insert(nid(rows[entry]))
    heap = [(-key(d, nid(rows[entry])), entry, d)]
    results, visited, evals = [], 0, 1
    while heap:
        nk, node, dd = heapq.heappop(heap)
        if visited >= ef:
            break
        visited += 1
        results.append((-nk, node, dd))
        for s in range(min(shortlist, nb)):
            nbl = int(adj[node, s])
            if nbl == EMPTY:
                continue
            row = rows[nbl]
            if fs.member(nid(row)):
                continue
            v = dist(row)
            evals += 1
            fs.insert(nid(row))
            heapq.heappush(heap, (-key(v, nid(row)), nbl, v))
    results.sort(key=lambda t: -t[0])
    return results[:final_len], evals, visited


def py_ann_search(fg, qcode, qmag, ef, shortlist):
    entry, out_rows, out_scores, evals, pops = fg.entry, [], [], 0, 0
    for level in range(fg.num_levels, -1, -1):
        fs = PyFixedSet(fg.nbrs(level))
        fs.insert(QUERY_ID)
        z, e, p = py_traverse(fg, level, entry, qcode, qmag, ef, shortlist, 100, fs)
        evals += e
        pops += p
        out_rows += [int(fg.node_row[level][node]) for _, node, _ in z]
        out_scores += [s for _, _, s in z]
        if level > 0:
            entry = int(fg.child[level][z[0][1]])
    return np.array(out_rows, np.uint32), np.array(out_scores, np.float32), evals, pops


@pytest.mark.parametrize("st,metric", [(orc.ST_F16, orc.METRIC_COSINE), (orc.ST_U8, orc.METRIC_COSINE),
                                        (orc.ST_SUB2, orc.METRIC_DOT), (orc.ST_F32, orc.METRIC_COSINE)])
def test_c_ann_search_equals_python_restatement(st, metric):
    fg, vecs = make_graph(st=st, metric=metric)
    queries = orc.synth_matrix(777, 6, fg.dim)
    for q in queries:
        qc, qm = orc.quantize(st, q)
        for ef in (8, 64):
            rc, rows, scores, ev, pp = pyhnsw.ann_search(fg, qc, qm, ef_search=ef, shortlist_size=64)
            prow, pscore, pev, ppp = py_ann_search(fg, qc, qm, ef, 64)
            assert rc == 0
            assert np.array_equal(rows, prow) and np.array_equal(scores.view(np.uint32), pscore.view(np.uint32))
            assert (ev, pp) == (pev, ppp)


def test_graph_structure_invariants():
    fg, _ = make_graph()
    n = fg.n
    assert fg.cnt[0] == n + 1 and np.array_equal(fg.node_row[0], np.arange(n + 1, dtype=np.uint32))
    for lv in range(fg.num_levels + 1):
        nb = fg.nbrs(lv)
        adj = fg.adj[lv].reshape(-1, nb)
        valid = adj[adj != EMPTY]
        assert valid.size == 0 or valid.max() < fg.cnt[lv]
        assert not np.any(adj == np.arange(adj.shape[0], dtype=np.uint32)[:, None]), "self loop"
        for i in range(adj.shape[0]):                         # no duplicate neighbours
            row = adj[i][adj[i] != EMPTY]
            assert row.size == np.unique(row).size
        if lv >= 1:
            assert fg.node_row[lv][0] == n                     # root is local 0 on every upper level
            ch = fg.child[lv]
            assert np.array_equal(fg.node_row[lv - 1][ch], fg.node_row[lv])   # child = same vector one level down
    # level population follows P(level >= L) = 4^-L (common.rs:421-429 probabilities): loose check
    assert 0.15 * n < fg.cnt[1] - 1 < 0.35 * n
    # every inserted row is reachable at level 0 from some neighbour or has outgoing edges
    deg0 = (fg.adj[0].reshape(-1, fg.nbrs(0)) != EMPTY).sum(1)
    assert (deg0[:n] > 0).mean() > 0.99


def test_fixed_set_aliasing_is_reproduced():
    # ids equal mod 64*len alias: at level 0 (len 16 -> 1024 bits) id 5 + 1024 is "visited" once 5 is
    fs = PyFixedSet(16)
    fs.insert(5)
    assert fs.member(5 + 1024) and not fs.member(6)
    # the query id u32::MAX-1 occupies bit (2^32-2) mod 1024 = 1022 -> node 1022 is unreachable at level 0
    fs = PyFixedSet(16)
    fs.insert(QUERY_ID)
    assert fs.member(1022)


def test_dedup_filter_matches_reference_rules():
    fg, _ = make_graph()
    rows = np.array([5, 9, 5, fg.n, 7, 9, 11], dtype=np.uint32)      # duplicates and the root row
    scores = np.array([0.5, 0.9, 0.4, 2.0, 0.9, 0.1, 0.3], dtype=np.float32)
    r, s = pyhnsw.dedup_filter(fg, rows, scores, 1)                   # truncate to 5*k = 5
    assert r.tolist() == [7, 9, 5, 11] and s.tolist() == [np.float32(0.9), np.float32(0.9), 0.5, np.float32(0.3)]
    r, s = pyhnsw.dedup_filter(fg, np.arange(20, dtype=np.uint32), np.linspace(0, 1, 20, dtype=np.float32), 2)
    assert r.size == 10 and r[0] == 19


def test_search_batch_recall_against_brute_force():
    # tests/test-dataset.py:695-772 procedure: recall of the ANN result vs brute-force ground truth
    n, dim, k = 3000, 48, 10
    rng = np.random.default_rng(1)
    centres = rng.normal(size=(32, dim)).astype(np.float32)
    vecs = (centres[rng.integers(0, 32, n)] + 0.35 * rng.normal(size=(n, dim))).astype(np.float32)
    vecs /= np.abs(vecs).max() * 1.01                                  # inside (-1,1) for the quantizers
    root = orc.synth_matrix(4242, 1, dim)[0]
    fg = pyhnsw.build(orc.METRIC_COSINE, orc.ST_F16, vecs, root, num_levels=5, neighbors_count=16,
                      level0_neighbors_count=32, ef_construction=64, seed=11)
    queries = vecs[rng.integers(0, n, 40)] + 0.05 * rng.normal(size=(40, dim)).astype(np.float32)
    ids, scores, counts, err, ev, pp = pyhnsw.search_batch(fg, vecs, queries, k, ef_search=64)
    gt, _ = orc.brute_topk_f32(vecs, queries, k)
    recall = np.mean([len(set(ids[i]) & set(gt[i])) / k for i in range(len(queries))])
    assert err.sum() == 0 and counts.min() == k
    assert recall > 0.85, recall
    # scores are the exact f32 re-rank values of the returned ids
    for i in range(5):
        want_ids, want_scores = orc.rerank_f32(vecs, queries[i], ids[i], k)
        assert np.array_equal(ids[i], want_ids) and np.array_equal(scores[i], want_scores)
