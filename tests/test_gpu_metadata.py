"""Metadata-filter arms on the GPU: pairwise replica-kind arms and the filtered HNSW search must reproduce
oracle/metadata_oracle.c (itself pinned against a pure-Python restatement in tests/test_oracle_metadata.py) bit for bit."""
import numpy as np
import pytest

import cosdata_b200 as cdb
import oracle as orc
from oracle import pymeta
from tests import mdgraph

pytestmark = pytest.mark.gpu
ST, MK = cdb.StorageType, cdb.DistanceMetricKind


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("st", [ST.HalfPrecisionFP, ST.UnsignedByte, ST.SubByte2, ST.FullPrecisionFP])
@pytest.mark.parametrize("md_dims", [5, 8, 19])
def test_pairwise_replica_arms_match_oracle(st, md_dims):
    dim, n = 40, 600
    rng = np.random.default_rng(10 + md_dims)
    xv, yv = orc.synth_matrix(1, n, dim), orc.synth_matrix(2, n, dim)
    yv[7] = 0.0                                                          # zero vector norm -> CalculationError in the vector arms
    xc, xm = orc.quantize_batch(int(st), xv)
    yc, ym = orc.quantize_batch(int(st), yv)
    pats = (rng.random((6, md_dims)) < 0.5).astype(np.int32)
    pats[0] = 0

    def side(seed, query_side):
        r = np.random.default_rng(seed)
        mb = pats[r.integers(0, 6, n)].copy()
        if query_side:
            mb = np.where(r.random((n, md_dims)) < 0.15, -mb, mb)        # filter dims are -1/0/1
        mm = np.array([pymeta.metadata_mag(b) for b in mb], dtype=np.float32)
        ids = r.integers(0, 5000, n).astype(np.uint32)
        special = r.random(n)
        ids = np.where(special < 0.25, 0xFFFFFFFF - 257 + r.integers(0, 256, n), ids).astype(np.uint32)   # pseudo range
        return {"ids": ids, "has_id": (r.random(n) < 0.7).astype(np.uint8), "md_bits": mb, "md_mags": mm,
                "has_md": (r.random(n) < 0.8).astype(np.uint8)}
    x = dict(side(3, True), codes=xc, mags=xm)
    y = dict(side(4, False), codes=yc, mags=ym)
    seen = set()
    for metric in (MK.Cosine, MK.DotProduct):
        if metric == MK.DotProduct and st == ST.FullPrecisionFP:
            continue
        val, status = cdb.DistanceMetric(metric).calculate_pairs_md(st, dim, md_dims, x, y)
        for i in range(n):
            vx = pymeta.VectorData(xc[i], xm[i], x["ids"][i] if x["has_id"][i] else None, x["md_bits"][i] if x["has_md"][i] else None, x["md_mags"][i])
            vy = pymeta.VectorData(yc[i], ym[i], y["ids"][i] if y["has_id"][i] else None, y["md_bits"][i] if y["has_md"][i] else None, y["md_mags"][i])
            rc, want = pymeta.distance_md(int(metric), int(st), dim, md_dims, vx, vy)
            assert status[i] == rc, (i, metric)
            if rc == 0:
                assert bits(val[i]) == bits(want), (i, metric)
            if metric == MK.Cosine:
                seen.add((pymeta.replica_kind(vy), pymeta.replica_kind(vx), rc))
    assert {(k[0], k[1]) for k in seen} == {(a, b) for a in range(3) for b in range(3)}      # all nine kind pairs exercised
    assert any(rc == 7 for _, _, rc in seen)


@pytest.mark.parametrize("st,metric", [(ST.HalfPrecisionFP, MK.Cosine), (ST.UnsignedByte, MK.Cosine), (ST.SubByte2, MK.DotProduct),
                                        (ST.FullPrecisionFP, MK.Cosine)])
@pytest.mark.parametrize("ef", [8, 40])
def test_filtered_hnsw_search_matches_oracle(st, metric, ef):
    vecs, mg = mdgraph.build(n=700, dim=32, md_dims=7, levels=4, nb=8, nb0=16, storage_type=int(st), metric=int(metric), seed=21)
    q, filters = mdgraph.make_queries(vecs, mg, 64, seed=9)
    filters[5] = []                                                      # Some(empty): the reference panics -> UNREACHABLE flag
    filters[6] = [np.zeros(mg.md_dims, np.int8)]                         # all-zero filter = Base query at the pseudo root: unreachable arm
    k = 10
    fg = mg.fg
    ix = cdb.DenseIndex(dim=32, storage_type=st, metric=metric, capacity=vecs.shape[0], keep_raw_f32=True)
    ix.append(vecs)
    ix.set_graph(fg.num_levels, fg.neighbors_count, fg.level0_neighbors_count, fg.entry, fg.n, fg.node_row, fg.adj, fg.child)
    ix.set_graph_metadata(mg.md_bits, mg.md_mags, mg.node_id, mg.node_md, mg.pseudo_entry)
    ev0, pp0 = ix.hnsw_counters()
    ids, scores, counts, err = ix.batch_search_filtered(q, filters, k, ef_search=ef, shortlist_size=64)
    ev1, pp1 = ix.hnsw_counters()
    want_ids, want_scores, want_counts, want_err, ev, pp = pymeta.search_batch_md(mg, vecs, q, filters, k, ef_search=ef)
    assert np.array_equal(err, want_err)
    assert np.array_equal(counts, want_counts)
    assert np.array_equal(ids, want_ids)
    assert np.array_equal(bits(scores), bits(want_scores))
    sel = np.flatnonzero(want_err == 0)                                  # same traversal node for node (error-free queries)
    ev0, pp0 = ix.hnsw_counters()
    ix.batch_search_filtered(q[sel], [filters[i] for i in sel], k, ef_search=ef, shortlist_size=64)
    ev1, pp1 = ix.hnsw_counters()
    _, _, _, _, ev, pp = pymeta.search_batch_md(mg, vecs, q[sel], [filters[i] for i in sel], k, ef_search=ef)
    assert (ev1 - ev0, pp1 - pp0) == (ev, pp)
    assert (want_err == 4).sum() >= (2 if metric == MK.Cosine else 1) and (want_counts > 0).sum() > 30
    # the plain entry point on the same graph = every query without a filter
    ids2, scores2, counts2, err2 = ix.batch_search(q, k, cdb.SearchMode.HNSW, ef_search=ef, shortlist_size=64)
    w_ids, w_scores, w_counts, w_err, _, _ = pymeta.search_batch_md(mg, vecs, q, [None] * q.shape[0], k, ef_search=ef)
    assert np.array_equal(ids2, w_ids) and np.array_equal(bits(scores2), bits(w_scores)) and np.array_equal(err2, w_err)
    ix.close()


def test_filter_api_errors():
    vecs, mg = mdgraph.build(n=100, dim=16, storage_type=4, metric=0, seed=2)
    fg = mg.fg
    ix = cdb.DenseIndex(dim=16, storage_type=ST.HalfPrecisionFP, capacity=vecs.shape[0], keep_raw_f32=True)
    ix.append(vecs)
    with pytest.raises(cdb.CosdataError):                               # metadata before a graph
        ix.set_graph_metadata(mg.md_bits, mg.md_mags, mg.node_id, mg.node_md, mg.pseudo_entry)
    ix.set_graph(fg.num_levels, fg.neighbors_count, fg.level0_neighbors_count, fg.entry, fg.n, fg.node_row, fg.adj, fg.child)
    ix.md_dims = mg.md_dims
    with pytest.raises(cdb.CosdataError):                               # filters without graph metadata
        ix.batch_search_filtered(vecs[:2], [None, [np.ones(mg.md_dims, np.int8)]], 3)
    bad = [a.copy() for a in mg.node_md]
    bad[0][3] = 10 ** 6
    with pytest.raises(cdb.CosdataError):
        ix.set_graph_metadata(mg.md_bits, mg.md_mags, mg.node_id, bad, mg.pseudo_entry)
    bad_adj = [a.copy() for a in fg.adj]
    bad_adj[0][5] = 10 ** 6
    with pytest.raises(cdb.CosdataError):                               # out-of-range adjacency is rejected at upload
        ix.set_graph(fg.num_levels, fg.neighbors_count, fg.level0_neighbors_count, fg.entry, fg.n, fg.node_row, bad_adj, fg.child)
    ix.close()
