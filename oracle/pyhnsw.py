"""ctypes binding of oracle/hnsw_oracle.c (TEST INFRASTRUCTURE ONLY).

FlatGraph mirrors the arrays cdb_index_set_graph takes (see hnsw_oracle.h)."""
import ctypes as C
import os

import numpy as np

from . import pyoracle as po


class _Graph(C.Structure):
    _fields_ = [
        ("num_levels", C.c_uint32), ("neighbors_count", C.c_uint32), ("level0_neighbors_count", C.c_uint32),
        ("n", C.c_uint32), ("entry", C.c_uint32),
        ("cnt", C.c_void_p), ("node_row", C.c_void_p), ("adj", C.c_void_p), ("child", C.c_void_p),
        ("metric", C.c_int), ("storage_type", C.c_int), ("dim", C.c_size_t),
        ("codes", C.c_void_p), ("mags", C.c_void_p),
    ]


class _TraverseParams(C.Structure):
    _fields_ = [("ef", C.c_uint32), ("shortlist_size", C.c_uint32), ("final_len", C.c_uint32), ("self_id", C.c_uint32)]


def _lib():
    L = po.lib()
    if not getattr(L, "_hnsw_declared", False):
        L.orc_hnsw_build.restype = C.c_void_p
        L.orc_hnsw_build.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_built_graph.restype = C.POINTER(_Graph)
        L.orc_built_graph.argtypes = [C.c_void_p]
        L.orc_built_free.argtypes = [C.c_void_p]
        L.orc_ann_search.restype = C.c_int
        L.orc_ann_search.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_dedup_filter.restype = C.c_size_t
        L.orc_dedup_filter.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.orc_hnsw_search_batch.restype = C.c_int
        L.orc_hnsw_search_batch.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float,
                                            C.c_uint32, C.c_uint32, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L._hnsw_declared = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class FlatGraph:
    """numpy copy of a flat HNSW graph + the vector table it refers to."""

    def __init__(self, metric, storage_type, dim, codes, mags, n, num_levels, neighbors_count, level0_neighbors_count,
                 entry, node_row, adj, child):
        self.metric, self.storage_type, self.dim = int(metric), int(storage_type), int(dim)
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.mags = np.ascontiguousarray(mags, dtype=np.float32)
        self.n, self.num_levels = int(n), int(num_levels)
        self.neighbors_count, self.level0_neighbors_count = int(neighbors_count), int(level0_neighbors_count)
        self.entry = int(entry)
        self.node_row = [np.ascontiguousarray(a, dtype=np.uint32) for a in node_row]
        self.adj = [np.ascontiguousarray(a, dtype=np.uint32) for a in adj]
        self.child = [np.ascontiguousarray(a, dtype=np.uint32) for a in child]
        self.cnt = np.array([a.size for a in self.node_row], dtype=np.uint32)

    def cstruct(self):
        L1 = self.num_levels + 1
        self._nr = (C.c_void_p * L1)(*[a.ctypes.data for a in self.node_row])
        self._ad = (C.c_void_p * L1)(*[a.ctypes.data for a in self.adj])
        self._ch = (C.c_void_p * L1)(*[a.ctypes.data for a in self.child])
        g = _Graph(self.num_levels, self.neighbors_count, self.level0_neighbors_count, self.n, self.entry,
                   self.cnt.ctypes.data, C.cast(self._nr, C.c_void_p), C.cast(self._ad, C.c_void_p),
                   C.cast(self._ch, C.c_void_p), self.metric, self.storage_type, self.dim,
                   self.codes.ctypes.data, self.mags.ctypes.data)
        return g

    def nbrs(self, level):
        return self.level0_neighbors_count if level == 0 else self.neighbors_count


def build(metric, storage_type, vectors, root_vector, lo=-1.0, hi=1.0, num_levels=9, neighbors_count=32,
          level0_neighbors_count=64, ef_construction=128, shortlist_size=64, seed=1):
    """quantize rows + root (row n) and run the deterministic single-threaded builder."""
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    n, dim = vectors.shape
    allv = np.concatenate([vectors, np.asarray(root_vector, dtype=np.float32)[None]], axis=0)
    codes, mags = po.quantize_batch(storage_type, allv, lo, hi)
    L = _lib()
    h = L.orc_hnsw_build(metric, storage_type, dim, _p(codes), _p(mags), n, num_levels, neighbors_count,
                         level0_neighbors_count, ef_construction, shortlist_size, seed)
    g = L.orc_built_graph(h).contents
    L1 = num_levels + 1
    cnt = np.ctypeslib.as_array(C.cast(g.cnt, C.POINTER(C.c_uint32)), shape=(L1,)).copy()
    nrp = C.cast(g.node_row, C.POINTER(C.c_void_p))
    adp = C.cast(g.adj, C.POINTER(C.c_void_p))
    chp = C.cast(g.child, C.POINTER(C.c_void_p))
    node_row, adj, child = [], [], []
    for lv in range(L1):
        c = int(cnt[lv])
        nb = level0_neighbors_count if lv == 0 else neighbors_count
        node_row.append(np.ctypeslib.as_array(C.cast(nrp[lv], C.POINTER(C.c_uint32)), shape=(c,)).copy())
        adj.append(np.ctypeslib.as_array(C.cast(adp[lv], C.POINTER(C.c_uint32)), shape=(c * nb,)).copy())
        child.append(np.ctypeslib.as_array(C.cast(chp[lv], C.POINTER(C.c_uint32)), shape=(c,)).copy())
    fg = FlatGraph(metric, storage_type, dim, codes, mags, n, num_levels, neighbors_count, level0_neighbors_count,
                   g.entry, node_row, adj, child)
    L.orc_built_free(h)
    return fg


def ann_search(fg, qcode, qmag, ef_search=256, shortlist_size=64):
    """-> (status, rows u32[], scores f32[], evals, pops): concatenated per-level results, top level first"""
    L = _lib()
    g = fg.cstruct()
    cap = (fg.num_levels + 1) * 100
    rows = np.zeros(cap, dtype=np.uint32)
    scores = np.zeros(cap, dtype=np.float32)
    n = C.c_size_t(0)
    ev, pp = C.c_uint64(0), C.c_uint64(0)
    qcode = np.ascontiguousarray(qcode, dtype=np.uint8)
    rc = L.orc_ann_search(C.byref(g), _p(qcode), float(qmag), ef_search, shortlist_size, _p(rows), _p(scores), cap,
                          C.byref(n), C.byref(ev), C.byref(pp))
    return rc, rows[: n.value].copy(), scores[: n.value].copy(), ev.value, pp.value


def dedup_filter(fg, rows, scores, k):
    L = _lib()
    g = fg.cstruct()
    rows = np.ascontiguousarray(rows, dtype=np.uint32).copy()
    scores = np.ascontiguousarray(scores, dtype=np.float32).copy()
    m = L.orc_dedup_filter(C.byref(g), _p(rows), _p(scores), rows.size, k)
    return rows[:m], scores[:m]


def search_batch(fg, raw, queries, k, lo=-1.0, hi=1.0, ef_search=256, shortlist_size=64, threads=None):
    """search_internal for a batch -> (ids, scores, counts, err, evals, pops)"""
    L = _lib()
    g = fg.cstruct()
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    nq = queries.shape[0]
    ids = np.zeros((nq, k), dtype=np.uint32)
    scores = np.zeros((nq, k), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    err = np.zeros(nq, dtype=np.uint8)
    ev, pp = C.c_uint64(0), C.c_uint64(0)
    rc = L.orc_hnsw_search_batch(C.byref(g), _p(raw), _p(queries), nq, lo, hi, ef_search, shortlist_size, k,
                                 threads or os.cpu_count() or 1, _p(ids), _p(scores), _p(counts), _p(err),
                                 C.byref(ev), C.byref(pp))
    assert rc == 0
    return ids, scores, counts, err, ev.value, pp.value
