"""ctypes binding of oracle/metadata_oracle.c (TEST INFRASTRUCTURE ONLY): metadata-filter arms of the cosine metric and
the filtered HNSW search on a flat graph with replica nodes (see metadata_oracle.h)."""
import ctypes as C
import os

import numpy as np

from . import pyoracle as po
from .pyhnsw import FlatGraph, _Graph

UNREACHABLE = 7
KIND_PSEUDO, KIND_BASE, KIND_METADATA = 0, 1, 2
EMPTY = 0xFFFFFFFF


class _VectorData(C.Structure):
    _fields_ = [("code", C.c_void_p), ("mag", C.c_float), ("has_id", C.c_int), ("id", C.c_uint32),
                ("md_bits", C.c_void_p), ("md_mag", C.c_float)]


class _MdGraph(C.Structure):
    _fields_ = [("g", _Graph), ("md_dims", C.c_size_t), ("md_bits", C.c_void_p), ("md_mags", C.c_void_p),
                ("node_id", C.c_void_p), ("node_md", C.c_void_p), ("pseudo_entry", C.c_uint32)]


class _ReplicaList(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("row", C.c_void_p), ("node_id", C.c_void_p), ("base_id", C.c_void_p),
                ("md_row", C.c_void_p), ("max_level", C.c_void_p), ("main_root_row", C.c_uint32), ("main_root_md", C.c_uint32),
                ("pseudo_root_row", C.c_uint32), ("pseudo_root_md", C.c_uint32)]


def _lib():
    L = po.lib()
    if not getattr(L, "_md_declared", False):
        L.orc_metadata_mag.restype = C.c_float
        L.orc_metadata_mag.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_query_filter_mag.restype = C.c_float
        L.orc_query_filter_mag.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_replica_kind.restype = C.c_int
        L.orc_replica_kind.argtypes = [C.POINTER(_VectorData)]
        L.orc_distance_md.restype = C.c_int
        L.orc_distance_md.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.POINTER(_VectorData), C.POINTER(_VectorData),
                                      C.POINTER(C.c_float)]
        L.orc_hnsw_search_batch_md.restype = C.c_int
        L.orc_hnsw_search_batch_md.argtypes = [C.POINTER(_MdGraph), C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                               C.POINTER(C.c_uint64)]
        L.orc_hnsw_build_md.restype = C.c_void_p
        L.orc_hnsw_build_md.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.POINTER(_ReplicaList), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_md_built_graph.restype = C.POINTER(_MdGraph)
        L.orc_md_built_graph.argtypes = [C.c_void_p]
        L.orc_md_built_free.restype = None
        L.orc_md_built_free.argtypes = [C.c_void_p]
        L._md_declared = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def metadata_mag(dims):
    d = np.ascontiguousarray(dims, dtype=np.int32)
    return np.float32(_lib().orc_metadata_mag(_p(d), d.size))


def query_filter_mag(dims):
    d = np.ascontiguousarray(dims, dtype=np.int8)
    return np.float32(_lib().orc_query_filter_mag(_p(d), d.size))


class VectorData:
    """VectorData { id, quantized_vec, metadata } (src/models/types.rs:203-212)"""

    def __init__(self, code, mag, vid=None, md_bits=None, md_mag=0.0):
        self.code = np.ascontiguousarray(code, dtype=np.uint8)
        self.md = None if md_bits is None else np.ascontiguousarray(md_bits, dtype=np.int32)
        self.c = _VectorData(self.code.ctypes.data, float(mag), 0 if vid is None else 1, 0 if vid is None else int(vid),
                             None if self.md is None else self.md.ctypes.data, float(md_mag))


def replica_kind(v):
    return _lib().orc_replica_kind(C.byref(v.c))


def distance_md(metric, st, dim, md_dims, x, y):
    out = C.c_float(0)
    rc = _lib().orc_distance_md(int(metric), int(st), dim, md_dims, C.byref(x.c), C.byref(y.c), C.byref(out))
    return rc, np.float32(out.value)


class MdGraph:
    """FlatGraph + per-node replica ids / metadata rows + metadata table + pseudo root"""

    def __init__(self, fg, md_bits, md_mags, node_id, node_md, pseudo_entry):
        self.fg = fg
        self.md_bits = np.ascontiguousarray(md_bits, dtype=np.int32)
        self.md_mags = np.ascontiguousarray(md_mags, dtype=np.float32)
        self.md_dims = self.md_bits.shape[1]
        self.node_id = [np.ascontiguousarray(a, dtype=np.uint32) for a in node_id]
        self.node_md = [np.ascontiguousarray(a, dtype=np.uint32) for a in node_md]
        self.pseudo_entry = int(pseudo_entry)

    def cstruct(self):
        L1 = self.fg.num_levels + 1
        self._ni = (C.c_void_p * L1)(*[a.ctypes.data for a in self.node_id])
        self._nm = (C.c_void_p * L1)(*[a.ctypes.data for a in self.node_md])
        return _MdGraph(self.fg.cstruct(), self.md_dims, self.md_bits.ctypes.data, self.md_mags.ctypes.data,
                        C.cast(self._ni, C.c_void_p), C.cast(self._nm, C.c_void_p), self.pseudo_entry)


def pack_filters(filters, md_dims):
    """filters: per query None (no filter) or a list of int8[md_dims] -> (offsets u32[nq+1], dims i8[total, md_dims], has u8[nq])"""
    offs, rows, has = [0], [], []
    for f in filters:
        has.append(0 if f is None else 1)
        for d in (f or []):
            rows.append(np.asarray(d, dtype=np.int8).reshape(md_dims))
        offs.append(len(rows))
    dims = np.stack(rows).astype(np.int8) if rows else np.zeros((0, md_dims), dtype=np.int8)
    return np.array(offs, dtype=np.uint32), np.ascontiguousarray(dims), np.array(has, dtype=np.uint8)


def search_batch_md(mg, raw, queries, filters, k, lo=-1.0, hi=1.0, ef_search=256, shortlist_size=64, threads=None):
    """search_internal with per-query filters -> (ids, scores, counts, err, evals, pops)"""
    g = mg.cstruct()
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    nq = queries.shape[0]
    offs, dims, has = pack_filters(filters, mg.md_dims)
    ids = np.zeros((nq, k), dtype=np.uint32)
    scores = np.zeros((nq, k), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    err = np.zeros(nq, dtype=np.uint8)
    ev, pp = C.c_uint64(0), C.c_uint64(0)
    rc = _lib().orc_hnsw_search_batch_md(C.byref(g), _p(raw), _p(queries), nq, lo, hi, _p(offs), _p(dims), _p(has), ef_search,
                                         shortlist_size, k, threads or os.cpu_count() or 1, _p(ids), _p(scores), _p(counts),
                                         _p(err), C.byref(ev), C.byref(pp))
    assert rc == 0
    return ids, scores, counts, err, ev.value, pp.value


PSEUDO_ROOT_ID = 0xFFFFFFFF - 257


class ReplicaList:
    """the flattened IndexableEmbeddings of preprocess_embedding, in insertion order (see metadata_oracle.h)"""

    def __init__(self, row, node_id, base_id, md_row, max_level, main_root_row, main_root_md, pseudo_root_row, pseudo_root_md):
        self.row = np.ascontiguousarray(row, dtype=np.uint32)
        self.node_id = np.ascontiguousarray(node_id, dtype=np.uint32)
        self.base_id = np.ascontiguousarray(base_id, dtype=np.uint32)
        self.md_row = np.ascontiguousarray(md_row, dtype=np.uint32)
        self.max_level = np.ascontiguousarray(max_level, dtype=np.uint8)
        self.main_root_row, self.main_root_md = int(main_root_row), int(main_root_md)
        self.pseudo_root_row, self.pseudo_root_md = int(pseudo_root_row), int(pseudo_root_md)
        n = self.row.size
        assert self.node_id.size == n and self.base_id.size == n and self.md_row.size == n and self.max_level.size == n

    def cstruct(self):
        return _ReplicaList(self.row.size, _p(self.row), _p(self.node_id), _p(self.base_id), _p(self.md_row), _p(self.max_level),
                            self.main_root_row, self.main_root_md, self.pseudo_root_row, self.pseudo_root_md)


def build_md(metric, storage_type, dim, codes, mags, md_bits, md_mags, rl, num_levels=9, neighbors_count=32, level0_neighbors_count=64,
             ef_construction=128, shortlist_size=64):
    """single-threaded index_embeddings over a ReplicaList -> (MdGraph, failed u8[n_nodes])"""
    L = _lib()
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    mags = np.ascontiguousarray(mags, dtype=np.float32)
    md_bits = np.ascontiguousarray(md_bits, dtype=np.int32)
    md_mags = np.ascontiguousarray(md_mags, dtype=np.float32)
    failed = np.zeros(rl.row.size, dtype=np.uint8)
    crl = rl.cstruct()
    h = L.orc_hnsw_build_md(int(metric), int(storage_type), dim, _p(codes), _p(mags), md_bits.shape[1], _p(md_bits), _p(md_mags),
                            C.byref(crl), num_levels, neighbors_count, level0_neighbors_count, ef_construction, shortlist_size,
                            _p(failed))
    g = L.orc_md_built_graph(h).contents
    L1 = num_levels + 1

    def arrs(tbl, per):
        tp = C.cast(tbl, C.POINTER(C.c_void_p))
        return [np.ctypeslib.as_array(C.cast(tp[lv], C.POINTER(C.c_uint32)), shape=(int(cnt[lv]) * per(lv),)).copy() for lv in range(L1)]
    cnt = np.ctypeslib.as_array(C.cast(g.g.cnt, C.POINTER(C.c_uint32)), shape=(L1,)).copy()
    one = lambda lv: 1
    node_row, child = arrs(g.g.node_row, one), arrs(g.g.child, one)
    adj = arrs(g.g.adj, lambda lv: level0_neighbors_count if lv == 0 else neighbors_count)
    node_id, node_md = arrs(g.node_id, one), arrs(g.node_md, one)
    fg = FlatGraph(metric, storage_type, dim, codes, mags, rl.main_root_row, num_levels, neighbors_count, level0_neighbors_count,
                   g.g.entry, node_row, adj, child)
    mg = MdGraph(fg, md_bits, md_mags, node_id, node_md, g.pseudo_entry)
    L.orc_md_built_free(h)
    return mg, failed
