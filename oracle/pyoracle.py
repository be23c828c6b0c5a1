"""ctypes binding of oracle/libcosdata_oracle.so (TEST INFRASTRUCTURE ONLY).

Builds the shared object on first use with oracle/Makefile (gcc, AVX2+FMA).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcosdata_oracle.so")

OK, STORAGE_MISMATCH, CALCULATION_ERROR, INVALID, UNIMPLEMENTED = 0, 1, 2, 3, 6
ST_U8, ST_SUB1, ST_SUB2, ST_SUB3, ST_F16, ST_F32 = range(6)
ST_BF16 = 6   # labelled extension (bfloat16); not a reference StorageType
METRIC_COSINE, METRIC_EUCLIDEAN, METRIC_HAMMING, METRIC_DOT = range(4)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


def _p(a, ty=None):
    return a.ctypes.data_as(C.c_void_p)


def _declare(L):
    f32p, u8p, u16p, u32p, vp = (C.c_void_p,) * 5
    sz = C.c_size_t
    L.orc_synth_value.restype = C.c_float
    L.orc_synth_value.argtypes = [C.c_uint64, C.c_uint64]
    L.orc_synth_fill.argtypes = [C.c_uint64, C.c_uint64, sz, f32p]
    L.orc_f32_to_f16.restype = C.c_uint16
    L.orc_f32_to_f16.argtypes = [C.c_float]
    L.orc_f16_to_f32.restype = C.c_float
    L.orc_f16_to_f32.argtypes = [C.c_uint16]
    for name in ("orc_dot_u8_scalar", "orc_dot_u8_avx2"):
        getattr(L, name).restype = C.c_uint64
        getattr(L, name).argtypes = [u8p, u8p, sz]
    for name in ("orc_dot_f16", "orc_dot_f32_scalar", "orc_dot_f32_simd",
                 "orc_dot_binary_scalar", "orc_dot_binary_avx2",
                 "orc_dot_quaternary_scalar", "orc_dot_quaternary_avx2",
                 "orc_dot_octal_scalar", "orc_dot_octal_avx2"):
        getattr(L, name).restype = C.c_float
        getattr(L, name).argtypes = [vp, vp, sz]
    L.orc_count_ones_256.restype = C.c_uint64
    L.orc_count_ones_256.argtypes = [u8p]
    L.orc_code_bytes.restype = sz
    L.orc_code_bytes.argtypes = [C.c_int, sz]
    L.orc_quantize.restype = C.c_int
    L.orc_quantize.argtypes = [C.c_int, C.c_float, C.c_float, f32p, sz, vp, f32p]
    L.orc_distance.restype = C.c_int
    L.orc_distance.argtypes = [C.c_int, C.c_int, sz, vp, C.c_float, vp, C.c_float, f32p]
    L.orc_order_key.restype = C.c_uint32
    L.orc_order_key.argtypes = [C.c_int, C.c_float]
    L.orc_mag_f32.restype = C.c_float
    L.orc_mag_f32.argtypes = [f32p, sz]
    L.orc_rerank_cosine.restype = C.c_float
    L.orc_rerank_cosine.argtypes = [f32p, C.c_float, f32p, sz]
    L.orc_brute_topk_f32.restype = C.c_int
    L.orc_brute_topk_f32.argtypes = [f32p, sz, sz, f32p, sz, sz, C.c_int, u32p, f32p]
    L.orc_brute_topk_codes.restype = C.c_int
    L.orc_brute_topk_codes.argtypes = [C.c_int, C.c_int, sz, vp, f32p, sz, vp, f32p, sz, sz,
                                       C.c_int, u32p, f32p, u8p]
    L.orc_sample_counts.restype = None
    L.orc_sample_counts.argtypes = [f32p, sz, vp]
    L.orc_values_range.restype = None
    L.orc_values_range.argtypes = [vp, C.c_uint64, C.c_float, f32p]
    L.orc_rerank_f32.restype = C.c_int
    L.orc_rerank_f32.argtypes = [f32p, sz, f32p, u32p, sz, sz, u32p, f32p]


# ----------------------------------------------------------------- wrappers

def synth(seed, first_idx, n):
    out = np.empty(n, dtype=np.float32)
    lib().orc_synth_fill(seed, first_idx, n, _p(out))
    return out


def synth_matrix(seed, n, dim, first_row=0):
    return synth(seed, first_row * dim, n * dim).reshape(n, dim)


def f32_to_f16_bits(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    L = lib()
    return np.array([L.orc_f32_to_f16(float(v)) for v in x.ravel()], dtype=np.uint16).reshape(x.shape)


def code_bytes(st, dim):
    return lib().orc_code_bytes(st, dim)


def quantize(st, v, lo=-1.0, hi=1.0):
    """Quantization::quantize for one vector -> (code bytes as uint8 array, mag)."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    code = np.zeros(code_bytes(st, v.size), dtype=np.uint8)
    mag = np.zeros(1, dtype=np.float32)
    rc = lib().orc_quantize(st, lo, hi, _p(v), v.size, _p(code), _p(mag))
    assert rc == OK
    return code, np.float32(mag[0])


def quantize_batch(st, m, lo=-1.0, hi=1.0):
    m = np.ascontiguousarray(m, dtype=np.float32)
    n, dim = m.shape
    cb = code_bytes(st, dim)
    codes = np.zeros((n, cb), dtype=np.uint8)
    mags = np.zeros(n, dtype=np.float32)
    for i in range(n):
        c, g = quantize(st, m[i], lo, hi)
        codes[i] = c
        mags[i] = g
    return codes, mags


def distance(metric, st, dim, x_code, x_mag, y_code, y_mag):
    """DistanceMetric::calculate, (Base,Base) arm -> (status, f32 value)."""
    x_code = np.ascontiguousarray(x_code)
    y_code = np.ascontiguousarray(y_code)
    out = np.zeros(1, dtype=np.float32)
    rc = lib().orc_distance(metric, st, dim, _p(x_code), float(x_mag), _p(y_code), float(y_mag), _p(out))
    return rc, np.float32(out[0])


def order_key(metric, value):
    return lib().orc_order_key(metric, float(value))


def mag_f32(v):
    v = np.ascontiguousarray(v, dtype=np.float32)
    return np.float32(lib().orc_mag_f32(_p(v), v.size))


def dot_f32_simd(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return np.float32(lib().orc_dot_f32_simd(_p(a), _p(b), a.size))


def brute_topk_f32(corpus, queries, k, threads=None):
    corpus = np.ascontiguousarray(corpus, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    n, dim = corpus.shape
    nq = queries.shape[0]
    ids = np.zeros((nq, k), dtype=np.uint32)
    scores = np.zeros((nq, k), dtype=np.float32)
    threads = threads or os.cpu_count() or 1
    rc = lib().orc_brute_topk_f32(_p(corpus), n, dim, _p(queries), nq, k, threads, _p(ids), _p(scores))
    assert rc == OK
    return ids, scores


def brute_topk_codes(metric, st, dim, codes, mags, qcodes, qmags, k, threads=None):
    codes = np.ascontiguousarray(codes)
    qcodes = np.ascontiguousarray(qcodes)
    mags = np.ascontiguousarray(mags, dtype=np.float32)
    qmags = np.ascontiguousarray(qmags, dtype=np.float32)
    n, nq = codes.shape[0], qcodes.shape[0]
    ids = np.zeros((nq, k), dtype=np.uint32)
    scores = np.zeros((nq, k), dtype=np.float32)
    err = np.zeros(nq, dtype=np.uint8)
    threads = threads or os.cpu_count() or 1
    rc = lib().orc_brute_topk_codes(metric, st, dim, _p(codes), _p(mags), n, _p(qcodes), _p(qmags),
                                    nq, k, threads, _p(ids), _p(scores), _p(err))
    return rc, ids, scores, err


def rerank_f32(corpus, q, cand, k):
    corpus = np.ascontiguousarray(corpus, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    cand = np.ascontiguousarray(cand, dtype=np.uint32)
    ids = np.zeros(k, dtype=np.uint32)
    scores = np.zeros(k, dtype=np.float32)
    rc = lib().orc_rerank_f32(_p(corpus), corpus.shape[1], _p(q), _p(cand), cand.size, k, _p(ids), _p(scores))
    assert rc == OK
    return ids, scores


def sample_values_range(vectors, clamp_margin_percent=1.0, prior_counts=None, prior_values=0):
    """sample_embedding over every vector + finalize_sampling -> (counts u64[14], (range_start, range_end))"""
    v = np.ascontiguousarray(vectors, dtype=np.float32)
    counts = np.zeros(14, dtype=np.uint64) if prior_counts is None else np.array(prior_counts, dtype=np.uint64)
    lib().orc_sample_counts(_p(v), v.size, _p(counts))
    rng = np.zeros(2, dtype=np.float32)
    lib().orc_values_range(_p(counts), v.size + prior_values, clamp_margin_percent, _p(rng))
    return counts, (np.float32(rng[0]), np.float32(rng[1]))
