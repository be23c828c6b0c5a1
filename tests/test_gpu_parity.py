"""GPU parity tests: every result comes back through the C ABI
(include/cosdata_b200.h) and is compared with the CPU oracle on the same seeded
inputs.  Bar: bit-exact ids and integer scores; f32 scores bit-identical too
(the kernels reproduce the reference's AVX2 reduction order), so comparisons use
array_equal on the raw bits, not a tolerance."""
import numpy as np
import pytest

import cosdata_b200 as cdb
import oracle as orc

pytestmark = pytest.mark.gpu

DIMS = [8, 31, 32, 33, 128, 768, 1024]
ST = cdb.StorageType
MK = cdb.DistanceMetricKind


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def edge_matrix(seed, n, dim):
    m = orc.synth_matrix(seed, n, dim).copy()
    edge = np.array([-1.0, 1.0, 0.0, -0.0, 0.99999994, -1.5, 1.5, 0.5, -0.5, 0.25], dtype=np.float32)
    m[0, : min(dim, edge.size)] = edge[: min(dim, edge.size)]
    if n > 1:
        m[1] = 0.0  # zero-norm row
    return m


# ------------------------------------------------------------------ quantize

@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("st", list(ST))
def test_quantize_batch_matches_oracle(st, dim):
    m = edge_matrix(100 + dim, 37, dim)
    lo, hi = (-0.8, 0.9) if st == ST.UnsignedByte else (-1.0, 1.0)
    codes, mags = cdb.ScalarQuantization().quantize_batch(m, st, (lo, hi))
    want_c, want_m = orc.quantize_batch(int(st), m, lo, hi)
    assert np.array_equal(codes, want_c)
    assert np.array_equal(bits(mags), bits(want_m))


# ------------------------------------------------------------------ pairwise distances

@pytest.mark.parametrize("dim", [8, 33, 128, 768])
@pytest.mark.parametrize("st", list(ST))
@pytest.mark.parametrize("metric", list(MK))
def test_distance_pairs_match_oracle(metric, st, dim):
    n = 24
    x = edge_matrix(200 + dim, n, dim)
    y = edge_matrix(300 + dim, n, dim)[::-1].copy()
    if st == ST.UnsignedByte:  # make large |a-b| so the i16 wrap of euclidean_distance_u8 is exercised
        x[2], y[2] = 1.0, -1.0
    xc, xm = orc.quantize_batch(int(st), x)
    yc, ym = orc.quantize_batch(int(st), y)
    got, status = cdb.DistanceMetric(metric).calculate_pairs(st, dim, xc, xm, yc, ym)
    for i in range(n):
        rc, v = orc.distance(int(metric), int(st), dim, xc[i], xm[i], yc[i], ym[i])
        assert status[i] == rc, (i, status[i], rc)
        if rc == 0:
            assert bits(got[i]) == bits(v), (i, got[i], v)


def test_distance_calculate_raises_like_the_trait():
    q = cdb.ScalarQuantization()
    v = orc.synth_matrix(1, 2, 64)
    a, b = q.quantize(v[0], ST.FullPrecisionFP), q.quantize(v[1], ST.FullPrecisionFP)
    got = cdb.DistanceMetric(MK.Cosine).calculate(a, b)
    rc, want = orc.distance(0, 5, 64, a.code, a.mag, b.code, b.mag)
    assert bits(got) == bits(want)
    with pytest.raises(cdb.DistanceError) as e:
        cdb.DistanceMetric(MK.DotProduct).calculate(a, b)           # dotproduct.rs:62
    assert e.value.status == cdb.Status.STORAGE_MISMATCH
    z = q.quantize(np.zeros(64, np.float32), ST.FullPrecisionFP)
    with pytest.raises(cdb.DistanceError) as e:
        cdb.DistanceMetric(MK.Cosine).calculate(a, z)               # cosine.rs:230-231
    assert e.value.status == cdb.Status.CALCULATION_ERROR
    with pytest.raises(cdb.DistanceError):
        cdb.DistanceMetric(MK.Cosine).calculate(a, q.quantize(v[1], ST.HalfPrecisionFP))   # cosine.rs:214


# ------------------------------------------------------------------ brute force, raw f32 (configs C1/C2 shape)

def _check_raw(corpus, queries, k, **kw):
    ix = cdb.DenseIndex(dim=corpus.shape[1], capacity=max(1, corpus.shape[0]))
    ix.append(corpus)
    ids, scores, counts, err = ix.batch_search(queries, k, cdb.SearchMode.BRUTE_RAW, **kw)
    want_ids, want_scores = orc.brute_topk_f32(corpus, queries, k)
    assert np.array_equal(ids, want_ids)
    assert np.array_equal(bits(scores), bits(want_scores))
    assert np.array_equal(counts, np.minimum(k, corpus.shape[0]) * np.ones(len(queries), np.uint32))
    ix.close()


def test_config1_100k_x_128_batch1_exact():
    # BASELINE.json configs[0]: brute-force cosine, 100k x 128 fp32, batch = 1
    corpus = orc.synth_matrix(0xC05DA7A + 1, 100_000, 128)
    q = orc.synth_matrix(0xC05DA7A + 101, 1, 128)
    _check_raw(corpus, q, 10)


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("nq", [1, 2, 3, 5, 8, 9, 17])
def test_brute_raw_small_shapes(dim, nq):
    corpus = orc.synth_matrix(400 + dim, 3001, dim)
    q = orc.synth_matrix(500 + dim + nq, nq, dim)
    _check_raw(corpus, q, 10)


@pytest.mark.parametrize("k", [1, 10, 100, 1000])
def test_brute_raw_k_sweep(k):
    corpus = orc.synth_matrix(600, 5000, 64)
    q = orc.synth_matrix(601, 4, 64)
    _check_raw(corpus, q, k)


def test_brute_raw_ties_short_corpus_and_zero_norm_rows():
    base = orc.synth_matrix(700, 40, 48)
    corpus = np.concatenate([base, base, base[:5]])      # exact duplicates -> ties broken by smaller id
    corpus[17] = 0.0                                     # 0/0 -> x86 "real indefinite" NaN sorts last
    q = base[:6].copy()
    _check_raw(corpus, q, 10)
    _check_raw(corpus[:3], q, 10)                        # n < k: padded with INVALID_ID
    _check_raw(corpus[:1], q[:1], 1)
    _check_raw(corpus, q, corpus.shape[0])               # k == n: every row incl. the NaN one, in order


def test_brute_raw_self_match_and_batch_independence():
    corpus = orc.synth_matrix(800, 20000, 96)
    ix = cdb.DenseIndex(dim=96, capacity=20000)
    ix.append(corpus)
    rows = np.array([0, 1, 777, 19999, 4096, 255, 256, 257, 12345], dtype=np.int64)
    ids, scores, _, _ = ix.batch_search(corpus[rows], 5)
    assert np.array_equal(ids[:, 0], rows.astype(np.uint32))
    assert np.allclose(scores[:, 0], 1.0, atol=2e-6)
    for i, r in enumerate(rows):                          # same query alone == same query inside a batch
        ids1, scores1, _, _ = ix.batch_search(corpus[r], 5)
        assert np.array_equal(ids1[0], ids[i]) and np.array_equal(bits(scores1[0]), bits(scores[i]))
    ix.close()


def test_synthetic_append_equals_host_rows_and_id_base():
    dim, n = 40, 3000
    ix = cdb.DenseIndex(dim=dim, capacity=n, id_base=1000)
    ix.append_synthetic(77, 1000)
    ix.append_synthetic(77, n - 1000)
    codes, mags = ix.read_codes(0, n)
    host = orc.synth_matrix(77, n, dim)
    assert np.array_equal(codes.view(np.float32).reshape(n, dim), host)
    q = orc.synth_matrix(78, 3, dim)
    ids, scores, _, _ = ix.batch_search(q, 7)
    want_ids, want_scores = orc.brute_topk_f32(host, q, 7)
    assert np.array_equal(ids, want_ids + 1000)
    assert np.array_equal(bits(scores), bits(want_scores))
    ix.close()


# ------------------------------------------------------------------ brute force over quantized codes (config C4 shape)

@pytest.mark.parametrize("metric", [MK.Cosine, MK.DotProduct, MK.Euclidean, MK.Hamming])
@pytest.mark.parametrize("st", list(ST))
def test_brute_codes_match_oracle(st, metric):
    dim, n, nq, k = 100, 2500, 5, 10
    corpus = orc.synth_matrix(900 + int(st), n, dim)
    q = orc.synth_matrix(950 + int(st), nq, dim)
    ix = cdb.DenseIndex(dim=dim, storage_type=st, metric=metric, capacity=n)
    ix.append(corpus)
    codes, mags = orc.quantize_batch(int(st), corpus)
    got_codes, got_mags = ix.read_codes(0, n)
    assert np.array_equal(got_codes, codes) and np.array_equal(bits(got_mags), bits(mags))
    qc, qm = orc.quantize_batch(int(st), q)
    rc, want_ids, want_scores, want_err = orc.brute_topk_codes(int(metric), int(st), dim, codes, mags, qc, qm, k)
    if rc != 0:                       # StorageMismatch / unimplemented arm: the whole search is an Err
        with pytest.raises(cdb.CosdataError) as e:
            ix.batch_search(q, k, cdb.SearchMode.BRUTE_CODES)
        assert int(e.value.status) == rc
    else:
        ids, scores, counts, err = ix.batch_search(q, k, cdb.SearchMode.BRUTE_CODES)
        assert np.array_equal(ids, want_ids)
        assert np.array_equal(bits(scores), bits(want_scores))
        assert np.array_equal(err, want_err)
    ix.close()


def test_quaternary_1024_dot_and_cosine_with_zero_norm_row():
    # config C4 shape at reduced N: SubByte(2), D = 1024
    dim, n, nq, k = 1024, 4000, 9, 10
    corpus = orc.synth_matrix(0xC05DA7A + 4, n, dim).copy()
    corpus[123] = 0.0
    q = orc.synth_matrix(0xC05DA7A + 104, nq, dim)
    codes, mags = orc.quantize_batch(2, corpus)
    qc, qm = orc.quantize_batch(2, q)
    for metric in (MK.DotProduct, MK.Cosine):
        ix = cdb.DenseIndex(dim=dim, storage_type=ST.SubByte2, metric=metric, capacity=n)
        ix.append(corpus)
        ids, scores, counts, err = ix.batch_search(q, k, cdb.SearchMode.BRUTE_CODES)
        rc, want_ids, want_scores, want_err = orc.brute_topk_codes(int(metric), 2, dim, codes, mags, qc, qm, k)
        assert rc == 0
        assert np.array_equal(ids, want_ids) and np.array_equal(bits(scores), bits(want_scores))
        assert np.array_equal(err, want_err)
        assert err.any() == (metric == MK.Cosine)         # cosine.rs:230-231 -> Err for every query
        ix.close()


# ------------------------------------------------------------------ S2 / S3

@pytest.mark.parametrize("st", list(ST))
def test_score_ids_matches_oracle(st):
    dim, n = 72, 600
    corpus = orc.synth_matrix(1000 + int(st), n, dim)
    q = orc.synth_matrix(1100 + int(st), 1, dim)[0]
    ix = cdb.DenseIndex(dim=dim, storage_type=st, metric=MK.Cosine, capacity=n)
    ix.append(corpus)
    ids = np.array([0, 599, 17, 17, 300, 42], dtype=np.uint32)
    got, status = ix.score_ids(q, ids)
    codes, mags = orc.quantize_batch(int(st), corpus)
    qc, qm = orc.quantize(int(st), q)
    for i, r in enumerate(ids):
        rc, v = orc.distance(0, int(st), dim, qc, qm, codes[r], mags[r])
        assert status[i] == rc and bits(got[i]) == bits(v)
    ix.close()


@pytest.mark.parametrize("st", [ST.FullPrecisionFP, ST.UnsignedByte, ST.HalfPrecisionFP, ST.SubByte2, ST.BFloat16])
def test_rerank_matches_oracle(st):
    dim, n = 768, 400
    corpus = orc.synth_matrix(1200, n, dim)
    q = orc.synth_matrix(1201, 1, dim)[0]
    ix = cdb.DenseIndex(dim=dim, storage_type=st, metric=MK.Cosine, capacity=n, keep_raw_f32=True)
    ix.append(corpus)
    cand = np.random.default_rng(5).permutation(n)[:50].astype(np.uint32)   # 5*k candidates, k = 10
    ids, scores, cnt = ix.rerank(q, cand, 10)
    want_ids, want_scores = orc.rerank_f32(corpus, q, cand, 10)
    assert cnt == 10
    assert np.array_equal(ids, want_ids) and np.array_equal(bits(scores), bits(want_scores))
    ix.close()


# ------------------------------------------------------------------ multi-shard merge on one GPU

def test_merge_topk_device_over_three_shards_equals_single_index():
    import torch
    from cosdata_b200.sharding import cuda_merge_fn, gather_and_merge, shard_range
    n, dim, nq, k, world = 30011, 96, 21, 10, 3
    corpus = orc.synth_matrix(3000, n, dim)
    q = orc.synth_matrix(3001, nq, dim)
    per_ids, per_scores = [], []
    for r in range(world):
        row0, nloc = shard_range(n, world, r)
        ix = cdb.DenseIndex(dim=dim, capacity=nloc, id_base=row0)
        ix.append_synthetic(3000, nloc, first_row=row0)
        ids, scores, _, _ = ix.batch_search(q, k)
        per_ids.append(torch.from_numpy(ids.view(np.int32)).cuda())
        per_scores.append(torch.from_numpy(scores).cuda())
        ix.close()
    g_ids, g_scores = torch.stack(per_ids).contiguous(), torch.stack(per_scores).contiguous()
    import cosdata_b200._lib as L
    merge = cuda_merge_fn(L.load(), 0, 0, None)
    m_ids, m_scores = merge(g_ids, g_scores)
    torch.cuda.synchronize()
    want_ids, want_scores = orc.brute_topk_f32(corpus, q, k)
    assert np.array_equal(m_ids.cpu().numpy().view(np.uint32), want_ids)
    assert np.array_equal(bits(m_scores.cpu().numpy()), bits(want_scores))


# ------------------------------------------------------------------ empty / degenerate inputs

def test_empty_index_empty_batch_and_k_larger_than_n():
    ix = cdb.DenseIndex(dim=24, capacity=8)
    q = orc.synth_matrix(5000, 3, 24)
    ids, scores, counts, err = ix.batch_search(q, 5)                    # empty index: nothing to return
    assert np.all(ids == cdb.INVALID_ID) and np.all(counts == 0) and np.all(scores == 0)
    ids, scores, counts, err = ix.batch_search(np.zeros((0, 24), np.float32), 5)   # empty batch
    assert ids.shape == (0, 5)
    ix.append(orc.synth_matrix(5001, 2, 24))
    ids, scores, counts, err = ix.batch_search(q, 5)                    # k > n
    want_ids, want_scores = orc.brute_topk_f32(orc.synth_matrix(5001, 2, 24), q, 5)
    assert np.array_equal(ids, want_ids) and np.array_equal(bits(scores), bits(want_scores)) and np.all(counts == 2)
    with pytest.raises(cdb.CosdataError):
        ix.append(orc.synth_matrix(5002, 7, 24))                        # exceeds capacity: refused, nothing written
    assert len(ix) == 2
    with pytest.raises(cdb.CosdataError):
        ix.batch_search(q, 0)                                           # k must be >= 1
    out, st = ix.score_ids(q[0], np.array([0, 1, 7], dtype=np.uint32))  # out-of-range id is reported per id
    assert st.tolist() == [0, 0, int(cdb.Status.INVALID_PARAMS)]
    ix.close()


def test_append_from_device_memory_and_timing_history():
    import torch
    dim, n = 56, 4000
    host = orc.synth_matrix(5100, n, dim)
    for st in (ST.UnsignedByte, ST.SubByte3, ST.HalfPrecisionFP):
        ix = cdb.DenseIndex(dim=dim, storage_type=st, metric=MK.Cosine, capacity=n, keep_raw_f32=True)
        d = torch.from_numpy(host).cuda()
        ix.append_device(d.data_ptr(), 1500)
        ix.append_device(d[1500:].contiguous().data_ptr(), n - 1500)
        codes, mags = ix.read_codes(0, n)
        want_c, want_m = orc.quantize_batch(int(st), host)
        assert np.array_equal(codes, want_c) and np.array_equal(bits(mags), bits(want_m))
        q = orc.synth_matrix(5101, 3, dim)
        ids, scores, _, _ = ix.batch_search(q, 5)                         # raw rows were kept too
        want_ids, want_scores = orc.brute_topk_f32(host, q, 5)
        assert np.array_equal(ids, want_ids) and np.array_equal(bits(scores), bits(want_scores))
        hist = ix.scan_ms_history(4)
        assert hist.size >= 1 and (hist > 0).all()
        ix.close()


def test_searches_on_one_handle_are_reentrant():
    """the ABI is re-entrant per handle.  Eight host threads search one index at the same time (every call leases its
    own scratch set: buffers, stream, events); each must get exactly what a serial call returns."""
    import threading
    dim, n, k = 96, 60000, 10
    corpus = orc.synth_matrix(7100, n, dim)
    ix = cdb.DenseIndex(dim=dim, capacity=n)
    ix.append(corpus)
    batches = [orc.synth_matrix(7200 + t, 5 + 17 * t, dim) for t in range(8)]        # 5 .. 124 queries: exact and prefilter paths
    want = [ix.batch_search(b, k) for b in batches]
    got, errors = [None] * 8, []

    def work(t):
        try:
            for _ in range(6):
                got[t] = ix.batch_search(batches[t], k)
        except Exception as e:                                                       # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(8):
        assert np.array_equal(got[t][0], want[t][0]) and np.array_equal(got[t][1].view(np.uint32), want[t][1].view(np.uint32))
        o_ids, o_scores = orc.brute_topk_f32(corpus, batches[t], k)
        assert np.array_equal(got[t][0], o_ids) and np.array_equal(got[t][1].view(np.uint32), o_scores.view(np.uint32))
    ix.close()
