"""tcgen05 prefilter + exact re-rank must return exactly what the exact scan / CPU oracle
return (bit-identical ids and scores), for every shape, and must actually have run."""
import numpy as np
import pytest

import cosdata_b200 as cdb
import oracle as orc

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def run_case(corpus, queries, k, expect_fallback=False, **kw):
    n, dim = corpus.shape
    ix = cdb.DenseIndex(dim=dim, capacity=n)
    ix.append(corpus)
    assert ix.stats()["has_shadow"]
    ids, scores, counts, err = ix.batch_search(queries, k, **kw)
    st = ix.stats()
    assert st["tensor_searches"] == 1, "the tcgen05 path did not run"
    assert (st["fallbacks"] == 1) == expect_fallback, st
    want_ids, want_scores = orc.brute_topk_f32(corpus, queries, k)
    assert np.array_equal(ids, want_ids)
    assert np.array_equal(bits(scores), bits(want_scores))
    assert np.array_equal(counts, np.full(len(queries), min(k, n), np.uint32))
    # and the pure FFMA scan agrees too
    ids2, scores2, _, _ = ix.batch_search(queries, k, exact_only=True)
    assert np.array_equal(ids2, ids) and np.array_equal(bits(scores2), bits(scores))
    cand = ix.last_candidate_counts(len(queries))
    ix.close()
    return cand


@pytest.mark.parametrize("dim,nq", [(768, 64), (768, 130), (128, 5), (100, 16), (33, 9), (1024, 256), (64, 128), (72, 4)])
def test_prefilter_equals_exact_on_uniform_data(dim, nq):
    n = 40000
    corpus = orc.synth_matrix(2000 + dim, n, dim)
    q = orc.synth_matrix(2100 + dim, nq, dim)
    cand = run_case(corpus, q, 10)
    assert cand.max() < 4096 and cand.min() >= 10


@pytest.mark.parametrize("k", [1, 10, 33, 64])
def test_prefilter_k_sweep(k):
    corpus = orc.synth_matrix(2200, 30000, 256)
    q = orc.synth_matrix(2201, 33, 256)
    run_case(corpus, q, k)


def test_prefilter_on_clustered_near_duplicates():
    # scores inside a cluster differ by far less than the prefilter's error bound:
    # the candidate lists get long but the final answer must still be exact
    rng = np.random.default_rng(3)
    dim, n = 256, 32768
    centres = rng.normal(size=(64, dim)).astype(np.float32)
    assign = rng.integers(0, 64, n)
    corpus = (centres[assign] + 1e-3 * rng.normal(size=(n, dim))).astype(np.float32)
    corpus[5000] = corpus[6000]                      # exact duplicate rows -> tie broken by id
    q = (centres[:24] + 1e-3 * rng.normal(size=(24, dim))).astype(np.float32)
    cand = run_case(corpus, q, 10)
    assert cand.max() > 100                          # the filter really had to keep whole clusters


def test_prefilter_overflow_falls_back_to_exact_scan():
    corpus = orc.synth_matrix(2300, 20000, 128)
    q = orc.synth_matrix(2301, 8, 128)
    run_case(corpus, q, 10, expect_fallback=True, prefilter_k=4)   # 4 candidate slots < k


def test_zero_norm_rows_and_queries():
    corpus = orc.synth_matrix(2400, 20000, 96).copy()
    corpus[[0, 77, 19999]] = 0.0
    q = orc.synth_matrix(2401, 12, 96).copy()
    run_case(corpus, q, 10)                                        # NaN rows sort last, never in the top-k
    q[3] = 0.0                                                      # a zero query scores NaN everywhere
    run_case(corpus, q, 10, expect_fallback=True)


def test_sharded_id_base_and_device_api_agree():
    import torch
    dim, n, nq, k = 384, 50000, 48, 10
    corpus = orc.synth_matrix(2500, n, dim)
    q = orc.synth_matrix(2501, nq, dim)
    ix = cdb.DenseIndex(dim=dim, capacity=n, id_base=7_000_000)
    ix.append_synthetic(2500, n)
    dq = torch.from_numpy(q).cuda()
    ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    scores = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ix.batch_search_device(dq.data_ptr(), nq, k, ids.data_ptr(), scores.data_ptr(), stream_ptr=s.cuda_stream)
    s.synchronize()
    want_ids, want_scores = orc.brute_topk_f32(corpus, q, k)
    assert np.array_equal(ids.cpu().numpy().view(np.uint32), want_ids + 7_000_000)
    assert np.array_equal(bits(scores.cpu().numpy()), bits(want_scores))
    assert ix.stats()["tensor_searches"] == 1 and ix.stats()["fallbacks"] == 0
    ix.close()
