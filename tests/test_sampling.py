"""`quantization: auto` value-range sampling (HNSWIndex::sample_embedding + finalize_sampling,
src/indexes/hnsw/mod.rs:202-351): oracle pinned against a numpy restatement on CPU; CUDA reduction == oracle on GPU."""
import numpy as np
import pytest

import oracle as orc

T = np.array([0.025, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5], dtype=np.float32)


def numpy_counts(v):
    v = np.asarray(v, dtype=np.float32).reshape(-1)
    return np.array([(v > t).sum() for t in T] + [(v < -t).sum() for t in T], dtype=np.uint64)


def numpy_range(counts, n_values, clamp):
    vc = np.float32(n_values)
    pct = (counts.astype(np.float32) / vc) * np.float32(100.0)
    hi = next((T[i] for i in range(7) if pct[i] <= np.float32(clamp)), np.float32(1.0))
    lo = next((-T[i] for i in range(7) if pct[7 + i] <= np.float32(clamp)), np.float32(-1.0))
    return np.float32(lo), np.float32(hi)


def cases():
    base = orc.synth_matrix(77, 300, 37)
    rng = np.random.default_rng(3)
    yield "uniform", base, 1.0                                      # -> (-1, 1)
    for scale in (0.02, 0.04, 0.09, 0.19, 0.29, 0.39, 0.49, 0.6):
        yield f"scaled{scale}", (base * np.float32(scale)).astype(np.float32), 1.0
    yield "gauss", rng.normal(scale=0.08, size=(500, 64)).astype(np.float32), 1.0
    yield "gauss_tight_margin", rng.normal(scale=0.08, size=(500, 64)).astype(np.float32), 0.01
    yield "asymmetric", np.abs(base * np.float32(0.3)).astype(np.float32), 0.5
    edge = np.tile(np.concatenate([T, -T, np.nextafter(T, np.float32(1)), np.nextafter(-T, np.float32(-1)),
                                   np.array([np.nan, np.inf, -np.inf, 0.0, -0.0], dtype=np.float32)]), 9)[:297]
    yield "thresholds_exact", edge.reshape(9, 33), 50.0             # v > t is strict: values equal to t do not count
    yield "one_value", np.array([[0.3]], dtype=np.float32), 1.0
    yield "odd_tail", base.reshape(-1)[:1001].reshape(7, 143), 1.0


@pytest.mark.parametrize("name,vecs,clamp", list(cases()), ids=[c[0] for c in cases()])
def test_oracle_sampling_matches_numpy_restatement(name, vecs, clamp):
    counts, rng = orc.sample_values_range(vecs, clamp)
    want = numpy_counts(vecs)
    assert np.array_equal(counts, want)
    assert rng == numpy_range(want, vecs.size, clamp)


def test_oracle_sampling_expected_ranges():
    base = orc.synth_matrix(77, 300, 37)
    assert orc.sample_values_range(base)[1] == (np.float32(-1.0), np.float32(1.0))
    assert orc.sample_values_range(base * np.float32(0.19))[1] == (np.float32(-0.2), np.float32(0.2))
    assert orc.sample_values_range(np.abs(base) * np.float32(0.04))[1] == (np.float32(-0.025), np.float32(0.05))
    # no embeddings: 0/0 = NaN compares false everywhere -> full range (mod.rs:273-277 with values_count = 0)
    assert orc.sample_values_range(np.zeros((0, 8), dtype=np.float32))[1] == (np.float32(-1.0), np.float32(1.0))
    # sampling in two batches == one batch (the reference accumulates atomics per embedding)
    c1, _ = orc.sample_values_range(base[:100] * np.float32(0.3))
    c2, r2 = orc.sample_values_range(base[100:] * np.float32(0.3), prior_counts=c1, prior_values=100 * 37)
    c, r = orc.sample_values_range(base * np.float32(0.3))
    assert np.array_equal(c2, c) and r2 == r


@pytest.mark.gpu
@pytest.mark.parametrize("name,vecs,clamp", list(cases()), ids=[c[0] for c in cases()])
def test_gpu_sampling_matches_oracle(name, vecs, clamp):
    import cosdata_b200 as cdb
    counts, rng = cdb.sample_values_range(vecs, clamp)
    want_counts, want_rng = orc.sample_values_range(vecs, clamp)
    assert np.array_equal(counts, want_counts) and rng == want_rng


@pytest.mark.gpu
def test_gpu_sampling_large_device_resident_and_batched():
    import torch
    import cosdata_b200 as cdb
    n, dim = 40000, 768                                              # 30.7M values: many CTAs, float4 body
    host = (orc.synth_matrix(0xC05DA7A, n, dim) * np.float32(0.27)).astype(np.float32)
    want_counts, want_rng = orc.sample_values_range(host)
    d = torch.from_numpy(host).cuda()
    counts, rng = cdb.sample_values_range_device(d.data_ptr(), n, dim)
    assert np.array_equal(counts, want_counts) and rng == want_rng
    flat = d.reshape(-1)[1:1 + (n - 1) * dim]                       # misaligned start (4-byte offset): head/tail path
    counts, rng = cdb.sample_values_range_device(flat.data_ptr(), n - 1, dim)
    w2, r2 = orc.sample_values_range(host.reshape(-1)[1:1 + (n - 1) * dim])
    assert np.array_equal(counts, w2) and rng == r2
    c1, _ = cdb.sample_values_range(host[:1000])
    c2, r2 = cdb.sample_values_range(host[1000:], prior_counts=c1, prior_values=1000 * dim)
    assert np.array_equal(c2, want_counts) and r2 == want_rng
    # empty sample -> full range, no launch
    assert cdb.sample_values_range(np.zeros((0, dim), dtype=np.float32))[1] == (np.float32(-1.0), np.float32(1.0))
