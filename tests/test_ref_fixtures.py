"""Reference-pinned parity: fixtures WRITTEN BY THE RUST REFERENCE (tools/ref_fixtures/b200_fixtures.rs, run inside the
cosdata crate on a machine with cargo) are compared with the oracle, the committed golden vectors and the CUDA path.

The fixture files live in tests/golden/ref/ and are absent until someone with a Rust toolchain generates them (this image
has none: DESIGN.md section 3) -- the reference tests skip then.  The pipeline itself is always exercised: an
oracle-written file in the same container format goes through exactly the same checks."""
import os

import numpy as np
import pytest

import oracle as orc
from tests.ref_fixture_io import load_cdbf, write_cdbf

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "golden", "ref")
REF_FILE = os.path.join(REF_DIR, "hotpath_ref_v1.cdbf")
DIMS = (8, 31, 32, 33, 128, 768, 1024)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def oracle_side(dim):
    """what the oracle computes for the inputs of make_golden.corpus_for(dim), keyed like the reference's file"""
    from tests.golden.make_golden import corpus_for
    m, q = corpus_for(dim)
    out = {f"d{dim}/corpus": m, f"d{dim}/queries": q}
    for st in range(6):
        codes, mags = orc.quantize_batch(st, m)
        qcodes, qmags = orc.quantize_batch(st, q)
        out[f"d{dim}/st{st}/codes"], out[f"d{dim}/st{st}/mag"] = codes, mags
        out[f"d{dim}/st{st}/qcodes"], out[f"d{dim}/st{st}/qmag"] = qcodes, qmags
        for metric in range(4):
            val = np.zeros((len(q), len(m)), np.float32)
            status = np.zeros(val.shape, np.int8)
            for i in range(len(q)):
                for j in range(len(m)):
                    rc, d = orc.distance(metric, st, dim, qcodes[i], qmags[i], codes[j], mags[j])
                    status[i, j], val[i, j] = rc, (d if rc == 0 else 0.0)
            out[f"d{dim}/st{st}/m{metric}/value"], out[f"d{dim}/st{st}/m{metric}/status"] = val, status
    codes, mags = orc.quantize_batch(0, m, -0.5, 0.75)
    out[f"d{dim}/st0_range/codes"], out[f"d{dim}/st0_range/mag"] = codes, mags
    return out


def compare(ref, mine, what):
    """every key of `mine` that the reference file holds must agree bit for bit (values only where the status is Ok)"""
    checked = 0
    for key, got in mine.items():
        if key not in ref:
            continue
        want = ref[key]
        assert want.shape == got.shape, (what, key, want.shape, got.shape)
        if key.endswith("/value"):
            ok = ref[key[:-5] + "status"] == 0
            assert np.array_equal(bits(want)[ok], bits(got)[ok]), (what, key)
        elif got.dtype == np.float32:
            assert np.array_equal(bits(want), bits(got)), (what, key)
        else:
            assert np.array_equal(want, got), (what, key)
        checked += 1
    return checked


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_FILE):
        pytest.skip("no reference-written fixtures (tools/ref_fixtures/README.md): parity stays pinned by the oracle only")
    return load_cdbf(REF_FILE)


def test_pipeline_with_an_oracle_written_file(tmp_path):
    # same container, same checks -- proves the consumer side works before a reference-written file exists
    arrays = {}
    for dim in (8, 33):
        arrays.update(oracle_side(dim))
    write_cdbf(tmp_path / "self.cdbf", arrays)
    back = load_cdbf(tmp_path / "self.cdbf")
    assert set(back) == set(arrays)
    assert compare(back, oracle_side(8), "self") > 40
    # and the committed golden vectors use the same inputs and answers
    z = np.load(os.path.join(HERE, "golden", "hotpath_v1.npz"))
    for st in range(6):
        assert np.array_equal(z[f"d8/st{st}/codes"], back[f"d8/st{st}/codes"])
        assert np.array_equal(z[f"d8/st{st}/mag_bits"], bits(back[f"d8/st{st}/mag"]))
        for metric in range(4):
            assert np.array_equal(z[f"d8/st{st}/m{metric}/status"], back[f"d8/st{st}/m{metric}/status"])
            assert np.array_equal(z[f"d8/st{st}/m{metric}/value_bits"], bits(back[f"d8/st{st}/m{metric}/value"]))


@pytest.mark.parametrize("dim", DIMS)
def test_reference_quantize_and_distance_vs_oracle(ref, dim):
    assert compare(ref, oracle_side(dim), "oracle") >= 50
    # raw f32 dot products (dot_product_f32 AVX2 + FMA path) and the finalize_ann_results formula built on them
    m, q = ref[f"d{dim}/corpus"], ref[f"d{dim}/queries"]
    ids, scores = orc.brute_topk_f32(m, q, len(m))
    dots, mags, qmags = ref[f"d{dim}/f32_dot"], ref[f"d{dim}/f32_mag"], ref[f"d{dim}/f32_qmag"]
    with np.errstate(all="ignore"):
        cs = (dots / (qmags[:, None] * mags[None, :])).astype(np.float32)
    for i in range(len(q)):
        got = cs[i][ids[i]]
        same = (bits(got) == bits(scores[i])) | (np.isnan(got) & np.isnan(scores[i]))
        assert same.all(), (dim, i)


def test_reference_metadata_arms_vs_oracle(ref):
    z = np.load(os.path.join(HERE, "golden", "hotpath_v1.npz"))
    want, got = ref["metadata/arm_table"], z["metadata/arm_table"]
    assert np.array_equal(want[..., 0], got[..., 0])
    ok = want[..., 0] == 0
    assert np.array_equal(want[..., 1][ok], got[..., 1][ok])


def test_reference_prop_file_and_itoe_readers(ref):
    import cosdata_b200 as cdb
    from tests.golden.make_golden import corpus_for
    m, _ = corpus_for(32)
    codes, mags = orc.quantize_batch(0, m)
    n, st, elems, cb = cdb.prop_file_scan(os.path.join(REF_DIR, "prop.data"))
    assert (n, st, elems, cb) == (len(m), 0, 32, 32)
    pf = cdb.prop_file_load(os.path.join(REF_DIR, "prop.data"))
    assert np.array_equal(pf["ids"], np.arange(len(m), dtype=np.uint32) * 3 + 1)
    assert np.array_equal(pf["codes"], codes) and np.array_equal(bits(pf["mags"]), bits(mags))
    locs = ref["prop/locations"]                                     # (offset, length) of every record the reference wrote
    mine = {(int(o), int(l)) for o, l in zip(pf["offsets"], pf["lengths"])}
    md = cdb.prop_file_load_metadata(os.path.join(REF_DIR, "prop.data"))
    mine |= {(int(o), int(l)) for o, l in zip(md["offsets"], md["lengths"])}
    assert mine == {(int(o), int(l)) for o, l in locs}
    assert md["mbits"].shape[1] == 5 and np.array_equal(md["replica_ids"], 1000 + np.arange(len(m))[1::4])
    entries, dim, max_id = cdb.itoe_scan(REF_DIR)
    assert dim == 32 and entries == len(m) - 1                       # one key deleted in version 1
    iids, vecs = cdb.itoe_load(REF_DIR)
    assert 4 * 3 + 1 not in iids.tolist()
    for iid, v in zip(iids, vecs):
        src = m[5] if iid == 2 * 3 + 1 else m[(iid - 1) // 3]        # key 7 was overwritten in version 1
        assert np.array_equal(bits(v), bits(src))


@pytest.mark.gpu
@pytest.mark.parametrize("dim", DIMS)
def test_reference_quantize_and_distance_vs_cuda(ref, dim):
    import cosdata_b200 as cdb
    m, q = ref[f"d{dim}/corpus"], ref[f"d{dim}/queries"]
    sq = cdb.ScalarQuantization()
    mine = {}
    for st in range(6):
        codes, mags = sq.quantize_batch(m, cdb.StorageType(st))
        qcodes, qmags = sq.quantize_batch(q, cdb.StorageType(st))
        mine[f"d{dim}/st{st}/codes"], mine[f"d{dim}/st{st}/mag"] = codes, mags
        mine[f"d{dim}/st{st}/qcodes"], mine[f"d{dim}/st{st}/qmag"] = qcodes, qmags
        for metric in range(4):
            xi, yi = np.repeat(np.arange(len(q)), len(m)), np.tile(np.arange(len(m)), len(q))
            val, status = cdb.DistanceMetric(cdb.DistanceMetricKind(metric)).calculate_pairs(
                cdb.StorageType(st), dim, qcodes[xi], qmags[xi], codes[yi], mags[yi])
            mine[f"d{dim}/st{st}/m{metric}/value"] = np.where(status == 0, val, 0).astype(np.float32).reshape(len(q), len(m))
            mine[f"d{dim}/st{st}/m{metric}/status"] = status.astype(np.int8).reshape(len(q), len(m))
    assert compare(ref, mine, "cuda") >= 48
