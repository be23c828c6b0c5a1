"""Reader for the reference's prop.data (src/models/file_persist.rs:58-108).

The writer side here is a test-only restatement of what serde_cbor 0.11.2 emits for
`NodePropValueSerialize { id: &InternalId, value: &Storage }` (Cargo.toml:28; the crate is not vendored, so its published
encoding rules are restated: structs = definite maps with text keys in declaration order, struct enum variant = 1-entry map,
newtype structs transparent, Vec<T> = definite array, shortest unsigned ints, f32 as half precision when lossless).  The
reader lives in the library (host code, no GPU): these tests run without a GPU except the index-append one."""
import os
import struct

import numpy as np
import pytest

import cosdata_b200 as cdb
import oracle as orc

ST = cdb.StorageType


# ------------------------------------------------------------------ serde_cbor-style writer (test infrastructure)

def head(major, arg):
    if arg < 24:
        return bytes([major << 5 | arg])
    for info, fmt, lim in ((24, ">B", 1 << 8), (25, ">H", 1 << 16), (26, ">I", 1 << 32), (27, ">Q", 1 << 64)):
        if arg < lim:
            return bytes([major << 5 | info]) + struct.pack(fmt, arg)
    raise ValueError(arg)


def enc_uint(v):
    return head(0, int(v))


def enc_text(s):
    b = s.encode()
    return head(3, len(b)) + b


def enc_f32(v):
    v = np.float32(v)
    if np.isinf(v):
        return b"\xf9\x7c\x00" if v > 0 else b"\xf9\xfc\x00"
    if np.isnan(v):
        return b"\xf9\x7e\x00"
    h = np.float16(v)
    if np.float32(h) == v:                                   # serde_cbor: half precision when lossless
        return b"\xf9" + struct.pack(">H", int(h.view(np.uint16)))
    return b"\xfa" + struct.pack(">f", float(v))


def enc_array(items):
    return head(4, len(items)) + b"".join(items)


def enc_map(pairs):
    return head(5, len(pairs)) + b"".join(enc_text(k) + v for k, v in pairs)


def enc_storage(st, mag, code, dim):
    st = int(st)
    if st == 0:
        return enc_map([("UnsignedByte", enc_map([("mag", enc_f32(mag)), ("quant_vec", enc_array([enc_uint(x) for x in code]))]))])
    if 1 <= st <= 3:
        pb = (dim + 7) // 8
        planes = [enc_array([enc_uint(x) for x in code[p * pb:(p + 1) * pb]]) for p in range(st)]
        return enc_map([("SubByte", enc_map([("mag", enc_f32(mag)), ("quant_vec", enc_array(planes)), ("resolution", enc_uint(st))]))])
    if st == 4:
        bits = np.frombuffer(bytes(code), dtype="<u2")
        return enc_map([("HalfPrecisionFP", enc_map([("mag", enc_f32(mag)), ("quant_vec", enc_array([enc_uint(x) for x in bits]))]))])
    vals = np.frombuffer(bytes(code), dtype="<f4")
    return enc_map([("FullPrecisionFP", enc_map([("mag", enc_f32(mag)), ("vec", enc_array([enc_f32(x) for x in vals]))]))])


def enc_record(node_id, st, mag, code, dim):
    return enc_map([("id", enc_uint(node_id)), ("value", enc_storage(st, mag, code, dim))])


def write_prop_file(path, st, vectors, ids):
    codes, mags = orc.quantize_batch(int(st), vectors)
    locs = []
    with open(path, "wb") as f:
        for i in range(vectors.shape[0]):
            rec = enc_record(ids[i], st, mags[i], codes[i], vectors.shape[1])
            locs.append((f.tell(), len(rec)))
            f.write(rec)
    return codes, mags, locs


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------ tests

def test_writer_matches_known_serde_cbor_bytes():
    # hand-assembled from RFC 8949: {"id": 7, "value": {"UnsignedByte": {"mag": 1.5, "quant_vec": [0, 23, 24, 255]}}}
    want = bytes.fromhex("a2" "626964" "07" "6576616c7565" "a1" "6c556e7369676e656442797465" "a2"
                         "636d6167" "f93e00" "697175616e745f766563" "84" "00" "17" "1818" "18ff")
    assert enc_record(7, ST.UnsignedByte, 1.5, [0, 23, 24, 255], 4) == want
    assert enc_f32(0.1) == b"\xfa" + struct.pack(">f", np.float32(0.1))     # not representable in half -> f32


@pytest.mark.parametrize("st", [s for s in ST if s != ST.BFloat16])
@pytest.mark.parametrize("dim", [8, 33, 100])
def test_prop_file_round_trip(tmp_path, st, dim):
    n = 57
    vecs = orc.synth_matrix(4000 + dim + int(st), n, dim)
    ids = np.arange(n, dtype=np.uint32) * 3 + 1
    ids[-1] = 0xFFFFFFFF                                                     # the root node's prop is in the file too
    path = str(tmp_path / "prop.data")
    codes, mags, locs = write_prop_file(path, st, vecs, ids)
    total, got_st, elems, cb = cdb.prop_file_scan(path)
    assert (total, got_st, cb) == (n, st, cdb.code_bytes(st, dim))
    assert elems == ((dim + 7) // 8 if 1 <= int(st) <= 3 else dim)
    rec = cdb.prop_file_load(path)
    assert np.array_equal(rec["ids"], ids)
    assert np.array_equal(rec["codes"], codes) and np.array_equal(bits(rec["mags"]), bits(mags))
    assert [(int(o), int(l)) for o, l in zip(rec["offsets"], rec["lengths"])] == locs   # == ProbNode prop_value.location
    part = cdb.prop_file_load(path, first_record=50, max_records=100)
    assert np.array_equal(part["ids"], ids[50:]) and np.array_equal(part["codes"], codes[50:])


def test_prop_file_tolerates_other_cbor_spellings(tmp_path):
    # f64 mag, indefinite-length containers, a byte string for Vec<u8>, an unknown extra field, swapped key order
    rec = (b"\xbf" + enc_text("value") + b"\xa1" + enc_text("UnsignedByte") + b"\xbf" + enc_text("quant_vec") + b"\x44\x01\x02\x03\xff"
           + enc_text("note") + enc_array([enc_text("x"), b"\xf6"]) + enc_text("mag") + b"\xfb" + struct.pack(">d", 2.5) + b"\xff"
           + enc_text("id") + enc_uint(70000) + b"\xff")
    rec2 = enc_map([("id", enc_uint(5)), ("value", enc_map([("UnsignedByte", enc_map([("mag", enc_f32(0.0)),
                   ("quant_vec", b"\x9f" + b"".join(enc_uint(x) for x in (9, 8, 7, 200)) + b"\xff")]))]))])
    path = str(tmp_path / "prop.data")
    open(path, "wb").write(rec + rec2)
    out = cdb.prop_file_load(path)
    assert out["ids"].tolist() == [70000, 5]
    assert out["codes"].tolist() == [[1, 2, 3, 255], [9, 8, 7, 200]] and out["mags"].tolist() == [2.5, 0.0]


def test_prop_file_errors(tmp_path):
    path = str(tmp_path / "prop.data")
    good = enc_record(1, ST.UnsignedByte, 1.0, [1, 2, 3, 4], 4)
    open(path, "wb").write(good + good[:-3])
    with pytest.raises(cdb.CosdataError) as e:
        cdb.prop_file_scan(path)
    assert e.value.status == cdb.Status.INVALID_PARAMS and "record 1" in str(e.value)
    open(path, "wb").write(good + enc_record(2, ST.UnsignedByte, 1.0, [1, 2, 3], 3))
    with pytest.raises(cdb.CosdataError) as e:
        cdb.prop_file_scan(path)
    assert e.value.status == cdb.Status.STORAGE_MISMATCH
    open(path, "wb").write(good + enc_record(2, ST.HalfPrecisionFP, 1.0, [0, 60, 0, 60, 0, 60, 0, 60], 4))
    with pytest.raises(cdb.CosdataError) as e:
        cdb.prop_file_load(path)
    assert e.value.status == cdb.Status.STORAGE_MISMATCH
    open(path, "wb").write(enc_map([("id", enc_uint(1)), ("value", enc_map([("Quux", enc_map([]))]))]))
    with pytest.raises(cdb.CosdataError):
        cdb.prop_file_scan(path)
    with pytest.raises(cdb.CosdataError):
        cdb.prop_file_scan(str(tmp_path / "missing.data"))
    open(path, "wb").write(b"")
    assert cdb.prop_file_scan(path) == (0, None, 0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("st,metric", [(ST.UnsignedByte, cdb.DistanceMetricKind.Cosine), (ST.SubByte2, cdb.DistanceMetricKind.DotProduct),
                                        (ST.HalfPrecisionFP, cdb.DistanceMetricKind.Cosine)])
def test_index_fed_from_prop_file_searches_like_index_fed_from_vectors(tmp_path, st, metric):
    n, dim, k = 3000, 40, 10
    vecs = orc.synth_matrix(4242, n, dim)
    path = str(tmp_path / "prop.data")
    codes, mags, _ = write_prop_file(path, st, vecs, np.arange(n, dtype=np.uint32))
    ix = cdb.DenseIndex(dim=dim, storage_type=st, metric=metric, capacity=n)
    got_n, ids = ix.append_prop_file(path, max_ids=n)
    assert got_n == n and np.array_equal(ids, np.arange(n, dtype=np.uint32)) and len(ix) == n
    c2, m2 = ix.read_codes(0, n)
    assert np.array_equal(c2, codes) and np.array_equal(bits(m2), bits(mags))
    q = orc.synth_matrix(4243, 4, dim)
    qc, qm = orc.quantize_batch(int(st), q)
    rc, want_ids, want_scores, want_err = orc.brute_topk_codes(int(metric), int(st), dim, codes, mags, qc, qm, k)
    got_ids, got_scores, _, err = ix.batch_search(q, k, cdb.SearchMode.BRUTE_CODES)
    assert rc == 0 and np.array_equal(got_ids, want_ids) and np.array_equal(bits(got_scores), bits(want_scores))
    wrong = cdb.DenseIndex(dim=dim + 8, storage_type=st, metric=metric, capacity=n)
    with pytest.raises(cdb.CosdataError) as e:
        wrong.append_prop_file(path)
    assert e.value.status == cdb.Status.STORAGE_MISMATCH
    ix.close(); wrong.close()


def enc_int(v):
    return head(0, int(v)) if v >= 0 else head(1, -1 - int(v))


def enc_metadata_record(replica_id, mag, mbits):
    """write_prop_metadata_to_file (file_persist.rs:110-139): { replica_id, vec: Metadata { mag, mbits } }"""
    return enc_map([("replica_id", enc_uint(replica_id)), ("vec", enc_map([("mag", enc_f32(mag)), ("mbits", enc_array([enc_int(b) for b in mbits]))]))])


def test_prop_file_with_interleaved_metadata_records(tmp_path):
    """collections with a metadata schema write the replicas' Metadata into the same file (vector_store.rs:560-585)"""
    from oracle import pymeta
    dim, n, M = 24, 40, 6
    vecs = orc.synth_matrix(77, n, dim)
    codes, mags = orc.quantize_batch(0, vecs)
    rng = np.random.default_rng(5)
    path = str(tmp_path / "prop.data")
    value_locs, md_want = [], []
    with open(path, "wb") as f:
        for i in range(n):
            rec = enc_record(i * 4, ST.UnsignedByte, mags[i], codes[i], dim)
            value_locs.append((f.tell(), len(rec)))
            f.write(rec)
            for j in range(int(rng.integers(0, 3))):                        # 0-2 replicas with metadata after the vector
                mb = rng.integers(-3, 1025, M).astype(np.int32)
                mg = pymeta.metadata_mag(mb)
                rec = enc_metadata_record(i * 4 + j, mg, mb)
                md_want.append((i * 4 + j, mg, mb, f.tell(), len(rec)))
                f.write(rec)
    total, st, elems, cb = cdb.prop_file_scan(path)
    assert (total, st, cb) == (n, ST.UnsignedByte, dim)
    rec = cdb.prop_file_load(path)
    assert rec["ids"].tolist() == [i * 4 for i in range(n)] and np.array_equal(rec["codes"], codes)
    assert [(int(o), int(l)) for o, l in zip(rec["offsets"], rec["lengths"])] == value_locs
    part = cdb.prop_file_load(path, first_record=30, max_records=5)            # ordinals count Storage records only
    assert part["ids"].tolist() == [i * 4 for i in range(30, 35)]
    md = cdb.prop_file_load_metadata(path)
    assert md["replica_ids"].tolist() == [w[0] for w in md_want]
    assert np.array_equal(bits(md["mags"]), bits(np.array([w[1] for w in md_want], dtype=np.float32)))
    assert np.array_equal(md["mbits"], np.stack([w[2] for w in md_want]))
    assert [(int(o), int(l)) for o, l in zip(md["offsets"], md["lengths"])] == [(w[3], w[4]) for w in md_want]
    open(path, "ab").write(enc_map([("replica_id", enc_uint(1)), ("value", enc_uint(3))]))      # neither kind of record
    with pytest.raises(cdb.CosdataError):
        cdb.prop_file_load_metadata(path)
