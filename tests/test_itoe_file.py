"""Reader for the reference's raw-embedding store itoe.dim / itoe.<version>.data (TreeMap<InternalId, RawVectorEmbedding>).

The writer below is a test-only restatement of TreeMap::serialize for a never-serialized map (the branch that appends
everything: src/models/serializer/tree_map/node.rs:72-87, quotients_map.rs:182-209, versioned_item.rs:52-74,
raw_vector_embedding.rs:16-72, serializer/mod.rs:25-39) including the node placement of tree_map.rs:512-518 /
models/utils.rs:3-24.  The reader lives in the library (host code); only the index-append test needs a GPU."""
import os
import struct

import numpy as np
import pytest

import cosdata_b200 as cdb
import oracle as orc

NONE32 = 0xFFFFFFFF


def write_len(n):                                            # serializer/mod.rs:25-39
    if n < 1 << 7:
        return bytes([n])
    if n < 1 << 14:
        return bytes([(n & 0x7F) | 0x80, n >> 7])
    return bytes([(n & 0x7F) | 0x80, ((n >> 7) & 0x7F) | 0x80, n >> 14])


def raw_embedding(ext_id, dense=None, doc_id=None, metadata=None, sparse=None, text=None):
    b = write_len(len(ext_id)) + ext_id.encode()
    b += write_len(len(doc_id)) + doc_id.encode() if doc_id else write_len(0)
    if dense is not None:
        b += write_len(len(dense)) + np.asarray(dense, dtype="<f4").tobytes()
    else:
        b += write_len(0)
    if metadata:
        b += write_len(len(metadata))
        for k, v in metadata.items():
            b += write_len(len(k)) + k.encode()
            b += (b"\x00" + struct.pack("<i", v)) if isinstance(v, int) else (b"\x01" + write_len(len(v)) + v.encode())
    else:
        b += write_len(0)
    if sparse:
        b += write_len(len(sparse)) + b"".join(struct.pack("<If", i, v) for i, v in sparse)
    else:
        b += write_len(0)
    b += write_len(len(text)) + text.encode() if text else write_len(0)
    return b


def calculate_path(pos):                                     # models/utils.rs:3-24
    path = []
    while pos > 0:
        power = (pos.bit_length() - 1) // 2
        path.append(power)
        pos -= 1 << (2 * power)
    return path


class Node:
    def __init__(self, node_idx):
        self.node_idx = node_idx
        self.children = [None] * 8
        self.quotients = []                                  # [(key, [(version, value bytes or None), ...])] in insertion order


class TreeMapWriter:
    def __init__(self):
        self.root = Node(0)

    def _node(self, key):
        cur = self.root
        for idx in calculate_path(key % 65536):
            if cur.children[idx] is None:
                cur.children[idx] = Node((cur.node_idx + (1 << (idx * 2))) & 0xFFFF)
            cur = cur.children[idx]
        return cur

    def insert(self, version, key, value):
        node = self._node(key)
        for k, chain in node.quotients:
            if k == key:
                chain.append((version, value))
                return
        node.quotients.append((key, [(version, value)]))

    def delete(self, version, key):
        self.insert(version, key, None)

    def serialize(self, directory):
        dim = bytearray(struct.pack("<I", NONE32))
        data = {}

        def item(chain):                                     # versioned_item.rs:52-74 (next first, then value, then header)
            version, value = chain[0]
            nxt_off, nxt_ver = (item(chain[1:]), chain[1][0]) if len(chain) > 1 else (NONE32, NONE32)
            f = data.setdefault(version, bytearray())
            value_off = NONE32
            if value is not None:
                value_off = len(f)
                f += value
            off = len(f)
            f += struct.pack("<IIII", nxt_off, nxt_ver, version, value_off)
            return off

        def quotients(qs):                                   # quotients_map.rs:182-209
            if not qs:
                return NONE32
            chunks = []
            for c in range(0, len(qs), 4):
                buf = b""
                for i in range(c, c + 4):
                    if i < len(qs):
                        key, chain = qs[i]
                        buf += struct.pack("<QII", key, item(chain), chain[0][0])
                    else:
                        buf += b"\xff" * 16
                chunks.append(buf)
            start = len(dim)
            dim.extend(struct.pack("<Q", len(qs)) + chunks[0] + b"\xff" * 4)
            prev = start + 8
            for buf in chunks[1:]:
                off = len(dim)
                dim.extend(buf + b"\xff" * 4)
                dim[prev + 64:prev + 68] = struct.pack("<I", off)
                prev = off
            return start

        def node(n):                                         # node.rs:72-87 (children, then quotients, then the node)
            buf = struct.pack("<H", n.node_idx)
            for c in n.children:
                buf += struct.pack("<I", node(c) if c is not None else NONE32)
            buf += struct.pack("<I", quotients(n.quotients))
            off = len(dim)
            dim.extend(buf)
            return off

        root = node(self.root)
        dim[0:4] = struct.pack("<I", root)
        os.makedirs(directory, exist_ok=True)
        open(os.path.join(directory, "itoe.dim"), "wb").write(bytes(dim))
        for v, f in data.items():
            open(os.path.join(directory, f"itoe.{v}.data"), "wb").write(bytes(f))


def sample_store(directory, dim=20, n=300):
    vecs = orc.synth_matrix(8080, n + 10, dim)
    w = TreeMapWriter()
    live = {}
    keys = list(range(n - 12)) + [65536 * j + 5 for j in range(1, 7)] + [65535, 65536 * 3 + 65535, 70000, 131071, 200000, 999999]
    for i, key in enumerate(keys):
        version = i % 3
        w.insert(version, key, raw_embedding(f"vec-{key}", vecs[i], doc_id="doc" if i % 5 == 0 else None,
                                             metadata={"age": i, "name": "x" * (i % 200)} if i % 7 == 0 else None,
                                             sparse=[(1, 0.5), (9, 1.5)] if i % 11 == 0 else None,
                                             text="lorem " * (i % 40) if i % 13 == 0 else None))
        live[key] = vecs[i]
    w.insert(3, 7, raw_embedding("vec-7", vecs[n]))                      # update: the newest version wins
    live[7] = vecs[n]
    w.insert(4, 7, raw_embedding("vec-7", vecs[n + 1]))
    live[7] = vecs[n + 1]
    w.delete(3, 11)                                                      # delete: key disappears
    del live[11]
    w.delete(2, 65536 + 5)
    del live[65536 + 5]
    w.delete(3, 13)
    w.insert(4, 13, raw_embedding("vec-13", vecs[n + 2]))                # delete then re-insert
    live[13] = vecs[n + 2]
    w.insert(1, 424242, raw_embedding("sparse-only", None, sparse=[(3, 1.0)]))   # no dense values: skipped
    w.serialize(directory)
    return live


def test_itoe_enumeration_and_lookup(tmp_path):
    d = str(tmp_path / "coll")
    live = sample_store(d)
    n, dim, mx = cdb.itoe_scan(d)
    assert (n, dim, mx) == (len(live), 20, max(live))
    ids, vecs = cdb.itoe_load(d)
    assert ids.tolist() == sorted(live)
    for i, key in enumerate(ids):
        assert np.array_equal(vecs[i].view(np.uint32), live[int(key)].view(np.uint32)), key
    ids2, vecs2 = cdb.itoe_load(d, first_entry=100, max_entries=50)
    assert np.array_equal(ids2, ids[100:150]) and np.array_equal(vecs2, vecs[100:150])
    for key in (0, 7, 13, 65535, 65536 * 3 + 5, 65536 * 3 + 65535, 999999):
        assert np.array_equal(cdb.itoe_get(d, key), live[key]), key
    for key in (11, 65536 + 5, 424242, 555, 65536 * 9 + 5):             # deleted, sparse-only, never inserted
        assert cdb.itoe_get(d, key) is None, key


def test_itoe_empty_and_damaged_stores(tmp_path):
    d = str(tmp_path / "empty")
    TreeMapWriter().serialize(d)
    assert cdb.itoe_scan(d) == (0, 0, 0) and cdb.itoe_get(d, 3) is None
    os.makedirs(tmp_path / "fresh")
    open(tmp_path / "fresh" / "itoe.dim", "wb").write(b"")                # created but never serialized (collection.rs:149-155)
    assert cdb.itoe_scan(str(tmp_path / "fresh")) == (0, 0, 0)
    with pytest.raises(cdb.CosdataError):
        cdb.itoe_scan(str(tmp_path / "missing"))
    d = str(tmp_path / "coll")
    sample_store(d)
    os.remove(os.path.join(d, "itoe.2.data"))
    with pytest.raises(cdb.CosdataError) as e:
        cdb.itoe_scan(d)
    assert "itoe.2.data" in str(e.value)
    d = str(tmp_path / "coll2")
    sample_store(d)
    blob = open(os.path.join(d, "itoe.dim"), "rb").read()
    open(os.path.join(d, "itoe.dim"), "wb").write(blob[: len(blob) // 2])
    with pytest.raises(cdb.CosdataError):
        cdb.itoe_scan(d)
    d = str(tmp_path / "mixed")
    w = TreeMapWriter()
    w.insert(0, 1, raw_embedding("a", np.zeros(4, dtype=np.float32)))
    w.insert(0, 2, raw_embedding("b", np.zeros(5, dtype=np.float32)))
    w.serialize(d)
    with pytest.raises(cdb.CosdataError) as e:
        cdb.itoe_scan(d)
    assert e.value.status == cdb.Status.STORAGE_MISMATCH


@pytest.mark.gpu
def test_index_fed_from_itoe_store_reranks_like_the_oracle(tmp_path):
    d = str(tmp_path / "coll")
    live = sample_store(d, dim=20)
    keys = sorted(live)
    corpus = np.stack([live[k] for k in keys])
    ix = cdb.DenseIndex(dim=20, storage_type=cdb.StorageType.UnsignedByte, capacity=len(keys), keep_raw_f32=True)
    n, ids = ix.append_itoe(d, max_ids=len(keys))
    assert n == len(keys) and ids.tolist() == keys
    q = orc.synth_matrix(8081, 5, 20)
    got_ids, got_scores, _, _ = ix.batch_search(q, 10)                    # exact f32 cosine over the raw rows
    want_ids, want_scores = orc.brute_topk_f32(corpus, q, 10)
    assert np.array_equal(got_ids, want_ids) and np.array_equal(got_scores.view(np.uint32), want_scores.view(np.uint32))
    codes, mags = ix.read_codes(0, n)
    want_codes, want_mags = orc.quantize_batch(0, corpus)
    assert np.array_equal(codes, want_codes) and np.array_equal(mags.view(np.uint32), want_mags.view(np.uint32))
    wrong = cdb.DenseIndex(dim=24, storage_type=cdb.StorageType.UnsignedByte, capacity=len(keys))
    with pytest.raises(cdb.CosdataError) as e:
        wrong.append_itoe(d)
    assert e.value.status == cdb.Status.STORAGE_MISMATCH
    ix.close(); wrong.close()


def test_itoe_cyclic_links_are_rejected(tmp_path):
    d = str(tmp_path / "coll")
    sample_store(d, dim=8, n=120)
    blob = bytearray(open(os.path.join(d, "itoe.dim"), "rb").read())
    root = struct.unpack("<I", blob[0:4])[0]
    first_child = next(c for c in struct.unpack("<8I", blob[root + 2:root + 34]) if c != NONE32)
    cyc = bytearray(blob)
    slot = next(i for i in range(8) if struct.unpack("<I", cyc[first_child + 2 + 4 * i:first_child + 6 + 4 * i])[0] != NONE32
                or True)
    cyc[first_child + 2 + 4 * slot:first_child + 6 + 4 * slot] = struct.pack("<I", root)     # a child pointing back at the root
    open(os.path.join(d, "itoe.dim"), "wb").write(bytes(cyc))
    with pytest.raises(cdb.CosdataError):
        cdb.itoe_scan(d)
    cyc = bytearray(blob)
    q = struct.unpack("<I", cyc[root + 34:root + 38])[0]                                      # root's quotient map: chunk -> itself
    assert q != NONE32
    cyc[q:q + 8] = struct.pack("<Q", 1 << 40)
    cyc[q + 8 + 64:q + 8 + 68] = struct.pack("<I", q + 8)
    open(os.path.join(d, "itoe.dim"), "wb").write(bytes(cyc))
    with pytest.raises(cdb.CosdataError):
        cdb.itoe_scan(d)
