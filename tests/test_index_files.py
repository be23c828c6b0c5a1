"""Reader for the reference's on-disk HNSW index (prop.data + nodes.ptr + <id>.index) -> flat graph arrays (CPU tests).

The writer below restates the reference's serializers for a finished index: ProbNode records (serializer/hnsw/node.rs:19-101),
neighbour entries (neighbors.rs:21-60), latest-version links (latest_node.rs:17-44, flat nodes.ptr image) and the prop.data
records (file_persist.rs:58-139).  The flattened graph must be the source graph up to the numbering of the nodes, and the
oracle's searches (filtered and unfiltered) must give identical results on both."""
import os
import struct

import numpy as np
import pytest

import cosdata_b200 as cdb
import oracle as orc
from oracle import pyhnsw, pymeta
from tests import mdgraph
from tests.test_prop_file import enc_metadata_record, enc_record

NONE32 = 0xFFFFFFFF


def write_index_dir(directory, vecs, mg, seed=0, with_metadata=True):
    """-> (root_link, pseudo_link).  Storage records are written in row order, so a node's row is its record ordinal."""
    os.makedirs(directory, exist_ok=True)
    fg = mg.fg
    rng = np.random.default_rng(seed)
    st, dim = fg.storage_type, fg.dim
    L1 = fg.num_levels + 1
    # ---- prop.data
    value_loc, md_loc = {}, {}
    row_id = {}
    for lv in range(L1):                                                  # prop id of a row = id of its base node
        for i in range(fg.cnt[lv]):
            row = int(fg.node_row[lv][i])
            nid, md = int(mg.node_id[lv][i]), int(mg.node_md[lv][i])
            if md == NONE32 or mg.md_mags[md] == 0.0 or row not in row_id:
                row_id.setdefault(row, nid)
    with open(os.path.join(directory, "prop.data"), "wb") as f:
        for row in range(fg.codes.shape[0]):
            rec = enc_record(row_id.get(row, 4_000_000 + row), st, fg.mags[row], fg.codes[row], dim)
            value_loc[row] = (f.tell(), len(rec))
            f.write(rec)
            if with_metadata:
                for lv in range(L1):                                      # metadata records of the nodes of this row, once per id
                    for i in np.flatnonzero(fg.node_row[lv] == row):
                        nid, md = int(mg.node_id[lv][i]), int(mg.node_md[lv][i])
                        if md != NONE32 and nid not in md_loc:
                            rec = enc_metadata_record(nid, mg.md_mags[md], mg.md_bits[md])
                            md_loc[nid] = (f.tell(), len(rec))
                            f.write(rec)
    # ---- link slots (shuffled) and record placement over two index files, with stale garbage in between
    nodes = [(lv, i) for lv in range(L1) for i in range(fg.cnt[lv])]
    slots = rng.permutation(len(nodes))
    link = {n: int(s) * 8 for n, s in zip(nodes, slots)}
    parent = {}
    for lv in range(1, L1):
        for j in range(fg.cnt[lv]):
            parent[(lv - 1, int(fg.child[lv][j]))] = (lv, j)
    files = {0: bytearray(), 1: bytearray()}
    place = {}
    for n in nodes:
        lv, i = n
        nb = fg.nbrs(lv)
        size = 31 + 13 * nb
        fid = int(rng.integers(0, 2))
        if rng.random() < 0.3:
            files[fid] += bytes(rng.integers(0, 256, size, dtype=np.uint8))   # an older version nobody links to
        place[n] = (fid, len(files[fid]))
        files[fid] += bytes(size)
    for n in nodes:
        lv, i = n
        nb = fg.nbrs(lv)
        row, nid, md = int(fg.node_row[lv][i]), int(mg.node_id[lv][i]), int(mg.node_md[lv][i])
        buf = struct.pack("<BI", lv, 7) + struct.pack("<II", *value_loc[row])
        buf += struct.pack("<II", *md_loc[nid]) if (with_metadata and md != NONE32) else b"\xff" * 8
        buf += struct.pack("<I", link[parent[n]] if n in parent else NONE32)
        buf += struct.pack("<I", link[(lv - 1, int(fg.child[lv][i]))] if lv > 0 else NONE32)
        buf += struct.pack("<H", nb)
        for s in range(nb):
            t = int(fg.adj[lv][i * nb + s])
            if t == NONE32:
                buf += b"\xff" * 13
            else:
                buf += struct.pack("<IIBf", int(mg.node_id[lv][t]), link[(lv, t)], 0, 0.5)
        fid, off = place[n]
        files[fid][off:off + len(buf)] = buf
    ptr = bytearray(8 * len(nodes))
    for n in nodes:
        fid, off = place[n]
        ptr[link[n]:link[n] + 8] = struct.pack("<II", off, fid)
    open(os.path.join(directory, "nodes.ptr"), "wb").write(bytes(ptr))
    for fid, b in files.items():
        open(os.path.join(directory, f"{fid}.index"), "wb").write(bytes(b))
    return link[(fg.num_levels, fg.entry)], link[(fg.num_levels, mg.pseudo_entry)]


def reachable(mg):
    fg = mg.fg
    L1 = fg.num_levels + 1
    parent = {}
    for lv in range(1, L1):
        for j in range(fg.cnt[lv]):
            parent[(lv - 1, int(fg.child[lv][j]))] = (lv, j)
    seen, todo = set(), [(fg.num_levels, fg.entry), (fg.num_levels, mg.pseudo_entry)]
    while todo:
        n = todo.pop()
        if n in seen:
            continue
        seen.add(n)
        lv, i = n
        nb = fg.nbrs(lv)
        todo += [(lv, int(t)) for t in fg.adj[lv][i * nb:(i + 1) * nb] if t != NONE32]
        if lv > 0:
            todo.append((lv - 1, int(fg.child[lv][i])))
        if n in parent:
            todo.append(parent[n])
    return seen


@pytest.mark.parametrize("st,metric", [(4, 0), (0, 0)])
def test_flattened_index_files_equal_the_source_graph(tmp_path, st, metric):
    vecs, mg = mdgraph.build(n=260, dim=16, md_dims=6, levels=3, nb=8, nb0=16, storage_type=st, metric=metric, seed=31)
    fg = mg.fg
    d = str(tmp_path / "idx")
    root_link, pseudo_link = write_index_dir(d, vecs, mg, seed=3)
    hf = cdb.HnswFiles(d, root_link, pseudo_link)
    assert (hf.num_levels, hf.neighbors_count, hf.level0_neighbors_count) == (fg.num_levels, 8, 16)
    assert hf.root_row == fg.n and hf.md_dims == mg.md_dims
    reach = reachable(mg)
    for lv in range(fg.num_levels + 1):
        src = {int(mg.node_id[lv][i]): i for i in range(fg.cnt[lv]) if (lv, i) in reach}
        assert sorted(hf.node_id[lv].tolist()) == sorted(src)             # exactly the reachable nodes, each once
        nb = fg.nbrs(lv)
        for j, nid in enumerate(hf.node_id[lv].tolist()):
            i = src[nid]
            assert hf.node_row[lv][j] == fg.node_row[lv][i]
            md_src, md_got = int(mg.node_md[lv][i]), int(hf.node_md[lv][j])
            assert (md_src == NONE32) == (md_got == NONE32)
            if md_src != NONE32:
                assert np.array_equal(hf.md_bits[md_got], mg.md_bits[md_src]) and hf.md_mags[md_got] == mg.md_mags[md_src]
            want = [NONE32 if t == NONE32 else int(mg.node_id[lv][t]) for t in fg.adj[lv][i * nb:(i + 1) * nb]]
            got = [NONE32 if t == NONE32 else int(hf.node_id[lv][t]) for t in hf.adj[lv][j * nb:(j + 1) * nb]]
            assert got == want                                            # slot order preserved
            if lv > 0:
                assert hf.node_id[lv - 1][hf.child[lv][j]] == mg.node_id[lv - 1][fg.child[lv][i]]
    assert hf.node_id[fg.num_levels][hf.entry] == 0xFFFFFFFF and hf.node_id[fg.num_levels][hf.pseudo_entry] == mdgraph.PSEUDO_ROOT_ID
    # the oracle's search gives the same answers on the flattened graph
    fg2 = pyhnsw.FlatGraph(metric, st, fg.dim, fg.codes, fg.mags, fg.n, hf.num_levels, hf.neighbors_count, hf.level0_neighbors_count,
                           hf.entry, hf.node_row, hf.adj, hf.child)
    mg2 = pymeta.MdGraph(fg2, hf.md_bits, hf.md_mags, hf.node_id, hf.node_md, hf.pseudo_entry)
    q, filters = mdgraph.make_queries(vecs, mg, 40, seed=4)
    a = pymeta.search_batch_md(mg, vecs, q, filters, 5, ef_search=16)
    b = pymeta.search_batch_md(mg2, vecs, q, filters, 5, ef_search=16)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    hf.close()


def test_index_files_without_metadata_and_damage(tmp_path):
    vecs, mg = mdgraph.build(n=120, dim=16, levels=2, nb=4, nb0=8, storage_type=4, metric=0, seed=8)
    d = str(tmp_path / "plain")
    root_link, pseudo_link = write_index_dir(d, vecs, mg, seed=1, with_metadata=False)
    hf = cdb.HnswFiles(d, root_link)                                      # no pseudo root: only the main component is read
    assert hf.n_md == 0 and hf.pseudo_entry == NONE32
    assert all((a == NONE32).all() for a in hf.node_md)
    main = {int(i) for lv in range(3) for i in mg.node_id[lv]}
    assert set(hf.node_id[0].tolist()) <= main and 0xFFFFFFFF in hf.node_id[0].tolist()
    assert mdgraph.PSEUDO_ROOT_ID not in hf.node_id[2].tolist()
    hf.close()
    with pytest.raises(cdb.CosdataError):                                 # a link outside nodes.ptr
        cdb.HnswFiles(d, 10 ** 8)
    os.remove(os.path.join(d, "1.index"))
    with pytest.raises(cdb.CosdataError) as e:
        cdb.HnswFiles(d, root_link)
    assert "1.index" in str(e.value)
    with pytest.raises(cdb.CosdataError):
        cdb.HnswFiles(str(tmp_path / "missing"), 0)
    d2 = str(tmp_path / "lowroot")
    rl, _ = write_index_dir(d2, vecs, mg, seed=2, with_metadata=False)
    blob = bytearray(open(os.path.join(d2, "nodes.ptr"), "rb").read())
    off, fid = struct.unpack("<II", blob[rl:rl + 8])
    idx = bytearray(open(os.path.join(d2, f"{fid}.index"), "rb").read())
    idx[off + 5:off + 9] = struct.pack("<I", 123456789)                   # the root's prop offset points nowhere
    open(os.path.join(d2, f"{fid}.index"), "wb").write(bytes(idx))
    with pytest.raises(cdb.CosdataError) as e:
        cdb.HnswFiles(d2, rl)
    assert "Storage record" in str(e.value)


def test_versioned_link_files_of_enable_context_history(tmp_path):
    # enable_context_history: no nodes.ptr; every flush writes each dirty 8192-byte region of the link image in full to
    # "<region>-<version>.ptr" and the loader takes the highest version <= the current one per region
    # (cache_loader.rs:91-113, buffered_io.rs:524-570)
    vecs, mg = mdgraph.build(n=1500, dim=16, md_dims=6, levels=3, nb=8, nb0=16, storage_type=4, metric=0, seed=13)
    flat_dir, ver_dir = str(tmp_path / "flat"), str(tmp_path / "versioned")
    root_link, pseudo_link = write_index_dir(flat_dir, vecs, mg, seed=5)
    write_index_dir(ver_dir, vecs, mg, seed=5)
    image = open(os.path.join(ver_dir, "nodes.ptr"), "rb").read()
    os.remove(os.path.join(ver_dir, "nodes.ptr"))
    assert len(image) > 2 * 8192                                          # several regions
    rng = np.random.default_rng(2)
    for r in range((len(image) + 8191) // 8192):
        chunk = image[r * 8192:(r + 1) * 8192]
        open(os.path.join(ver_dir, f"{r}-3.ptr"), "wb").write(chunk)                      # the current state
        stale = bytearray(chunk)
        stale[: len(stale) // 2] = bytes(rng.integers(0, 256, len(stale) // 2, dtype=np.uint8))
        open(os.path.join(ver_dir, f"{r}-1.ptr"), "wb").write(bytes(stale))               # an older flush of the region
        if r % 2 == 0:
            open(os.path.join(ver_dir, f"{r}-9.ptr"), "wb").write(bytes(len(chunk)))      # a version after "current"
    open(os.path.join(ver_dir, "notes-1.ptr.bak"), "wb").write(b"x")                      # names the loader ignores
    open(os.path.join(ver_dir, "a-b.ptr"), "wb").write(b"x")
    a = cdb.HnswFiles(flat_dir, root_link, pseudo_link)
    b = cdb.HnswFiles(ver_dir, root_link, pseudo_link, latest_version=3)
    assert (a.num_levels, a.entry, a.pseudo_entry, a.root_row) == (b.num_levels, b.entry, b.pseudo_entry, b.root_row)
    for lv in range(a.num_levels + 1):
        for name in ("node_row", "node_id", "node_md", "adj", "child"):
            assert np.array_equal(getattr(a, name)[lv], getattr(b, name)[lv]), (lv, name)
    a.close(); b.close()
    with pytest.raises(cdb.CosdataError):                                 # version 9 zeroed the even regions: links lead nowhere valid
        cdb.HnswFiles(ver_dir, root_link, pseudo_link)
    with pytest.raises(cdb.CosdataError):                                 # nothing at or below version 0
        cdb.HnswFiles(ver_dir, root_link, pseudo_link, latest_version=0)


@pytest.mark.gpu
@pytest.mark.parametrize("st,metric", [(0, 0), (4, 0)])
def test_cold_start_from_a_collection_directory_end_to_end(tmp_path, st, metric):
    """prop.data -> rows (quantized payloads), itoe store -> their raw f32 rows, index files -> graph;
    then a filtered HNSW search on the GPU must equal the oracle's on the source graph.  Storage is U8 / F16, i.e. the raw
    rows are NOT derivable from the codes -- the case the documented cold start could not serve before."""
    from tests.test_itoe_file import TreeMapWriter, raw_embedding
    vecs, mg = mdgraph.build(n=500, dim=24, md_dims=6, levels=3, nb=8, nb0=16, storage_type=st, metric=metric, seed=17)
    fg = mg.fg
    idx_dir, coll_dir = str(tmp_path / "coll" / "idx"), str(tmp_path / "coll")
    os.makedirs(coll_dir)
    root_link, pseudo_link = write_index_dir(idx_dir, vecs, mg, seed=2)
    n = fg.n
    w = TreeMapWriter()
    for row in range(n):                                                  # raw embeddings are keyed by the base id of the row
        w.insert(row % 3, row * 4, raw_embedding(f"vec-{row}", vecs[row]))
    w.serialize(coll_dir)

    ix = cdb.DenseIndex(dim=24, storage_type=cdb.StorageType(st), metric=cdb.DistanceMetricKind(metric), capacity=vecs.shape[0],
                        keep_raw_f32=True)
    cnt, ids = ix.append_prop_file(os.path.join(idx_dir, "prop.data"), max_ids=vecs.shape[0])
    assert cnt == vecs.shape[0] and ix.raw_missing == cnt
    with pytest.raises(cdb.CosdataError):                                 # no raw rows yet: the re-ranking modes are refused
        ix.batch_search(vecs[:2], 3)
    row_ids = ids.copy()
    row_ids[n:] = cdb.INVALID_ID                                          # the two root vectors have no embedding by design
    assert np.array_equal(row_ids[:n], np.arange(n, dtype=np.uint32) * 4)
    filled, missing = ix.fill_raw_from_itoe(coll_dir, row_ids)
    assert (filled, missing) == (vecs.shape[0], 0) and ix.raw_missing == 0
    hf = cdb.HnswFiles(idx_dir, root_link, pseudo_link)
    hf.apply(ix)
    hf.close()
    q, filters = mdgraph.make_queries(vecs, mg, 48, seed=5)
    got = ix.batch_search_filtered(q, filters, 5, ef_search=24, shortlist_size=64)
    want = pymeta.search_batch_md(mg, vecs, q, filters, 5, ef_search=24)
    for g, w_ in zip(got, want[:4]):
        assert np.array_equal(np.asarray(g).view(np.uint8), np.asarray(w_).view(np.uint8))
    # a row whose id is unknown to the store is reported, not silently zero-filled
    bad = row_ids.copy()
    bad[3] = 999_999
    ix2 = cdb.DenseIndex(dim=24, storage_type=cdb.StorageType(st), metric=cdb.DistanceMetricKind(metric), capacity=vecs.shape[0],
                         keep_raw_f32=True)
    ix2.append_prop_file(os.path.join(idx_dir, "prop.data"))
    with pytest.raises(cdb.CosdataError):
        ix2.fill_raw_from_itoe(coll_dir, bad)
    ix.close(); ix2.close()
