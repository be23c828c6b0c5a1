"""Pins oracle/metadata_oracle.c (metadata-filter arms, filtered ann_search) against a pure-Python restatement of
src/distance/cosine.rs:34-102, 243-259, src/models/types.rs:111-147, 223-243 and src/vector_store.rs:256-402."""
import numpy as np
import pytest

import oracle as orc
from oracle import pyhnsw, pymeta
from tests import f32emu, mdgraph

EMPTY = 0xFFFFFFFF


def bits(x):
    return np.float32(x).view(np.uint32)


def py_kind(vid, md_bits, md_mag):
    if md_bits is None or md_mag == 0.0:
        return "base"
    if vid is not None and 0xFFFFFFFF - 257 <= vid <= 0xFFFFFFFF - 2:
        return "pseudo"
    return "metadata"


def py_mdims_cos(xb, xm, yb, ym):
    dot = f32emu.dot_f32_simd_order([float(v) for v in xb], [float(v) for v in yb])
    den = f32emu.mul32(float(xm), float(ym))
    return None if den == 0.0 else np.float32(f32emu.div32(dot, den))


def py_distance(metric, st, dim, x, y):
    """x, y = (code, mag, id or None, md_bits or None, md_mag) -> (rc, value)"""
    if metric != 0:
        return orc.distance(metric, st, dim, x[0], x[1], y[0], y[1])
    xk, yk = py_kind(x[2], x[3], x[4]), py_kind(y[2], y[3], y[4])
    if (yk, xk) == ("pseudo", "pseudo"):
        c = py_mdims_cos(x[3], x[4], y[3], y[4])
        return (2, np.float32(0)) if c is None else (0, c)
    if (yk, xk) == ("pseudo", "metadata"):
        return 0, np.float32(1.0 if list(x[3]) == list(y[3]) else -1.0)
    if (yk, xk) == ("base", "base"):
        return orc.distance(metric, st, dim, x[0], x[1], y[0], y[1])
    if (yk, xk) == ("metadata", "metadata"):
        c = py_mdims_cos(x[3], x[4], y[3], y[4])
        if c is None:
            return 2, np.float32(0)
        return orc.distance(metric, st, dim, x[0], x[1], y[0], y[1]) if c > np.float32(0.99) else (0, np.float32(-1.0))
    if (yk, xk) == ("base", "metadata"):
        return 0, np.float32(0.0)
    return 7, np.float32(0)


def test_metadata_magnitudes():
    assert pymeta.metadata_mag([1, 1, 0, 1]) == np.float32(np.sqrt(np.float32(3)))
    assert pymeta.metadata_mag([0, 0, 0]) == 0.0
    assert pymeta.metadata_mag([2 ** 31 - 1] * 64) == np.float32(2.0 ** 34)       # (i32::MAX as f32) = 2^31; 64 * 2^62 = 2^68, far below f32::MAX
    assert pymeta.query_filter_mag([1, -1, 0, 1]) == np.float32(np.sqrt(np.float32(3)))
    big = np.arange(1, 40, dtype=np.int32) * 1000
    assert pymeta.metadata_mag(big) == np.float32(f32emu.sqrt32(f32emu.sumsq_sequential([float(v) for v in big])))


def test_replica_kinds_and_every_arm():
    dim, st, M = 16, 4, 5
    v = orc.synth_matrix(1, 3, dim)
    codes, mags = orc.quantize_batch(st, v)
    pat_a, pat_b = np.array([1, 0, 1, 1, 0], np.int32), np.array([0, 1, 1, 0, 0], np.int32)
    zeros = np.zeros(M, np.int32)
    sides = {
        "none": (None, None, 0.0), "noid_md": (None, pat_a, pymeta.metadata_mag(pat_a)), "zero_md": (5, zeros, 0.0),
        "meta_a": (8, pat_a, pymeta.metadata_mag(pat_a)), "meta_b": (9, pat_b, pymeta.metadata_mag(pat_b)),
        "pseudo_lo": (0xFFFFFFFF - 257, pat_a, pymeta.metadata_mag(pat_a)), "pseudo_hi": (0xFFFFFFFF - 2, pat_b, pymeta.metadata_mag(pat_b)),
        "just_below": (0xFFFFFFFF - 258, pat_a, pymeta.metadata_mag(pat_a)), "query_id": (0xFFFFFFFF - 1, pat_a, pymeta.metadata_mag(pat_a)),
        "root_id_md": (0xFFFFFFFF, pat_a, pymeta.metadata_mag(pat_a)),
    }
    want_kind = {"none": 1, "noid_md": 2, "zero_md": 1, "meta_a": 2, "meta_b": 2, "pseudo_lo": 0, "pseudo_hi": 0, "just_below": 2,
                 "query_id": 2, "root_id_md": 2}
    for name, (vid, mb, mm) in sides.items():
        assert pymeta.replica_kind(pymeta.VectorData(codes[0], mags[0], vid, mb, mm)) == want_kind[name], name
    seen = set()
    for xn, (xid, xb, xm) in sides.items():
        for yn, (yid, yb, ym) in sides.items():
            for metric in (0, 3):
                x = pymeta.VectorData(codes[0], mags[0], xid, xb, xm)
                y = pymeta.VectorData(codes[1], mags[1], yid, yb, ym)
                rc, val = pymeta.distance_md(metric, st, dim, M, x, y)
                wrc, wval = py_distance(metric, st, dim, (codes[0], mags[0], xid, xb, xm), (codes[1], mags[1], yid, yb, ym))
                assert rc == wrc and (rc != 0 or bits(val) == bits(wval)), (xn, yn, metric)
                if metric == 0:
                    seen.add((want_kind[yn], want_kind[xn], rc))
    assert {(0, 0, 0), (0, 2, 0), (1, 1, 0), (2, 2, 0), (1, 2, 0), (0, 1, 7), (1, 0, 7), (2, 0, 7), (2, 1, 7)} <= seen


# ------------------------------------------------------------------ filtered ann_search, pure Python

def py_search(mg, vecs, q, filt, k, ef, shortlist):
    fg = mg.fg
    st, metric, dim, M = fg.storage_type, fg.metric, fg.dim, mg.md_dims
    qc, qm = orc.quantize(st, q)

    def node(level, i):
        row = int(fg.node_row[level][i])
        md = int(mg.node_md[level][i])
        return (fg.codes[row], fg.mags[row], int(mg.node_id[level][i]), None if md == EMPTY else mg.md_bits[md],
                0.0 if md == EMPTY else mg.md_mags[md])

    def key(score, vid):
        return (orc.order_key(metric, score) << 32) | (~vid & 0xFFFFFFFF)

    def traverse(level, entry, x, fs, nb):
        adj = fg.adj[level].reshape(-1, nb)
        take = min(shortlist, nb)
        y = node(level, entry)
        rc, d = py_distance(metric, st, dim, x, y)
        if rc:
            return rc, []
        fs.add(((y[2] >> 6) & (nb - 1), y[2] & 63))
        heap, res, visited = [(key(d, y[2]), entry, d)], [], 0
        while heap:
            heap.sort(reverse=True)
            cur = heap.pop(0)
            if visited >= ef:
                break
            visited += 1
            res.append(cur)
            for s in range(take):
                nbl = int(adj[cur[1]][s])
                if nbl == EMPTY:
                    continue
                y = node(level, nbl)
                bit = ((y[2] >> 6) & (nb - 1), y[2] & 63)
                if bit in fs:
                    continue
                rc, d = py_distance(metric, st, dim, x, y)
                if rc:
                    return rc, []
                fs.add(bit)
                heap.append((key(d, y[2]), nbl, d))
        res.sort(reverse=True)
        return 0, res[:100]

    entry = mg.pseudo_entry if filt is not None else fg.entry
    out = []
    for level in range(fg.num_levels, -1, -1):
        nb = fg.nbrs(level)
        qid = 0xFFFFFFFE
        fs = {((qid >> 6) & (nb - 1), qid & 63)}
        if filt is not None:
            z = []
            for f in filt:
                x = (qc, qm, None, np.asarray(f, np.int32), pymeta.query_filter_mag(f))
                rc, r = traverse(level, entry, x, fs, nb)
                if rc:
                    return rc, [], []
                z += [t for t in r if not (metric == 0 and t[2] == np.float32(-1.0))]
            z.sort(reverse=True)
            z = z[:100]
        else:
            rc, z = traverse(level, entry, (qc, qm, None, None, 0.0), fs, nb)
            if rc:
                return rc, [], []
        if not z:
            y = node(level, entry)
            if filt is not None:
                ds = []
                for f in filt:
                    rc, d = py_distance(metric, st, dim, (qc, qm, None, np.asarray(f, np.int32), pymeta.query_filter_mag(f)), y)
                    if rc:
                        return rc, [], []
                    ds.append(d)
                if not ds:
                    return 7, [], []
                d = max(ds, key=lambda v: orc.order_key(metric, v))
            else:
                rc, d = py_distance(metric, st, dim, (qc, qm, None, None, 0.0), y)
                if rc:
                    return rc, [], []
            z = [(key(d, y[2]), entry, d)]
        for _, i, d in z:
            y = node(level, i)
            out.append((y[2], EMPTY if py_kind(y[2], y[3], y[4]) == "pseudo" else int(fg.node_row[level][i]), d))
        if level > 0:
            entry = int(fg.child[level][z[0][1]])
    seen, cand = set(), []
    for vid, row, d in out:
        if vid in seen:
            continue
        seen.add(vid)
        if vid == 0xFFFFFFFF or row == EMPTY:
            continue
        cand.append((key(d, vid), vid, row))
    cand.sort(reverse=True)
    cand = cand[: 5 * k]
    mag_q = orc.mag_f32(q)
    final = []
    for _, vid, row in cand:
        cs = orc.dot_f32_simd(q, vecs[row]) / (mag_q * orc.mag_f32(vecs[row]))
        final.append((((orc.order_key(0, cs) << 32) | (~vid & 0xFFFFFFFF)), vid, np.float32(cs)))
    final.sort(reverse=True)
    return 0, [f[1] for f in final[:k]], [f[2] for f in final[:k]]


@pytest.mark.parametrize("st,metric", [(4, 0), (0, 0), (2, 3)])
def test_filtered_search_matches_python_restatement(st, metric):
    vecs, mg = mdgraph.build(n=150, dim=16, storage_type=st, metric=metric, seed=3 + st)
    q, filters = mdgraph.make_queries(vecs, mg, 24, seed=5)
    filters[5] = []                                                     # Some(empty vec): max().unwrap() panics in the reference
    k, ef = 5, 12
    ids, scores, counts, err, _, _ = pymeta.search_batch_md(mg, vecs, q, filters, k, ef_search=ef, shortlist_size=64)
    outcomes = set()
    for i in range(q.shape[0]):
        rc, wids, wsc = py_search(mg, vecs, q[i], filters[i], k, ef, 64)
        outcomes.add(rc)
        if rc:
            assert err[i] == {2: 1, 7: 4}.get(rc, 2) and counts[i] == 0, i
            continue
        assert err[i] == 0 and counts[i] == len(wids), i
        assert ids[i, : len(wids)].tolist() == wids, i
        assert [bits(s) for s in scores[i, : len(wids)]] == [bits(s) for s in wsc], i
    assert 0 in outcomes
    if metric == 0:
        assert 7 in outcomes                                            # the Some(empty) query


# ----------------------------------------------------------------------------- builder with replica nodes
def _oracle_replica_build(pop, st, metric, levels=4, nb=8, nb0=16, efc=32):
    n, dim = pop["vecs"].shape
    root = orc.synth_matrix(99, 1, dim)[0]
    allv = np.concatenate([pop["vecs"], root[None], np.zeros((1, dim), np.float32)])     # rows n = main root, n+1 = pseudo root
    codes, mags = orc.quantize_batch(st, allv)
    rows = np.where(pop["row"] == mdgraph.EMPTY, n + 1, pop["row"]).astype(np.uint32)
    rl = pymeta.ReplicaList(rows, pop["node_id"], pop["base_id"], pop["md_row"], pop["max_level"], n, pop["main_root_md"], n + 1,
                            pop["pseudo_root_md"])
    return pymeta.build_md(metric, st, dim, codes, mags, pop["md_bits"], pop["md_mags"], rl, num_levels=levels, neighbors_count=nb,
                           level0_neighbors_count=nb0, ef_construction=efc), allv


def _kinds(mg, lv):
    ids, mds = mg.node_id[lv], mg.node_md[lv]
    has = (mds != mdgraph.EMPTY) & (mg.md_mags[np.minimum(mds, mg.md_mags.size - 1)] != 0)
    pseudo = (ids >= 0xFFFFFFFF - 257) & (ids <= 0xFFFFFFFF - 2)
    return np.where(~has, pymeta.KIND_BASE, np.where(pseudo, pymeta.KIND_PSEUDO, pymeta.KIND_METADATA))


def test_replica_builder_applies_the_edge_rules():
    # create_node_edges, vector_store.rs:1014-1040: a Metadata node and a Pseudo node are linked only on a perfect match
    # (identical dims -> cs == 1.0), two Metadata nodes never on cs == -1.0 (dims cosine <= 0.99), and the traversal arms keep
    # Base nodes (main root) apart from the pseudo component.
    pop = mdgraph.replica_population(n=300, levels=4)
    (mg, failed), _ = _oracle_replica_build(pop, 4, 0)
    assert failed.sum() == 0
    seen = set()
    for lv in range(5):
        nbc = mg.fg.nbrs(lv)
        adj = mg.fg.adj[lv].reshape(-1, nbc)
        ks = _kinds(mg, lv)
        assert mg.node_id[lv][0] == 0xFFFFFFFF and mg.node_id[lv][1] == pymeta.PSEUDO_ROOT_ID
        for a in range(adj.shape[0]):
            for b in adj[a][adj[a] != mdgraph.EMPTY]:
                pair = (int(ks[a]), int(ks[b]))
                seen.add(pair)
                assert (pair[0] == pymeta.KIND_BASE) == (pair[1] == pymeta.KIND_BASE), (lv, a, b)
                if pymeta.KIND_METADATA in pair and pair != (pymeta.KIND_BASE,) * 2:
                    assert np.array_equal(mg.md_bits[mg.node_md[lv][a]], mg.md_bits[mg.node_md[lv][b]]), (lv, a, b, pair)
        if lv:
            assert np.array_equal(mg.fg.node_row[lv - 1][mg.fg.child[lv]], mg.fg.node_row[lv])
            assert np.array_equal(mg.node_id[lv - 1][mg.fg.child[lv]], mg.node_id[lv])
    assert seen == {(1, 1), (0, 0), (0, 2), (2, 0), (2, 2)}
    # every created node is listed on levels 0..max_level, in list order behind the two roots
    for lv in range(5):
        want = pop["node_id"][pop["max_level"] >= lv]
        assert np.array_equal(mg.node_id[lv][2:], want)


@pytest.mark.parametrize("st,metric", [(4, 0), (0, 0), (2, 3)])
def test_replica_builder_without_metadata_equals_the_plain_builder(st, metric):
    # a replica list whose nodes carry no metadata is the plain collection: same level draws -> same graph as orc_hnsw_build,
    # up to the layout (the replica builder lists an unused pseudo root at index 1 and level 0 in list order)
    n, dim, levels, nb, nb0, efc, seed = 500, 24, 3, 8, 16, 24, 4
    pop = mdgraph.replica_population(n=n, dim=dim, levels=levels)
    vecs = pop["vecs"]
    root = orc.synth_matrix(99, 1, dim)[0]
    fg = pyhnsw.build(metric, st, vecs, root, num_levels=levels, neighbors_count=nb, level0_neighbors_count=nb0,
                      ef_construction=efc, seed=seed)
    max_level = np.zeros(n, np.uint8)
    for lv in range(1, levels + 1):
        max_level[fg.node_row[lv][1:]] = lv
    allv = np.concatenate([vecs, root[None], np.zeros((1, dim), np.float32)])
    codes, mags = orc.quantize_batch(st, allv)
    ids = np.arange(n, dtype=np.uint32)
    rl = pymeta.ReplicaList(ids, ids, ids, np.full(n, mdgraph.EMPTY, np.uint32), max_level, n, mdgraph.EMPTY, n + 1, 1)
    mg, failed = pymeta.build_md(metric, st, dim, codes, mags, pop["md_bits"], pop["md_mags"], rl, num_levels=levels,
                                 neighbors_count=nb, level0_neighbors_count=nb0, ef_construction=efc)
    assert failed.sum() == 0
    for lv in range(levels + 1):
        nbc = nb0 if lv == 0 else nb
        plain = fg.adj[lv].reshape(-1, nbc).astype(np.int64)
        if lv == 0:     # plain level 0: node i = row i, root = n ; replica layout: root 0, pseudo root 1, row i -> 2 + i
            to_new = np.concatenate([np.arange(n) + 2, [0]])
        else:           # plain: root 0 then rows ; replica layout: root 0, pseudo root 1, rows shifted by one
            to_new = np.concatenate([[0], np.arange(1, plain.shape[0]) + 1])
        want = np.full((plain.shape[0] + 1, nbc), mdgraph.EMPTY, dtype=np.int64)
        mapped = np.where(plain == mdgraph.EMPTY, mdgraph.EMPTY, to_new[np.minimum(plain, plain.shape[0] - 1)])
        want[to_new] = mapped
        assert np.array_equal(mg.fg.adj[lv].reshape(-1, nbc).astype(np.int64), want), lv


def test_level_draw_helpers_match_the_reference_vectors():
    from cosdata_b200.api import level_probs, max_insert_level, pseudo_level_probs
    # src/metadata/mod.rs:286-301 (test_pseudo_level_probs)
    assert pseudo_level_probs(9, 128) == [(0.999, 9), (0.99, 8), (0.9, 7), (0.0, 6), (0.0, 5), (0.0, 4), (0.0, 3), (0.0, 2),
                                          (0.0, 1), (0.0, 0)]
    assert pseudo_level_probs(2, 1000) == [(0.0, 2), (0.0, 1), (0.0, 0)]          # more higher levels than levels: all lower
    lp = level_probs(9)                                                          # generate_level_probs(4.0, 9), common.rs:421-429
    assert lp[0] == (1.0 - 4.0 ** -9, 9) and lp[-1] == (0.0, 0) and len(lp) == 10
    assert max_insert_level(0.0, lp) == 0 and max_insert_level(0.75, lp) == 1 and max_insert_level(0.99999999, lp) == 9


# ------------------------------------------------------------------ index_embeddings with replica nodes, pure Python
def py_build_md(pop, codes, mags, st, metric, levels, nb, nb0, efc, shortlist=64):
    """index_embeddings / index_embedding / create_node_edges / ProbNode::add_neighbor (vector_store.rs:714-1070,
    prob_node.rs:210-283) for a replica list, written independently of oracle/metadata_oracle.c: dict-free lists, a heap as a
    sorted list, floats compared through the order key.  Level layout as the oracle's: [0] main root, [1] pseudo root."""
    n, dim = pop["vecs"].shape
    md_bits, md_mags = pop["md_bits"], pop["md_mags"]
    okey = lambda v: orc.order_key(metric, v)
    min_key, max_key = (okey(np.float32(-1.0)), okey(np.float32(2.0))) if metric == 0 else (okey(np.float32(-np.inf)), okey(np.float32(np.inf)))
    L1 = levels + 1
    nbs = [nb0 if lv == 0 else nb for lv in range(L1)]
    # per level: parallel lists
    rows = [[n, n + 1] for _ in range(L1)]
    ids = [[0xFFFFFFFF, pymeta.PSEUDO_ROOT_ID] for _ in range(L1)]
    mds = [[pop["main_root_md"], pop["pseudo_root_md"]] for _ in range(L1)]
    adj = [[[EMPTY] * nbs[lv] for _ in range(2)] for lv in range(L1)]
    sim = [[[0] * nbs[lv] for _ in range(2)] for lv in range(L1)]
    low = [[[0, min_key] for _ in range(2)] for lv in range(L1)]
    child = [[(0 if lv else EMPTY), (1 if lv else EMPTY)] for lv in range(L1)]

    def vd(lv, i):
        m = mds[lv][i]
        return (codes[rows[lv][i]], mags[rows[lv][i]], ids[lv][i], None if m == EMPTY else md_bits[m], 0.0 if m == EMPTY else md_mags[m])

    def kind_of(lv, i):
        y = vd(lv, i)
        return py_kind(y[2], y[3], y[4])

    def add_neighbor(lv, node, nbr, dkey):
        lidx, lkey = low[lv][node]
        if dkey <= lkey:
            return -1
        ok, old = False, EMPTY
        if adj[lv][node][lidx] == EMPTY:
            adj[lv][node][lidx], sim[lv][node][lidx], ok = nbr, dkey, True
        elif dkey > sim[lv][node][lidx]:
            old = adj[lv][node][lidx]
            adj[lv][node][lidx], sim[lv][node][lidx], ok = nbr, dkey, True
        nidx, nkey = 0, max_key
        for s in range(nbs[lv]):
            if adj[lv][node][s] == EMPTY:
                nidx, nkey = s, min_key
                break
            if sim[lv][node][s] < nkey:
                nidx, nkey = s, sim[lv][node][s]
        low[lv][node] = [nidx, nkey]
        if not ok:
            return -1
        if old != EMPTY:
            for s in range(nbs[lv]):
                if adj[lv][old][s] == node:
                    adj[lv][old][s] = EMPTY
                    break
        return lidx

    def traverse(lv, entry, x, self_id):
        nbv = nbs[lv]
        take = min(shortlist, nbv)
        fs = {((self_id >> 6) & (nbv - 1), self_id & 63)}
        y = vd(lv, entry)
        rc, d = py_distance(metric, st, dim, x, y)
        if rc:
            return rc, []
        fs.add(((y[2] >> 6) & (nbv - 1), y[2] & 63))
        heap, res, visited = [((okey(d) << 32) | (~y[2] & 0xFFFFFFFF), entry, d)], [], 0
        while heap:
            heap.sort(reverse=True)
            cur = heap.pop(0)
            if visited >= efc:
                break
            visited += 1
            res.append(cur)
            for s in range(take):
                nbl = adj[lv][cur[1]][s]
                if nbl == EMPTY:
                    continue
                y = vd(lv, nbl)
                bit = ((y[2] >> 6) & (nbv - 1), y[2] & 63)
                if bit in fs:
                    continue
                rc, d = py_distance(metric, st, dim, x, y)
                if rc:
                    return rc, []
                fs.add(bit)
                heap.append(((okey(d) << 32) | (~y[2] & 0xFFFFFFFF), nbl, d))
        res.sort(reverse=True)
        return 0, res[:64]

    failed = []
    for t in range(len(pop["row"])):
        row = n + 1 if pop["row"][t] == EMPTY else int(pop["row"][t])
        m = int(pop["md_row"][t])
        x = (codes[row], mags[row], int(pop["base_id"][t]), None if m == EMPTY else md_bits[m], 0.0 if m == EMPTY else md_mags[m])
        under_pseudo = x[3] is not None and x[4] != 0.0
        entry, parent, node_at, zs, bad = (1 if under_pseudo else 0), EMPTY, {}, {}, False
        max_level = int(pop["max_level"][t])
        for lv in range(levels, -1, -1):
            rc, z = traverse(lv, entry, x, int(pop["node_id"][t]))
            if rc:
                bad = True
                break
            zs[lv] = z
            nxt = child[lv][z[0][1]] if lv else 0
            if lv <= max_level:
                idx = len(rows[lv])
                rows[lv].append(row); ids[lv].append(int(pop["node_id"][t])); mds[lv].append(m)
                adj[lv].append([EMPTY] * nbs[lv]); sim[lv].append([0] * nbs[lv]); low[lv].append([0, min_key]); child[lv].append(EMPTY)
                if parent != EMPTY:
                    child[lv + 1][parent] = idx
                node_at[lv], parent = idx, idx
            entry = nxt
        failed.append(1 if bad else 0)
        if bad:
            continue
        for lv in range(0, min(max_level, levels) + 1):
            node, good = node_at[lv], 0
            nk = kind_of(lv, node)
            for _, nbr, d in zs[lv]:
                if good >= nbs[lv]:
                    break
                if metric == 0:
                    bk = kind_of(lv, nbr)
                    if bk == "pseudo" and nk == "metadata" and d != np.float32(1.0):
                        continue
                    if bk == "metadata" and nk == "metadata" and d == np.float32(-1.0):
                        continue
                i = add_neighbor(lv, node, nbr, okey(d))
                if i >= 0:
                    if add_neighbor(lv, nbr, node, okey(d)) >= 0:
                        good += 1
                    elif adj[lv][node][i] == nbr:
                        adj[lv][node][i] = EMPTY
    return rows, ids, mds, adj, child, failed


@pytest.mark.parametrize("st,metric", [(4, 0), (0, 0)])
def test_replica_builder_matches_python_restatement(st, metric):
    levels, nb, nb0, efc = 3, 4, 8, 12
    pop = mdgraph.replica_population(n=90, dim=16, md_dims=5, levels=levels, n_patterns=5, seed=41)
    (mg, failed), allv = _oracle_replica_build(pop, st, metric, levels=levels, nb=nb, nb0=nb0, efc=efc)
    codes, mags = orc.quantize_batch(st, allv)
    rows, ids, mds, adj, child, pfailed = py_build_md(pop, codes, mags, st, metric, levels, nb, nb0, efc)
    assert failed.tolist() == pfailed
    edges = 0
    for lv in range(levels + 1):
        assert mg.fg.node_row[lv].tolist() == rows[lv] and mg.node_id[lv].tolist() == ids[lv] and mg.node_md[lv].tolist() == mds[lv], lv
        assert mg.fg.adj[lv].reshape(-1, nb0 if lv == 0 else nb).tolist() == adj[lv], lv
        if lv:
            assert mg.fg.child[lv].tolist() == child[lv], lv
        edges += sum(v != EMPTY for r in adj[lv] for v in r)
    assert edges > 3 * len(pop["row"])
