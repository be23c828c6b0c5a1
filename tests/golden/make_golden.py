"""Generates tests/golden/hotpath_v1.npz -- the committed known-answer vectors of the hot path.

The reference cannot be built or imported in this image (Rust) and its own tests hold properties,
not golden data, so these vectors come from the CPU oracle (oracle/), after tests/test_oracle_arith.py has pinned it
against the re-expressed reference properties and the exact-integer f32 emulation (tests/f32emu.py).  They freeze the
oracle's answers: tests/test_golden.py checks the oracle (CPU, with an independent f32emu re-derivation of a subset) and
the CUDA path through the C ABI (GPU) against the same bytes, so a later change to either side cannot drift silently.

    python tests/golden/make_golden.py        # rewrites the .npz; review the diff of tests/golden/MANIFEST.txt

Cases: D in {8, 31, 32, 33, 128, 768, 1024}; edge values -1, 1.0, +-0, out-of-range clamps,
denormals; a zero-norm row; duplicated rows (ties); every StorageType x DistanceMetric arm incl. the error arms;
one small HNSW graph (oracle builder) with search results.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as orc  # noqa: E402
from oracle import pyhnsw, pymeta  # noqa: E402

DIMS = (8, 31, 32, 33, 128, 768, 1024)
STORAGES = (0, 1, 2, 3, 4, 5)       # UnsignedByte, SubByte1..3, HalfPrecisionFP, FullPrecisionFP
METRICS = (0, 1, 2, 3)              # Cosine, Euclidean, Hamming, DotProduct
K = 5
EDGE = np.array([-1.0, 1.0, 0.0, -0.0, 1.5, -1.5, 0.99999994, -0.99999994, 1e-40, -1e-40, 0.5, -0.5, 0.25, 0.75,
                 2.0 ** -24, -(2.0 ** -24)], dtype=np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def corpus_for(dim):
    n = 16 if dim <= 128 else 8
    m = orc.synth_matrix(0x601D + dim, n, dim).copy()
    m[0, :] = 0.0                                             # zero norm -> CalculationError in the cosine arms
    reps = (dim + EDGE.size - 1) // EDGE.size
    m[1, :] = np.tile(EDGE, reps)[:dim]                       # edge values
    m[3, :] = m[2, :]                                         # exact duplicate -> tie broken by the smaller id
    m[4, :] = -m[2, :]
    q = orc.synth_matrix(0x9E57 + dim, 3, dim).copy()
    q[1, :] = m[2, :]                                         # a query equal to a corpus row
    q[2, : min(dim, EDGE.size)] = EDGE[: min(dim, EDGE.size)]
    return m, q


def distance_table(metric, st, dim, codes, mags, qcodes, qmags):
    val = np.zeros((qcodes.shape[0], codes.shape[0]), dtype=np.float32)
    status = np.zeros(val.shape, dtype=np.int8)
    for i in range(qcodes.shape[0]):
        for j in range(codes.shape[0]):
            rc, d = orc.distance(metric, st, dim, qcodes[i], qmags[i], codes[j], mags[j])
            status[i, j] = rc
            val[i, j] = d if rc == 0 else 0.0
    return val, status


def generate():
    out = {}
    for dim in DIMS:
        m, q = corpus_for(dim)
        out[f"d{dim}/corpus"] = m
        out[f"d{dim}/queries"] = q
        ids, scores = orc.brute_topk_f32(m, q, K)              # finalize_ann_results formula over the whole corpus
        out[f"d{dim}/f32_topk_ids"] = ids
        out[f"d{dim}/f32_topk_score_bits"] = bits(scores)
        for st in STORAGES:
            codes, mags = orc.quantize_batch(st, m)
            qcodes, qmags = orc.quantize_batch(st, q)
            out[f"d{dim}/st{st}/codes"] = codes
            out[f"d{dim}/st{st}/mag_bits"] = bits(mags)
            out[f"d{dim}/st{st}/qcodes"] = qcodes
            out[f"d{dim}/st{st}/qmag_bits"] = bits(qmags)
            for metric in METRICS:
                val, status = distance_table(metric, st, dim, codes, mags, qcodes, qmags)
                out[f"d{dim}/st{st}/m{metric}/value_bits"] = bits(val)
                out[f"d{dim}/st{st}/m{metric}/status"] = status
                if (status == 6).all() or (status == 1).all():
                    continue                                    # arm the reference does not implement
                rc, tids, tscores, terr = orc.brute_topk_codes(metric, st, dim, codes, mags, qcodes, qmags, K)
                assert rc == 0
                out[f"d{dim}/st{st}/m{metric}/topk_ids"] = tids
                out[f"d{dim}/st{st}/m{metric}/topk_score_bits"] = bits(tscores)
                out[f"d{dim}/st{st}/m{metric}/topk_err"] = terr
        # u8 with a non-default values_range (clamp + scale arm of scalar.rs:17-24)
        codes, mags = orc.quantize_batch(0, m, -0.5, 0.75)
        out[f"d{dim}/st0_range/codes"] = codes
        out[f"d{dim}/st0_range/mag_bits"] = bits(mags)

    # ---- quantization:auto value-range sampling (hnsw/mod.rs:202-351)
    base = orc.synth_matrix(0x5A4D, 64, 48)
    for tag, arr, clamp in (("uniform", base, 1.0), ("x019", (base * np.float32(0.19)).astype(np.float32), 1.0),
                            ("pos004", (np.abs(base) * np.float32(0.04)).astype(np.float32), 1.0),
                            ("x03_tight", (base * np.float32(0.3)).astype(np.float32), 0.05)):
        counts, rng_ = orc.sample_values_range(arr, clamp)
        out[f"sampling/{tag}/values"] = arr
        out[f"sampling/{tag}/clamp_counts"] = np.concatenate([np.array([clamp], dtype=np.float64), counts.astype(np.float64)])
        out[f"sampling/{tag}/range_bits"] = bits(np.array(rng_, dtype=np.float32))

    # ---- metadata replica-kind arms of the cosine metric (cosine.rs:34-102), f16 storage, D = 16, 5 metadata dims
    mv = orc.synth_matrix(0x3D7A, 2, 16)
    mcodes, mmags = orc.quantize_batch(4, mv)
    pat_a, pat_b, zeros = np.array([1, 0, 1, 1, 0], np.int32), np.array([0, 1, 1, 0, 0], np.int32), np.zeros(5, np.int32)
    sides = [(None, None), (None, pat_a), (5, zeros), (8, pat_a), (9, pat_b), (0xFFFFFFFF - 257, pat_a), (0xFFFFFFFF - 2, pat_b),
             (0xFFFFFFFF - 258, pat_a), (0xFFFFFFFF, pat_a)]
    table = np.zeros((len(sides), len(sides), 2), dtype=np.uint32)          # [x][y] = (status, value bits)
    for xi, (xid, xb) in enumerate(sides):
        for yi, (yid, yb) in enumerate(sides):
            x = pymeta.VectorData(mcodes[0], mmags[0], xid, xb, 0.0 if xb is None else pymeta.metadata_mag(xb))
            y = pymeta.VectorData(mcodes[1], mmags[1], yid, yb, 0.0 if yb is None else pymeta.metadata_mag(yb))
            rc, val = pymeta.distance_md(0, 4, 16, 5, x, y)
            table[xi, yi] = (rc, bits(np.array([val if rc == 0 else 0.0], dtype=np.float32))[0])
    out["metadata/vectors"] = mv
    out["metadata/arm_table"] = table

    # ---- HNSW: oracle builder (deterministic restatement of index_embedding) + ann_search + re-rank
    n, dim, seed = 400, 24, 7
    rng = np.random.default_rng(2024)
    centres = rng.normal(size=(12, dim)).astype(np.float32)
    vecs = (centres[rng.integers(0, 12, n)] + 0.3 * rng.normal(size=(n, dim))).astype(np.float32)
    vecs = (vecs / (np.abs(vecs).max() * 1.01)).astype(np.float32)
    queries = (vecs[rng.integers(0, n, 6)] + 0.05 * rng.normal(size=(6, dim))).astype(np.float32).clip(-0.999, 0.999)
    root = orc.synth(seed ^ 0x526F6F74, 0, dim)
    out["hnsw/vectors"] = vecs
    out["hnsw/queries"] = queries
    out["hnsw/root_vector"] = root
    out["hnsw/params"] = np.array([4, 8, 16, 32, 64, seed, 24, 5], dtype=np.uint32)  # levels, nb, nb0, efc, shortlist, seed, ef_search, k
    for st, metric in ((4, 0), (0, 0), (2, 3)):
        fg = pyhnsw.build(metric, st, vecs, root, num_levels=4, neighbors_count=8, level0_neighbors_count=16,
                          ef_construction=32, shortlist_size=64, seed=seed)
        tag = f"hnsw/st{st}_m{metric}"
        out[f"{tag}/entry"] = np.array([fg.entry], dtype=np.uint32)
        for lv in range(5):
            out[f"{tag}/L{lv}/node_row"] = fg.node_row[lv]
            out[f"{tag}/L{lv}/adj"] = fg.adj[lv]
            out[f"{tag}/L{lv}/child"] = fg.child[lv]
        ids, scores, counts, err, evals, pops = pyhnsw.search_batch(fg, vecs, queries, 5, ef_search=24, shortlist_size=64)
        out[f"{tag}/result_ids"] = ids
        out[f"{tag}/result_score_bits"] = bits(scores)
        out[f"{tag}/result_counts"] = counts
        out[f"{tag}/evals_pops"] = np.array([evals, pops], dtype=np.uint64)
    return out


def manifest(arrays):
    lines = []
    for k in sorted(arrays):
        a = np.ascontiguousarray(arrays[k])
        lines.append(f"{hashlib.sha256(a.tobytes()).hexdigest()[:16]}  {a.dtype.str:>4} {str(a.shape):>14}  {k}")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    arrays = generate()
    np.savez_compressed(os.path.join(HERE, "hotpath_v1.npz"), **arrays)
    with open(os.path.join(HERE, "MANIFEST.txt"), "w") as f:
        f.write(manifest(arrays))
    sz = os.path.getsize(os.path.join(HERE, "hotpath_v1.npz"))
    print(f"{len(arrays)} arrays, {sz / 1024:.0f} KiB")
