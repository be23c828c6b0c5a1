#!/usr/bin/env python
"""Builds an HNSW graph with the CPU oracle's deterministic builder (reference defaults:
nbrs 32/64, ef_construction 128, 9 layers, config.toml:19-25) and stores the flat arrays for the
HNSW bench / large parity test.  TEST/BENCH INFRASTRUCTURE (uses oracle/).

    python tools/build_hnsw_graph.py --rows 100000 --dim 128 --out bench_data/hnsw_100k_128_f16.npz
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import oracle as orc
from oracle import pyhnsw


def clustered(n, dim, seed, centres=256, sigma=0.35):
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(centres, dim)).astype(np.float32)
    v = (c[rng.integers(0, centres, n)] + sigma * rng.normal(size=(n, dim))).astype(np.float32)
    return (v / (np.abs(v).max() * 1.01)).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--storage", type=int, default=orc.ST_F16)
    ap.add_argument("--out", default="bench_data/hnsw_100k_128_f16.npz")
    a = ap.parse_args()
    vecs = clustered(a.rows, a.dim, 2024)
    root = orc.synth_matrix(31337, 1, a.dim)[0]
    t0 = time.time()
    fg = pyhnsw.build(orc.METRIC_COSINE, a.storage, vecs, root, num_levels=9, neighbors_count=32,
                      level0_neighbors_count=64, ef_construction=128, shortlist_size=64, seed=7)
    dt = time.time() - t0
    print(f"built {a.rows}x{a.dim} in {dt:.1f}s; level counts {fg.cnt.tolist()}")
    d = {"vecs": vecs, "root": root, "entry": fg.entry, "storage": a.storage, "build_s": dt}
    for lv in range(10):
        d[f"node_row{lv}"], d[f"adj{lv}"], d[f"child{lv}"] = fg.node_row[lv], fg.adj[lv], fg.child[lv]
    np.savez(a.out, **d)


if __name__ == "__main__":
    main()
