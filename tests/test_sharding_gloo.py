"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard ranges, global id offsets, the packed-key
all-gather layout and the merge rule.  Per-shard searches are done by the CPU oracle and the merge by the numpy mirror in
cosdata_b200/sharding.py (test infrastructure); on GPUs the same steps run inside cdb_search_batch_sharded
(csrc/shard_group.cu, NCCL; tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cosdata_b200.sharding import gather_and_merge, shard_range

INVALID = 0xFFFFFFFF


def _worker(rank, world, port, n, dim, b, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    corpus = orc.synth_matrix(42, n, dim)                 # every rank can regenerate any row (counter RNG)
    queries = orc.synth_matrix(43, b, dim)
    row0, nloc = shard_range(n, world, rank)
    ids, scores = orc.brute_topk_f32(corpus[row0:row0 + nloc], queries, k, threads=2)
    ids = np.where(ids == INVALID, INVALID, ids + row0).astype(np.uint32)      # id_base = row0

    calls = []

    def all_gather(x):                                    # ONE collective of B*k packed u64 keys per rank
        assert x.dtype == np.uint64
        calls.append(x.shape)
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.int64))
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.stack([o.numpy().view(np.uint64) for o in out])

    m_ids, m_scores = gather_and_merge(ids, scores, world, all_gather)
    assert len(calls) == 1
    want_ids, want_scores = orc.brute_topk_f32(corpus, queries, k, threads=2)
    ok = np.array_equal(m_ids, want_ids) and np.array_equal(m_scores.view(np.uint32), want_scores.view(np.uint32))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_ranges_cover_the_corpus_exactly():
    for n in (1, 7, 10_000_000, 100_000_003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, world, i) for i in range(world)]
            assert r[0][0] == 0 and sum(c for _, c in r) == n
            assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in r) - min(c for _, c in r) <= 1


def test_two_rank_sharded_search_equals_single_index():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 3001, 24, 5, 10, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_short_shards_with_padding():
    # shards smaller than k produce INVALID padding that the merge must ignore
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 13, 8, 3, 10, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
