"""Row-sharding of a corpus over the ranks of one node and the per-shard top-k merge.

The reference is single-process; the natural B200 form is one index shard per GPU (contiguous id
ranges), the same query batch on every rank, a local search, ONE all-gather of the per-shard
top-k (B*k*(u32 id, f32 score) per rank -- latency bound on NVLink) and a k-way merge with the
same ordering rule (better score, then smaller id).  The merge itself is the CUDA kernel behind
cdb_merge_topk_device; `merge_fn` exists so the host logic can be exercised on CPU with gloo.
"""
import numpy as np


def shard_range(n_rows, world, rank):
    """contiguous row range [row0, row0+n) of `rank`; sizes differ by at most one row"""
    base, rem = divmod(int(n_rows), int(world))
    n = base + (1 if rank < rem else 0)
    row0 = rank * base + min(rank, rem)
    return row0, n


def gather_and_merge(local_ids, local_scores, world, all_gather_fn, merge_fn):
    """local_ids/local_scores: [B, k] per-shard results with GLOBAL ids (id_base already added).
    all_gather_fn(x) -> [world, ...] stacked over ranks; merge_fn(ids[world,B,k], scores[world,B,k]) -> (ids[B,k], scores[B,k])"""
    if world == 1:
        return local_ids, local_scores
    g_ids = all_gather_fn(local_ids)
    g_scores = all_gather_fn(local_scores)
    return merge_fn(g_ids, g_scores)


def cuda_merge_fn(lib, device, metric, stream_ptr):
    """merge through cdb_merge_topk_device (torch tensors on `device`)"""
    import torch

    def merge(g_ids, g_scores):
        world, b, k = g_ids.shape
        out_ids = torch.empty((b, k), dtype=torch.int32, device=g_ids.device)
        out_scores = torch.empty((b, k), dtype=torch.float32, device=g_ids.device)
        rc = lib.cdb_merge_topk_device(device, int(metric), g_ids.data_ptr(), g_scores.data_ptr(), world, b, k,
                                       out_ids.data_ptr(), out_scores.data_ptr(), stream_ptr)
        if rc != 0:
            raise RuntimeError(f"cdb_merge_topk_device failed: {rc}")
        return out_ids, out_scores

    return merge
