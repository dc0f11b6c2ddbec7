"""Mirror of VectorUtil.hs (VectorUtil.hs:11-19) for device-resident clouds keyed by a coordinate.

kthSmallestBy / kthLargestBy :: (a -> b) -> Int -> v a -> a     with the key restricted to x/y/z of a Vec3,
which is the only use in the reference (removeCeiling, Main.hs:2652-2654).  k is 1-based; k < 1 or k > n
raise with the reference's messages."""
from __future__ import annotations

import numpy as np

X, Y, Z = 0, 1, 2


def kthLargestBy(axis: int, k: int, cloud):
    return cloud.ctx.kth_largest(cloud, axis, k)


def kthSmallestBy(axis: int, k: int, cloud):
    return cloud.ctx.kth_smallest(cloud, axis, k)


# ---- point ranges on several GPUs --------------------------------------------------------------------------------
SHARD_PASSES = ((21, 11), (10, 11), (0, 10))  # (shift, bits) of the three MSB radix passes over the order-preserving key image


def kth_sharded(hist_fn, k: int, largest: bool, group=None):
    """k-th order statistic (1-based) of a key that is sharded over the ranks of `group`.  `hist_fn(pass_no, prefix, mask)` returns
    THIS rank's 2048-bin histogram of the pass's digit over its keys with (key & mask) == prefix (on the GPU:
    Context.kth_shard_pass).  Per pass the histograms are summed over the ranks (the path's only exchange: 2048 counters), every
    rank picks the same digit, and after three passes the prefix is the key of the answer.  Returns the float32 value.
    k out of range raises with the reference's messages (VectorUtil.hs:13-14)."""
    import torch
    import torch.distributed as dist

    from . import _lib as L

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    prefix = mask = 0
    k_rem = int(k)
    for pass_no, (shift, bits) in enumerate(SHARD_PASSES):
        hist = torch.from_numpy(np.asarray(hist_fn(pass_no, prefix, mask)).astype(np.int64))
        if multi:
            dist.all_reduce(hist, group=group)
        h = hist.numpy()[: 1 << bits]
        if pass_no == 0:
            n = int(h.sum())
            if k < 1:
                raise ValueError("kLargestBy: k must be >= 1 if the vector is not empty")   # VectorUtil.hs:13
            if k > n:
                raise ValueError("kLargestBy: k must bet be > length of the vector")        # VectorUtil.hs:14 (sic)
        order = h[::-1] if largest else h
        cum = np.cumsum(order)
        pos = int(np.searchsorted(cum, k_rem, side="left"))  # first bin (in scan order) whose cumulative count reaches k_rem
        k_rem -= int(cum[pos - 1]) if pos > 0 else 0
        digit = ((1 << bits) - 1 - pos) if largest else pos
        prefix |= digit << shift
        mask |= ((1 << bits) - 1) << shift
    return np.float32(L.load().hs_kth_float_of_key(prefix))


def kthLargestBySharded(axis: int, k: int, cloud, group=None):
    """kthLargestBy over a cloud whose point ranges live on the ranks of `group` (each rank passes its own shard)"""
    return kth_sharded(lambda p, pre, m: cloud.ctx.kth_shard_pass(cloud, axis, p, pre, m), k, True, group)


def kthSmallestBySharded(axis: int, k: int, cloud, group=None):
    return kth_sharded(lambda p, pre, m: cloud.ctx.kth_shard_pass(cloud, axis, p, pre, m), k, False, group)


def remove_ceiling_sharded(n_local: int, hist_fn, filter_fn, group=None):
    """removeCeiling (Main.hs:2643-2664) over a cloud sharded by point range: k = n `quot` 5 of the GLOBAL point count, yLimit = the
    k-th largest y over all ranks (kth_sharded), then every rank keeps its own points with y <= yLimit in input order
    (`filter_fn(y_limit)` -> whatever the rank keeps, e.g. Context.filter_le).  Returns (kept, y_limit, first_index): first_index is the
    position of this rank's first kept point in the concatenated output (exclusive scan of the kept counts over the ranks), so the
    ranks can write one file / one buffer without another pass.  An empty global cloud returns (None, None, 0) like the V.null guard."""
    import torch
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    n = torch.tensor([int(n_local)], dtype=torch.int64)
    if multi:
        dist.all_reduce(n, group=group)
    n_global = int(n.item())
    if n_global == 0:
        return None, None, 0
    y_limit = kth_sharded(hist_fn, n_global // 5, True, group)  # n < 5 => k = 0 => raises like the reference
    kept = filter_fn(float(y_limit))
    first = 0
    if multi:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(kept)], dtype=torch.int64), group=group)
        first = int(sum(int(c.item()) for c in counts[:rank]))
    return kept, y_limit, first


def removeCeilingSharded(cloud, colors=None, group=None):
    """this rank's shard of the cloud in, this rank's kept points out: (Cloud, colours or None, y_limit, first_index)"""
    ctx = cloud.ctx
    res = {}

    def flt(y_limit):
        out, cout = ctx.filter_le(cloud, Y, y_limit, colors)
        res["c"] = cout
        return out

    kept, y_limit, first = remove_ceiling_sharded(len(cloud), lambda p, pre, m: ctx.kth_shard_pass(cloud, Y, p, pre, m), flt, group)
    return kept, res.get("c"), y_limit, first
