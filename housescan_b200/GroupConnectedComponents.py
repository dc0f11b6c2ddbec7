"""Mirror of GroupConnectedComponents.hs (GroupConnectedComponents.hs:16-54).

groupConnectedComponents :: Ord node => [((node,node), a)] -> [[((node,node), a)]]

Vertices are bijected on the host (Bijection.biject), component labels (minimum vertex id per component)
come from the GPU union-find kernels (hs_cc_label), and the edge regrouping reproduces the reference's
output order: components by ascending minimum vertex, edges inside a component in reverse input order
(IntMap.fromListWith (++)), payload of duplicated edges = last one given (Map.fromList)."""
from __future__ import annotations

import numpy as np

from . import default_context
from .Bijection import biject


def groupCCContiguous(edges, ctx=None):
    """[(Int,Int)] -> [[(Int,Int)]] for vertices in a contiguous range (GroupConnectedComponents.hs:39-54)."""
    if not edges:
        return []
    ctx = ctx or default_context()
    src = np.array([e[0] for e in edges], dtype=np.int64)
    dst = np.array([e[1] for e in edges], dtype=np.int64)
    lo = int(min(src.min(), dst.min()))
    n = int(max(src.max(), dst.max())) - lo + 1
    comp, order, ncomp = ctx.group_cc((src - lo).astype(np.uint32), (dst - lo).astype(np.uint32), n)
    out = [[] for _ in range(ncomp)]
    for e in order:
        out[comp[e]].append(edges[e])
    return out


def groupConnectedComponents(edgesData, ctx=None):
    index_of, a_of_index = biject([v for ((i, j), _) in edgesData for v in (i, j)])
    bij = [((index_of[i], index_of[j]), a) for ((i, j), a) in edgesData]
    data = {}
    for e, a in bij:
        data[e] = a
    comps = groupCCContiguous([e for e, _ in bij], ctx)
    return [[((a_of_index[i], a_of_index[j]), data[(i, j)]) for (i, j) in comp] for comp in comps]


# ---- multi-GPU labelling: vertices sharded by contiguous id range, one exchange of cut edges ----------
def split_edges_for_shard(src, dst, lo: int, hi: int):
    """Edges a shard is responsible for (those whose FIRST endpoint it owns), split into interior edges (both endpoints in
    [lo, hi), rebased to the shard) and cut edges (global ids)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    mine = (src >= lo) & (src < hi)
    inside = mine & (dst >= lo) & (dst < hi)
    cut = mine & ~inside
    return (src[inside] - lo).astype(np.uint32), (dst[inside] - lo).astype(np.uint32), np.stack([src[cut], dst[cut]], axis=1)


def merge_shard_labels(local_label, lo: int, cut_edges, root_of):
    """Final labels of one shard.  local_label: this shard's labels as GLOBAL ids (lo + min local id of the local component);
    cut_edges: every shard's cut edges (global ids, K x 2); root_of: {endpoint: its shard-local root} for every cut-edge
    endpoint (gathered from the owning shards).  The contracted graph (one vertex per touched local root) is solved by a
    small union-find keeping the minimum as representative, so the result is the canonical min-index label."""
    parent = {}

    def find(x):
        while parent.setdefault(x, x) != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for u, v in np.asarray(cut_edges, dtype=np.int64).reshape(-1, 2):
        a, b = find(int(root_of[int(u)])), find(int(root_of[int(v)]))
        if a != b:
            parent[max(a, b)] = min(a, b)
    out = np.asarray(local_label, dtype=np.int64).copy()
    touched = np.array(sorted(r for r in parent if lo <= r < lo + len(out)), dtype=np.int64)
    if len(touched):
        final = np.array([find(int(r)) for r in touched], dtype=np.int64)
        idx = np.searchsorted(touched, out)
        idx[idx >= len(touched)] = 0
        hit = touched[idx] == out
        out[hit] = final[idx[hit]]
    return out.astype(np.uint32)


def cc_label_sharded(label_fn, src, dst, n: int, rank: int, world: int, group=None):
    """Canonical labels of the vertices [lo, hi) owned by `rank`.  label_fn(src_u32, dst_u32, n_local) -> local min-id labels
    (Context.cc_label on the GPU).  Exchange: one all-gather of the cut edges and one of the (endpoint, local root) pairs."""
    import torch.distributed as dist

    from .rooms import shard_range

    lo, hi = shard_range(n, rank, world, align=1)
    s, d, cut = split_edges_for_shard(src, dst, lo, hi)
    local = label_fn(s, d, hi - lo).astype(np.int64) + lo
    cuts = [None] * world
    dist.all_gather_object(cuts, cut, group=group)
    all_cut = np.concatenate([c.reshape(-1, 2) for c in cuts]) if cuts else np.zeros((0, 2), np.int64)
    ends = np.unique(all_cut)
    own = ends[(ends >= lo) & (ends < hi)]
    pairs = [None] * world
    dist.all_gather_object(pairs, (own, local[own - lo]), group=group)
    root_of = {}
    for e, r in pairs:
        root_of.update(zip(e.tolist(), r.tolist()))
    return merge_shard_labels(local, lo, all_cut, root_of), (lo, hi)
