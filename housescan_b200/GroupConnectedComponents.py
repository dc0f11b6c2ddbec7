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
