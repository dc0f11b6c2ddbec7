"""Mirror of TranslationOptimizer.hs.

lstSqDistances :: Ord a => Map (a,a) Double -> Maybe (Map a Double, RMSE)   (TranslationOptimizer.hs:36-42)

The <= 25-unknown solve stays on the host as in the reference (hs_lstsq_distances, C++ Householder QR);
`None` plays `Nothing` for singular systems.  The per-room inputs it is fed at scale (wall offsets from
millions of inlier points) come from the GPU reductions, see housescan_b200.rooms."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .Bijection import biject


def lstSqDistancesI(dist_map: dict):
    """Map (Int,Int) Double -> Maybe ([Double], RMSE)   (TranslationOptimizer.hs:48-72)"""
    dists = sorted(dist_map.items())  # Map.toList
    n = 1 + max(max(i, j) for (i, j), _ in dists)
    ii = np.array([i for (i, _), _ in dists], np.int32)
    jj = np.array([j for (_, j), _ in dists], np.int32)
    d = np.array([v for _, v in dists], np.float64)
    pos = np.empty(n, np.float64)
    rmse = C.c_double()
    rc = L.load().hs_lstsq_distances(L.ptr(ii), L.ptr(jj), L.ptr(d), len(dists), n, L.ptr(pos), C.byref(rmse))
    if rc == L.HS_ESINGULAR:
        return None
    if rc:
        raise L.HsError(rc, "hs_lstsq_distances")
    return list(pos), rmse.value


def lstSqDistances(dist_map: dict):
    keys = sorted(dist_map.keys())  # Map.keys
    index_of, a_of_index = biject([v for (a, b) in keys for v in (a, b)])
    imap = {}
    for (a, b) in keys:  # Map.mapKeys
        imap[(index_of[a], index_of[b])] = dist_map[(a, b)]
    res = lstSqDistancesI(imap)
    if res is None:
        return None
    pos, rmse = res
    return {a_of_index[i]: pos[i] for i in range(len(pos))}, rmse
