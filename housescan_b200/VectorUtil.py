"""Mirror of VectorUtil.hs (VectorUtil.hs:11-19) for device-resident clouds keyed by a coordinate.

kthSmallestBy / kthLargestBy :: (a -> b) -> Int -> v a -> a     with the key restricted to x/y/z of a Vec3,
which is the only use in the reference (removeCeiling, Main.hs:2652-2654).  k is 1-based; k < 1 or k > n
raise with the reference's messages."""
from __future__ import annotations

X, Y, Z = 0, 1, 2


def kthLargestBy(axis: int, k: int, cloud):
    return cloud.ctx.kth_largest(cloud, axis, k)


def kthSmallestBy(axis: int, k: int, cloud):
    return cloud.ctx.kth_smallest(cloud, axis, k)
