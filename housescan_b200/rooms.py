"""Room-level host logic around the GPU reductions: sharding of the concatenated room clouds over ranks and
the wall-alignment driver of Main.optimizeRoomPositions (Main.hs:2089-2168).

Sharding: contiguous POINT ranges, not rooms, so 12 rooms spread evenly over 8 GPUs; every rank
reduces one HS_REC record per room over its range and the records are summed (they are plain sums).  The only
exchange is that all-reduce of nrooms x 24 doubles."""
from __future__ import annotations

import numpy as np

from .GroupConnectedComponents import groupConnectedComponents
from .TranslationOptimizer import lstSqDistances


def shard_range(n: int, rank: int, world: int, align: int = 4):
    """Point range [lo, hi) of `rank`; boundaries are multiples of `align` points (48 B = 3 x float4) so every
    shard starts 16-byte aligned inside the parent buffer."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi


def local_room_offsets(room_offsets, lo: int, hi: int) -> np.ndarray:
    """Room offsets of the global cloud clipped to [lo, hi) and rebased to the shard (rooms outside become empty)."""
    ro = np.asarray(room_offsets, dtype=np.int64)
    return (np.clip(ro, lo, hi) - lo).astype(np.int64)


def shard_frames(nframes: int, rank: int, world: int):
    """Frame range [lo, hi) of `rank` for a replayed depth stream (BASELINE configs[4]): frames are independent, so the stream
    is cut into contiguous, near-equal ranges (the first nframes % world ranks take one frame more)."""
    base, extra = divmod(nframes, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def depth_stream_records_sharded(reduce_fn, frames, poses, rank: int, world: int, group=None):
    """Per-frame 6x6 records of a depth stream sharded by frame range.  `reduce_fn(frames, poses)` evaluates
    a block of frames (on the GPU: Context.backproject_reduce6x6) and returns [n, HS_NE] doubles.  No collective on the data path:
    every rank reduces its own frames; the 232-byte records are gathered in frame order at the end (all_gather of ragged blocks).
    Returns the full [nframes, HS_NE] array on every rank."""
    import torch
    import torch.distributed as dist

    nframes = len(frames)
    lo, hi = shard_frames(nframes, rank, world)
    local = np.asarray(reduce_fn(frames[lo:hi], None if poses is None else poses[lo:hi]), np.float64).reshape(hi - lo, -1)
    if world == 1:
        return local
    width = local.shape[1]
    per = -(-nframes // world)  # pad every block to the largest range so the gather is one fixed-size collective
    buf = torch.zeros(per, width, dtype=torch.float64)
    buf[: hi - lo] = torch.from_numpy(local)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.empty((nframes, width), np.float64)
    for r in range(world):
        a, b = shard_frames(nframes, r, world)
        out[a:b] = parts[r][: b - a].numpy()
    return out


def export_room_ply_sharded(transform_fn, write_part_fn, path: str, n_total: int, rank: int, world: int, has_rgb: bool = False, group=None):
    """Per-room rigid transform + full-resolution .ply export sharded by point range (the room's `roomProj`
    of Main.hs:1716-1730 applied to the FULL-resolution cloud, README.md:16 step 4, Main.hs:2305-2313 delegates this to external
    tools).  Rank r holds the points [lo, hi) = shard_range(n_total, r, world) of the room:
        transform_fn(lo, hi)          -> this rank's transformed shard (on the GPU: Context.transform of its cloud)
        write_part_fn(shard, lo)      -> writes it at its place in the file (Context.write_ply_part / write_ply_part_host)
    Rank 0 creates the file with the header at its final size, a barrier publishes it, then every rank writes its own bytes: no
    collective on the data path, and the file is byte-identical to the single-rank export.  Returns (lo, hi)."""
    from .core import write_ply_begin

    lo, hi = shard_range(n_total, rank, world)
    if rank == 0:
        write_ply_begin(path, n_total, has_rgb)
    if world > 1:
        import torch.distributed as dist

        dist.barrier(group=group)
    write_part_fn(transform_fn(lo, hi), lo)
    if world > 1:
        import torch.distributed as dist

        dist.barrier(group=group)
    return lo, hi


X, Y, Z = 0, 1, 2
SAME = ("Same",)


def Opposite(d: float):
    return ("Opposite", float(d))


def rotationBetweenPlaneEqs(plane1, plane2) -> np.ndarray:
    """Main.hs:1553-1560: the row-vector rotation R (p' = p .* R) that turns plane1's normal into the direction of plane2's"""
    from . import _lib as L

    R = np.empty(9, np.float32)
    L.load().hs_rotation_between_plane_eqs(L.ptr(L.as_f32(plane1, (4,))), L.ptr(L.as_f32(plane2, (4,))), L.ptr(R))
    return R.reshape(3, 3)


def rotatePlaneEqAround(center, R, plane) -> np.ndarray:
    """Main.hs:1571-1578 (re-normalised by mkPlaneEq)"""
    from . import _lib as L

    out = np.empty(4, np.float32)
    L.load().hs_rotate_plane_eq_around(L.ptr(L.as_f32(center, (3,))), L.ptr(L.as_f32(R, (9,))), L.ptr(L.as_f32(plane, (4,))), L.ptr(out))
    return out


def translatePlaneEq(offset, plane) -> np.ndarray:
    """Main.hs:1681-1688"""
    from . import _lib as L

    out = np.empty(4, np.float32)
    L.load().hs_translate_plane_eq(L.ptr(L.as_f32(offset, (3,))), L.ptr(L.as_f32(plane, (4,))), L.ptr(out))
    return out


# ---- roomProj bookkeeping and the room movers (Main.hs:1665-1735) -------------------------------------------------------------------
def _proj(fn_name, *args):
    from . import _lib as L

    out = np.empty(16, np.float32)
    getattr(L.load(), fn_name)(*[L.ptr(a) for a in args], L.ptr(out))
    return out.reshape(4, 4)


def projCompose(a, b):
    """`a .*. b` in Float (Main.hs:1720)"""
    from . import _lib as L
    return _proj("hs_proj_compose", L.as_f32(a, (16,)), L.as_f32(b, (16,)))


def projTranslate(proj, off):
    """`translate4 off proj` (Main.hs:1708)"""
    from . import _lib as L
    return _proj("hs_proj_translate", L.as_f32(proj, (16,)), L.as_f32(off, (3,)))


def projRotateAround(proj, center, R):
    """`translate4 c . (.*. linear R) . translate4 (neg c) $ proj` (Main.hs:1674)"""
    from . import _lib as L
    return _proj("hs_proj_rotate_around", L.as_f32(proj, (16,)), L.as_f32(center, (3,)), L.as_f32(R, (9,)))


class Room:
    """The moving parts of the reference's `Room` (Main.hs:308-316): planes (PlaneEq rows), the cloud, corners and roomProj, moved
    together exactly as rotateRoomAround / translateRoom / projectRoom do (Main.hs:1665-1730).  The cloud lives wherever `engine`
    keeps it: on the GPU the engine is a `Context` and the cloud a device `Cloud` (hs_rotate_around / hs_translate / hs_transform /
    hs_mean_extent); anything with the same four methods will do.  Plane hull points are not carried (the GUI draws them; nothing on
    the compute path reads them)."""

    def __init__(self, engine, cloud, planes, corners=(), proj=None, name="ANON"):
        self.engine, self.cloud, self.name = engine, cloud, name
        self.planes = np.array(planes, np.float32).reshape(-1, 4)
        self.corners = np.array(corners, np.float32).reshape(-1, 3)
        self.proj = np.eye(4, dtype=np.float32) if proj is None else np.array(proj, np.float32).reshape(4, 4)

    def roomMean(self) -> np.ndarray:
        mean, _ = self.engine.mean_extent(self.cloud)  # cloudMean: Double sums on the GPU, then Float
        return np.asarray(mean, np.float64).astype(np.float32)

    def rotateRoomAround(self, center, R):
        c, R = np.asarray(center, np.float32), np.asarray(R, np.float32).reshape(3, 3)
        self.planes = np.stack([rotatePlaneEqAround(c, R, p) for p in self.planes]) if len(self.planes) else self.planes
        self.cloud = self.engine.rotate_around(self.cloud, c, R)
        self.corners = np.stack([_rotate_around(c, R, v) for v in self.corners]) if len(self.corners) else self.corners
        self.proj = projRotateAround(self.proj, c, R)
        return self

    def rotateRoom(self, R):
        return self.rotateRoomAround(self.roomMean(), R)

    def translateRoom(self, off):
        off = np.asarray(off, np.float32)
        self.planes = np.stack([translatePlaneEq(off, p) for p in self.planes]) if len(self.planes) else self.planes
        self.cloud = self.engine.translate(self.cloud, off)
        self.corners = (self.corners + off).astype(np.float32)
        self.proj = projTranslate(self.proj, off)
        return self

    def roomAutoAlignAxis(self, axis):
        """Main.roomAutoAlignAxis (Main.hs:1895-1906): the plane whose normal is most parallel to `axis` (maximumBy: the LAST of equal
        maxima) is turned onto `axis` by rotating the whole room about its mean.  autoAlignFloor = roomAutoAlignAxis (0, 1, 0)."""
        if len(self.planes) == 0:
            return "room has no planes"
        ax = np.asarray(axis, np.float32)
        f32 = np.float32
        dots = [f32(f32(f32(ax[0] * p[0]) + f32(ax[1] * p[1])) + f32(ax[2] * p[2])) for p in self.planes]
        best = 0
        for i in range(1, len(dots)):
            if dots[i] >= dots[best]:  # Data.List.maximumBy keeps the last maximum
                best = i
        from . import _lib as L

        target = np.empty(4, np.float32)  # mkPlaneEq axis 1
        nrm = f32(np.sqrt(f32(f32(f32(ax[0] * ax[0]) + f32(ax[1] * ax[1])) + f32(ax[2] * ax[2]))))
        inv = f32(f32(1.0) / nrm)
        target[:3] = [f32(ax[0] * inv), f32(ax[1] * inv), f32(ax[2] * inv)]
        target[3] = f32(f32(1.0) / nrm)
        self.rotateRoom(rotationBetweenPlaneEqs(self.planes[best], target))
        return None

    def autoAlignFloor(self):
        return self.roomAutoAlignAxis([0.0, 1.0, 0.0])

    def suggestPoints(self, cutoff_factor=1.2):
        """Main.suggestPoints (Main.hs:1521-1538): corners of every plane triple p < q < s (planeCorner), kept when they lie within
        cutoffFactor x (largest distance of a cloud point from the room mean) of the room mean.  Mean and extent are the GPU
        reductions of hs_mean_extent.  The reference filters `p < q, q < s` on the planes' Ord instance (their ids): `self.planes`
        is kept in ascending plane-id order (loadRoom / addPlane append with fresh, growing ids), which makes the index triples
        i < j < k the same enumeration in the same order.  Returns (kept corners [m, 3], number of triples)."""
        mean, maxdist = self.engine.mean_extent(self.cloud)
        m = np.asarray(mean, np.float64).astype(np.float32)
        cutoff = np.float32(np.float32(cutoff_factor) * np.float32(maxdist))
        K, kept, triples = len(self.planes), [], 0
        for i in range(K):
            for j in range(i + 1, K):
                for k in range(j + 1, K):
                    triples += 1
                    c = planeCorner(self.planes[i], self.planes[j], self.planes[k])
                    if c is None:
                        continue
                    with np.errstate(over="ignore", invalid="ignore"):  # nearly parallel planes meet far away: inf / nan never pass the cutoff
                        d = (c - m).astype(np.float32)
                        dist = np.float32(np.sqrt(np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2]))))
                    if dist <= cutoff:
                        kept.append(c)
        return (np.array(kept, np.float32).reshape(-1, 3), triples)

    def projectRoom(self, proj):
        """rotate about the origin, then translate, with R and t read from the rows of `proj`; pattern-fails like the reference
        unless the last column is exactly (0, 0, 0, 1) (Main.hs:1725-1728)"""
        P = np.asarray(proj, np.float32).reshape(4, 4)
        if not np.array_equal(P[:, 3], np.array([0, 0, 0, 1], np.float32)):
            raise ValueError("projectRoom: last column of the projection is not (0,0,0,1)")
        R, off, zero = P[:3, :3], P[3, :3], np.zeros(3, np.float32)
        self.planes = np.stack([translatePlaneEq(off, rotatePlaneEqAround(zero, R, p)) for p in self.planes]) if len(self.planes) else self.planes
        self.cloud = self.engine.transform(self.cloud, P)
        if len(self.corners):  # corners go through the full 4-vector product: (x y z 1) .* proj
            self.corners = np.stack([_row_times_proj(v, P) for v in self.corners])
        self.proj = projCompose(self.proj, P)
        return self


def fitCuboidToRoom(room: "Room", conns=(), room_id=None):
    """Main.fitCuboidToRoom (Main.hs:1814-1849): fit the 10-parameter cuboid to the room's 8 corners (fitCuboidFromCenterFirst, the
    reference's Nelder-Mead), replace the room's planes by the cuboid's six walls (makePlanesFromCuboid) and its corners by the
    cuboid's (corner order re-used positionally), and drop the wall connections that referred to the old planes of this room
    (connections are (axis, relation, (room, wall), (room, wall)) as in connectWalls).
    Returns (message lines, params, steps, err, remaining connections); with fewer than 8 corners nothing changes."""
    from . import FitCuboidBFGS
    from .core import planes_from_cuboid

    log = [f"fitting cuboid to room {room_id if room_id is not None else room.name}"]
    if len(room.corners) < 8:
        log.append("not enough room corners; need 8")
        return log, None, 0, None, list(conns)
    params, steps, err, _ = FitCuboidBFGS.fitCuboidFromCenterFirst(np.asarray(room.corners[:8], np.float64))
    log.append(f"fit cuboid in {steps} steps, RMSE: {float(np.sqrt(err))}")  # sqrt err, not divided by 8 (Main.hs:1827)
    pf = np.asarray(params, np.float64).astype(np.float32).astype(np.float64)  # `map toFloat params` before the planes are made
    room.planes = np.array(planes_from_cuboid(pf), np.float32).reshape(6, 4)
    new = FitCuboidBFGS.cuboidFromParams(params).astype(np.float32)
    k = min(len(new), len(room.corners))
    room.corners = new[:k].copy()  # parallel list comprehension: as many corners as there were ids
    kept = [w for w in conns if not (room_id is not None and (w[2][0] == room_id or w[3][0] == room_id))]
    return log, np.asarray(params, np.float64), int(steps), float(err), kept


def _rotate_around(c, R, p):
    """rotateAround c R p = ((p - c) .* R) + c in Float, left-to-right sums (Main.hs:1582-1583)"""
    f32 = np.float32
    d = (np.asarray(p, f32) - c).astype(f32)
    out = np.empty(3, f32)
    for j in range(3):
        out[j] = f32(f32(f32(d[0] * R[0, j]) + f32(d[1] * R[1, j])) + f32(d[2] * R[2, j]))
    return (out + c).astype(f32)


def _row_times_proj(v, P):
    f32 = np.float32
    x = np.array([v[0], v[1], v[2], 1.0], f32)
    out = np.empty(4, f32)
    for j in range(4):
        acc = f32(x[0] * P[0, j])
        for k in range(1, 4):
            acc = f32(acc + f32(x[k] * P[k, j]))
        out[j] = acc
    if out[3] != f32(1.0):
        raise ValueError("myTrim")  # Main.hs:1723-1724
    return out[:3].copy()


def planeCorner(plane1, plane2, plane3):
    """Main.hs:1413-1430: the point where three planes meet, or None when the 3x3 system is singular (`Nothing`)"""
    from . import _lib as L

    out = np.empty(3, np.float32)
    rc = L.load().hs_plane_corner(L.ptr(L.as_f32(plane1, (4,))), L.ptr(L.as_f32(plane2, (4,))), L.ptr(L.as_f32(plane3, (4,))), L.ptr(out))
    return None if rc == L.HS_ESINGULAR else out


def bestAxis(normal) -> int:
    """`snd $ maximum [(abs (n `dotprod` v), ax) | (v, ax) <- [(vec3X, X), (vec3Y, Y), (vec3Z, Z)]]` (Main.hs:2051): the axis the
    wall normal is most parallel to; on equal components the tuple maximum prefers the LATER axis (Z over Y over X)."""
    n = np.asarray(normal, np.float32)
    return max((abs(np.float32(n[a])), a) for a in (X, Y, Z))[1]


def connectWalls(conns, relation, wall1, wall2, normal1, normal2):
    """Main.connectWalls (Main.hs:2039-2068) as a pure function on the connection list.  wall1 / wall2 identify the two selected
    wall planes (any hashable id; the reference uses plane IDs), normal1 / normal2 are their PlaneEq normals.  Returns
    (new_conns, message): the connection is consed in front (newest first — this is the edge order groupConnectedComponents sees,
    Main.hs:2061) unless the two walls disagree on the axis or the pair is already connected in either order."""
    a1, a2 = bestAxis(normal1), bestAxis(normal2)
    if a1 != a2:
        return list(conns), "Could not guess axis of wall connection"
    if any((wa, wb) in ((wall1, wall2), (wall2, wall1)) for (_, _, wa, wb) in conns):
        return list(conns), None
    return [(a1, relation, wall1, wall2)] + list(conns), None


def optimizeRoomPositions(room_ids, conns, plane_mean, corner_mean, ctx=None):
    """Main.optimizeRoomPositions (Main.hs:2089-2168) as a pure function.

    room_ids   : room ids in the order of `Map.elems sRooms` (ascending id)
    conns      : sConnectedWalls, newest first: [(axis, relation, (room1, wall1), (room2, wall2))]
    plane_mean : (room, wall) -> 3-vector  (reference: mean of the wall's 4 bound corners, Main.hs:1608; at scale:
                 sum p / count of the wall's inlier points from hs_plane_sums / hs_rooms_cuboid_sums)
    corner_mean: room -> 3-vector          (reference: mean of the 8 room corners, Main.hs:2183-2184)
    Returns ({room: translation 3-vector (float32)}, log lines).  Quirks kept: offsets are `o + signum o * wallDistance`;
    the shift uses the first room OF THE AXIS for every component (Main.hs:2121-2123, 2161); Float arithmetic for the
    room-centre bookkeeping; lstSqDistances' "rmse" (TranslationOptimizer.hs:70)."""
    f32 = np.float32
    cm = {r: np.asarray(corner_mean(r), f32) for r in room_ids}
    moved = {r: np.zeros(3, f32) for r in room_ids}
    log = []
    for axis in (X, Y, Z):
        desired = []
        axis_rooms = []
        for (ax, relation, (r1, w1), (r2, w2)) in conns:
            if ax != axis:
                continue
            pm1, pm2 = np.asarray(plane_mean(r1, w1), f32), np.asarray(plane_mean(r2, w2), f32)
            o = f32(f32(pm1[axis] - cm[r1][axis]) - f32(pm2[axis] - cm[r2][axis]))  # rooms as captured before the loop (Main.hs:2091)
            wall_d = f32(relation[1]) if relation[0] == "Opposite" else f32(0)
            desired.append(((r1, r2), float(f32(o + f32(np.sign(o)) * wall_d))))
            axis_rooms.append(r1)
        if not axis_rooms:
            log.append(f"Don't need to align along {'XYZ'[axis]} axis")
            continue
        first_room = axis_rooms[0]
        comps = groupConnectedComponents(desired, ctx)
        log.append(f"Aligning the {'XYZ'[axis]} ({len(comps)} components)")
        for comp in comps:
            res = lstSqDistances(dict(comp))  # Map.fromList: last duplicate wins
            if res is None:
                log.append("WARNING: optimizeRoomPositions singularity error")
                continue
            centers, rmse = res
            log.append(f"Aligned component of {'XYZ'[axis]} axis with RMSE {rmse:.3f}")
            first_c = cm[first_room][axis]  # the captured firstRoom, not the moved one (Main.hs:2161)
            for rid in sorted(centers):
                new_c = f32(f32(centers[rid]) + first_c)
                moved[rid][axis] = f32(new_c - cm[rid][axis])  # a room is in one component per axis
    return moved, log
