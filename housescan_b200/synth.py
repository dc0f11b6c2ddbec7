"""Seeded synthetic inputs for the five BASELINE.json configurations.
numpy only; used by tests/ and bench.py.  Nothing here is on the measured path."""
from __future__ import annotations

import math

import numpy as np


def quat_from_axis_angle(axis, deg):
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    h = math.radians(deg) / 2
    return np.array([math.cos(h), *(math.sin(h) * a)])


def rot_rows_from_quat(q):
    """Rows of R = rightOrthoU(mkU q) (transpose of the standard quaternion matrix), float64."""
    a, b, c, d = np.asarray(q, float) / np.linalg.norm(q)
    L = np.array(
        [
            [a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c],
            [2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b],
            [2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d],
        ]
    )
    return L.T


def cuboid_room_cloud(n, params, sigma=0.005, seed=2, rng=None):
    """n points uniform on the 6 faces (area weighted) of the cuboid `params`, normal noise sigma (metres)."""
    rng = rng or np.random.default_rng(seed)
    p = np.asarray(params, float)
    c, dims, R = p[:3], p[3:6], rot_rows_from_quat(p[6:])
    a, b, cc = dims
    areas = np.array([b * cc, b * cc, a * cc, a * cc, a * b, a * b])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    u = rng.uniform(-0.5, 0.5, size=(n, 3)) * dims
    axis = face // 2
    sign = np.where(face % 2 == 0, 1.0, -1.0)
    u[np.arange(n), axis] = sign * dims[axis] / 2 + rng.normal(0, sigma, size=n)
    return (u @ R + c).astype(np.float32), face.astype(np.uint8)


C1_PARAMS = np.concatenate([[0.3, -0.2, 4.0], [5.0, 2.6, 4.0], quat_from_axis_angle([1, 2, 3], 20.0)])
KINFU_INTR = np.array([525.0, 525.0, 319.5, 239.5], np.float32)


def render_depth_frame(params=C1_PARAMS, w=640, h=480, cam_pos=None, cam_R=None, intr=KINFU_INTR, sigma_mm=3.0, invalid_frac=0.02, seed=1, rng=None):
    """C1: ray-cast the inside of the cuboid room from a pinhole camera -> uint16 depth (mm).
    cam_R rows = camera axes in world (row-vector convention: p_world = p_cam @ cam_R + cam_pos)."""
    rng = rng or np.random.default_rng(seed)
    p = np.asarray(params, float)
    c, dims, R = p[:3], p[3:6], rot_rows_from_quat(p[6:])
    cam_pos = c if cam_pos is None else np.asarray(cam_pos, float)
    cam_R = np.eye(3) if cam_R is None else np.asarray(cam_R, float)
    fx, fy, cx, cy = [float(v) for v in intr]
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    rays_cam = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs, float)], -1).reshape(-1, 3)
    rays = rays_cam @ cam_R  # world directions (z_cam = 1 => t is camera depth)
    o_loc = (cam_pos - c) @ R.T
    d_loc = rays @ R.T
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (dims / 2 - o_loc) / d_loc
        t2 = (-dims / 2 - o_loc) / d_loc
    t = np.where(d_loc > 0, t1, t2)
    t = np.where(d_loc == 0, np.inf, t)
    depth = t.min(axis=1)  # first wall hit from inside
    mm = depth * 1000.0 + rng.normal(0, sigma_mm, size=depth.shape)
    mm = np.clip(np.rint(mm), 1, 65535)
    mm[rng.random(mm.shape) < invalid_frac] = 0
    return mm.astype(np.uint16).reshape(h, w)


def diagonal_pairs(n):
    out, k = [], 1
    while len(out) < n:
        for a in range(k):
            out.append((a, k - 1 - a))
            if len(out) == n:
                break
        k += 1
    return out


def apartment(n_rooms=12, pts_per_room=8_333_334, seed=3, sigma=0.005, room_dims=(5.0, 2.6, 4.0)):
    """C3: rooms on the first Cantor pairs x 6 m (Main.hs:2330-2331, :2504), per-room jitter U[-0.2,0.2] m.
    Returns (xyz [N,3] f32, room_offsets int64[n_rooms+1], params [n_rooms,10])."""
    rng = np.random.default_rng(seed)
    params = np.zeros((n_rooms, 10))
    clouds = []
    offs = [0]
    for r, (gx, gz) in enumerate(diagonal_pairs(n_rooms)):
        jit = rng.uniform(-0.2, 0.2, size=3)
        params[r, :3] = np.array([6.0 * gx, 0.0, 6.0 * gz]) + jit
        params[r, 3:6] = np.asarray(room_dims) + rng.uniform(-0.3, 0.3, size=3)
        params[r, 6:] = quat_from_axis_angle([0, 1, 0], rng.uniform(-3, 3))
        pts, _ = cuboid_room_cloud(pts_per_room, params[r], sigma=sigma, rng=rng)
        clouds.append(pts)
        offs.append(offs[-1] + pts.shape[0])
    return np.concatenate(clouds), np.array(offs, np.int64), params


def voxel_building_graph(nx=64, ny=24, nz=64, storeys=3, seed=4):
    """C4 (small-scale): plane-inlier voxels of a multi-storey building (floors + outer walls of each storey),
    edges = 6-neighbour adjacency between occupied voxels with the same plane id, emitted in lexicographic voxel
    order.  Returns (src, dst uint32, n_nodes, plane_id per node)."""
    rng = np.random.default_rng(seed)
    occ = -np.ones((nx, ny, nz), np.int32)
    sh = ny // storeys
    for s in range(storeys):
        y0 = s * sh
        occ[:, y0, :] = 10 * s + 0  # floor slab
        occ[0, y0 + 1 : y0 + sh, :] = 10 * s + 1
        occ[nx - 1, y0 + 1 : y0 + sh, :] = 10 * s + 2
        occ[1 : nx - 1, y0 + 1 : y0 + sh, 0] = 10 * s + 3
        occ[1 : nx - 1, y0 + 1 : y0 + sh, nz - 1] = 10 * s + 4
    holes = rng.random(occ.shape) < 0.03  # scan dropouts split some components
    occ[holes] = -1
    idx = -np.ones(occ.shape, np.int64)
    coords = np.argwhere(occ >= 0)  # lexicographic (x, y, z)
    idx[tuple(coords.T)] = np.arange(coords.shape[0])
    src, dst = [], []
    for dx, dy, dz in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
        a = occ[: nx - dx, : ny - dy, : nz - dz]
        b = occ[dx:, dy:, dz:]
        m = (a >= 0) & (a == b)
        ia = idx[: nx - dx, : ny - dy, : nz - dz][m]
        ib = idx[dx:, dy:, dz:][m]
        src.append(ia)
        dst.append(ib)
    src = np.concatenate(src)
    dst = np.concatenate(dst)
    order = np.lexsort((dst, src))
    return src[order].astype(np.uint32), dst[order].astype(np.uint32), int(coords.shape[0]), occ[tuple(coords.T)]


def depth_stream(n_frames, w=640, h=480, params=C1_PARAMS, seed=5):
    """C5: the C1 renderer along a circular trajectory inside the room.  Returns (frames u16 [n,h,w], poses f32 [n,16])."""
    rng = np.random.default_rng(seed)
    p = np.asarray(params, float)
    frames = np.empty((n_frames, h, w), np.uint16)
    poses = np.empty((n_frames, 16), np.float32)
    for i in range(n_frames):
        ang = 2 * math.pi * i / max(n_frames, 1)
        pos = p[:3] + np.array([0.8 * math.cos(ang), 0.1 * math.sin(2 * ang), 0.8 * math.sin(ang)])
        ca, sa = math.cos(ang), math.sin(ang)
        cam_R = np.array([[ca, 0, -sa], [0, 1, 0], [sa, 0, ca]])  # yaw about Y
        frames[i] = render_depth_frame(p, w, h, cam_pos=pos, cam_R=cam_R, rng=rng)
        M = np.eye(4)
        M[:3, :3] = cam_R
        M[3, :3] = pos
        poses[i] = M.reshape(-1).astype(np.float32)
    return frames, poses
