"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the nh2/housescan data-parallel point-cloud path.
Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this package; ``housescan_b200`` never does.

Parity status: the reference is Haskell, GHC is absent from this image and the reference does
not build as mounted (``HmatrixUtils`` missing), so there is no ``oracle/_ref``.  The oracle is
pinned against the checks the reference's own sources hold (tests/test_oracle_golden.py); the
point-cloud-scale generalisations (A4, A6 over clouds) have no reference code and are
"parity unpinned by reference".  Third-party arithmetic restated from published definitions:
``vect >= 0.4.7`` (vect_restate.h), GSL ``nmsimplex2`` as driven by ``hmatrix >= 0.15.2``
``Numeric.GSL.Minimization.minimize`` (nm_simplex2 below), ``containers`` Data.Graph/IntMap.

Per-point arithmetic lives in oracle.c (gcc, -ffp-contract=off); list/graph-level semantics
(biject, groupConnectedComponents, lstSqDistances, Nelder-Mead, export formats) are here.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
REC = 24  # orc_cuboid_sums record length
PS = 10  # orc_plane_sums per-plane record length


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "vect_restate.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_backproject_ref.restype = C.c_int64
        _LIB.orc_filter_le.restype = C.c_int64
        _LIB.orc_errfun.restype = C.c_double
        _LIB.orc_errfun_closest.restype = C.c_double
        _LIB.orc_max_distance.restype = C.c_float
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------- A1-A3
def backproject_ref(depth: np.ndarray, w: int, h: int):
    """Main.hs:1296-1313.  Returns (xyz[n_valid,3] f32, mask[w*h] u8)."""
    depth = np.ascontiguousarray(depth, dtype=np.uint16).reshape(-1)
    assert depth.size == w * h
    xyz = np.empty((w * h, 3), np.float32)
    mask = np.empty(w * h, np.uint8)
    n = lib().orc_backproject_ref(_p(depth), w, h, _p(xyz), _p(mask))
    return xyz[:n].copy(), mask


# ---------------------------------------------------------------- A5
def plane_assign(xyz, planes):
    xyz = _f32(xyz).reshape(-1, 3)
    planes = _f32(planes).reshape(-1, 4)
    n = xyz.shape[0]
    a = np.empty(n, np.uint8)
    r = np.empty(n, np.float32)
    lib().orc_plane_assign(_p(xyz), C.c_int64(n), _p(planes), planes.shape[0], _p(a), _p(r))
    return a, r


def project_to_plane(xyz, plane):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty_like(xyz)
    lib().orc_project_to_plane(_p(xyz), C.c_int64(xyz.shape[0]), _p(_f32(plane)), _p(out))
    return out


def mk_plane_eq(abc, d):
    out = np.empty(4, np.float32)
    lib().orc_mk_plane_eq(_p(_f32(abc)), C.c_float(d), _p(out))
    return out


def rotate_plane_eq_around(c, R, eq):
    out = np.empty(4, np.float32)
    lib().orc_rotate_plane_eq_around(_p(_f32(c)), _p(_f32(R).reshape(9)), _p(_f32(eq)), _p(out))
    return out


def translate_plane_eq(off, eq):
    out = np.empty(4, np.float32)
    lib().orc_translate_plane_eq(_p(_f32(off)), _p(_f32(eq)), _p(out))
    return out


# ---------------------------------------------------------------- A6
def planes_from_cuboid(params):
    out = np.empty((6, 4), np.float32)
    lib().orc_planes_from_cuboid(_p(_f64(params)), _p(out))
    return out


def cuboid_from_params(params):
    out = np.empty((8, 3), np.float64)
    lib().orc_cuboid_from_params(_p(_f64(params)), _p(out))
    return out


def cuboid_from_params_rotate_around(params):
    out = np.empty((8, 3), np.float64)
    lib().orc_cuboid_from_params_rotate_around(_p(_f64(params)), _p(out))
    return out


def errfun(pts, params):
    return lib().orc_errfun(_p(_f64(pts).reshape(24)), _p(_f64(params)))


def errfun_closest(pts, params):
    pts = _f64(pts).reshape(-1, 3)
    return lib().orc_errfun_closest(_p(pts), pts.shape[0], _p(_f64(params)))


def guess_dims(pts):
    out = np.empty(3, np.float64)
    lib().orc_guess_dims(_p(_f64(pts).reshape(24)), _p(out))
    return out


def rot_matrix3(axis, ang, dtype=np.float64):
    if dtype == np.float64:
        out = np.empty(9, np.float64)
        lib().orc_rot_matrix3_d(_p(_f64(axis)), C.c_double(ang), _p(out))
    else:
        out = np.empty(9, np.float32)
        lib().orc_rot_matrix3_f(_p(_f32(axis)), C.c_float(ang), _p(out))
    return out.reshape(3, 3)


def rotation_between_normals(n1, n2):
    out = np.empty(9, np.float32)
    lib().orc_rotation_between_normals(_p(_f32(n1)), _p(_f32(n2)), _p(out))
    return out.reshape(3, 3)


def cuboid_sums(xyz, params):
    xyz = _f32(xyz).reshape(-1, 3)
    rec = np.empty(REC, np.float64)
    lib().orc_cuboid_sums(_p(xyz), C.c_int64(xyz.shape[0]), _p(_f64(params)), _p(rec))
    return rec


def cuboid_residual_grad(xyz, params):
    """-> (f, grad[10], counts[6], gscale[10])"""
    xyz = _f32(xyz).reshape(-1, 3)
    f = C.c_double()
    g = np.empty(10, np.float64)
    gs = np.empty(10, np.float64)
    cnt = np.empty(6, np.int64)
    lib().orc_cuboid_residual_grad(_p(xyz), C.c_int64(xyz.shape[0]), _p(_f64(params)), C.byref(f), _p(g), _p(cnt), _p(gs))
    return f.value, g, cnt, gs


def plane_sums(xyz, room_offsets, planes, K):
    xyz = _f32(xyz).reshape(-1, 3)
    ro = np.ascontiguousarray(room_offsets, dtype=np.int64)
    nrooms = ro.size - 1
    planes = _f32(planes).reshape(nrooms, K, 4)
    out = np.empty((nrooms, K, PS), np.float64)
    lib().orc_plane_sums(_p(xyz), _p(ro), nrooms, _p(planes), K, _p(out))
    return out


# ---------------------------------------------------------------- A4
def backproject_reduce6x6(frames, w, h, planes, intr=None, poses=None):
    frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, h * w)
    nf = frames.shape[0]
    planes = _f32(planes).reshape(-1, 4)
    intr_a = _f32(intr) if intr is not None else None
    poses_a = _f32(poses).reshape(nf, 16) if poses is not None else None
    out = np.empty((nf, 29), np.float64)
    lib().orc_backproject_reduce6x6(_p(frames), C.c_int64(nf), w, h, _p(intr_a), _p(poses_a), _p(planes), planes.shape[0], _p(out))
    return out


# ---------------------------------------------------------------- A7 / A9
def point_mean_f32seq(xyz):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty(3, np.float32)
    lib().orc_point_mean_f32seq(_p(xyz), C.c_int64(xyz.shape[0]), _p(out))
    return out


def point_mean_f64(xyz):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty(3, np.float64)
    lib().orc_point_mean_f64(_p(xyz), C.c_int64(xyz.shape[0]), _p(out))
    return out


def max_distance(xyz, m):
    xyz = _f32(xyz).reshape(-1, 3)
    return float(lib().orc_max_distance(_p(xyz), C.c_int64(xyz.shape[0]), _p(_f32(m))))


def scatter3x3(xyz, mean_mode=1):
    xyz = _f32(xyz).reshape(-1, 3)
    m = np.empty(3, np.float32)
    sc = np.empty(6, np.float64)
    lib().orc_scatter3x3(_p(xyz), C.c_int64(xyz.shape[0]), mean_mode, _p(m), _p(sc))
    return m, sc


def fit_plane(xyz, mean_mode=0):
    """Main.hs:1436-1450.  eigSH = LAPACK dsyev, eigenvalues descending; normal = last column
    (smallest eigenvalue); eigenvector sign is arbitrary (parity up to sign)."""
    xyz = _f32(xyz).reshape(-1, 3)
    if xyz.shape[0] < 3:
        raise ValueError(f"fitPlane: {xyz.shape[0]} points given, need at least 3")
    m, sc = scatter3x3(xyz, mean_mode)
    S = np.array([[sc[0], sc[1], sc[2]], [sc[1], sc[3], sc[4]], [sc[2], sc[4], sc[5]]])
    w, v = np.linalg.eigh(S)  # ascending
    nvec = v[:, 0].astype(np.float32)
    eq0 = np.concatenate([mk_plane_eq(nvec, 0.0)[:3], [0.0]]).astype(np.float32)
    # PlaneEq (mkNormal n) d with d = signedDistanceToPlaneEq (PlaneEq (mkNormal n) 0) m
    d = np.float32(np.float32(np.float32(eq0[0] * m[0]) + np.float32(eq0[1] * m[1])) + np.float32(eq0[2] * m[2])) - np.float32(0)
    return np.array([eq0[0], eq0[1], eq0[2], d], np.float32)


# ---------------------------------------------------------------- A8
def rotate_cloud_around(xyz, c, R):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty_like(xyz)
    lib().orc_rotate_cloud_around(_p(xyz), C.c_int64(xyz.shape[0]), _p(_f32(c)), _p(_f32(R).reshape(9)), _p(out))
    return out


def translate_cloud(xyz, off):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty_like(xyz)
    lib().orc_translate_cloud(_p(xyz), C.c_int64(xyz.shape[0]), _p(_f32(off)), _p(out))
    return out


def project_cloud(xyz, M):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty_like(xyz)
    rc = lib().orc_project_cloud(_p(xyz), C.c_int64(xyz.shape[0]), _p(_f32(M).reshape(16)), _p(out))
    if rc:
        raise ValueError("projectRoom: last column of the projection is not (0,0,0,1)")  # Main.hs:1725-1728 pattern failure
    return out


# Proj4 algebra (Float, right-multiply; Main.hs:1674,1708,1720).  4x4 row-major, translation in row 3.
def proj_identity():
    return np.eye(4, dtype=np.float32)


def proj_linear(R):
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = _f32(R).reshape(3, 3)
    return M


def _matmul_f32(A, B):
    """(.*.) in Float: entry = ((a0*b0 + a1*b1) + a2*b2) + a3*b3, no FMA."""
    A = A.astype(np.float32)
    B = B.astype(np.float32)
    out = np.zeros((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            acc = np.float32(A[i, 0] * B[0, j])
            for k in range(1, 4):
                acc = np.float32(acc + np.float32(A[i, k] * B[k, j]))
            out[i, j] = acc
    return out


def proj_compose(A, B):
    return _matmul_f32(A, B)


def proj_translate4(v, M):
    """translate4 v: post-translation (row 3 += v)."""
    T = np.eye(4, dtype=np.float32)
    T[3, :3] = _f32(v)
    return _matmul_f32(M, T)


def proj_rotate_around(c, R, M):
    """translate4 c . (.*. linear R) . translate4 (neg c)   Main.hs:1674"""
    c = _f32(c)
    return proj_translate4(c, _matmul_f32(proj_translate4(-c, M), proj_linear(R)))


# ---------------------------------------------------------------- A12
def kth_largest(keys, k):
    keys = np.asarray(keys)
    base = keys
    stride = keys.strides[0] // 4
    out = C.c_float()
    rc = lib().orc_kth_largest_f32(C.c_void_p(base.ctypes.data), C.c_int64(keys.shape[0]), C.c_int64(stride), C.c_int64(k), C.byref(out))
    if rc == 1:
        raise ValueError("kLargestBy: k must be >= 1 if the vector is not empty")  # VectorUtil.hs:13
    if rc == 2:
        raise ValueError("kLargestBy: k must bet be > length of the vector")  # VectorUtil.hs:14
    return np.float32(out.value)


def filter_le(xyz, axis, limit, rgb=None):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty_like(xyz)
    rgb_a = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3) if rgb is not None else None
    rgb_o = np.empty_like(rgb_a) if rgb_a is not None else None
    m = lib().orc_filter_le(_p(xyz), C.c_int64(xyz.shape[0]), axis, C.c_float(limit), _p(out), _p(rgb_a), _p(rgb_o))
    return (out[:m].copy(), rgb_o[:m].copy() if rgb_o is not None else None)


def remove_ceiling(xyz, rgb=None):
    """Main.hs:2643-2664: nDiscard = n quot 5, yLimit = k-th largest y, keep y <= yLimit."""
    xyz = _f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    if n == 0:
        return xyz.copy(), (rgb.copy() if rgb is not None else None)
    ylim = kth_largest(xyz[:, 1], n // 5)
    return filter_le(xyz, 1, ylim, rgb)


# ---------------------------------------------------------------- A10 / A11
def biject(xs):
    """Bijection.hs:16-32: dense ints in first-occurrence order."""
    index_of = {}
    a_of_index = []
    for x in xs:
        if x not in index_of:
            index_of[x] = len(a_of_index)
            a_of_index.append(x)
    return index_of, a_of_index


def cc_label(src, dst, n_nodes):
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    lab = np.empty(n_nodes, np.uint32)
    lib().orc_cc_label(_p(src), _p(dst), C.c_int64(src.size), C.c_uint32(n_nodes), _p(lab))
    return lab


def group_cc_contiguous(edges):
    """GroupConnectedComponents.hs:39-54 in pure Python (small inputs).  Data.Graph.components =
    dff of the undirected graph, roots in ascending vertex order; compToEdges = IntMap.fromListWith (++)
    => within a component edges appear in REVERSE input order; components in ascending index."""
    if not edges:
        return []
    verts = [v for e in edges for v in e]
    lo, hi = min(verts), max(verts)
    adj = {v: [] for v in range(lo, hi + 1)}
    for a, b in edges:
        adj[a].append(b)
        adj[b].append(a)
    comp_of = {}
    ncomp = 0
    for root in range(lo, hi + 1):
        if root in comp_of:
            continue
        stack = [root]
        comp_of[root] = ncomp
        while stack:
            v = stack.pop()
            for u in adj[v]:
                if u not in comp_of:
                    comp_of[u] = ncomp
                    stack.append(u)
        ncomp += 1
    comp_to_edges = {}
    for (i, j) in edges:
        c = comp_of[i]
        comp_to_edges[c] = [(i, j)] + comp_to_edges.get(c, [])  # fromListWith (++): new ++ old
    return [comp_to_edges[c] for c in sorted(comp_to_edges)]


def group_connected_components(edges_data):
    """GroupConnectedComponents.hs:16-32.  edges_data: list of ((node,node), payload)."""
    idx, unb = biject([v for ((i, j), _) in edges_data for v in (i, j)])
    bij = [((idx[i], idx[j]), a) for ((i, j), a) in edges_data]
    data_map = {}
    for e, a in bij:
        data_map[e] = a  # Map.fromList: last duplicate wins
    comps = group_cc_contiguous([e for e, _ in bij])
    return [[((unb[i], unb[j]), data_map[(i, j)]) for (i, j) in comp] for comp in comps]


# ---------------------------------------------------------------- A13
def lst_sq_distances(dist_map: dict):
    """TranslationOptimizer.hs:36-72.  dist_map: {(a,b): d}.  Returns (positions dict, rmse) or None.
    Rows in Map.toList (sorted key) order AFTER re-keying to bijected indices; biject order is over
    `Map.keys distMap` (sorted original keys)."""
    keys = sorted(dist_map.keys())
    idx, unb = biject([v for (a, b) in keys for v in (a, b)])
    imap = {}
    for (a, b) in keys:  # Map.mapKeys: later (in ascending original-key order) wins on collisions
        imap[(idx[a], idx[b])] = dist_map[(a, b)]
    dists = sorted(imap.items())
    n = 1 + max(max(i, j) for (i, j), _ in dists)
    A = np.zeros((len(dists), n))
    for r, ((i, j), _) in enumerate(dists):
        for p in range(n):
            A[r, p] = -1.0 if p == i else (1.0 if p == j else 0.0)  # multiway-if: i checked first
    A = A[:, 1:]
    b = np.array([d for _, d in dists])
    if A.shape[1] == 0:
        x = np.zeros(0)
    else:
        if np.linalg.matrix_rank(A) < A.shape[1] or A.shape[0] < A.shape[1]:
            return None  # safeLinearSolveLS -> Nothing on singular systems
        x = np.linalg.lstsq(A, b, rcond=None)[0]
    pts = [0.0] + list(x)
    resid = (A @ x if A.shape[1] else np.zeros(len(b))) - b
    rmse = math.sqrt(np.linalg.norm(resid) / len(b))  # quirk: 2-norm not squared (TranslationOptimizer.hs:70)
    return {unb[i]: pts[i] for i in range(n)}, rmse


# ---------------------------------------------------------------- Nelder-Mead (GSL nmsimplex2 as hmatrix drives it)
def nm_simplex2(f, x0, step, eps=1e-8, maxit=2000):
    """Restatement of gsl_multimin_fminimizer_nmsimplex2 (GSL multimin/simplex2.c) driven like hmatrix's
    `minimize NMSimplex2 eps maxit xi f sz` (FitCuboidBFGS.hs:184,201,233): iterate; size = sqrt(S2);
    stop when size < eps or after maxit iterations.  Returns (x_best, path rows [iter, f, size, x...]).
    Trajectory parity with GSL is UNPINNED (GSL absent); end states are what tests compare."""
    x0 = np.asarray(x0, float)
    n = x0.size
    P = n + 1
    x1 = np.tile(x0, (P, 1))
    y1 = np.empty(P)
    y1[0] = f(x0)
    for i in range(n):
        x1[i + 1, i] += step[i]
        y1[i + 1] = f(x1[i + 1])
    center = x1.mean(axis=0)
    S2 = float(np.mean(np.sum((x1 - center) ** 2, axis=1)))

    def corner_move(coeff, corner):
        alpha = (1 - coeff) * P / (P - 1.0)
        beta = (P * coeff - 1.0) / (P - 1.0)
        xc = alpha * center + beta * x1[corner]
        return xc, f(xc)

    def update_point(i, x, val):
        nonlocal center, S2
        delta = x - x1[i]
        xmc = x1[i] - center
        d = float(np.linalg.norm(delta))
        S2 += (2.0 / P) * float(xmc @ delta) + ((P - 1.0) / P) * (d * d / P)
        center = center + delta / P
        x1[i] = x
        y1[i] = val

    path = []
    it = 0
    while True:
        it += 1
        hi = lo = 0
        dhi = dlo = y1[0]
        s_hi, ds_hi = 1, y1[1]
        for i in range(1, P):
            v = y1[i]
            if v < dlo:
                dlo, lo = v, i
            elif v > dhi:
                ds_hi, s_hi = dhi, hi
                dhi, hi = v, i
            elif v > ds_hi:
                ds_hi, s_hi = v, i
        xc, val = corner_move(-1.0, hi)
        if math.isfinite(val) and val < y1[lo]:
            xc2, val2 = corner_move(-2.0, hi)
            if math.isfinite(val2) and val2 < y1[lo]:
                update_point(hi, xc2, val2)
            else:
                update_point(hi, xc, val)
        elif (not math.isfinite(val)) or val > y1[s_hi]:
            if math.isfinite(val) and val <= y1[hi]:
                update_point(hi, xc, val)
            xc2, val2 = corner_move(0.5, hi)
            if math.isfinite(val2) and val2 <= y1[hi]:
                update_point(hi, xc2, val2)
            else:
                for i in range(P):
                    if i != lo:
                        x1[i] = 0.5 * (x1[i] + x1[lo])
                        y1[i] = f(x1[i])
                center = x1.mean(axis=0)
                S2 = float(np.mean(np.sum((x1 - center) ** 2, axis=1)))
        else:
            update_point(hi, xc, val)
        lo = int(np.argmin(y1))
        size = math.sqrt(S2) if S2 > 0 else math.sqrt(float(np.mean(np.sum((x1 - center) ** 2, axis=1))))
        path.append([it, y1[lo], size] + list(x1[lo]))
        if size < eps or it >= maxit:
            break
    return x1[lo].copy(), np.array(path)


def point_mean_d(points):
    """FitCuboidBFGS.hs:80-84 (Double, sequential)."""
    pts = np.asarray(points, float).reshape(-1, 3)
    acc = np.zeros(3)
    for p in pts:
        acc = acc + p
    return acc * (1 / pts.shape[0])


def fit_cuboid_from_center(points, maxit=2000):
    """FitCuboidBFGS.hs:172-184 -> (params[10], steps, err, path)"""
    pts = np.asarray(points, float).reshape(8, 3)
    c = point_mean_d(pts)
    a = guess_dims(pts)[0]
    errf = lambda s: errfun_closest(pts, np.concatenate([c, s]))
    sol, path = nm_simplex2(errf, [a, a, a, 0.1, 0.1, 0.1, 0.1], [a / 10, a / 10, a / 10, 0.1, 0.1, 0.1, 0.1], 1e-8, maxit)
    return np.concatenate([c, sol]), path.shape[0], errf(sol), path


def fit_cuboid_from_center_first(points, maxit=2000):
    """FitCuboidBFGS.hs:188-201"""
    pts = np.asarray(points, float).reshape(8, 3)
    a = guess_dims(pts)[0]
    initial, steps1, _, _ = fit_cuboid_from_center(pts, maxit)
    errf = lambda s: errfun_closest(pts, s)
    sol, path = nm_simplex2(errf, initial, [0.01, 0.01, 0.01, a / 10, a / 10, a / 10, 0.1, 0.1, 0.1, 0.1], 1e-8, maxit)
    return sol, steps1 + path.shape[0], errf(sol), path


def fit_cuboid(points, maxit=2000):
    """FitCuboidBFGS.hs:205-233"""
    pts = np.asarray(points, float).reshape(8, 3)
    a, b, c = guess_dims(pts)
    x, y, z = point_mean_d(pts)
    errf = lambda s: errfun(pts, s)
    sol, path = nm_simplex2(errf, [x, y, z, a, b, c, 0.1, 0.1, 0.1, 0.1], [0.01, 0.01, 0.01, a / 10, a / 10, a / 10, 0.1, 0.1, 0.1, 0.1], 1e-8, maxit)
    return sol, path.shape[0], errf(sol), path


# ---------------------------------------------------------------- export formats (Main.hs:2271-2302)
def haskell_show_float(x) -> str:
    """`show :: Float -> String`: shortest digits that round-trip (floatToDigits), fixed notation for
    0.1 <= |x| < 10^7, otherwise d.ddde<exp>."""
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    sign = "-" if x < 0 else ""
    s = np.format_float_scientific(abs(x), unique=True, trim="-")  # e.g. '9.671569e-02'
    mant, exp = s.split("e")
    digits = mant.replace(".", "")
    e = int(exp) + 1  # value = 0.d1d2... * 10^e
    if e == 0:
        return f"{sign}0.{digits}"
    if 0 < e <= 7:
        ip = digits[:e].ljust(e, "0")
        fp = digits[e:] or "0"
        return f"{sign}{ip}.{fp}"
    d0, rest = digits[0], digits[1:] or "0"
    return f"{sign}{d0}.{rest}e{e - 1}"


def room_projection_to_string(M):
    """Main.hs:2271-2284: transpose to the left-multiplicative form, 16 comma-separated `show`s."""
    T = _f32(M).reshape(4, 4).T
    return ",".join(haskell_show_float(v) for v in T.reshape(-1))


def room_projection_to_xf(M):
    """Main.hs:2289-2302: 4 lines of 4 space-separated `show`s (unlines => trailing newline)."""
    T = _f32(M).reshape(4, 4).T
    return "".join(" ".join(haskell_show_float(v) for v in row) + "\n" for row in T)


def diagonal_pairs(n):
    """Main.hs:2330-2331 Cantor pairs."""
    out = []
    k = 1
    while len(out) < n:
        for a in range(k):
            out.append((a, k - 1 - a))
            if len(out) == n:
                break
        k += 1
    return out


# ------------------------------------------------------------------ room input formats
# Restatement of planeEqsFromFile (Main.hs:1379-1389), loadPCDFileXyzFloat / loadPCDFileXyzRgbNormalFloat (Main.hs:1318-1329)
# and makeInwardFacing (Main.hs:1746-1751).  pcd-loader and attoparsec are not mounted: PCD follows the published v0.7 layout,
# numbers follow attoparsec's documented `double` grammar; parity unpinned by reference fixtures (the reference ships none).
import re as _re

# attoparsec >= 0.11 (`scientifically`: after the integer part a '.' is consumed whenever present, `anyWord8 *> takeWhile isDigit`,
# so "1." is 1); housescan.cabal:31 only asks for >= 0.10.4.0, whose `floaty` backtracked over a dot without digits - a build
# today resolves to the newer grammar, which is the one restated here.
_ATTO_DOUBLE = _re.compile(rb"[+-]?[0-9]+(?:\.[0-9]*)?(?:[eE][+-]?[0-9]+)?")
_ATTO_SPACE = _re.compile(rb"[ \t\n\v\f\r]*")


def plane_eqs_from_text(text):
    if isinstance(text, str):
        text = text.encode()
    i, out = 0, []
    while True:
        vals, j, ok = [], i, True
        for c in range(4):
            m = _ATTO_DOUBLE.match(text, j)
            if not m:
                ok = False
                break
            tok = m.group()
            if b"." in tok and not tok.split(b".")[1][:1].isdigit():
                tok = tok.replace(b".", b".0", 1)  # "1." / "1.e5"
            vals.append(float(tok))
            j = m.end()
            if c < 3:
                j = _ATTO_SPACE.match(text, j).end()
        if not ok:
            break
        out.append(mk_plane_eq(np.array(vals[:3], np.float32), -np.float32(vals[3])))
        i = j
        if text[i:i + 1] == b"\n":
            i += 1
        elif text[i:i + 2] == b"\r\n":
            i += 2
        else:
            break
    if not out:
        raise ValueError("Could not load planes")
    return np.stack(out)


def make_inward_facing(center, plane_means, planes):
    planes = np.array(planes, np.float32).reshape(-1, 4)
    c = np.asarray(center, np.float32)
    for k in range(len(planes)):
        inward = c - np.asarray(plane_means[k], np.float32)
        n = planes[k, :3]
        d = np.float32(np.float32(np.float32(inward[0] * n[0]) + np.float32(inward[1] * n[1])) + np.float32(inward[2] * n[2]))
        if not d > 0:
            planes[k] = -planes[k]
    return planes


def _lzf_decompress(data, out_len):
    out = bytearray()
    i = 0
    while i < len(data):
        ctrl = data[i]
        i += 1
        if ctrl < 32:
            out += data[i:i + ctrl + 1]
            i += ctrl + 1
        else:
            ln = ctrl >> 5
            if ln == 7:
                ln += data[i]
                i += 1
            back = ((ctrl & 0x1F) << 8) + data[i] + 1
            i += 1
            for _ in range(ln + 2):
                out.append(out[-back])
    assert len(out) == out_len
    return bytes(out)


def pcd_load(path):
    """-> (xyz float32 [n,3], colours float32 [n,3] or None)"""
    raw = open(path, "rb").read()
    hdr, pos = {}, 0
    while True:
        e = raw.index(b"\n", pos)
        line = raw[pos:e].decode().strip()
        pos = e + 1
        if not line or line.startswith("#"):
            continue
        k, *v = line.split()
        hdr[k] = v
        if k == "DATA":
            break
    fields, size, typ = hdr["FIELDS"], [int(x) for x in hdr["SIZE"]], hdr["TYPE"]
    count = [int(x) for x in hdr.get("COUNT", ["1"] * len(fields))]
    n = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr["HEIGHT"][0])
    kind = hdr["DATA"][0]
    rgb_name = "rgb" if "rgb" in fields else ("rgba" if "rgba" in fields else None)
    if kind == "ascii":
        toks = raw[pos:].split()
        ncol = sum(count)
        col0 = np.cumsum([0] + count)[:-1]
        tab = [toks[i * ncol:(i + 1) * ncol] for i in range(n)]
        xyz = np.array([[np.float32(float(r[col0[fields.index(a)]])) for a in "xyz"] for r in tab], np.float32).reshape(n, 3)
        bits = None
        if rgb_name:
            f = fields.index(rgb_name)
            if typ[f] == "U":
                bits = np.array([int(r[col0[f]]) for r in tab], np.uint32)
            else:
                bits = np.array([np.float32(float(r[col0[f]])) for r in tab], np.float32).view(np.uint32)
    else:
        dt = np.dtype({"names": fields, "formats": [({"F": "f", "U": "u", "I": "i"}[t] + str(s), c) if c != 1 else {"F": "f", "U": "u", "I": "i"}[t] + str(s)
                                                    for t, s, c in zip(typ, size, count)]})
        if kind == "binary":
            rec = np.frombuffer(raw, dt, n, pos)
            get = lambda name: rec[name]
        else:
            csz, usz = np.frombuffer(raw, np.uint32, 2, pos)
            soa = _lzf_decompress(raw[pos + 8:pos + 8 + int(csz)], int(usz))
            offs = np.cumsum([0] + [s * c * n for s, c in zip(size, count)])
            get = lambda name: np.frombuffer(soa, dt[name], n, int(offs[fields.index(name)]))
        xyz = np.stack([get(a).astype(np.float32) for a in "xyz"], axis=1)
        bits = get(rgb_name).view(np.uint32) if rgb_name else None
    cols = None
    if bits is not None:
        cols = np.stack([((bits >> 16) & 255).astype(np.float32) / np.float32(255), ((bits >> 8) & 255).astype(np.float32) / np.float32(255),
                         (bits & 255).astype(np.float32) / np.float32(255)], axis=1).astype(np.float32)
    return np.ascontiguousarray(xyz), cols


def load_room(directory):
    import os as _os
    xyz, cols = pcd_load(_os.path.join(directory, "cloud_downsampled.pcd"))
    planes = plane_eqs_from_text(open(_os.path.join(directory, "planes.txt"), "rb").read())
    means = np.stack([point_mean_f32seq(pcd_load(_os.path.join(directory, f"cloud_plane_hull{k}.pcd"))[0]) for k in range(len(planes))])
    center = point_mean_f64(xyz).astype(np.float32)
    return xyz, cols, make_inward_facing(center, means, planes)


# ------------------------------------------------------------------ sharded k-th: one radix pass over a shard
def kth_key_image(keys):
    """order-preserving uint32 image of float32 keys (negative: all bits flipped; non-negative: sign bit set)"""
    b = np.ascontiguousarray(keys, np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def kth_shard_hist(keys, pass_no, prefix, mask):
    shift, bits = ((21, 11), (10, 11), (0, 10))[pass_no]
    u = kth_key_image(keys)
    sel = u[(u & np.uint32(mask)) == np.uint32(prefix)]
    h = np.zeros(2048, np.uint32)
    d = (sel >> np.uint32(shift)) & np.uint32((1 << bits) - 1)
    np.add.at(h, d, 1)
    return h


# ------------------------------------------------------------------ planeCorner (Main.hs:1413-1430)
def plane_corner(p1, p2, p3):
    """3x3 solve in Double through LAPACK dgesv (numpy.linalg.solve = the routine behind hmatrix's linearSolve), result in Float;
    None for an exactly singular system (safeLinearSolve -> Nothing)."""
    A = np.array([np.asarray(p, np.float32)[:3] for p in (p1, p2, p3)], np.float64)
    b = np.array([np.float32(p[3]) for p in (p1, p2, p3)], np.float64)
    try:
        return np.linalg.solve(A, b).astype(np.float32)
    except np.linalg.LinAlgError:
        return None
