"""Mirror of FitCuboidBFGS.hs (export list FitCuboidBFGS.hs:3-12).

type Cuboid = [Vec3] (8 corners).  The 8-corner functions run in the C++ host mirror (they are O(64) flops in
the reference as well); `fitCuboidToCloudBFGS` is the north-star addition that evaluates the objective and its
gradient over the whole room cloud on the GPU (hs_cuboid_residual_grad) inside a BFGS loop."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _c(points):
    a = np.ascontiguousarray(points, dtype=np.float64).reshape(-1)
    if a.size != 24:
        raise ValueError("Cuboid must have 8 corners")
    return a


def cuboidFromParams(params):
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
    if p.size != 10:
        raise ValueError("bad arguments passed to cuboidFromParams")  # FitCuboidBFGS.hs:112
    out = np.empty((8, 3), np.float64)
    L.load().hs_cuboid_from_params(L.ptr(p), L.ptr(out))
    return out


def errfun(ps, params):
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
    if p.size != 10:
        raise ValueError("bad arguments passed to cuboidFromParams")
    return L.load().hs_errfun(L.ptr(_c(ps)), L.ptr(p))


def errfunClosest(ps, params):
    pts = np.ascontiguousarray(ps, dtype=np.float64).reshape(-1, 3)
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
    if p.size != 10:
        raise ValueError("errfunClosest: bad params")  # FitCuboidBFGS.hs:70
    return L.load().hs_errfun_closest(L.ptr(pts), pts.shape[0], L.ptr(p))


def guessDims(ps):
    out = np.empty(3, np.float64)
    L.load().hs_guess_dims(L.ptr(_c(ps)), L.ptr(out))
    return tuple(out)


def _fit(points, variant, want_path=True):
    pts = _c(points)
    params = np.empty(10, np.float64)
    steps, err = C.c_int32(), C.c_double()
    cap = 4001
    path = np.zeros((cap, 13), np.float64) if want_path else None
    rc = L.load().hs_fit_cuboid(L.ptr(pts), variant, L.ptr(params), C.byref(steps), C.byref(err), L.ptr(path), cap)
    if rc:
        raise L.HsError(rc, "hs_fit_cuboid")
    if want_path:
        cols = 10 if variant == 1 else 13
        flat = path.reshape(-1)
        rows = steps.value if variant != 2 else None
        mat = flat[: cap * cols].reshape(cap, cols)
        n_rows = int(np.count_nonzero(mat[:, 0]))
        path = mat[:n_rows].copy()
    return list(params), steps.value, err.value, path


def fitCuboid(points):
    """-> ([Double], Int, Double, Matrix Double)   FitCuboidBFGS.hs:205-233"""
    return _fit(points, 0)


def fitCuboidFromCenter(points):
    """FitCuboidBFGS.hs:172-184"""
    return _fit(points, 1)


def fitCuboidFromCenterFirst(points):
    """FitCuboidBFGS.hs:188-201"""
    return _fit(points, 2)


def fitCuboidFromCenterFirstError(ps):
    """[Vec3] -> (Double, Int)   FitCuboidBFGS.hs:236-237"""
    _, steps, err, _ = _fit(ps, 2, want_path=False)
    return err, steps


def fitCuboidToCloudBFGS(cloud, initial, max_iter=200, gtol=1e-6):
    """North-star addition: BFGS on f(params) = sum over the room cloud of squared distance to the nearest cuboid
    plane (planes as makePlanesFromCuboid, Main.hs:1852-1874), objective and gradient reduced on the GPU.
    -> (params, f, iterations, evaluations)"""
    return cloud.ctx.fit_cuboid_cloud_bfgs(cloud, initial, max_iter, gtol)


def bfgsMinimize(objective, x0, max_iter=200, gtol=1e-6):
    """The library's BFGS (`hs_bfgs_minimize`) over a Python objective `x -> (f, grad)`: the identical optimiser that
    fitCuboidToCloudBFGS runs over the GPU objective.  -> (x, f, iterations, evaluations)"""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.size
    CB = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))

    def cb(_user, xp, fp, gp):
        try:
            f, g = objective(np.ctypeslib.as_array(xp, shape=(n,)).copy())
            fp[0] = float(f)
            g = np.asarray(g, dtype=np.float64)
            for i in range(n):
                gp[i] = g[i]
            return 0
        except Exception:  # noqa: BLE001
            return 1

    cbk = CB(cb)
    out = np.zeros(n, np.float64)
    f, it, ev = C.c_double(), C.c_int32(), C.c_int32()
    rc = L.load().hs_bfgs_minimize(C.cast(cbk, C.c_void_p), None, L.ptr(x0), n, max_iter, gtol, L.ptr(out), C.byref(f), C.byref(it), C.byref(ev))
    if rc:
        raise L.HsError(rc, "hs_bfgs_minimize")
    return out, f.value, it.value, ev.value


def nmMinimize(objective, x0, step, eps=1e-8, maxit=2000):
    """The library's Nelder-Mead (`hs_nm_minimize`: GSL nmsimplex2's rules as `minimize NMSimplex2 eps maxit x0 f step` drives them,
    FitCuboidBFGS.hs:184,201,233) over a Python objective `x -> f`.  -> (x, f, iterations, evaluations)"""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    n = x0.size
    CB = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double))

    def cb(_user, xp):
        try:
            return float(objective(np.ctypeslib.as_array(xp, shape=(n,)).copy()))
        except Exception:  # noqa: BLE001
            return float("nan")

    cbk = CB(cb)
    out = np.zeros(n, np.float64)
    f, it, ev = C.c_double(), C.c_int32(), C.c_int32()
    rc = L.load().hs_nm_minimize(C.cast(cbk, C.c_void_p), None, L.ptr(x0), L.ptr(step), n, eps, maxit, L.ptr(out), C.byref(f), C.byref(it), C.byref(ev))
    if rc:
        raise L.HsError(rc, "hs_nm_minimize")
    return out, f.value, it.value, ev.value


def fitCuboidToCloudNM(cloud, initial, step, eps=1e-8, maxit=2000):
    """The reference's optimiser over the whole room cloud: NMSimplex2 on f(params) = sum of squared distances to the nearest wall,
    every evaluation one pass of the resident session kernel (initial simplex and shrink steps posted as batches).
    -> (params, f, iterations, evaluations)"""
    return cloud.ctx.fit_cuboid_cloud_nm(cloud, initial, step, eps, maxit)
