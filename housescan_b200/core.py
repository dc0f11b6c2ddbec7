"""Context / Cloud objects over the C ABI.  Everything numeric happens in libhousescan_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import HS_NE, HS_PS, HS_REC, HsError, as_f32, as_f64, ptr


class Cloud:
    """Device-resident `Vector Vec3` (Main.hs:117-121)."""

    def __init__(self, ctx: "Context", handle: int, keepalive=None):
        self.ctx, self.h, self._keep = ctx, handle, keepalive

    def __len__(self):
        return int(L.load().hs_cloud_size(self.h))

    @property
    def device_ptr(self) -> int:
        return int(L.load().hs_cloud_device_ptr(self.h) or 0)

    def download(self) -> np.ndarray:
        out = np.empty((len(self), 3), np.float32)
        self.ctx._chk(L.load().hs_cloud_download(self.ctx.h, self.h, ptr(out)))
        return out

    def free(self):
        if self.h:
            L.load().hs_cloud_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.free()
        except Exception:
            pass


class EvalSession:
    """A resident multi-evaluation kernel (`hs_eval_session_*`): the cloud and the room offsets are fixed, parameter sets are
    posted one after the other (or many at once) and every evaluation's records come back without a relaunch."""

    def __init__(self, ctx: "Context", cloud: Cloud, room_offsets, allreduce: bool = False):
        self.ctx = ctx
        self._ro = np.ascontiguousarray(room_offsets, dtype=np.int64)
        self.nrooms = self._ro.size - 1
        self._keep = cloud
        self.posted = 0
        h = C.c_void_p()
        ctx._chk(ctx.lib.hs_eval_session_begin(ctx.h, cloud.h, ptr(self._ro), self.nrooms, 1 if allreduce else 0, C.byref(h)))
        self.h = h

    def post(self, params) -> int:
        """enqueue evaluations (params: count x nrooms x 10); returns the sequence number of the last one"""
        p = as_f64(params).reshape(-1, self.nrooms, 10)
        self.ctx._chk(self.ctx.lib.hs_eval_session_post(self.h, ptr(p), p.shape[0]))
        self.posted += p.shape[0]
        return self.posted - 1

    def wait(self, seq: int, want_record: bool = True):
        rec = np.empty((self.nrooms, HS_REC), np.float64) if want_record else None
        self.ctx._chk(self.ctx.lib.hs_eval_session_wait(self.h, seq, ptr(rec)))
        return rec

    def eval(self, params) -> np.ndarray:
        p = as_f64(params, (self.nrooms, 10))
        rec = np.empty((self.nrooms, HS_REC), np.float64)
        self.ctx._chk(self.ctx.lib.hs_eval_session_eval(self.h, ptr(p), ptr(rec)))
        self.posted += 1
        return rec

    def times(self, seq: int):
        """device clock (ns) of evaluation seq: (command seen by the device, records committed)"""
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx._chk(self.ctx.lib.hs_eval_session_times(self.h, seq, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def done(self) -> int:
        return int(self.ctx.lib.hs_eval_session_done(self.h))

    @property
    def device_results_ptr(self) -> int:
        return int(self.ctx.lib.hs_eval_session_device_results(self.h) or 0)

    def stop(self):
        self.ctx._chk(self.ctx.lib.hs_eval_session_stop(self.h))

    def close(self):
        if self.h:
            h, self.h = self.h, None
            self.ctx._chk(self.ctx.lib.hs_eval_session_end(h))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.h:
            h, self.h = self.h, None
            rc = self.ctx.lib.hs_eval_session_end(h)
            if exc[0] is None:
                self.ctx._chk(rc)
        return False


def peer_group_local(ctxs) -> None:
    """ranks = the given contexts (one per device, all in this process): `hs_peer_group_create_local`"""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = L.load().hs_peer_group_create_local(arr, len(ctxs))
    if rc:
        raise HsError(rc, (L.load().hs_last_error(ctxs[0].h) or b"").decode())


class Context:
    """One CUDA device + stream (`hs_ctx`).  Raises HsError(HS_ECUDA) when no sm_100 GPU is present."""

    def __init__(self, device: int = 0):
        lib = L.load()
        h = C.c_void_p()
        rc = lib.hs_ctx_create(device, C.byref(h))
        if rc != L.HS_OK:
            raise HsError(rc, (lib.hs_last_error(None) or b"").decode())
        self.h = h
        self.lib = lib

    def _chk(self, rc: int):
        if rc != L.HS_OK:
            raise HsError(rc, (self.lib.hs_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.lib.hs_ctx_destroy(self.h)
            self.h = None

    # -- plumbing
    def set_stream(self, cuda_stream: int | None):
        self._chk(self.lib.hs_ctx_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def sync(self):
        self._chk(self.lib.hs_ctx_sync(self.h))

    @property
    def sm_count(self) -> int:
        return self.lib.hs_ctx_sm_count(self.h)

    @property
    def launch_count(self) -> int:
        return int(self.lib.hs_ctx_launch_count(self.h))

    def set_mode(self, key: int, value: int):
        self._chk(self.lib.hs_ctx_set_mode(self.h, key, value))

    # -- clouds
    def upload(self, xyz) -> Cloud:
        xyz = as_f32(xyz).reshape(-1, 3)
        h = C.c_void_p()
        self._chk(self.lib.hs_cloud_upload(self.h, ptr(xyz), xyz.shape[0], C.byref(h)))
        return Cloud(self, h)

    def alloc(self, n: int) -> Cloud:
        h = C.c_void_p()
        self._chk(self.lib.hs_cloud_alloc(self.h, n, C.byref(h)))
        return Cloud(self, h)

    def wrap(self, device_ptr: int, n: int, keepalive=None) -> Cloud:
        h = C.c_void_p()
        self._chk(self.lib.hs_cloud_wrap_device(self.h, C.c_void_p(device_ptr), n, C.byref(h)))
        return Cloud(self, h, keepalive)

    def write(self, cloud: Cloud, host_ptr: int, n: int):
        self._chk(self.lib.hs_cloud_write(self.h, cloud.h, C.c_void_p(host_ptr), n))

    # -- (1) depth
    def backproject_ref(self, depth, w: int, h: int):
        depth = np.ascontiguousarray(depth, dtype=np.uint16).reshape(-1)
        if depth.size != w * h:
            raise ValueError("depth frame size != w*h")
        xyz = np.empty((w * h, 3), np.float32)
        mask = np.empty(w * h, np.uint8)
        n = C.c_int64()
        self._chk(self.lib.hs_backproject_ref(self.h, ptr(depth), w, h, ptr(xyz), ptr(mask), C.byref(n)))
        return xyz[: n.value].copy(), mask

    def backproject_reduce6x6(self, frames, w: int, h: int, planes, intr=None, poses=None):
        frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, w * h)
        nf = frames.shape[0]
        planes = as_f32(planes).reshape(-1, 4)
        intr_a = as_f32(intr) if intr is not None else None
        poses_a = as_f32(poses).reshape(nf, 16) if poses is not None else None
        out = np.empty((nf, HS_NE), np.float64)
        self._chk(self.lib.hs_backproject_reduce6x6(self.h, ptr(frames), nf, w, h, ptr(intr_a), ptr(poses_a), ptr(planes), planes.shape[0], ptr(out)))
        return out

    # -- (2) planes
    def plane_assign(self, cloud: Cloud, planes, want_resid=True):
        planes = as_f32(planes).reshape(-1, 4)
        n = len(cloud)
        a = np.empty(n, np.uint8)
        r = np.empty(n, np.float32) if want_resid else None
        self._chk(self.lib.hs_plane_assign(self.h, cloud.h, ptr(planes), planes.shape[0], ptr(a), ptr(r)))
        return a, r

    def cuboid_residual_grad(self, cloud: Cloud, params):
        p = as_f64(params, (10,))
        f = C.c_double()
        g = np.empty(10, np.float64)
        cnt = np.empty(6, np.int64)
        self._chk(self.lib.hs_cuboid_residual_grad(self.h, cloud.h, ptr(p), C.byref(f), ptr(g), ptr(cnt)))
        return f.value, g, cnt

    def rooms_cuboid_sums(self, cloud: Cloud, room_offsets, params):
        ro = np.ascontiguousarray(room_offsets, dtype=np.int64)
        nrooms = ro.size - 1
        p = as_f64(params, (nrooms, 10))
        rec = np.empty((nrooms, HS_REC), np.float64)
        self._chk(self.lib.hs_rooms_cuboid_sums(self.h, cloud.h, ptr(ro), nrooms, ptr(p), ptr(rec)))
        return rec

    def rooms_cuboid_sums_async(self, cloud: Cloud, room_offsets: np.ndarray, params: np.ndarray, d_rec_ptr: int):
        self._chk(self.lib.hs_rooms_cuboid_sums_async(self.h, cloud.h, ptr(room_offsets), room_offsets.size - 1, ptr(params), C.c_void_p(d_rec_ptr)))

    def peer_connect(self, rank: int, world: int, group=None):
        """Join the NVLink peer-memory all-reduce group of this node (one process per GPU): exchanges the CUDA IPC handles
        of the mailboxes through torch.distributed and maps the peers' mailboxes."""
        import torch.distributed as dist

        handle = np.zeros(64, np.uint8)
        self._chk(self.lib.hs_peer_mailbox_create(self.h, rank, world, ptr(handle)))
        gathered = [None] * world
        dist.all_gather_object(gathered, handle.tobytes(), group=group)
        allh = np.frombuffer(b"".join(gathered), dtype=np.uint8).copy()
        self._chk(self.lib.hs_peer_mailbox_connect(self.h, ptr(allh)))
        dist.barrier(group=group)

    def rooms_cuboid_sums_allreduce_async(self, cloud: Cloud, room_offsets: np.ndarray, params: np.ndarray, d_rec_ptr: int):
        self._chk(self.lib.hs_rooms_cuboid_sums_allreduce_async(self.h, cloud.h, ptr(room_offsets), room_offsets.size - 1, ptr(params), C.c_void_p(d_rec_ptr)))

    def eval_session(self, cloud: Cloud, room_offsets, allreduce: bool = False) -> EvalSession:
        return EvalSession(self, cloud, room_offsets, allreduce)

    def plane_sums(self, cloud: Cloud, room_offsets, planes, K: int):
        ro = np.ascontiguousarray(room_offsets, dtype=np.int64)
        nrooms = ro.size - 1
        pl = as_f32(planes).reshape(nrooms, K, 4)
        out = np.empty((nrooms, K, HS_PS), np.float64)
        self._chk(self.lib.hs_plane_sums(self.h, cloud.h, ptr(ro), nrooms, ptr(pl), K, ptr(out)))
        return out

    def scatter3x3(self, cloud: Cloud):
        mean = np.empty(3, np.float64)
        sc = np.empty(6, np.float64)
        self._chk(self.lib.hs_scatter3x3(self.h, cloud.h, ptr(mean), ptr(sc)))
        return mean, sc

    def fit_plane(self, cloud: Cloud):
        out = np.empty(4, np.float32)
        self._chk(self.lib.hs_fit_plane(self.h, cloud.h, ptr(out)))
        return out

    # -- (3) transforms
    def transform(self, cloud: Cloud, m, out: Cloud | None = None) -> Cloud:
        out = out if out is not None else self.alloc(len(cloud))
        self._chk(self.lib.hs_transform(self.h, cloud.h, ptr(as_f32(m, (16,))), out.h))
        return out

    def rotate_around(self, cloud: Cloud, center, R, out: Cloud | None = None) -> Cloud:
        out = out if out is not None else self.alloc(len(cloud))
        self._chk(self.lib.hs_rotate_around(self.h, cloud.h, ptr(as_f32(center, (3,))), ptr(as_f32(R, (9,))), out.h))
        return out

    def translate(self, cloud: Cloud, off, out: Cloud | None = None) -> Cloud:
        out = out if out is not None else self.alloc(len(cloud))
        self._chk(self.lib.hs_translate(self.h, cloud.h, ptr(as_f32(off, (3,))), out.h))
        return out

    def mean_extent(self, cloud: Cloud):
        mean = np.empty(3, np.float64)
        md = C.c_float()
        self._chk(self.lib.hs_mean_extent(self.h, cloud.h, ptr(mean), C.byref(md)))
        return mean, np.float32(md.value)

    def write_ply(self, cloud: Cloud, path: str, rgb=None):
        rgb_a = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3) if rgb is not None else None
        if rgb_a is not None and rgb_a.shape[0] != len(cloud):
            raise ValueError("rgb must have one row per point")
        self._chk(self.lib.hs_write_ply(self.h, cloud.h, ptr(rgb_a), path.encode()))

    def write_ply_part(self, cloud: Cloud, path: str, first: int, n_total: int, rgb=None):
        """this rank's points [first, first + len(cloud)) of a file created by `write_ply_begin` (any rank order)"""
        rgb_a = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3) if rgb is not None else None
        if rgb_a is not None and rgb_a.shape[0] != len(cloud):
            raise ValueError("rgb must have one row per point")
        self._chk(self.lib.hs_write_ply_part(self.h, cloud.h, ptr(rgb_a), path.encode(), first, n_total))

    def write_pcd(self, cloud: Cloud, path: str, rgb=None):
        rgb_a = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3) if rgb is not None else None
        if rgb_a is not None and rgb_a.shape[0] != len(cloud):
            raise ValueError("rgb must have one row per point")
        self._chk(self.lib.hs_write_pcd(self.h, cloud.h, ptr(rgb_a), path.encode()))

    # -- (4) connected components
    def cc_label(self, src, dst, n_nodes: int):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.ascontiguousarray(dst, dtype=np.uint32)
        lab = np.empty(n_nodes, np.uint32)
        self._chk(self.lib.hs_cc_label(self.h, ptr(src), ptr(dst), src.size, n_nodes, ptr(lab)))
        return lab

    def group_cc(self, src, dst, n_nodes: int):
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.ascontiguousarray(dst, dtype=np.uint32)
        comp = np.empty(src.size, np.int32)
        order = np.empty(src.size, np.int64)
        nc = C.c_int32()
        self._chk(self.lib.hs_group_cc(self.h, ptr(src), ptr(dst), src.size, n_nodes, ptr(comp), ptr(order), C.byref(nc)))
        return comp, order, nc.value

    # -- VectorUtil
    def kth_largest(self, cloud: Cloud, axis: int, k: int):
        out = C.c_float()
        self._chk(self.lib.hs_kth_largest(self.h, cloud.h, axis, k, C.byref(out)))
        return np.float32(out.value)

    def kth_smallest(self, cloud: Cloud, axis: int, k: int):
        out = C.c_float()
        self._chk(self.lib.hs_kth_smallest(self.h, cloud.h, axis, k, C.byref(out)))
        return np.float32(out.value)

    def kth_shard_pass(self, cloud: Cloud, axis: int, pass_no: int, prefix: int, mask: int) -> np.ndarray:
        """one radix pass of the sharded k-th over this rank's points: local histogram (2048 bins) of the pass's digit"""
        hist = np.zeros(2048, np.uint32)
        self._chk(self.lib.hs_kth_shard_pass(self.h, cloud.h, axis, pass_no, prefix, mask, ptr(hist)))
        return hist

    def filter_le(self, cloud: Cloud, axis: int, limit: float, colors: Cloud | None = None):
        out = self.alloc(len(cloud))
        cout = self.alloc(len(cloud)) if colors is not None else None
        n = C.c_int64()
        self._chk(self.lib.hs_filter_le(self.h, cloud.h, axis, limit, colors.h if colors is not None else None, out.h, cout.h if cout is not None else None, C.byref(n)))
        return out, cout

    def remove_ceiling(self, cloud: Cloud, colors: Cloud | None = None):
        out = self.alloc(len(cloud))
        cout = self.alloc(len(cloud)) if colors is not None else None
        n = C.c_int64()
        yl = C.c_float()
        self._chk(self.lib.hs_remove_ceiling(self.h, cloud.h, colors.h if colors is not None else None, out.h, cout.h if cout is not None else None, C.byref(n), C.byref(yl)))
        return out, cout, np.float32(yl.value)

    # -- optimiser on the cloud
    def fit_cuboid_cloud_bfgs(self, cloud: Cloud, init, max_iter=200, gtol=1e-6):
        p0 = as_f64(init, (10,))
        out = np.empty(10, np.float64)
        f = C.c_double()
        it = C.c_int32()
        ev = C.c_int32()
        self._chk(self.lib.hs_fit_cuboid_cloud_bfgs(self.h, cloud.h, ptr(p0), max_iter, gtol, ptr(out), C.byref(f), C.byref(it), C.byref(ev)))
        return out, f.value, it.value, ev.value


    def fit_cuboid_cloud_nm(self, cloud: Cloud, init, step, eps=1e-8, maxit=2000):
        """the reference's optimiser (NMSimplex2, FitCuboidBFGS.hs:184,201,233) over the cloud objective, through a session"""
        p0, st = as_f64(init, (10,)), as_f64(step, (10,))
        out = np.empty(10, np.float64)
        f = C.c_double()
        it = C.c_int32()
        ev = C.c_int32()
        self._chk(self.lib.hs_fit_cuboid_cloud_nm(self.h, cloud.h, ptr(p0), ptr(st), eps, maxit, ptr(out), C.byref(f), C.byref(it), C.byref(ev)))
        return out, f.value, it.value, ev.value


# ---- host-only helpers (no device needed) -------------------------------------------------------------------
def write_ply_begin(path: str, n_total: int, has_rgb: bool = False) -> None:
    """create the .ply with its header at the final size (ONE caller, before any `write_ply_part`)"""
    rc = L.load().hs_write_ply_begin(path.encode(), n_total, 1 if has_rgb else 0)
    if rc:
        raise HsError(rc, f"hs_write_ply_begin: cannot write {path}")


def write_ply_part_host(path: str, xyz, first: int, n_total: int, rgb=None) -> None:
    xyz = as_f32(xyz).reshape(-1, 3)
    rgb_a = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3) if rgb is not None else None
    rc = L.load().hs_write_ply_part_host(path.encode(), ptr(xyz), ptr(rgb_a), first, xyz.shape[0], n_total)
    if rc:
        raise HsError(rc, f"hs_write_ply_part_host: cannot write {path}")


def planes_from_cuboid(params) -> np.ndarray:
    out = np.empty((6, 4), np.float32)
    rc = L.load().hs_planes_from_cuboid(ptr(as_f64(params, (10,))), ptr(out))
    if rc:
        raise HsError(rc, "hs_planes_from_cuboid")
    return out


def cuboid_grad_from_sums(params, rec):
    f = C.c_double()
    g = np.empty(10, np.float64)
    cnt = np.empty(6, np.int64)
    rc = L.load().hs_cuboid_grad_from_sums(ptr(as_f64(params, (10,))), ptr(as_f64(rec, (HS_REC,))), C.byref(f), ptr(g), ptr(cnt))
    if rc:
        raise HsError(rc, "hs_cuboid_grad_from_sums")
    return f.value, g, cnt


def proj_to_string(m) -> str:
    buf = C.create_string_buffer(1024)
    rc = L.load().hs_proj_to_string(ptr(as_f32(m, (16,))), buf, 1024)
    if rc:
        raise HsError(rc, "hs_proj_to_string")
    return buf.value.decode()


def proj_to_xf(m) -> str:
    buf = C.create_string_buffer(1024)
    rc = L.load().hs_proj_to_xf(ptr(as_f32(m, (16,))), buf, 1024)
    if rc:
        raise HsError(rc, "hs_proj_to_xf")
    return buf.value.decode()
