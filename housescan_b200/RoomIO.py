"""Mirror of the reference's room input functions (they live in Main.hs, not in a helper module):

    planeEqsFromFile :: FilePath -> IO [PlaneEq]                      Main.hs:1379-1389
    cloudFromFile    :: State -> FilePath -> IO Cloud                 Main.hs:1332-1345
    planesFromDir    :: State -> FilePath -> IO [Plane]               Main.hs:1391-1404
    loadRoom         :: State -> FilePath -> IO Room                  Main.hs:1740-1765

The `State` argument of the reference only hands out object IDs for the GUI and is dropped here.  Clouds come back as device
handles (`Cloud`), planes as float32 rows `nx ny nz d`.  Errors keep the reference's messages."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from ._lib import HsError, ptr
from .core import Cloud, Context

MAX_PLANES = 4096


def planeEqsFromText(text: bytes | str) -> np.ndarray:
    if isinstance(text, str):
        text = text.encode()
    out = np.empty((MAX_PLANES, 4), np.float32)
    n = C.c_int32()
    rc = L.load().hs_plane_eqs_from_text(text, len(text), ptr(out), MAX_PLANES, C.byref(n))
    if rc == L.HS_EIO:
        raise HsError(rc, "Could not load planes")  # Main.hs:1388
    if rc != L.HS_OK:
        raise HsError(rc, "planeEqsFromFile: too many planes")
    return out[: n.value].copy()


def planeEqsFromFile(path: str) -> np.ndarray:
    with open(path, "rb") as fh:
        return planeEqsFromText(fh.read())


def cloudFromFile(ctx: Context, path: str):
    """-> (Cloud, colours Cloud or None): `ManyColors` when the file carries rgb, else the reference's `OneColor` red."""
    cl, col = C.c_void_p(), C.c_void_p()
    ctx._chk(ctx.lib.hs_cloud_from_pcd(ctx.h, os.fsencode(path), C.byref(cl), C.byref(col)))
    return Cloud(ctx, cl), (Cloud(ctx, col) if col.value else None)


def pcdInfo(path: str):
    n, rgb, kind = C.c_int64(), C.c_int32(), C.c_int32()
    rc = L.load().hs_pcd_info(os.fsencode(path), C.byref(n), C.byref(rgb), C.byref(kind))
    if rc != L.HS_OK:
        raise HsError(rc, f"{path}: not a readable PCD file")
    return n.value, bool(rgb.value), ("ascii", "binary", "binary_compressed")[kind.value]


def makeInwardFacing(room_center, plane_means, planes) -> np.ndarray:
    planes = np.array(planes, np.float32).reshape(-1, 4)
    means = np.ascontiguousarray(plane_means, np.float32).reshape(-1, 3)
    center = np.ascontiguousarray(room_center, np.float32).reshape(3)
    rc = L.load().hs_make_inward_facing(ptr(center), ptr(means), ptr(planes), len(planes))
    if rc != L.HS_OK:
        raise HsError(rc, "makeInwardFacing: bad arguments")
    return planes


def loadRoom(ctx: Context, directory: str):
    """-> (Cloud, colours or None, planes [K, 4] made inward facing)"""
    cl, col = C.c_void_p(), C.c_void_p()
    planes = np.empty((MAX_PLANES, 4), np.float32)
    k = C.c_int32()
    ctx._chk(ctx.lib.hs_load_room(ctx.h, os.fsencode(directory), C.byref(cl), C.byref(col), ptr(planes), MAX_PLANES, C.byref(k)))
    return Cloud(ctx, cl), (Cloud(ctx, col) if col.value else None), planes[: k.value].copy()


# ---- transform export compatibility (Main.hs:2271-2325) -------------------------------------------------------
def cloudFromPly(ctx: Context, path: str):
    """-> (Cloud, colours Cloud or None) from a PLY vertex cloud (ascii / binary_little_endian, float x y z [+ uchar red green blue])"""
    cl, col = C.c_void_p(), C.c_void_p()
    ctx._chk(ctx.lib.hs_cloud_from_ply(ctx.h, os.fsencode(path), C.byref(cl), C.byref(col)))
    return Cloud(ctx, cl), (Cloud(ctx, col) if col.value else None)


def transformFromText(text: bytes | str) -> np.ndarray:
    """Inverse of roomProjectionToXfFormat / roomProjectionToString (Main.hs:2271-2302): the file holds the left-multiplicative
    matrix; the result is roomProj (row vectors, right-multiplied, translation in row 3)."""
    if isinstance(text, str):
        text = text.encode()
    m = np.empty(16, np.float32)
    rc = L.load().hs_transform_from_text(text, len(text), ptr(m))
    if rc != L.HS_OK:
        raise HsError(rc, "not a 4x4 transform (16 numbers expected)")
    return m.reshape(4, 4)


def transformCloudFile(ctx: Context, src: str, matrix, dst: str):
    """What `plyxform` / `pcl_transform_point_cloud -matrix` do with the reference's exports (Main.hs:2305-2325), at full resolution
    on the GPU: read src (.ply or .pcd), apply roomProj, write dst (.ply or .pcd).  `matrix` is a 4x4 roomProj array, the text of an
    .xf file / -matrix string, or the path of an .xf file.  Returns the number of points."""
    if isinstance(matrix, (str, bytes)) and os.path.exists(matrix):
        with open(matrix, "rb") as fh:
            matrix = fh.read()
    m = transformFromText(matrix) if isinstance(matrix, (str, bytes)) else np.asarray(matrix, np.float32).reshape(4, 4)
    cl, col = (cloudFromPly if src.lower().endswith(".ply") else cloudFromFile)(ctx, src)
    out = ctx.transform(cl, m)
    rgb = None
    if col is not None:  # colours are u8 / 255 in Float: * 255 and rounding gives the bytes back exactly
        rgb = np.rint(col.download() * np.float32(255.0)).astype(np.uint8)
    (ctx.write_ply if dst.lower().endswith(".ply") else ctx.write_pcd)(out, dst, rgb)
    return len(out)


def plyInfo(path: str):
    """(vertices, has colours, "ascii" | "binary_little_endian") of a PLY file's header; raises on anything the reader would refuse"""
    n, rgb, asc = C.c_int64(), C.c_int32(), C.c_int32()
    rc = L.load().hs_ply_info(os.fsencode(path), C.byref(n), C.byref(rgb), C.byref(asc))
    if rc != L.HS_OK:
        raise HsError(rc, f"{path}: not a readable PLY vertex cloud")
    return n.value, bool(rgb.value), ("ascii" if asc.value else "binary_little_endian")
