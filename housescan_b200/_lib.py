"""ctypes binding of libhousescan_b200.so — the same C ABI a Haskell `foreign import ccall` binds
(include/housescan_b200.h).  There is no CPU fallback: if the library is missing this raises, and
creating a context without an sm_100 GPU raises HsError(HS_ECUDA)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libhousescan_b200.so")

HS_OK, HS_EINVAL, HS_ECUDA, HS_ENCCL, HS_ENOMEM, HS_ESINGULAR, HS_EIO = range(7)
HS_REC, HS_PS, HS_NE = 24, 10, 29
_STATUS = {1: "HS_EINVAL", 2: "HS_ECUDA", 3: "HS_ENCCL", 4: "HS_ENOMEM", 5: "HS_ESINGULAR", 6: "HS_EIO"}

vp, i32, i64, u32, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol declared in include/housescan_b200.h
SIGNATURES = {
    "hs_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "hs_ctx_destroy": (i32, [vp]),
    "hs_last_error": (C.c_char_p, [vp]),
    "hs_ctx_set_stream": (i32, [vp, vp]),
    "hs_ctx_sync": (i32, [vp]),
    "hs_ctx_device": (i32, [vp]),
    "hs_ctx_sm_count": (i32, [vp]),
    "hs_ctx_launch_count": (i64, [vp]),
    "hs_ctx_set_mode": (i32, [vp, i32, i32]),
    "hs_cloud_upload": (i32, [vp, vp, i64, C.POINTER(vp)]),
    "hs_cloud_alloc": (i32, [vp, i64, C.POINTER(vp)]),
    "hs_cloud_wrap_device": (i32, [vp, vp, i64, C.POINTER(vp)]),
    "hs_cloud_write": (i32, [vp, vp, vp, i64]),
    "hs_cloud_download": (i32, [vp, vp, vp]),
    "hs_cloud_size": (i64, [vp]),
    "hs_cloud_device_ptr": (vp, [vp]),
    "hs_cloud_free": (i32, [vp, vp]),
    "hs_backproject_ref": (i32, [vp, vp, i32, i32, vp, vp, C.POINTER(i64)]),
    "hs_backproject_ref_dev": (i32, [vp, vp, i32, i32, vp, vp, C.POINTER(i64)]),
    "hs_backproject_reduce6x6": (i32, [vp, vp, i64, i32, i32, vp, vp, vp, i32, vp]),
    "hs_backproject_reduce6x6_dev": (i32, [vp, vp, i64, i32, i32, vp, vp, vp, i32, vp]),
    "hs_plane_assign": (i32, [vp, vp, vp, i32, vp, vp]),
    "hs_plane_assign_dev": (i32, [vp, vp, vp, i32, vp, vp]),
    "hs_planes_from_cuboid": (i32, [vp, vp]),
    "hs_cuboid_residual_grad": (i32, [vp, vp, vp, C.POINTER(f64), vp, vp]),
    "hs_rooms_cuboid_sums": (i32, [vp, vp, vp, i32, vp, vp]),
    "hs_rooms_cuboid_sums_async": (i32, [vp, vp, vp, i32, vp, vp]),
    "hs_peer_mailbox_create": (i32, [vp, i32, i32, vp]),
    "hs_peer_mailbox_connect": (i32, [vp, vp]),
    "hs_rooms_cuboid_sums_allreduce_async": (i32, [vp, vp, vp, i32, vp, vp]),
    "hs_peer_group_create_local": (i32, [vp, i32]),
    "hs_eval_session_begin": (i32, [vp, vp, vp, i32, i32, C.POINTER(vp)]),
    "hs_eval_session_post": (i32, [vp, vp, i32]),
    "hs_eval_session_wait": (i32, [vp, i64, vp]),
    "hs_eval_session_eval": (i32, [vp, vp, vp]),
    "hs_eval_session_done": (i64, [vp]),
    "hs_eval_session_device_results": (vp, [vp]),
    "hs_eval_session_times": (i32, [vp, i64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hs_eval_session_stop": (i32, [vp]),
    "hs_eval_session_end": (i32, [vp]),
    "hs_eval_plan": (i32, [i64, vp, i32, i32, i32, C.POINTER(i32), vp, vp, vp, vp, vp]),
    "hs_nm_minimize": (i32, [vp, vp, vp, vp, i32, f64, i32, vp, C.POINTER(f64), C.POINTER(i32), C.POINTER(i32)]),
    "hs_fit_cuboid_cloud_nm": (i32, [vp, vp, vp, vp, f64, i32, vp, C.POINTER(f64), C.POINTER(i32), C.POINTER(i32)]),
    "hs_bfgs_minimize": (i32, [vp, vp, vp, i32, i32, f64, vp, C.POINTER(f64), C.POINTER(i32), C.POINTER(i32)]),
    "hs_cuboid_grad_from_sums": (i32, [vp, vp, C.POINTER(f64), vp, vp]),
    "hs_plane_sums": (i32, [vp, vp, vp, i32, vp, i32, vp]),
    "hs_scatter3x3": (i32, [vp, vp, vp, vp]),
    "hs_fit_plane": (i32, [vp, vp, vp]),
    "hs_transform": (i32, [vp, vp, vp, vp]),
    "hs_rotate_around": (i32, [vp, vp, vp, vp, vp]),
    "hs_translate": (i32, [vp, vp, vp, vp]),
    "hs_mean_extent": (i32, [vp, vp, vp, C.POINTER(f32)]),
    "hs_write_ply": (i32, [vp, vp, vp, C.c_char_p]),
    "hs_write_ply_begin": (i32, [C.c_char_p, i64, i32]),
    "hs_write_ply_part": (i32, [vp, vp, vp, C.c_char_p, i64, i64]),
    "hs_write_ply_part_host": (i32, [C.c_char_p, vp, vp, i64, i64, i64]),
    "hs_proj_to_string": (i32, [vp, C.c_char_p, i32]),
    "hs_proj_to_xf": (i32, [vp, C.c_char_p, i32]),
    "hs_cc_label": (i32, [vp, vp, vp, i64, u32, vp]),
    "hs_cc_label_dev": (i32, [vp, vp, vp, i64, u32, vp]),
    "hs_kth_largest": (i32, [vp, vp, i32, i64, C.POINTER(f32)]),
    "hs_kth_smallest": (i32, [vp, vp, i32, i64, C.POINTER(f32)]),
    "hs_filter_le": (i32, [vp, vp, i32, f32, vp, vp, vp, C.POINTER(i64)]),
    "hs_remove_ceiling": (i32, [vp, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(f32)]),
    "hs_cuboid_from_params": (i32, [vp, vp]),
    "hs_errfun": (f64, [vp, vp]),
    "hs_errfun_closest": (f64, [vp, i32, vp]),
    "hs_guess_dims": (i32, [vp, vp]),
    "hs_fit_cuboid": (i32, [vp, i32, vp, C.POINTER(i32), C.POINTER(f64), vp, i32]),
    "hs_fit_cuboid_cloud_bfgs": (i32, [vp, vp, vp, i32, f64, vp, C.POINTER(f64), C.POINTER(i32), C.POINTER(i32)]),
    "hs_lstsq_distances": (i32, [vp, vp, vp, i32, i32, vp, C.POINTER(f64)]),
    "hs_group_cc": (i32, [vp, vp, vp, i64, u32, vp, vp, C.POINTER(i32)]),
    "hs_plane_eqs_from_text": (i32, [C.c_char_p, i64, vp, i32, C.POINTER(i32)]),
    "hs_plane_eqs_from_file": (i32, [C.c_char_p, vp, i32, C.POINTER(i32)]),
    "hs_cloud_from_pcd": (i32, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(vp)]),
    "hs_pcd_info": (i32, [C.c_char_p, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
    "hs_make_inward_facing": (i32, [vp, vp, vp, i32]),
    "hs_load_room": (i32, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(vp), vp, i32, C.POINTER(i32)]),
    "hs_transform_from_text": (i32, [C.c_char_p, i64, vp]),
    "hs_cloud_from_ply": (i32, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(vp)]),
    "hs_write_pcd": (i32, [vp, vp, vp, C.c_char_p]),
    "hs_ply_info": (i32, [C.c_char_p, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
    "hs_kth_shard_pass": (i32, [vp, vp, i32, i32, u32, u32, vp]),
    "hs_kth_key_of_float": (u32, [f32]),
    "hs_kth_float_of_key": (f32, [u32]),
    "hs_rotation_between_plane_eqs": (i32, [vp, vp, vp]),
    "hs_rotate_plane_eq_around": (i32, [vp, vp, vp, vp]),
    "hs_translate_plane_eq": (i32, [vp, vp, vp]),
    "hs_plane_corner": (i32, [vp, vp, vp, vp]),
    "hs_proj_compose": (i32, [vp, vp, vp]),
    "hs_proj_translate": (i32, [vp, vp, vp]),
    "hs_proj_rotate_around": (i32, [vp, vp, vp, vp]),
    "hs_version": (C.c_char_p, []),
}


class HsError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"{_STATUS.get(status, status)}: {msg}")
        self.status = status
        self.message = msg


_lib = None


def load():
    """Load the shared library (raises if it has not been built — no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -m housescan_b200.build` (nvcc, sm_100a). "
                "housescan_b200 has no CPU fallback."
            )
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(vp)
    return vp(a)


def as_f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a
