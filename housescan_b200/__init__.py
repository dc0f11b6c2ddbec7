"""housescan_b200 — B200-native (sm_100a CUDA) implementation of the data-parallel point-cloud path of
nh2/housescan behind the reference's own module API.

Layout: ``csrc/`` CUDA kernels + the C ABI (include/housescan_b200.h), ``host/`` the C++ mirror of the
reference's host-side helper modules, and thin Python mirrors of the Haskell modules
(``HoniHelper``, ``FitCuboidBFGS``, ``TranslationOptimizer``, ``GroupConnectedComponents``,
``VectorUtil``, ``Bijection``) that call the C ABI through ctypes exactly as a Haskell
``foreign import ccall`` would.  There is no CPU fallback.
"""
from __future__ import annotations

import os

from ._lib import HS_NE, HS_PS, HS_REC, HsError, SO_PATH, load  # noqa: F401
from .core import Cloud, Context, EvalSession, peer_group_local, write_ply_begin, write_ply_part_host, cuboid_grad_from_sums, planes_from_cuboid, proj_to_string, proj_to_xf  # noqa: F401

_default_ctx = None


def default_context() -> Context:
    """Process-wide context on cuda:LOCAL_RANK (one process per GPU under torchrun)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
