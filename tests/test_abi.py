"""The drop-in boundary: the C-ABI library loads, exports every symbol include/housescan_b200.h declares, the ctypes
binding covers all of them, and there is no CPU fallback (creating a context without a GPU fails loudly)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "housescan_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"HS_API\s+[\w\s\*]+?\b(hs_\w+)\s*\(", src)))


def test_header_declares_the_surveyed_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 45
    for must in ("hs_ctx_create", "hs_cloud_upload", "hs_backproject_ref", "hs_backproject_reduce6x6", "hs_plane_assign",
                 "hs_cuboid_residual_grad", "hs_rooms_cuboid_sums", "hs_plane_sums", "hs_scatter3x3", "hs_transform",
                 "hs_mean_extent", "hs_write_ply", "hs_cc_label", "hs_kth_largest", "hs_filter_le", "hs_last_error"):
        assert must in syms
    # every entry cites the reference interface it replaces
    assert len(re.findall(r"\w+\.hs:\d+", open(HEADER).read())) >= 30


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (hs_\w+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    # nothing but the C ABI leaks out (no C++ symbols, no oracle symbols)
    leaked = [l for l in out.splitlines() if " T " in l and " T hs_" not in l]
    assert not leaked, leaked[:5]
    assert "orc_" not in out


def test_ctypes_binding_covers_the_header(built_lib):
    from housescan_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.hs_version().startswith(b"housescan_b200")


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "--list-elf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_gpu(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is exercised on the CPU-only builder")
    import housescan_b200 as hb

    with pytest.raises(hb.HsError) as e:
        hb.Context(0)
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)  # HS_ECUDA


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "housescan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_haskell_ffi_binds_only_declared_symbols():
    """haskell/HouseScanB200/FFI.hs is the reference-side binding a maintainer adds (no GHC here to compile it): every
    `foreign import ccall` in it must name an entry point that the header declares and the library exports."""
    src = open(os.path.join(ROOT, "haskell", "HouseScanB200", "FFI.hs")).read()
    bound = re.findall(r'foreign import ccall (?:safe|unsafe) "(\w+)"', src)
    assert len(bound) >= 30
    declared = set(declared_symbols())
    assert not [b for b in bound if b not in declared]


SHIMS = ["FitCuboidBFGS", "TranslationOptimizer", "GroupConnectedComponents", "VectorUtil", "HoniHelper"]


def _export_list(src: str):
    """names between `module X (` and `) where`, comments dropped, in order"""
    m = re.search(r"^module\s+\w+\s*\((.*?)\)\s*where", src, re.S | re.M)
    assert m, "no export list"
    body = re.sub(r"--.*", "", m.group(1))
    return [t.strip() for t in body.split(",") if t.strip()]


@pytest.mark.parametrize("mod", SHIMS)
def test_haskell_shim_keeps_the_reference_export_list(mod):
    """haskell/<Module>.hs drops in for housescan/<Module>.hs: its export list starts with the reference's, name for name and in
    the reference's order; what follows is additive.  Every c_* it calls is bound in HouseScanB200/FFI.hs."""
    shim = open(os.path.join(ROOT, "haskell", mod + ".hs")).read()
    ours = _export_list(shim)
    ref_path = os.path.join("/root/reference/housescan", mod + ".hs")
    if os.path.exists(ref_path):  # the reference is mounted in the build container only
        theirs = _export_list(open(ref_path).read())
        assert ours[: len(theirs)] == theirs, (mod, ours, theirs)
    ffi = open(os.path.join(ROOT, "haskell", "HouseScanB200", "FFI.hs")).read()
    bound = set(re.findall(r"\b(c_\w+)\s*::", ffi))
    dev = open(os.path.join(ROOT, "haskell", "HouseScanB200", "Device.hs")).read()
    used = set(re.findall(r"\b(c_\w+)\b", shim)) | set(re.findall(r"\b(c_\w+)\b", dev))
    assert not (used - bound), (mod, sorted(used - bound))
