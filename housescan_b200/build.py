"""Build libhousescan_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m housescan_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libhousescan_b200.so")
CU = ["hs_api.cu", "k_planes.cu", "k_eval.cu", "k_transform.cu", "k_depth.cu", "k_graph.cu", "k_select.cu", "k_pcd.cu"]
HOST = ["hs_host.cpp", "hs_roomio.cpp"]
HEADERS = [
    os.path.join(ROOT, "include", "housescan_b200.h"),
    os.path.join(HERE, "csrc", "hs_internal.cuh"),
    os.path.join(HERE, "csrc", "k_common.cuh"),
    os.path.join(HERE, "csrc", "k_ring.cuh"),
    os.path.join(HERE, "csrc", "k_eval.cuh"),
    os.path.join(HERE, "csrc", "k_eval_point.cuh"),
    os.path.join(HERE, "csrc", "k_peer.cuh"),
    os.path.join(HERE, "host", "hs_host.hpp"),
    os.path.join(HERE, "host", "vec.hpp"),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # never contract mul+add into fma behind our back: the reference's Float arithmetic has no FMA, and ptxas does fuse
    # mul.rn.f32x2 + add.rn.f32x2 into FFMA2 without this (seen in SASS); every FMA we want is written as fma().
    "--fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
]


def sources():
    return [os.path.join(HERE, "csrc", f) for f in CU] + [os.path.join(HERE, "host", f) for f in HOST]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in sources() + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libhousescan_b200.so cannot be built (there is no CPU fallback)")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    log = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-ccbin", "/usr/bin/g++", "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(out.stdout)
    if out.returncode != 0:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("link failed")
    with open(os.path.join(objdir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
