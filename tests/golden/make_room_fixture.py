"""Regenerates the committed room fixture tests/golden/room/ (a KinFu-style room directory as loadRoom reads it, Main.hs:1740-1765):

    cloud_downsampled.pcd      1 203 points, DATA binary_compressed, FIELDS x y z rgb normal_x normal_y normal_z curvature
    planes.txt                 six PCL planes `a b c d` (ax + by + cz + d = 0) with arbitrary signs and scales
    cloud_plane_hull<i>.pcd    12 hull points per plane, DATA ascii
    expected.json              what the ORACLE's restatement reads out of it (uint32 images of the floats: exact)

The reference ships no such files and pcd-loader / attoparsec are not mounted, so this freezes the oracle's reading of the
published formats; the C-ABI parsers and the GPU unpack path are tested against it.

usage: python tests/golden/make_room_fixture.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O  # noqa: E402
from housescan_b200 import synth  # noqa: E402
from pcd_util import write_pcd  # noqa: E402

ROOM = os.path.join(HERE, "room")


def main():
    os.makedirs(ROOM, exist_ok=True)
    rng = np.random.default_rng(2024)
    n = 1203
    xyz, face = synth.cuboid_room_cloud(n, synth.C1_PARAMS, sigma=0.004, seed=77)
    rgb = rng.integers(0, 256, (n, 3))
    nrm = rng.normal(size=(n, 4))
    write_pcd(os.path.join(ROOM, "cloud_downsampled.pcd"), xyz, rgb, nrm, kind="binary_compressed")
    planes = O.planes_from_cuboid(synth.C1_PARAMS)
    lines = []
    for k in range(6):
        a, b, c, d = (float(v) for v in planes[k])
        s = -1.0 if k in (1, 2, 5) else 1.0
        scale = (1.0, 2.5, 0.5, 3.0, 1.25, 0.75)[k]
        lines.append(f"{s * a * scale:.9g} {s * b * scale:.9g} {s * c * scale:.9g} {-s * d * scale:.9g}")
        write_pcd(os.path.join(ROOM, f"cloud_plane_hull{k}.pcd"), xyz[face == k][:12], kind="ascii")
    with open(os.path.join(ROOM, "planes.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    cloud, cols, pl = O.load_room(ROOM)
    raw_planes = O.plane_eqs_from_text(open(os.path.join(ROOM, "planes.txt"), "rb").read())
    exp = {
        "n": int(len(cloud)),
        "cloud_u32_sum": int(cloud.view(np.uint32).astype(np.uint64).sum()),
        "cloud_first": cloud[:3].view(np.uint32).tolist(), "cloud_last": cloud[-2:].view(np.uint32).tolist(),
        "colors_first": cols[:3].view(np.uint32).tolist(), "colors_u32_sum": int(cols.view(np.uint32).astype(np.uint64).sum()),
        "planes_raw": raw_planes.view(np.uint32).tolist(),
        "planes_inward": pl.view(np.uint32).tolist(),
    }
    with open(os.path.join(ROOM, "expected.json"), "w") as fh:
        json.dump(exp, fh, indent=1)
    print("wrote", ROOM, {k: (v if isinstance(v, int) else "...") for k, v in exp.items()})


if __name__ == "__main__":
    main()
