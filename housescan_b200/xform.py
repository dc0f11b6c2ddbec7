"""Full-resolution transform of a room's .ply / .pcd on the GPU — the step the reference leaves to external tools
(`plyxform` with the .xf files of exportAllRoomXfFiles, Main.hs:2315-2325; `pcl_transform_point_cloud -matrix` with the strings of
exportAllRoomPCLTransforms, Main.hs:2305-2313; README.md:16 step 4).

    python -m housescan_b200.xform IN.{ply,pcd} ROOM.xf OUT.{ply,pcd}
    python -m housescan_b200.xform IN.pcd -matrix a,b,c,...,p OUT.pcd
"""
from __future__ import annotations

import sys

from . import RoomIO, default_context


def main(argv=None) -> int:
    a = list(sys.argv[1:] if argv is None else argv)
    if len(a) == 4 and a[1] == "-matrix":
        src, matrix, dst = a[0], a[2], a[3]
    elif len(a) == 3:
        src, matrix, dst = a
    else:
        sys.stderr.write(__doc__)
        return 2
    n = RoomIO.transformCloudFile(default_context(), src, matrix, dst)
    print(f"{src} -> {dst}: {n} points")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
