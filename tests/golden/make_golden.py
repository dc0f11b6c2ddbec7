"""Regenerates the committed golden fixtures.

  room_corners.json  — golden INPUTS copied as data from the reference: the six real rooms of devSetup
                       (housescan/Main.hs:2346-2413) and the test room of loadTestRoom1WithCorners (Main.hs:2531-2540).
                       The reference holds no golden OUTPUTS for them.
  oracle_vectors.npz — small seeded inputs + the oracle's outputs.  The reference (Haskell, no GHC in this image,
                       HmatrixUtils missing) cannot be run, so these freeze the ORACLE, not the reference: they catch
                       accidental drift of oracle/ and give the GPU tests a fixture that travels to the GPU box.

usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402
from housescan_b200 import synth  # noqa: E402

ROOMS = {
    "elabathroom1": [[-0.80041015, -0.9884287, -1.5198468], [-0.80076337, 1.5652194, -1.6966621], [-0.96627235, 1.6066637, 1.6987381], [-0.95966613, -0.98843944, 1.7501512],
                     [0.5775635, -0.98843974, 1.8941374], [0.69802976, 1.5778129, -1.5970236], [0.6918931, -0.9884289, -1.4197717], [0.5793414, 1.6201758, 1.843245]],
    "elakitchen1": [[9.671569e-2, -0.91251373, -2.2253428], [-2.1025066, -0.91251373, -1.0994778], [-2.2102604, 1.7036777, -1.1300573], [9.584594e-2, 1.7212467, -2.3112164],
                    [2.0891232, 1.725184, 1.2726974], [-0.30790687, 1.707284, 2.3521976], [-0.23941708, -0.9125142, 2.3106794], [2.046792, -0.912514, 1.2810183]],
    "elamiddle1": [[0.4379654, 1.5980418, -0.26340306], [-0.4265194, 1.6302242, -0.19807673], [-0.38968277, 1.6646035, 0.5294547], [0.4116578, 1.6347461, 0.4683745],
                   [-0.41408634, -0.97689533, 0.7326498], [-0.4582219, -0.9768954, -0.14981699], [0.41233778, -0.9768954, -0.21617222], [0.38007212, -0.97689533, 0.66986275]],
    "elaroom1": [[2.1304765, -1.0610049, -1.6002798], [2.380793, 1.5803287, -1.5484123], [-1.6485546, 2.0091648, -1.9336071], [-1.9732126, -0.5630518, -1.9919395],
                 [-2.0735986, -0.7954781, 2.0725145], [-1.732965, 1.9079933, 2.1676683], [1.7069712, -1.187545, 1.8982229], [1.9569516, 1.5288324, 2.014926]],
    "elarooma2": [[-1.2394748, -0.9991329, -1.9200139], [-1.3578, 1.6094978, -1.8529172], [-1.1683455, 1.2875693, 2.6794243], [-1.0703187, -0.9991331, 2.472702],
                  [0.8354454, 1.5010788, 2.7371817], [0.9719229, -0.9991331, 2.5117178], [1.0528631, -0.9991329, -1.8371677], [0.91312027, 1.618686, -1.7705941]],
    "elaroomb3": [[2.2902393, -1.1796348, -2.025272], [2.3693638, -1.1879312, 1.463068], [-2.0558214, -0.76278543, 2.231657], [-2.467492, -0.7224221, -1.7942874],
                  [-2.4088001, 1.9028702, -1.733478], [-2.0076504, 1.806385, 2.200717], [2.601904, 1.4829392, 1.3990588], [2.5302696, 1.5456867, -1.9704242]],
    "testroom1": [[0.5213087, 1.3714368, 0.9477334], [0.6015281, 0.7033132, 4.419407], [4.8369703, 1.2523801, 4.0971937], [4.4101005, 1.8874655, 0.5908974],
                  [0.3593011, 4.1540117, 0.914716], [4.14219, 4.488981, 1.1421864], [4.5736876, 3.750552, 4.565998], [0.46467793, 3.254958, 4.8851647]],
}


def main():
    with open(os.path.join(HERE, "room_corners.json"), "w") as fh:
        json.dump(ROOMS, fh, indent=1)
    w, h = 80, 60
    depth = synth.render_depth_frame(w=w, h=h, intr=synth.KINFU_INTR / 8, seed=101)
    bp_xyz, mask = O.backproject_ref(depth, w, h)
    params = synth.C1_PARAMS + 0.01 * np.random.default_rng(7).normal(size=10)
    cloud, _ = synth.cuboid_room_cloud(4099, synth.C1_PARAMS, sigma=0.01, seed=102)
    planes = O.planes_from_cuboid(params)
    assign, resid = O.plane_assign(cloud, planes)
    f, grad, counts, _ = O.cuboid_residual_grad(cloud, params)
    src, dst, n, _ = synth.voxel_building_graph(12, 9, 12, seed=103)
    np.savez_compressed(
        os.path.join(HERE, "oracle_vectors.npz"),
        depth=depth, w=w, h=h, mask=mask, bp_xyz=bp_xyz, params=params, cloud=cloud, planes=planes, assign=assign, resid=resid,
        f=f, grad=grad, counts=counts, cc_src=src, cc_dst=dst, cc_n=n, cc_label=O.cc_label(src, dst, n),
        kth=O.kth_largest(cloud[:, 1], len(cloud) // 5), ne29=O.backproject_reduce6x6(depth, w, h, planes),
    )
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
