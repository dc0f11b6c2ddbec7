"""Pin the CPU oracle against every check the reference's own sources hold for this path.
The reference is Haskell and cannot be built here, so these are its self-consistency properties, asserts,
doc-contracts and known answers restated as tests — plus the committed golden fixtures (tests/golden/)."""
import json
import math
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- FitCuboidBFGS.hs:134-140  cuboidFromParamsIdentityCheck (QuickCheck property)
@settings(max_examples=300, deadline=None)
@given(
    st.tuples(*[st.floats(-100, 100, allow_nan=False) for _ in range(3)]),
    st.tuples(*[st.floats(-100, 100, allow_nan=False) for _ in range(3)]),
    st.tuples(*[st.floats(-100, 100, allow_nan=False) for _ in range(4)]),
)
def test_cuboid_from_params_identity_check(c, d, q):
    if abs(sum(q)) < 1e-3 or math.sqrt(sum(v * v for v in q)) < 1e-3:
        return  # "Unit quaternion can't be all 0" (FitCuboidBFGS.hs:137)
    p = list(c) + list(d) + list(q)
    a, b = O.cuboid_from_params(p), O.cuboid_from_params_rotate_around(p)
    # the reference bound is 1e-6 on QuickCheck-sized values; scale it with the magnitude of the inputs
    assert np.sum(np.linalg.norm(a - b, axis=1)) < 1e-6 * max(1.0, np.abs(p[:6]).max())


def _example_points():
    """FitCuboidBFGS.hs:29-41: 2x1x1 box rotated 20 degrees about (1,2,3)"""
    pts = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [2, 0, 0], [2, 0, 1], [2, 1, 0], [2, 1, 1]], float)
    return pts @ O.rot_matrix3([1, 2, 3], 20 / 180 * math.pi)


def test_example_box_known_answer():
    """derived KAT of FitCuboidBFGS.main (:260-269): dims {2,1,1}, err ~ 0, centre = mean of the points"""
    pts = _example_points()
    for fit in (O.fit_cuboid, O.fit_cuboid_from_center_first):
        sol, steps, err, path = fit(pts)
        assert err < 1e-12
        assert np.allclose(sorted(sol[3:6]), [1, 1, 2], atol=1e-5)
        assert np.allclose(sol[:3], pts.mean(axis=0), atol=1e-5)
        assert 0 < steps <= 4001 and path.shape[1] == 13
    assert np.allclose(sorted(O.guess_dims(pts)), [1, 1, 2], atol=1e-12)


def test_cuboid_gen_diagnostic():
    """cuboidGen (FitCuboidBFGS.hs:143-168).  The reference's driver only PRINTS (err, steps) and dumps the input when
    err > 1 (:277-282) — it asserts nothing, and the closest-corner objective started from a cube does get stuck in local
    minima (two points claiming one corner).  What is pinned: exact boxes are recovered in a good share of the draws, every
    run terminates within the iteration budget, and err is never worse than the starting simplex."""
    rng = np.random.default_rng(0)
    errs = []
    for _ in range(12):
        a, b, c = rng.uniform(1, 10, 3)
        ps = np.array([[x, y, z] for x in (0, a) for y in (0, b) for z in (0, c)], float)
        R = O.rot_matrix3(rng.uniform(0.05, 3, 3), math.radians(rng.uniform(0, 360)))
        pts = ps @ R
        sol, steps, err, path = O.fit_cuboid_from_center(pts)
        g = O.guess_dims(pts)[0]
        start = O.errfun_closest(pts, np.concatenate([O.point_mean_d(pts), [g, g, g, 0.1, 0.1, 0.1, 0.1]]))
        assert steps <= 2000 and np.isfinite(err) and err <= start + 1e-12
        errs.append(err)
    assert np.mean(np.array(errs) < 1e-9) >= 0.4


def test_planes_hold_four_corners_each():
    """Main.hs:1881 assert: every cuboid plane contains exactly 4 of the 8 corners within 1e-4"""
    rng = np.random.default_rng(1)
    for _ in range(50):
        p = np.concatenate([rng.normal(size=3) * 3, rng.uniform(1, 10, 3), rng.normal(size=4)])
        planes = O.planes_from_cuboid(p)
        corners = O.cuboid_from_params(p).astype(np.float32)
        for k in range(6):
            _, r = O.plane_assign(corners, planes[k : k + 1])
            assert int(np.sum(np.abs(r) < 1e-4)) == 4
        for j in range(3):  # the antiparallel-pair structure the GPU kernel relies on is exact
            assert np.array_equal(planes[2 * j, :3], -planes[2 * j + 1, :3])


def test_rotation_between_normals_doc_comment():
    """Main.hs:1548-1552: the matrix rotates plane1's normal into plane2's direction (right multiplication)"""
    rng = np.random.default_rng(2)
    for _ in range(20):
        n1, n2 = rng.normal(size=3), rng.normal(size=3)
        n1, n2 = (n1 / np.linalg.norm(n1)).astype(np.float32), (n2 / np.linalg.norm(n2)).astype(np.float32)
        R = O.rotation_between_normals(n1, n2)
        assert np.allclose(n1 @ R, n2, atol=2e-6)


def test_room_proj_is_linear_r_assert():
    """Main.hs:2637 (projTest6): projectRoom (linear R) on a fresh room leaves roomProj == linear R exactly"""
    R = O.rot_matrix3([1, 0, 0], np.float32(math.radians(10)), np.float32)
    proj = O.proj_linear(R)
    assert np.array_equal(O.proj_compose(O.proj_identity(), proj), proj)


ROOM1_CORNERS = np.array(  # Main.hs:2531-2540 (golden input)
    [[0.5213087, 1.3714368, 0.9477334], [0.6015281, 0.7033132, 4.419407], [4.8369703, 1.2523801, 4.0971937], [4.4101005, 1.8874655, 0.5908974],
     [0.3593011, 4.1540117, 0.914716], [4.14219, 4.488981, 1.1421864], [4.5736876, 3.750552, 4.565998], [0.46467793, 3.254958, 4.8851647]], np.float32)


@pytest.mark.parametrize("case", ["projTest", "projTest2", "projTest4", "projTest5"])
def test_proj_replay_equivalence(case):
    """projTest..projTest5 (Main.hs:2543-2616): replaying the accumulated roomProj with projectRoom on a freshly loaded
    room coincides with the incrementally transformed room (Float tolerance).  Pins right-multiplication with the
    translation in row 3."""
    rng = np.random.default_rng(5)
    cloud = np.concatenate([ROOM1_CORNERS, rng.uniform(0, 5, size=(500, 3)).astype(np.float32)])
    proj, cur = O.proj_identity(), cloud.copy()

    def translate(v):
        nonlocal proj, cur
        cur, proj = O.translate_cloud(cur, v), O.proj_translate4(v, proj)

    def rotate_around(c, R):
        nonlocal proj, cur
        cur, proj = O.rotate_cloud_around(cur, c, R), O.proj_rotate_around(c, R, proj)

    rx = lambda deg: O.rot_matrix3([1, 0, 0], np.float32(math.radians(deg)), np.float32)
    if case == "projTest":
        translate([6, 0, 0]); rotate_around(O.point_mean_f32seq(cur), rx(90))  # rotateRoom = about roomMean (Main.hs:1677-1678)
    elif case == "projTest2":
        rotate_around(O.point_mean_f32seq(cur), rx(10))
    elif case == "projTest4":
        translate([0, 0, 6]); rotate_around(np.zeros(3, np.float32), rx(10))
    else:
        translate([1, 2, 6])
    assert np.allclose(O.project_cloud(cloud, proj), cur, atol=2e-5)
    assert proj[3, 3] == 1 and not proj[:3, 3].any()  # projectRoom's pattern match (Main.hs:1725-1728)


def test_project_cloud_rejects_bad_last_column():
    M = O.proj_identity()
    M[1, 3] = 1e-3
    with pytest.raises(ValueError):
        O.project_cloud(np.zeros((1, 3), np.float32), M)


# ---- golden inputs: the 6 real rooms of devSetup (Main.hs:2346-2413) + loadTestRoom1WithCorners
def test_real_room_corner_sets_are_regression_inputs():
    """The reference holds no golden OUTPUTS for these hand-picked noisy corner sets (its fit prints, asserts nothing), so
    only sanity is pinned here: the two-stage fit terminates, never ends above the err of its starting point, and yields
    positive room-sized dimensions.  tests/test_host_logic.py checks the product's C++ optimiser reaches the same end state."""
    rooms = json.load(open(os.path.join(GOLDEN, "room_corners.json")))
    assert len(rooms) == 7
    for name, corners in rooms.items():
        pts = np.array(corners, float)
        sol, steps, err, _ = O.fit_cuboid_from_center_first(pts)
        assert steps <= 4000 and np.isfinite(err), name
        assert np.all(np.array(sol[3:6]) > 0.3) and np.all(np.array(sol[3:6]) < 8), name
        assert err < np.sum((pts - pts.mean(axis=0)) ** 2), name  # better than collapsing the cuboid to its centre


# ---- Bijection.hs:10-15, TranslationOptimizer.hs:22-35, GroupConnectedComponents.hs:54 contracts
def test_biject_first_occurrence_order():
    idx, unb = O.biject(["c", "a", "c", "b", "a"])
    assert idx == {"c": 0, "a": 1, "b": 2} and unb == ["c", "a", "b"]


def test_lst_sq_distances_contract():
    res = O.lst_sq_distances({("b", "c"): 2.0, ("a", "b"): -1.5})  # negative d allowed; first node of first (sorted) edge at 0
    pos, rmse = res
    assert pos["a"] == 0.0 and abs(pos["b"] + 1.5) < 1e-12 and abs(pos["c"] - 0.5) < 1e-12 and rmse < 1e-7
    pos, rmse = O.lst_sq_distances({(1, 2): 1.0, (2, 3): 1.0, (1, 3): 2.6})  # inconsistent triangle
    assert abs(pos[2] - 1.2) < 1e-12 and abs(pos[3] - 2.4) < 1e-12
    resid = np.array([1.2 - 1.0, 2.4 - 2.6, 1.2 - 1.0])
    assert abs(rmse - math.sqrt(np.linalg.norm(resid) / 3)) < 1e-12  # the quirk: sqrt(||r||_2 / m)
    assert O.lst_sq_distances({(1, 2): 1.0, (3, 4): 1.0}) is None  # disconnected => singular => Nothing


def test_group_connected_components_order():
    edges = [((5, 6), "a"), ((1, 2), "b"), ((6, 7), "c"), ((2, 1), "d"), ((9, 9), "e"), ((5, 6), "f")]
    comps = O.group_connected_components(edges)
    # biject order 5,6,1,2,7,9 -> components by ascending bijected minimum: {5,6,7}, {1,2}, {9}
    # inside a component: reverse input order; duplicate edge (5,6) carries the LAST payload twice (Map.fromList)
    assert comps == [[((5, 6), "f"), ((6, 7), "c"), ((5, 6), "f")], [((2, 1), "d"), ((1, 2), "b")], [((9, 9), "e")]]
    assert O.group_connected_components([]) == []


def test_remove_ceiling_semantics():
    rng = np.random.default_rng(3)
    xyz = rng.normal(size=(1003, 3)).astype(np.float32)
    kept, _ = O.remove_ceiling(xyz)
    k = 1003 // 5
    ylim = np.sort(xyz[:, 1])[::-1][k - 1]
    assert np.array_equal(kept, xyz[xyz[:, 1] <= ylim]) and len(kept) == 1003 - (k - 1)
    with pytest.raises(ValueError):
        O.remove_ceiling(xyz[:4])  # n < 5 => k = 0 => error (Main.hs:2650 + VectorUtil.hs:13)


def test_haskell_show_float_matches_reference_literals():
    """the corner literals in Main.hs are GHC `show` output: they must round-trip through our formatter"""
    for lit in ["9.671569e-2", "-0.80041015", "1.5652194", "0.5213087", "9.584594e-2", "-2.2253428", "4.8369703", "0.46467793"]:
        assert O.haskell_show_float(np.float32(float(lit))) == lit
    assert O.haskell_show_float(1.0) == "1.0" and O.haskell_show_float(0.1) == "0.1" and O.haskell_show_float(12345678.0) == "1.2345678e7"
    assert O.haskell_show_float(5e-2) == "5.0e-2" and O.haskell_show_float(-0.0) == "-0.0"
    M = O.proj_translate4([6, 0, -1.5], O.proj_identity())
    assert O.room_projection_to_string(M) == "1.0,0.0,0.0,6.0,0.0,1.0,0.0,0.0,0.0,0.0,1.0,-1.5,0.0,0.0,0.0,1.0"
    assert O.room_projection_to_xf(M).splitlines()[0] == "1.0 0.0 0.0 6.0"


# ---- committed golden vectors (generated by tests/golden/make_golden.py from the oracle; they freeze the oracle)
def test_oracle_matches_committed_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    xyz, mask = O.backproject_ref(g["depth"], int(g["w"]), int(g["h"]))
    assert np.array_equal(mask, g["mask"]) and np.array_equal(xyz.view(np.uint32), g["bp_xyz"].view(np.uint32))
    planes = O.planes_from_cuboid(g["params"])
    assert np.array_equal(planes.view(np.uint32), g["planes"].view(np.uint32))
    a, r = O.plane_assign(g["cloud"], planes)
    assert np.array_equal(a, g["assign"]) and np.array_equal(r.view(np.uint32), g["resid"].view(np.uint32))
    f, grad, cnt, _ = O.cuboid_residual_grad(g["cloud"], g["params"])
    assert np.array_equal(cnt, g["counts"]) and abs(f - float(g["f"])) <= 1e-13 * f
    assert np.allclose(grad, g["grad"], rtol=1e-11, atol=1e-11 * np.abs(g["grad"]).max())
    assert np.array_equal(O.cc_label(g["cc_src"], g["cc_dst"], int(g["cc_n"])), g["cc_label"])
    assert O.kth_largest(g["cloud"][:, 1], len(g["cloud"]) // 5) == g["kth"]
    assert np.allclose(O.backproject_reduce6x6(g["depth"], int(g["w"]), int(g["h"]), planes), g["ne29"], rtol=1e-12)
