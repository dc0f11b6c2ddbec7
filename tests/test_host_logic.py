"""Host-side logic of the product (C++ mirror of the Haskell modules, Python module shims) against the oracle.
No GPU needed: plane construction, chain rule, Nelder-Mead, least squares, export formats, sharding helpers."""
import json
import math
import os

import numpy as np
import pytest

import housescan_b200 as hb
import oracle as O
from housescan_b200 import FitCuboidBFGS as F
from housescan_b200 import TranslationOptimizer as T
from housescan_b200 import synth
from housescan_b200.Bijection import biject
from housescan_b200.rooms import local_room_offsets, shard_range

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_planes_from_cuboid_bit_exact_vs_oracle():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        p = np.concatenate([rng.normal(size=3) * 10, rng.uniform(0.5, 12, 3), rng.normal(size=4)])
        a, b = hb.planes_from_cuboid(p), O.planes_from_cuboid(p)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_chain_rule_from_sums_matches_per_point_gradient():
    rng = np.random.default_rng(1)
    for _ in range(5):
        true = np.concatenate([rng.normal(size=3), rng.uniform(2, 6, 3), rng.normal(size=4)])
        xyz, _ = synth.cuboid_room_cloud(20_000, true, sigma=0.01, rng=rng)
        p = true + 0.03 * rng.normal(size=10)
        f, g, c = hb.cuboid_grad_from_sums(p, O.cuboid_sums(xyz, p))
        f_o, g_o, c_o, gs = O.cuboid_residual_grad(xyz, p)
        assert np.array_equal(c, c_o) and abs(f - f_o) <= 1e-13 * f_o
        assert np.max(np.abs(g - g_o) / gs) < 1e-12
        assert abs(np.dot(g, np.concatenate([np.zeros(6), p[6:]]))) < 1e-9 * gs[6:].max()  # scale of q is a gauge direction


def test_eight_corner_functions_match_oracle():
    rng = np.random.default_rng(2)
    for _ in range(200):
        p = np.concatenate([rng.normal(size=3) * 5, rng.uniform(0.5, 9, 3), rng.normal(size=4)])
        pts = rng.normal(size=(8, 3)) * 3
        assert np.array_equal(F.cuboidFromParams(p), O.cuboid_from_params(p))
        assert F.errfun(pts, p) == O.errfun(pts, p)
        assert F.errfunClosest(pts, p) == O.errfun_closest(pts, p)
        assert np.array_equal(np.array(F.guessDims(pts)), O.guess_dims(pts), equal_nan=True)  # sqrt of a negative for non-cuboids: NaN in both
    with pytest.raises(ValueError, match="bad arguments passed to cuboidFromParams"):
        F.cuboidFromParams([1, 2, 3])


def test_nelder_mead_fit_reaches_the_oracle_end_state():
    pts = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [2, 0, 0], [2, 0, 1], [2, 1, 0], [2, 1, 1]], float) @ O.rot_matrix3([1, 2, 3], math.radians(20))
    for fit_p, fit_o in ((F.fitCuboid, O.fit_cuboid), (F.fitCuboidFromCenter, O.fit_cuboid_from_center), (F.fitCuboidFromCenterFirst, O.fit_cuboid_from_center_first)):
        p, steps, err, path = fit_p(pts)
        po, so, eo, patho = fit_o(pts)
        assert err < 1e-12 and eo < 1e-12
        assert np.allclose(sorted(p[3:6]), [1, 1, 2], atol=1e-5) and np.allclose(p[:3], po[:3], atol=1e-6)
        assert abs(steps - so) <= max(5, 0.05 * so)  # same algorithm; last-ulp summation order may shift a few iterations
        assert path.shape[1] == patho.shape[1] and path[0, 0] == 1
    err, steps = F.fitCuboidFromCenterFirstError(pts)
    assert err < 1e-12 and steps > 0
    rooms = json.load(open(os.path.join(GOLDEN, "room_corners.json")))
    for name, corners in rooms.items():
        p, steps, err, _ = F.fitCuboidFromCenterFirst(np.array(corners, float))
        po, so, eo, _ = O.fit_cuboid_from_center_first(np.array(corners, float))
        assert abs(err - eo) <= 1e-6 * max(1.0, eo), name


def test_lstsq_distances_matches_oracle_incl_quirks():
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = int(rng.integers(2, 9))
        names = [f"r{i}" for i in rng.permutation(12)[:n]]
        edges = {}
        for i in range(n - 1):  # a spanning chain keeps it solvable
            edges[(names[i], names[i + 1])] = float(rng.normal())
        for _ in range(int(rng.integers(0, 6))):
            a, b = rng.choice(n, 2, replace=False)
            edges[(names[a], names[b])] = float(rng.normal())
        res_p, res_o = T.lstSqDistances(edges), O.lst_sq_distances(edges)
        assert (res_p is None) == (res_o is None)
        if res_p:
            assert res_p[0].keys() == res_o[0].keys()
            assert all(abs(res_p[0][k] - res_o[0][k]) < 1e-9 for k in res_o[0])
            assert abs(res_p[1] ** 2 - res_o[1] ** 2) < 1e-12  # rmse = sqrt(||r||_2 / m): compare before the sqrt amplifies 1e-16
    assert T.lstSqDistances({(1, 2): 1.0, (3, 4): 1.0}) is None  # Nothing


def test_proj_export_formats_match_oracle():
    rng = np.random.default_rng(4)
    for _ in range(200):
        M = O.proj_translate4(rng.normal(size=3) * 10 ** rng.uniform(-3, 3), O.proj_linear(O.rot_matrix3(rng.normal(size=3), float(rng.uniform(0, 6)), np.float32)))
        M[rng.integers(0, 3), rng.integers(0, 3)] *= 10 ** rng.uniform(-9, 9)
        assert hb.proj_to_string(M) == O.room_projection_to_string(M)
        assert hb.proj_to_xf(M) == O.room_projection_to_xf(M)


def test_biject_and_sharding_helpers():
    idx, unb = biject(["c", "a", "c", "b"])
    assert idx == O.biject(["c", "a", "c", "b"])[0] and unb == ["c", "a", "b"]
    n = 100_000_008
    cover = 0
    for world in (1, 2, 3, 4, 8):
        prev = 0
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            assert lo == prev and lo % 4 == 0 and hi <= n
            prev = hi
        assert prev == n
    offs = np.arange(13) * 8_333_334
    lo, hi = shard_range(n, 3, 8)
    loc = local_room_offsets(offs, lo, hi)
    assert loc[0] == 0 and loc[-1] == hi - lo and np.all(np.diff(loc) >= 0)
    assert np.sum(np.diff(loc)) == hi - lo


def test_optimize_room_positions_reference_semantics():
    """Main.optimizeRoomPositions (Main.hs:2089-2168) with two rooms: an `Opposite 0.1` wall pair along X ends 0.1 m apart"""
    from housescan_b200.rooms import X, Opposite, optimizeRoomPositions

    corner_mean = {1: np.array([0.0, 0, 0]), 2: np.array([5.3, 0, 0])}
    plane_mean = {(1, 0): np.array([2.0, 0, 0]), (2, 1): np.array([5.3 - 2.5, 0, 0])}  # +x wall of room 1, -x wall of room 2

    class FakeCtx:  # groupConnectedComponents needs GPU labels; on CPU use the oracle's labels through the same interface
        def group_cc(self, src, dst, n):
            lab = O.cc_label(src, dst, n)
            uniq = sorted(set(lab[src]))
            comp = np.array([uniq.index(lab[s]) for s in src], np.int32)
            order = np.concatenate([np.flatnonzero(comp == c)[::-1] for c in range(len(uniq))]).astype(np.int64)
            return comp, order, len(uniq)

    moved, log = optimizeRoomPositions([1, 2], [(X, Opposite(0.1), (1, 0), (2, 1))], lambda r, w: plane_mean[(r, w)], lambda r: corner_mean[r], ctx=FakeCtx())
    # o = (2.0 - 0) - (2.8 - 5.3) = 4.5; desired centre offset = 4.5 + 0.1 => room 2 sits at x = 4.6
    assert abs(moved[2][0] - (4.6 - 5.3)) < 1e-6 and moved[1][0] == 0 and not moved[2][1:].any()
    assert any("Aligning the X (1 components)" in l for l in log)


# ------------------------------------------------------------------ exact constant-divisor division used by the depth kernels
def _fma32(a, b, c):
    """RN_f32(a*b + c) for float32 arrays: the product of two floats is exact in float64 and the sum is rounded once there;
    the second rounding to float32 can differ from a fused one only on exact float64 ties, which the checks below would catch."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def test_constant_division_sequence():
    """k_common.cuh div_rn_small: q0 = RN(a y), q = RN(q0 + RN(a - b q0) y) with y = RN(1/b) equals a / b for EVERY integer
    a in [0, 2^24) and b in {10, 20} (scalePoints, Main.hs:1311-1313: x/10, y/10, d/20; a tall raster of many frames has y > 65535)."""
    a = np.arange(1 << 24, dtype=np.float32)
    for b in (np.float32(10.0), np.float32(20.0)):
        y = np.float32(1.0) / b
        q0 = a * y
        q = _fma32(_fma32(np.full_like(a, -b), q0, a), np.full_like(a, y), q0)
        assert np.array_equal(q.view(np.uint32), (a / b).view(np.uint32))
        assert (q0 != a / b).any()  # the multiplication alone is NOT enough


def test_general_division_sequence_random():
    """k_common.cuh div_rn_by (two residual steps, Markstein): equals IEEE division on random pinhole-style operands"""
    rng = np.random.default_rng(7)
    for b in (np.float32(525.0), np.float32(131.25), np.float32(570.3422), np.float32(3.0), np.float32(0.7071)):
        a = ((rng.integers(0, 640, 2_000_00).astype(np.float32) - np.float32(319.5)) * (rng.integers(1, 65536, 2_000_00).astype(np.float32) * np.float32(0.001))).astype(np.float32)
        y = np.float32(1.0) / b
        nb, yy = np.full_like(a, -b), np.full_like(a, y)
        q0 = a * y
        q1 = _fma32(_fma32(nb, q0, a), yy, q0)
        q2 = _fma32(_fma32(nb, q1, a), yy, q1)
        assert np.array_equal(q2.view(np.uint32), (a / b).view(np.uint32)), b


def test_connect_walls_reference_semantics():
    """Main.connectWalls (Main.hs:2039-2068): axis = the one the normals are most parallel to (ties go to the later axis), walls
    that disagree are refused, duplicates in either order are ignored, new connections go in front"""
    from housescan_b200.rooms import X, Y, Z, Opposite, SAME, bestAxis, connectWalls

    assert bestAxis([0.9, 0.1, -0.3]) == X and bestAxis([0.1, -0.8, 0.3]) == Y and bestAxis([0, 0.2, -0.7]) == Z
    assert bestAxis([0.5, 0.5, 0.0]) == Y and bestAxis([0.5, 0.5, 0.5]) == Z  # tuple maximum: the later axis wins a tie
    conns = []
    conns, msg = connectWalls(conns, Opposite(0.1), (1, 0), (2, 1), [1, 0, 0], [-0.99, 0.1, 0])
    assert msg is None and conns == [(X, Opposite(0.1), (1, 0), (2, 1))]
    conns, msg = connectWalls(conns, SAME, (2, 4), (3, 5), [0, 0.1, 1], [0, 0, -1])
    assert conns[0] == (Z, SAME, (2, 4), (3, 5)) and len(conns) == 2  # newest first
    again, _ = connectWalls(conns, SAME, (2, 1), (1, 0), [1, 0, 0], [1, 0, 0])
    assert again == conns  # already connected (reversed order counts)
    refused, msg = connectWalls(conns, SAME, (5, 0), (6, 2), [1, 0, 0], [0, 1, 0])
    assert refused == conns and msg == "Could not guess axis of wall connection"


def test_plane_algebra_matches_oracle_bit_for_bit(built_lib):
    """rotationBetweenPlaneEqs / rotatePlaneEqAround / translatePlaneEq (Main.hs:1553-1578, :1681-1688) through the C ABI against
    the oracle's restatement; and the property the reference's doc comment states: n1 .* R points along n2"""
    from housescan_b200.rooms import rotatePlaneEqAround, rotationBetweenPlaneEqs, translatePlaneEq

    rng = np.random.default_rng(8)
    for _ in range(50):
        p1 = O.mk_plane_eq(rng.normal(size=3), rng.normal())
        p2 = O.mk_plane_eq(rng.normal(size=3), rng.normal())
        R = rotationBetweenPlaneEqs(p1, p2)
        assert np.array_equal(R.view(np.uint32), O.rotation_between_normals(p1[:3], p2[:3]).view(np.uint32))
        assert np.allclose(p1[:3] @ R, p2[:3], atol=2e-6)
        c = rng.normal(size=3).astype(np.float32)
        a = rotatePlaneEqAround(c, R, p1)
        assert np.array_equal(a.view(np.uint32), O.rotate_plane_eq_around(c, R, p1).view(np.uint32))
        off = rng.normal(size=3).astype(np.float32)
        t = translatePlaneEq(off, p2)
        assert np.array_equal(t.view(np.uint32), O.translate_plane_eq(off, p2).view(np.uint32))
    with np.errstate(invalid="ignore"):
        assert np.isnan(rotationBetweenPlaneEqs([0, 0, 1, 0], [0, 0, 1, 5])).any()  # parallel normals: NaN, as in the reference


def test_plane_corner_matches_lapack_and_the_cuboid_corners(built_lib):
    """planeCorner (Main.hs:1413-1430) against numpy's dgesv on random plane triples, and on a cuboid's own walls: every corner of
    cuboidFromParams is where its three walls meet (the reference's fitCuboidToRoom relies on that, Main.hs:1805-1812)"""
    from housescan_b200 import FitCuboidBFGS
    from housescan_b200.rooms import planeCorner

    rng = np.random.default_rng(12)
    worst = 0
    for _ in range(300):
        pl = [O.mk_plane_eq(rng.normal(size=3), rng.normal() * 3) for _ in range(3)]
        got, exp = planeCorner(*pl), O.plane_corner(*pl)
        ulps = np.abs(got.view(np.int32).astype(np.int64) - exp.view(np.int32).astype(np.int64)).max()
        worst = max(worst, int(ulps))
    assert worst <= 1, worst  # same algorithm as LAPACK's; at most the last Float bit may differ (measured: 0)
    assert planeCorner([1, 0, 0, 1], [1, 0, 0, 2], [0, 1, 0, 0]) is None and O.plane_corner([1, 0, 0, 1], [1, 0, 0, 2], [0, 1, 0, 0]) is None
    params = np.array([0.3, -0.2, 4.0, 5.0, 2.6, 4.0, 0.9, 0.1, 0.3, 0.2])
    planes = hb.planes_from_cuboid(params)
    corners = FitCuboidBFGS.cuboidFromParams(params)
    found = np.array([planeCorner(planes[i], planes[2 + j], planes[4 + k]) for i in (0, 1) for j in (0, 1) for k in (0, 1)])
    for c in corners:
        assert np.abs(found - c.astype(np.float32)).sum(axis=1).min() < 1e-5


class _OracleEngine:
    """CPU stand-in for the Context methods a Room uses (the oracle is the checker's cloud engine here)"""

    def mean_extent(self, cloud):
        m = O.point_mean_f64(cloud)
        d = (np.asarray(cloud, np.float32) - m.astype(np.float32)).astype(np.float64)
        return m, np.float32(np.sqrt((d * d).sum(axis=1).max()))

    def rotate_around(self, cloud, c, R):
        return O.rotate_cloud_around(cloud, c, R)

    def translate(self, cloud, off):
        return O.translate_cloud(cloud, off)

    def transform(self, cloud, P):
        return O.project_cloud(cloud, P)


def test_proj_algebra_and_room_movers(built_lib):
    """hs_proj_* against the oracle's Float matrix algebra (bit for bit), and rotateRoomAround / translateRoom / projectRoom
    (Main.hs:1665-1730) moving planes, cloud, corners and roomProj consistently: replaying the accumulated roomProj on the fresh
    room reproduces the incrementally moved one (the reference's projTest property, Main.hs:2543-2634)"""
    from housescan_b200 import FitCuboidBFGS
    from housescan_b200.rooms import Room, projCompose, projRotateAround, projTranslate

    rng = np.random.default_rng(21)
    A = rng.normal(size=(4, 4)).astype(np.float32)
    B = rng.normal(size=(4, 4)).astype(np.float32)
    c, off = rng.normal(size=3).astype(np.float32), rng.normal(size=3).astype(np.float32)
    R = O.rot_matrix3([1, 2, 3], 0.7, np.float32)
    assert np.array_equal(projCompose(A, B).view(np.uint32), O.proj_compose(A, B).view(np.uint32))
    assert np.array_equal(projTranslate(A, off).view(np.uint32), O.proj_translate4(off, A).view(np.uint32))
    assert np.array_equal(projRotateAround(A, c, R).view(np.uint32), O.proj_rotate_around(c, R, A).view(np.uint32))

    params = np.array([0.3, -0.2, 4.0, 5.0, 2.6, 4.0, 0.9, 0.1, 0.3, 0.2])
    xyz, _ = synth.cuboid_room_cloud(5_000, params, sigma=0.0, seed=3)
    planes = hb.planes_from_cuboid(params)
    corners = FitCuboidBFGS.cuboidFromParams(params).astype(np.float32)
    room = Room(_OracleEngine(), xyz.copy(), planes, corners)
    room.rotateRoom(O.rot_matrix3([1, 0, 0], math.radians(90), np.float32)).translateRoom([6, 0, -1.5]).rotateRoomAround([1, 2, 3], R)
    fresh = Room(_OracleEngine(), xyz.copy(), planes, corners).projectRoom(room.proj)
    assert np.allclose(room.cloud, fresh.cloud, atol=5e-5) and np.allclose(room.corners, fresh.corners, atol=5e-5)
    assert np.allclose(room.planes, fresh.planes, atol=5e-5)
    assert np.array_equal(fresh.proj.view(np.uint32), O.proj_compose(np.eye(4, dtype=np.float32), room.proj).view(np.uint32))
    # the moved planes still carry the moved points: every point of the noiseless cloud lies on its nearest moved wall
    a, r = O.plane_assign(room.cloud, room.planes)
    assert np.abs(r).max() < 1e-4
    # the moved corners are where the moved walls meet
    from housescan_b200.rooms import planeCorner
    found = np.array([planeCorner(room.planes[i], room.planes[2 + j], room.planes[4 + k]) for i in (0, 1) for j in (0, 1) for k in (0, 1)])
    for cc in room.corners:
        assert np.abs(found - cc).sum(axis=1).min() < 1e-4
    with pytest.raises(ValueError, match="last column"):
        bad = np.eye(4, dtype=np.float32)
        bad[0, 3] = 0.5
        room.projectRoom(bad)


def test_fit_cuboid_to_room_glue(built_lib):
    """Main.fitCuboidToRoom (Main.hs:1814-1849) on one of the reference's real rooms: planes and corners are replaced by the fitted
    cuboid's, every plane holds exactly 4 of the 8 new corners (the reference's own assert, Main.hs:1881), connections of the
    room's old walls disappear, fewer than 8 corners change nothing"""
    from housescan_b200.rooms import X, Opposite, Room, fitCuboidToRoom

    corners = np.array(json.load(open(os.path.join(os.path.dirname(__file__), "golden", "room_corners.json")))["testroom1"], np.float32)
    room = Room(_OracleEngine(), np.zeros((4, 3), np.float32), np.zeros((0, 4), np.float32), corners)
    conns = [(X, Opposite(0.1), (7, 0), (9, 1)), (X, Opposite(0.1), (3, 0), (4, 1))]
    log, params, steps, err, kept = fitCuboidToRoom(room, conns, room_id=7)
    assert steps > 50 and err < 1.0 and "RMSE" in log[1]  # the author's bar for these rooms (FitCuboidBFGS.hs:278)
    assert kept == [(X, Opposite(0.1), (3, 0), (4, 1))]
    assert room.planes.shape == (6, 4) and room.corners.shape == (8, 3)
    on = np.abs(room.corners @ room.planes[:, :3].T - room.planes[:, 3]) < 1e-4
    assert (on.sum(axis=0) == 4).all() and (on.sum(axis=1) == 3).all()  # 4 corners per wall, 3 walls per corner
    # the fitted box sits on the picked corners as a SET: the ids are re-used positionally (Main.hs:1839), so id <-> place is the
    # reference's to scramble; the closest-corner objective does not care about order
    d = np.linalg.norm(room.corners[:, None, :] - corners[None, :, :], axis=2)
    assert d.min(axis=1).max() < 0.6 and len(set(d.argmin(axis=1))) == 8
    few = Room(_OracleEngine(), np.zeros((4, 3), np.float32), np.zeros((0, 4), np.float32), corners[:5])
    log, params, steps, err, kept = fitCuboidToRoom(few, conns, room_id=7)
    assert params is None and "need 8" in log[1] and kept == conns and len(few.corners) == 5


def test_auto_align_floor_and_corner_suggestions(built_lib):
    """roomAutoAlignAxis / autoAlignFloor (Main.hs:1895-1910) and suggestPoints (Main.hs:1521-1538) on a tilted cuboid room: after the
    alignment one wall normal is the up axis; the 6 planes have 20 triples of which exactly the 8 real corners survive the cutoff"""
    from housescan_b200 import FitCuboidBFGS
    from housescan_b200.rooms import Room

    params = np.array([0.3, -0.2, 4.0, 5.0, 2.6, 4.0, 0.95, 0.05, 0.2, 0.1])  # a slightly tilted room
    xyz, _ = synth.cuboid_room_cloud(4_000, params, sigma=0.0, seed=5)
    room = Room(_OracleEngine(), xyz.copy(), hb.planes_from_cuboid(params), FitCuboidBFGS.cuboidFromParams(params).astype(np.float32))
    before = room.planes.copy()
    assert room.autoAlignFloor() is None
    up = room.planes[:, 1]
    assert np.isclose(up.max(), 1.0, atol=1e-6) and np.isclose(up.min(), -1.0, atol=1e-6)  # floor and ceiling normals are +-Y now
    assert not np.allclose(before, room.planes)
    a, r = O.plane_assign(room.cloud, room.planes)
    assert np.abs(r).max() < 1e-4  # cloud and planes moved together
    sugg, triples = room.suggestPoints(1.2)
    assert triples == 20 and len(sugg) == 8
    d = np.linalg.norm(sugg[:, None, :] - room.corners[None, :, :], axis=2)
    assert d.min(axis=1).max() < 1e-4  # the suggestions are the room's corners
    assert Room(_OracleEngine(), xyz, np.zeros((0, 4), np.float32)).autoAlignFloor() == "room has no planes"


def test_bfgs_minimize_callback_entry_point(built_lib):
    """`hs_bfgs_minimize`: the library's BFGS over a caller-supplied objective (quadratic with a known minimum; an objective that
    fails is reported, not swallowed)."""
    from housescan_b200 import FitCuboidBFGS as F
    from housescan_b200 import HsError

    A = np.diag([1.0, 4.0, 9.0, 0.5])
    b = np.array([1.0, 2.0, 3.0, -1.0])
    x, f, it, ev = F.bfgsMinimize(lambda x: (0.5 * x @ A @ x - b @ x, A @ x - b), np.zeros(4), 200, 1e-10)
    assert np.allclose(x, np.linalg.solve(A, b), atol=1e-8) and it > 0 and ev >= it

    def bad(x):
        raise RuntimeError("objective failed")

    with pytest.raises(HsError):
        F.bfgsMinimize(bad, np.zeros(4))


def _eval_plan(n, offs, sm_count=148, seg_cost=0):
    import ctypes as C

    from housescan_b200 import _lib as L

    offs = np.ascontiguousarray(offs, dtype=np.int64)
    nr = offs.size - 1
    nb = C.c_int32()
    g0 = np.zeros(257, np.int64)
    rf, rl = np.zeros(256, np.int32), np.zeros(256, np.int32)
    blo, nbr = np.zeros(nr, np.int32), np.zeros(nr, np.int32)
    rc = L.load().hs_eval_plan(n, L.ptr(offs), nr, sm_count, seg_cost, C.byref(nb), L.ptr(g0), L.ptr(rf), L.ptr(rl), L.ptr(blo), L.ptr(nbr))
    assert rc == 0
    return nb.value, g0[: nb.value + 1], rf[: nb.value], rl[: nb.value], blo, nbr


def test_eval_plan_partition_invariants(built_lib):
    """The evaluation kernel's static partition (`hs_eval_plan`, host only).  Every 4-point group belongs to exactly one block, block
    ranges are contiguous and ordered, a room's ticket count equals the number of blocks that hold points of it (rooms without
    points take none - the bug the random-layout GPU test found), and a block that keeps a room boundary inside its range streams
    fewer groups than a block without one."""
    rng = np.random.default_rng(5)
    for trial in range(300):
        nrooms = int(rng.integers(1, 13))
        scale = int(rng.choice([20, 3_000, 400_000, 9_000_000]))
        sizes = rng.integers(0, scale, size=nrooms)
        sizes[rng.random(nrooms) < 0.2] = 0
        lead, trail = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        offs = np.concatenate([[0], np.cumsum(sizes)]) + lead
        n = int(offs[-1]) + trail
        if n == 0:
            continue
        nb, g0, rf, rl, blo, nbr = _eval_plan(n, offs, sm_count=int(rng.choice([1, 7, 148])))
        G = (n + 3) // 4
        assert g0[0] == 0 and g0[-1] == G and np.all(np.diff(g0) >= 0), (trial, g0)
        holds = np.zeros((nb, nrooms), bool)
        for b in range(nb):
            p0, p1 = 4 * g0[b], min(4 * g0[b + 1], n)
            for r in range(nrooms):
                holds[b, r] = offs[r] < p1 and offs[r + 1] > p0 and offs[r + 1] > offs[r] and p1 > p0
            rooms = np.flatnonzero(holds[b])
            if rooms.size:
                assert rf[b] == rooms[0] and rl[b] == rooms[-1], (trial, b)
            else:
                assert rl[b] < rf[b]
        for r in range(nrooms):
            blocks = np.flatnonzero(holds[:, r])
            assert nbr[r] == blocks.size, (trial, r, nbr[r], blocks)
            if blocks.size:
                assert blo[r] == blocks[0] and np.array_equal(blocks, np.arange(blocks[0], blocks[0] + blocks.size))  # contiguous
    # the apartment of the bench on one of eight GPUs: blocks with a boundary inside stream ~1536 groups fewer, nobody more than the target
    per = 8_333_334
    offs = np.clip(np.arange(13) * per, 0, 12_500_004)
    nb, g0, rf, rl, blo, nbr = _eval_plan(12_500_004, offs)
    sizes = np.diff(g0)
    two = rl > rf
    assert nb == 148 and two.sum() == 1
    assert sizes[two][0] <= sizes[~two].max() - 1000 and sizes.max() - np.median(sizes) <= 16


def test_nm_minimize_callback_entry_point_matches_the_oracle_simplex(built_lib, oracle_lib):
    """`hs_nm_minimize` = the library's NMSimplex2 over a caller-supplied objective: the same end state as the oracle's restatement
    of GSL's nmsimplex2 on the reference's own fitting objective (errfun over 8 corners, FitCuboidBFGS.hs:51-66) and on Rosenbrock;
    iteration counts equal (same rules, same comparisons)."""
    import oracle as O
    from housescan_b200 import FitCuboidBFGS as F

    box = O.cuboid_from_params(np.array([0.3, -0.2, 1.0, 2.0, 1.0, 3.0, 0.9, 0.1, -0.2, 0.3]))
    x0 = np.array([0.0, 0.0, 0.0, 1.5, 1.5, 1.5, 0.1, 0.1, 0.1, 0.1])
    step = np.array([0.01, 0.01, 0.01, 0.15, 0.15, 0.15, 0.1, 0.1, 0.1, 0.1])
    obj = lambda p: O.errfun(box, p)
    x_o, path = O.nm_simplex2(obj, x0, step, 1e-8, 600)
    x_g, f_g, it_g, ev_g = F.nmMinimize(obj, x0, step, 1e-8, 600)
    assert it_g == int(path[-1][0]) and ev_g > it_g
    assert np.allclose(x_g, x_o, rtol=0, atol=1e-9) and abs(f_g - obj(x_o)) <= 1e-12 + 1e-9 * abs(f_g)
    rosen = lambda x: (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    x_o, path = O.nm_simplex2(rosen, np.array([-1.2, 1.0]), np.array([0.5, 0.5]), 1e-10, 2000)
    x_g, f_g, it_g, _ = F.nmMinimize(rosen, [-1.2, 1.0], [0.5, 0.5], 1e-10, 2000)
    assert it_g == int(path[-1][0]) and np.allclose(x_g, x_o, atol=1e-9) and np.allclose(x_g, [1.0, 1.0], atol=1e-6)
