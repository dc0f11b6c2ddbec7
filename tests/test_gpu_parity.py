"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for indices / masks / labels / Float geometry; Double sums within the tolerance written in each test
(the north-star bar is 1e-6 relative; measured agreement is far tighter and asserted as such)."""
import math

import numpy as np
import pytest

import oracle as O
from housescan_b200 import synth

pytestmark = pytest.mark.gpu


def _rel(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = np.maximum(np.abs(b), 1e-300) if scale is None else np.maximum(np.asarray(scale, float), 1e-300)
    return float(np.max(np.abs(a - b) / s))


@pytest.fixture(scope="module")
def c1():
    """config 1: one 640x480 frame of a cuboid room -> reference back-projection -> room cloud in its own frame."""
    depth = synth.render_depth_frame()
    return depth


@pytest.fixture(scope="module")
def room_small():
    params = synth.C1_PARAMS.copy()
    xyz, face = synth.cuboid_room_cloud(307_200, params, sigma=0.005, seed=11)
    return xyz, params


# ------------------------------------------------------------------ (1) back-projection
# mode key 10: 0 = single pass (decoupled look-back), 1 = count pass + scatter pass
@pytest.fixture(params=[0, 1], ids=["onepass", "twopass"])
def bp_mode(ctx, request):
    ctx.set_mode(10, request.param)
    yield request.param
    ctx.set_mode(10, 0)


def test_backproject_long_raster_many_tiles(ctx, bp_mode):
    """40 frames as one tall raster (6000 tiles: several look-back windows deep), runs of invalid tiles, repeated launches"""
    rng = np.random.default_rng(77)
    w, h = 640, 480 * 40
    depth = rng.integers(1, 65536, size=(h, w), dtype=np.uint16)
    depth[rng.random((h, w)) < 0.1] = 0
    depth[1000:1500] = 0      # ~150 consecutive tiles without a valid pixel
    depth[9000:9001, :7] = 0
    xyz_o, mask_o = O.backproject_ref(depth, w, h)
    for _ in range(3):
        xyz_g, mask_g = ctx.backproject_ref(depth, w, h)
        assert np.array_equal(mask_g, mask_o)
        assert xyz_g.shape == xyz_o.shape and np.array_equal(xyz_g.view(np.uint32), xyz_o.view(np.uint32))


@pytest.mark.parametrize("w,h", [(640, 480), (64, 48), (37, 5), (1, 1), (2048, 3), (2047, 2)])
def test_backproject_mask_and_points_bit_exact(ctx, w, h, bp_mode):
    rng = np.random.default_rng(w * 1000 + h)
    depth = rng.integers(0, 65536, size=(h, w), dtype=np.uint16)
    depth[rng.random((h, w)) < 0.3] = 0
    xyz_o, mask_o = O.backproject_ref(depth, w, h)
    xyz_g, mask_g = ctx.backproject_ref(depth, w, h)
    assert np.array_equal(mask_g, mask_o)
    assert xyz_g.shape == xyz_o.shape
    assert np.array_equal(xyz_g.view(np.uint32), xyz_o.view(np.uint32))


def test_backproject_all_invalid_and_all_valid(ctx, bp_mode):
    z = np.zeros((48, 64), np.uint16)
    xyz, mask = ctx.backproject_ref(z, 64, 48)
    assert xyz.shape == (0, 3) and mask.sum() == 0
    f = np.full((48, 64), 65535, np.uint16)
    xyz, mask = ctx.backproject_ref(f, 64, 48)
    xo, mo = O.backproject_ref(f, 64, 48)
    assert np.array_equal(xyz.view(np.uint32), xo.view(np.uint32)) and mask.all()


def test_backproject_every_depth_value_divides_exactly(ctx):
    """d/20 - 30 for every uint16 and x/10 for every column index: IEEE division on the device == host."""
    depth = np.arange(65536, dtype=np.uint16).reshape(16, 4096)
    xg, mg = ctx.backproject_ref(depth, 4096, 16)
    xo, mo = O.backproject_ref(depth, 4096, 16)
    assert np.array_equal(mg, mo) and np.array_equal(xg.view(np.uint32), xo.view(np.uint32))


def test_config1_frame(ctx, c1):
    xyz_o, mask_o = O.backproject_ref(c1, 640, 480)
    xyz_g, mask_g = ctx.backproject_ref(c1, 640, 480)
    assert np.array_equal(mask_g, mask_o) and np.array_equal(xyz_g.view(np.uint32), xyz_o.view(np.uint32))
    assert 0.97 < mask_g.mean() < 0.99  # 2 % invalid pixels


# ------------------------------------------------------------------ (2) planes
@pytest.mark.parametrize("n", [307_200, 1, 3, 4, 5, 1023, 4097])
def test_plane_assign_bit_exact(ctx, room_small, n):
    xyz, params = room_small
    xyz = xyz[:n]
    planes = O.planes_from_cuboid(params)
    cl = ctx.upload(xyz)
    a_g, r_g = ctx.plane_assign(cl, planes)
    a_o, r_o = O.plane_assign(xyz, planes)
    assert np.array_equal(a_g, a_o)
    assert np.array_equal(r_g.view(np.uint32), r_o.view(np.uint32))


def test_plane_assign_ties_pick_lowest_index(ctx):
    """points on the bisector planes of an axis-aligned unit cube: exact Float ties -> first minimum"""
    params = np.array([0, 0, 0, 2, 2, 2, 1, 0, 0, 0], float)
    planes = O.planes_from_cuboid(params)
    g = np.linspace(-1, 1, 9, dtype=np.float32)
    xyz = np.array([[x, y, z] for x in g for y in g for z in g], np.float32)
    a_g, _ = ctx.plane_assign(ctx.upload(xyz), planes)
    a_o, _ = O.plane_assign(xyz, planes)
    assert np.array_equal(a_g, a_o)
    assert a_g[np.all(xyz == 0, axis=1)][0] == 0  # centre: six-way tie -> plane 0


def test_plane_assign_generic_k(ctx):
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(50_000, 3)).astype(np.float32) * 3
    for K in (1, 2, 7, 16):
        pl = np.stack([O.mk_plane_eq(rng.normal(size=3), rng.normal()) for _ in range(K)])
        a_g, r_g = ctx.plane_assign(ctx.upload(xyz), pl)
        a_o, r_o = O.plane_assign(xyz, pl)
        assert np.array_equal(a_g, a_o) and np.array_equal(r_g.view(np.uint32), r_o.view(np.uint32))


# evaluation kernels: "exact" = every product in Double (k_planes.cu); "fast" = the product default, the throughput kernel
# (k_eval.cuh: scalar Float products, 256-point Float chains, transposing warp sums, Double accumulation).  Assignment (counts)
# is bit-exact in both; sums: exact 1e-11 of the magnitude sum, fast 1e-6 (the north-star bar; measured ~1e-7).
EVAL_MODES = {"exact": (1, 1e-11), "fast": (2, 1e-6)}


@pytest.fixture(params=["exact", "fast"])
def eval_mode(request, ctx):
    mode, tol = EVAL_MODES[request.param]
    ctx.set_mode(0, mode)
    yield tol
    ctx.set_mode(0, 0)


def _record_scale(xyz, params):
    """magnitude sums the tolerances are relative to: sum |term| of every record entry"""
    a, r = O.plane_assign(xyz, O.planes_from_cuboid(params))
    r = r.astype(np.float64)
    sc = np.zeros(22)
    sc[0] = np.sum(r * r)
    for k in range(6):
        sc[1 + k] = np.sum(np.abs(r[a == k]))
    for j in range(3):
        m = (a >> 1) == j
        sc[7 + 3 * j : 10 + 3 * j] = np.sum(np.abs(r[m, None] * xyz[m].astype(np.float64)), axis=0)
    sc[16:22] = 1.0
    return np.maximum(sc, 1e-300)


def test_cuboid_sums_config1(ctx, room_small, eval_mode):
    xyz, params = room_small
    cl = ctx.upload(xyz)
    rec_g = ctx.rooms_cuboid_sums(cl, [0, len(xyz)], params[None])[0]
    rec_o = O.cuboid_sums(xyz, params)
    assert np.array_equal(rec_g[16:22], rec_o[16:22])  # counts exact
    assert _rel(rec_g[:22], rec_o[:22], _record_scale(xyz, params)) < eval_mode


@pytest.mark.parametrize("perturb", [0.0, 0.05])
def test_cuboid_residual_grad_config1(ctx, room_small, perturb, eval_mode):
    """f, gradient (10) and counts (6) vs the oracle's direct per-point Double accumulation, relative to the gradient's
    magnitude sum (sum |2 r dr/dtheta|)."""
    xyz, params = room_small
    rng = np.random.default_rng(3)
    p = params + perturb * rng.normal(size=10)
    cl = ctx.upload(xyz)
    f_g, g_g, c_g = ctx.cuboid_residual_grad(cl, p)
    f_o, g_o, c_o, gs = O.cuboid_residual_grad(xyz, p)
    assert np.array_equal(c_g, c_o)
    assert abs(f_g - f_o) <= eval_mode * f_o
    assert _rel(g_g, g_o, gs) < max(eval_mode, 1e-9)


def test_cuboid_gradient_matches_finite_differences(ctx, room_small):
    """fixed-assignment finite differences of the oracle objective in Double (self-check of the chain rule)."""
    xyz, params = room_small
    p = params + 0.02 * np.random.default_rng(1).normal(size=10)
    cl = ctx.upload(xyz[:50_000])
    f0, g, _ = ctx.cuboid_residual_grad(cl, p)
    pts = xyz[:50_000].astype(np.float64)
    a0, _ = O.plane_assign(xyz[:50_000], O.planes_from_cuboid(p))

    def fobj(q):
        R = synth.rot_rows_from_quat(q[6:])
        j, sg = a0 >> 1, np.where(a0 & 1, -1.0, 1.0)
        r = sg * np.einsum("ij,ij->i", pts - q[:3], R[j]) - q[3:6][j] / 2
        return float(np.sum(r * r))

    for m in range(10):
        h = 1e-6
        e = np.zeros(10)
        e[m] = h
        fd = (fobj(p + e) - fobj(p - e)) / (2 * h)
        assert abs(fd - g[m]) <= 1e-4 * abs(g[m]) + 5e-4 * abs(f0), (m, fd, g[m])  # Float residuals vs the Double model


def test_rooms_sums_ragged_rooms(ctx, eval_mode):
    """several rooms with offsets that are not multiples of 4, an empty room, points outside every room"""
    rng = np.random.default_rng(9)
    sizes = [1, 0, 5, 1023, 4096, 7, 20_001, 3]
    params = np.stack([np.concatenate([rng.normal(size=3), rng.uniform(1, 5, 3), rng.normal(size=4)]) for _ in sizes])
    clouds = [synth.cuboid_room_cloud(s, params[i], sigma=0.01, rng=rng)[0] for i, s in enumerate(sizes)]
    lead, trail = rng.normal(size=(6, 3)).astype(np.float32), rng.normal(size=(9, 3)).astype(np.float32)
    xyz = np.concatenate([lead] + clouds + [trail])
    offs = np.concatenate([[0], np.cumsum(sizes)]) + 6
    rec_g = ctx.rooms_cuboid_sums(ctx.upload(xyz), offs, params)
    for r, s in enumerate(sizes):
        rec_o = O.cuboid_sums(xyz[offs[r] : offs[r + 1]], params[r])
        assert np.array_equal(rec_g[r, 16:22], rec_o[16:22]), r
        if s:
            assert _rel(rec_g[r, :16], rec_o[:16], _record_scale(xyz[offs[r] : offs[r + 1]], params[r])[:16]) < eval_mode, r
        assert rec_g[r, 22] == 0 and rec_g[r, 23] == 0


def test_rooms_sums_many_rooms_chunked(ctx, eval_mode):
    """more rooms than one launch's table (HS_MAX_ROOMS = 32) + a multi-block cloud"""
    rng = np.random.default_rng(10)
    nrooms = 40
    sizes = rng.integers(1000, 40_000, size=nrooms)
    params = np.stack([np.concatenate([rng.normal(size=3), rng.uniform(1, 5, 3), rng.normal(size=4)]) for _ in sizes])
    xyz = np.concatenate([synth.cuboid_room_cloud(int(s), params[i], sigma=0.01, rng=rng)[0] for i, s in enumerate(sizes)])
    offs = np.concatenate([[0], np.cumsum(sizes)])
    rec_g = ctx.rooms_cuboid_sums(ctx.upload(xyz), offs, params)
    for r in (0, 13, 31, 32, 39):
        rec_o = O.cuboid_sums(xyz[offs[r] : offs[r + 1]], params[r])
        assert np.array_equal(rec_g[r, 16:22], rec_o[16:22])
        assert _rel(rec_g[r, :16], rec_o[:16], _record_scale(xyz[offs[r] : offs[r + 1]], params[r])[:16]) < eval_mode


def test_rooms_sums_additive_over_point_shards(ctx, room_small, eval_mode):
    """size-independent property used by the multi-GPU path: records of point shards add up to the whole"""
    from housescan_b200.rooms import local_room_offsets, shard_range

    xyz, params = room_small
    n = len(xyz)
    offs = np.array([0, 100_003, n])
    pp = np.stack([params, params + 0.01])
    whole = ctx.rooms_cuboid_sums(ctx.upload(xyz), offs, pp)
    acc = np.zeros_like(whole)
    for rank in range(3):
        lo, hi = shard_range(n, rank, 3)
        acc += ctx.rooms_cuboid_sums(ctx.upload(xyz[lo:hi]), local_room_offsets(offs, lo, hi), pp)
    assert np.array_equal(acc[:, 16:22], whole[:, 16:22])
    for r in range(2):
        assert _rel(acc[r, :16], whole[r, :16], _record_scale(xyz[offs[r] : offs[r + 1]], pp[r])[:16]) < 2 * eval_mode


# mode key 7: 0 = ring form (bulk-async tiles, Float chains of 64 points summed across the warp, then Double),
# 1 = all-Double form, 2 = direct loads with Float chains of 32 points and per-thread Doubles
@pytest.mark.parametrize("ps_mode,tol", [(0, 1e-6), (1, 1e-11), (2, 1e-6)])
def test_plane_sums_generic(ctx, room_small, ps_mode, tol):
    xyz, params = room_small
    offs = np.array([0, 100_000, 100_000, len(xyz)])
    planes = np.stack([O.planes_from_cuboid(params)] * 3)
    ctx.set_mode(7, ps_mode)
    try:
        out_g = ctx.plane_sums(ctx.upload(xyz), offs, planes, 6)
    finally:
        ctx.set_mode(7, 0)
    out_o = O.plane_sums(xyz, offs, planes, 6)
    assert np.array_equal(out_g[..., 0], out_o[..., 0])
    assert np.array_equal(out_g[..., 9], out_o[..., 9])  # max |r| exact
    assert np.allclose(out_g, out_o, rtol=tol, atol=1e-9 if ps_mode == 1 else 1e-6 * np.abs(out_o).max())


@pytest.mark.parametrize("ps_mode", [0, 1, 2])
def test_plane_sums_ragged_offsets_and_small_k(ctx, room_small, ps_mode):
    """room offsets that are not multiples of 4 (head / tail points), rooms shorter than a group, K from 1 to 8"""
    xyz, params = room_small
    xyz = xyz[:50_003]
    cl = ctx.upload(xyz)
    base = O.planes_from_cuboid(params)
    offs = np.array([0, 1, 3, 6, 1001, 1001, 20_002, 50_003])
    ctx.set_mode(7, ps_mode)
    try:
        for K in (1, 2, 4, 5, 6, 8):
            pl = np.concatenate([base, base[:2] + np.float32(0.25)])[:K]
            planes = np.stack([pl] * (len(offs) - 1))
            out_g = ctx.plane_sums(cl, offs, planes, K)
            out_o = O.plane_sums(xyz, offs, planes, K)
            assert np.array_equal(out_g[..., 0], out_o[..., 0]), K
            assert np.array_equal(out_g[..., 9], out_o[..., 9]), K
            mag = np.abs(out_o).max()
            assert np.allclose(out_g, out_o, rtol=1e-6, atol=1e-6 * mag), K
    finally:
        ctx.set_mode(7, 0)


def test_scatter_and_fit_plane(ctx):
    rng = np.random.default_rng(2)
    n_true = np.array([0.3, -0.5, 0.81])
    n_true /= np.linalg.norm(n_true)
    basis = np.linalg.svd(n_true[None])[2][1:]
    pts = (rng.uniform(-3, 3, size=(200_000, 2)) @ basis + 1.7 * n_true + rng.normal(0, 0.004, size=(200_000, 1)) * n_true).astype(np.float32)
    cl = ctx.upload(pts)
    mean_g, sc_g = ctx.scatter3x3(cl)
    m_o, sc_o = O.scatter3x3(pts, mean_mode=1)
    assert np.array_equal(mean_g.astype(np.float32), m_o)
    assert np.allclose(sc_g, sc_o, rtol=1e-12, atol=1e-9)
    eq = ctx.fit_plane(cl)
    eq_o = O.fit_plane(pts, mean_mode=1)
    s = 1.0 if np.dot(eq[:3], eq_o[:3]) > 0 else -1.0  # eigenvector sign is arbitrary (LAPACK)
    assert np.allclose(s * eq, eq_o, atol=2e-6)
    assert abs(abs(np.dot(eq[:3], n_true)) - 1) < 1e-5 and abs(abs(eq[3]) - 1.7) < 1e-3
    with pytest.raises(Exception, match="need at least 3"):
        ctx.fit_plane(ctx.upload(pts[:2]))


# ------------------------------------------------------------------ (3) transforms
@pytest.mark.parametrize("n", [100_003, 1, 4, 7])
def test_rigid_transforms_bit_exact(ctx, n):
    rng = np.random.default_rng(n)
    xyz = (rng.normal(size=(n, 3)) * 4).astype(np.float32)
    R = O.rot_matrix3([1, 2, 3], 0.7, np.float32)
    c = np.array([0.5, -1.25, 2.0], np.float32)
    cl = ctx.upload(xyz)
    out = ctx.rotate_around(cl, c, R).download()
    assert np.array_equal(out.view(np.uint32), O.rotate_cloud_around(xyz, c, R).view(np.uint32))
    out = ctx.translate(cl, c).download()
    assert np.array_equal(out.view(np.uint32), O.translate_cloud(xyz, c).view(np.uint32))
    M = O.proj_translate4([6, 0, -3], O.proj_rotate_around(c, R, O.proj_identity()))
    out = ctx.transform(cl, M).download()
    assert np.array_equal(out.view(np.uint32), O.project_cloud(xyz, M).view(np.uint32))


def test_transform_rejects_projective_last_column(ctx):
    M = np.eye(4, dtype=np.float32)
    M[0, 3] = 0.5
    cl = ctx.upload(np.zeros((4, 3), np.float32))
    with pytest.raises(Exception, match="last column"):
        ctx.transform(cl, M)


def test_proj_replay_equivalence_on_gpu(ctx):
    """projTest..projTest5 (Main.hs:2543-2616): replaying the accumulated roomProj on the fresh cloud coincides with
    the incrementally transformed cloud (Float tolerance: the two paths round differently)."""
    rng = np.random.default_rng(4)
    xyz = (rng.uniform(0, 5, size=(20_000, 3))).astype(np.float32)
    cl = ctx.upload(xyz)
    Rx90 = O.rot_matrix3([1, 0, 0], math.radians(90), np.float32)
    mean, _ = ctx.mean_extent(ctx.translate(cl, [6, 0, 0]))
    m = mean.astype(np.float32)
    inc = ctx.rotate_around(ctx.translate(cl, [6, 0, 0]), m, Rx90).download()
    from housescan_b200 import rooms

    proj = rooms.projRotateAround(rooms.projTranslate(np.eye(4, dtype=np.float32), [6, 0, 0]), m, Rx90)  # the product's own roomProj algebra
    assert np.array_equal(proj.view(np.uint32), O.proj_rotate_around(m, Rx90, O.proj_translate4([6, 0, 0], O.proj_identity())).view(np.uint32))
    rep = ctx.transform(cl, proj).download()
    assert np.allclose(inc, rep, atol=5e-5)


def test_mean_extent(ctx):
    rng = np.random.default_rng(6)
    xyz = (rng.uniform(1, 6, size=(300_001, 3))).astype(np.float32)
    mean, md = ctx.mean_extent(ctx.upload(xyz))
    mo = O.point_mean_f64(xyz)
    assert np.allclose(mean, mo, rtol=1e-13)
    assert md == np.float32(O.max_distance(xyz, mo.astype(np.float32)))
    # the reference's sequential Float fold drifts from this by ~1e-5 at this size (SURVEY §7 hard part 2)
    assert np.allclose(O.point_mean_f32seq(xyz), mean, rtol=1e-3)
    with pytest.raises(Exception, match="pointMean: empty"):
        ctx.mean_extent(ctx.upload(np.zeros((0, 3), np.float32)))


def test_write_ply_roundtrip(ctx, tmp_path):
    rng = np.random.default_rng(7)
    xyz = rng.normal(size=(1001, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(1001, 3), dtype=np.uint8)
    for colors in (None, rgb):
        path = str(tmp_path / "room.ply")
        ctx.write_ply(ctx.upload(xyz), path, colors)
        raw = open(path, "rb").read()
        head, body = raw.split(b"end_header\n", 1)
        assert b"format binary_little_endian 1.0" in head and b"element vertex 1001" in head
        if colors is None:
            assert np.array_equal(np.frombuffer(body, np.float32).reshape(-1, 3), xyz)
        else:
            rec = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "u1", 3)]))
            assert np.array_equal(rec["p"], xyz) and np.array_equal(rec["c"], rgb)


# ------------------------------------------------------------------ (4) connected components
def test_cc_labels_voxel_building_bit_exact(ctx):
    src, dst, n, _ = synth.voxel_building_graph()
    lab_g = ctx.cc_label(src, dst, n)
    lab_o = O.cc_label(src, dst, n)
    assert np.array_equal(lab_g, lab_o)
    assert np.all(lab_g <= np.arange(n)) and np.array_equal(lab_g[lab_g], lab_g)  # canonical min-index, idempotent


def test_cc_labels_random_graphs_and_edge_cases(ctx):
    rng = np.random.default_rng(8)
    for n, e in [(1, 0), (10, 0), (2, 1), (1000, 300), (1000, 5000), (200_000, 150_000)]:
        src = rng.integers(0, n, size=e).astype(np.uint32)
        dst = rng.integers(0, n, size=e).astype(np.uint32)
        assert np.array_equal(ctx.cc_label(src, dst, n), O.cc_label(src, dst, n))
    # one long chain (worst case for pointer jumping), self loops, duplicate edges
    n = 100_000
    src = np.arange(n - 1, dtype=np.uint32)[::-1].copy()
    dst = src + 1
    assert np.all(ctx.cc_label(src, dst, n) == 0)
    src = np.array([3, 3, 5, 5, 7], np.uint32)
    dst = np.array([3, 4, 4, 4, 7], np.uint32)
    assert np.array_equal(ctx.cc_label(src, dst, 9), O.cc_label(src, dst, 9))
    with pytest.raises(Exception, match="out of range"):
        ctx.cc_label(np.array([9], np.uint32), np.array([0], np.uint32), 9)


def test_group_connected_components_matches_reference_order(ctx):
    from housescan_b200.GroupConnectedComponents import groupConnectedComponents

    rng = np.random.default_rng(12)
    names = [f"room{i}" for i in range(14)]
    for _ in range(20):
        m = int(rng.integers(1, 25))
        edges = [((names[int(rng.integers(14))], names[int(rng.integers(14))]), float(rng.normal())) for _ in range(m)]
        assert groupConnectedComponents(edges, ctx) == O.group_connected_components(edges)
    assert groupConnectedComponents([], ctx) == []


# ------------------------------------------------------------------ VectorUtil / removeCeiling
# mode key 8: 0 = three passes (11/11/10 bits) over compacted keys, 1 = four 8-bit passes over the cloud
@pytest.fixture(params=[0, 1], ids=["keys3", "cloud4"])
def sel_mode(ctx, request):
    ctx.set_mode(8, request.param)
    yield request.param
    ctx.set_mode(8, 0)


def test_kth_small_and_degenerate(ctx, sel_mode):
    rng = np.random.default_rng(113)
    for n in (1, 2, 3, 4, 5, 7, 1023, 4099):
        xyz = rng.normal(size=(n, 3)).astype(np.float32)
        cl = ctx.upload(xyz)
        for k in sorted({1, (n + 1) // 2, n}):
            assert ctx.kth_largest(cl, 0, k) == np.sort(xyz[:, 0])[::-1][k - 1], (n, k)
            assert ctx.kth_smallest(cl, 1, k) == np.sort(xyz[:, 1])[k - 1], (n, k)
    xyz = np.full((10_001, 3), np.float32(-2.5))  # every key equal
    cl = ctx.upload(xyz)
    assert ctx.kth_largest(cl, 2, 5000) == np.float32(-2.5) and ctx.kth_smallest(cl, 2, 1) == np.float32(-2.5)


def test_kth_and_remove_ceiling(ctx, sel_mode, flt_mode):
    rng = np.random.default_rng(13)
    n = 250_007
    xyz = rng.normal(size=(n, 3)).astype(np.float32) * 2
    xyz[rng.integers(0, n, 5000), 1] = np.float32(1.5)  # heavy ties around the cut
    xyz[:10, 1] = [0.0, -0.0, 1e-30, -1e-30, 3e38, -3e38, 1.5, 1.5, 1.5, 1.5]
    cl = ctx.upload(xyz)
    for k in (1, 2, n // 5, n // 2, n - 1, n):
        assert ctx.kth_largest(cl, 1, k) == O.kth_largest(xyz[:, 1], k), k
        assert ctx.kth_smallest(cl, 2, k) == np.sort(xyz[:, 2])[k - 1], k
    with pytest.raises(Exception, match="k must be >= 1"):
        ctx.kth_largest(cl, 1, 0)
    with pytest.raises(Exception, match="length of the vector"):
        ctx.kth_largest(cl, 1, n + 1)
    colors = rng.random((n, 3)).astype(np.float32)
    out, cout, ylim = ctx.remove_ceiling(cl, ctx.upload(colors))
    keep_o, col_o = O.remove_ceiling(xyz, None)[0], colors[xyz[:, 1] <= O.kth_largest(xyz[:, 1], n // 5)]
    assert ylim == O.kth_largest(xyz[:, 1], n // 5)
    assert np.array_equal(out.download().view(np.uint32), keep_o.view(np.uint32))
    assert np.array_equal(cout.download(), col_o)
    assert len(out) >= n - (n // 5 - 1)  # ties at the limit are kept: removes at most k-1 points
    o2, _, _ = ctx.remove_ceiling(ctx.upload(np.zeros((0, 3), np.float32)))
    assert len(o2) == 0
    with pytest.raises(Exception, match="k must be >= 1"):  # n < 5 => k = 0: the reference errors as well
        ctx.remove_ceiling(ctx.upload(xyz[:4]))


# mode key 12: 0 = count pass + scatter pass, 1 = single pass (decoupled look-back)
@pytest.fixture(params=[0, 1], ids=["twopass", "onepass"])
def flt_mode(ctx, request):
    ctx.set_mode(12, request.param)
    yield request.param
    ctx.set_mode(12, 0)


def test_filter_le_large_with_colours(ctx, flt_mode):
    """1.5 M points = 367 claims of 4 tiles: look-back windows several deep; long runs where nothing / everything is kept"""
    rng = np.random.default_rng(15)
    n = 1_500_003
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    xyz[200_000:330_000, 1] = 9.0   # nothing kept
    xyz[700_000:900_000, 1] = -9.0  # everything kept
    col = rng.random((n, 3)).astype(np.float32)
    cl, cc = ctx.upload(xyz), ctx.upload(col)
    for _ in range(2):
        out, cout = ctx.filter_le(cl, 1, 0.25, cc)
        keep = xyz[:, 1] <= np.float32(0.25)
        assert np.array_equal(out.download().view(np.uint32), xyz[keep].view(np.uint32))
        assert np.array_equal(cout.download(), col[keep])


def test_filter_le_order_preserving(ctx, flt_mode):
    rng = np.random.default_rng(14)
    for n in (1, 5, 1024, 1025, 4096, 4097, 70_001):
        xyz = rng.normal(size=(n, 3)).astype(np.float32)
        for axis, lim in ((0, 0.0), (2, -5.0), (1, 5.0)):
            out, _ = ctx.filter_le(ctx.upload(xyz), axis, lim)
            assert np.array_equal(out.download(), O.filter_le(xyz, axis, lim)[0])


# ------------------------------------------------------------------ A4 fused per-frame normal equations
def _ne_scale(rec):
    """per-element magnitude scale of a 6x6 record: |sum J_a J_b| <= sqrt(A_aa A_bb) (Cauchy-Schwarz), |sum J_a r| <= sqrt(A_aa sum r^2)"""
    iu = [(a, b) for a in range(6) for b in range(a, 6)]
    diag = np.stack([rec[:, iu.index((a, a))] for a in range(6)], axis=1)
    sc = np.ones_like(rec)
    for t, (a, b) in enumerate(iu):
        sc[:, t] = np.sqrt(diag[:, a] * diag[:, b])
    for a in range(6):
        sc[:, 21 + a] = np.sqrt(diag[:, a] * rec[:, 27])
    sc[:, 27] = rec[:, 27]
    return np.maximum(sc, 1e-300)


# mode key 9: 0 = throughput form (Float chains of 16 pixels, then Double; frames cut into row bands), 1 = all-Double form
@pytest.mark.parametrize("ne_mode", [0, 1], ids=["f32chains", "double"])
@pytest.mark.parametrize("use_intr,use_pose", [(False, False), (True, True), (True, False), (False, True)])
def test_backproject_reduce6x6(ctx, use_intr, use_pose, ne_mode):
    frames, poses = synth.depth_stream(3, 160, 120)
    intr = (synth.KINFU_INTR * 0.25).astype(np.float32) if use_intr else None
    planes = O.planes_from_cuboid(synth.C1_PARAMS) if use_intr else np.stack(
        [O.mk_plane_eq([0, 0, 1], 40.0), O.mk_plane_eq([1, 0, 0], 3.0), O.mk_plane_eq([0, 1, 0], 2.0), O.mk_plane_eq([0.6, 0, 0.8], 60.0)])
    ps = poses if use_pose else None
    ctx.set_mode(9, ne_mode)
    try:
        out_g = ctx.backproject_reduce6x6(frames, 160, 120, planes, intr, ps)
    finally:
        ctx.set_mode(9, 0)
    out_o = O.backproject_reduce6x6(frames, 160, 120, planes, intr, ps)
    assert np.array_equal(out_g[:, 28], out_o[:, 28])
    if ne_mode == 1:
        scale = np.maximum(np.abs(out_o), 1e-9 * np.abs(out_o).max(axis=1, keepdims=True))
        assert _rel(out_g, out_o, scale) < 1e-10
    else:  # the north-star bar is 1e-6; measured ~1e-8
        assert _rel(out_g[:, :28], out_o[:, :28], _ne_scale(out_o)[:, :28]) < 1e-6


@pytest.mark.parametrize("w,h,nf", [(640, 480, 2), (104, 50, 5), (100, 60, 4), (8, 1, 3)])
def test_backproject_reduce6x6_shapes(ctx, w, h, nf):
    """full-size frames (8 row bands per frame), widths that are / are not multiples of 8, a one-group frame; an all-invalid frame"""
    frames, poses = synth.depth_stream(nf, w, h)
    frames = frames.copy()
    frames[-1] = 0  # no valid pixel: an all-zero record
    intr = (synth.KINFU_INTR * (w / 640.0)).astype(np.float32)
    planes = O.planes_from_cuboid(synth.C1_PARAMS)
    out_g = ctx.backproject_reduce6x6(frames, w, h, planes, intr, poses)
    out_o = O.backproject_reduce6x6(frames, w, h, planes, intr, poses)
    assert np.array_equal(out_g[:, 28], out_o[:, 28])
    assert not out_g[-1].any()
    assert _rel(out_g[:, :28], out_o[:, :28], _ne_scale(out_o)[:, :28]) < 1e-6
    out_g2 = ctx.backproject_reduce6x6(frames, w, h, planes, intr, poses)  # counters re-armed: a second launch gives the same bits
    assert np.array_equal(out_g, out_g2)


# ------------------------------------------------------------------ optimiser on the GPU objective
def test_bfgs_on_cloud_recovers_cuboid(ctx):
    true = np.concatenate([[0.4, -0.3, 3.5], [4.0, 2.5, 3.0], synth.quat_from_axis_angle([0.2, 1, 0.1], 12.0)])
    xyz, _ = synth.cuboid_room_cloud(200_000, true, sigma=0.003, seed=21)
    init = true + np.concatenate([[0.05, -0.04, 0.03], [0.08, -0.06, 0.05], 0.02 * np.array([1, -1, 1, -1])])
    cl = ctx.upload(xyz)
    f0, _, _ = ctx.cuboid_residual_grad(cl, init)
    p, f, iters, evals = ctx.fit_cuboid_cloud_bfgs(cl, init, 100, 1e-7)
    assert f < 0.05 * f0 and f < 200_000 * (0.003 ** 2) * 1.3
    assert np.allclose(p[:6], true[:6], atol=2e-3)
    qa, qb = p[6:] / np.linalg.norm(p[6:]), true[6:] / np.linalg.norm(true[6:])
    assert min(np.linalg.norm(qa - qb), np.linalg.norm(qa + qb)) < 2e-3


def test_bfgs_fit_matches_the_same_optimiser_over_the_oracle_objective(ctx):
    """North-star bar: fitted cuboid parameters within 1e-6 relative of the reference's.  The reference side is the identical BFGS
    (`hs_bfgs_minimize`) driven by the oracle's objective / gradient on the CPU; the product side drives the GPU session."""
    from housescan_b200 import FitCuboidBFGS as F

    true = np.concatenate([[0.4, -0.3, 3.5], [4.0, 2.5, 3.0], synth.quat_from_axis_angle([0.2, 1, 0.1], 12.0)])
    xyz, _ = synth.cuboid_room_cloud(120_000, true, sigma=0.003, seed=22)
    init = true + np.concatenate([[0.05, -0.04, 0.03], [0.08, -0.06, 0.05], 0.02 * np.array([1, -1, 1, -1])])
    p_g, f_g, it_g, ev_g = ctx.fit_cuboid_cloud_bfgs(ctx.upload(xyz), init, 100, 1e-7)

    def objective(x):
        f, g, _, _ = O.cuboid_residual_grad(xyz, x)
        return f, g

    p_o, f_o, it_o, ev_o = F.bfgsMinimize(objective, init, 100, 1e-7)
    assert abs(f_g - f_o) <= 1e-6 * f_o, (f_g, f_o)
    assert np.max(np.abs(p_g - p_o) / np.maximum(np.abs(p_o), 1e-2)) < 1e-6, (p_g, p_o, it_g, it_o)


def test_nm_fit_over_the_cloud_matches_the_same_simplex_over_the_oracle_objective(ctx):
    """The reference's optimiser (NMSimplex2) over the cloud objective through a session (`hs_fit_cuboid_cloud_nm`: initial simplex
    and shrink steps posted as batches) against the identical simplex (`hs_nm_minimize`) over the oracle's f on the CPU."""
    from housescan_b200 import FitCuboidBFGS as F

    true = np.concatenate([[0.4, -0.3, 3.5], [4.0, 2.5, 3.0], synth.quat_from_axis_angle([0.2, 1, 0.1], 12.0)])
    xyz, _ = synth.cuboid_room_cloud(60_000, true, sigma=0.003, seed=23)
    init = true + np.concatenate([[0.05, -0.04, 0.03], [0.08, -0.06, 0.05], 0.02 * np.array([1, -1, 1, -1])])
    step = np.array([0.01, 0.01, 0.01, 0.1, 0.1, 0.1, 0.02, 0.02, 0.02, 0.02])
    p_g, f_g, it_g, ev_g = ctx.fit_cuboid_cloud_nm(ctx.upload(xyz), init, step, 1e-7, 400)
    p_o, f_o, it_o, ev_o = F.nmMinimize(lambda x: O.cuboid_residual_grad(xyz, x)[0], init, step, 1e-7, 400)
    f0 = O.cuboid_residual_grad(xyz, init)[0]
    assert f_g < 0.2 * f0 and abs(f_g - f_o) <= 1e-3 * f_o, (f_g, f_o, f0, it_g, it_o)
    assert ev_g >= it_g + 11  # the initial simplex alone is 11 evaluations
    f_check = O.cuboid_residual_grad(xyz, p_g)[0]  # the GPU's end state under the oracle's objective
    assert abs(f_check - f_g) <= 1e-6 * f_g


# ------------------------------------------------------------------ six planes that are NOT a cuboid's antiparallel pairs
def test_six_unpaired_planes_take_the_generic_path(ctx, room_small):
    """K == 6 picks the unrolled kernels; the shared-dot-product shortcut applies only when planes 2j / 2j+1 have exactly negated
    normals.  Break the pairing (one normal perturbed by an ulp, one pair re-ordered) and check every K == 6 kernel again."""
    xyz, params = room_small
    xyz = xyz[:60_001]
    cl = ctx.upload(xyz)
    base = O.planes_from_cuboid(params)
    ulp = base.copy()
    ulp[1, 0] = np.nextafter(ulp[1, 0], np.float32(2.0))
    swapped = base[[0, 2, 1, 3, 4, 5]]
    frames, poses = synth.depth_stream(2, 160, 120)
    intr = (synth.KINFU_INTR * 0.25).astype(np.float32)
    for planes in (ulp, swapped):
        a_g, r_g = ctx.plane_assign(cl, planes)
        a_o, r_o = O.plane_assign(xyz, planes)
        assert np.array_equal(a_g, a_o) and np.array_equal(r_g.view(np.uint32), r_o.view(np.uint32))
        offs = np.array([0, 30_002, len(xyz)])
        out_g = ctx.plane_sums(cl, offs, np.stack([planes] * 2), 6)
        out_o = O.plane_sums(xyz, offs, np.stack([planes] * 2), 6)
        assert np.array_equal(out_g[..., 0], out_o[..., 0]) and np.array_equal(out_g[..., 9], out_o[..., 9])
        assert np.allclose(out_g, out_o, rtol=1e-6, atol=1e-6 * np.abs(out_o).max())
        ne_g = ctx.backproject_reduce6x6(frames, 160, 120, planes, intr, poses)
        ne_o = O.backproject_reduce6x6(frames, 160, 120, planes, intr, poses)
        assert np.array_equal(ne_g[:, 28], ne_o[:, 28])
        assert _rel(ne_g[:, :28], ne_o[:, :28], _ne_scale(ne_o)[:, :28]) < 1e-6


def test_graft_entry_smoke():
    """the driver's smoke(): one small invocation of the hot path on cuda:0 checked against the oracle"""
    import __graft_entry__

    __graft_entry__.smoke()


# ------------------------------------------------------------------ config 3 end to end: wall alignment of an apartment
def test_wall_alignment_pipeline_on_gpu(ctx):
    """BASELINE configs[2] in small: per-room, per-wall point reductions on the GPU (hs_plane_sums) feed the reference's
    optimizeRoomPositions (Main.hs:2089-2168: desired offsets -> connected components (GPU) -> lstSqDistances).  The room
    translations must agree with the same pipeline fed by the oracle's sums to 1e-6, and the aligned rooms end `Opposite 0.1`
    apart, i.e. the facing walls of neighbouring rooms sit 0.1 m from each other."""
    from housescan_b200 import FitCuboidBFGS
    from housescan_b200.rooms import X, Z, Opposite, optimizeRoomPositions

    rng = np.random.default_rng(33)
    grid = [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1)]  # an L-shaped flat: X neighbours and Z neighbours
    params = np.zeros((len(grid), 10))
    clouds, offs = [], [0]
    for r, (gx, gz) in enumerate(grid):
        params[r, :3] = np.array([6.0 * gx, 0.0, 6.0 * gz]) + rng.uniform(-0.2, 0.2, size=3)
        params[r, 3:6] = np.array([5.0, 2.6, 4.0]) + rng.uniform(-0.3, 0.3, size=3)
        params[r, 6:] = [1.0, 0.0, 0.0, 0.0]
        pts, _ = synth.cuboid_room_cloud(20_003, params[r], sigma=0.003, rng=rng)
        clouds.append(pts)
        offs.append(offs[-1] + len(pts))
    xyz, offs = np.concatenate(clouds), np.array(offs, np.int64)
    planes = np.stack([O.planes_from_cuboid(params[r]) for r in range(len(grid))])
    sums_g = ctx.plane_sums(ctx.upload(xyz), offs, planes, 6)
    sums_o = O.plane_sums(xyz, offs, planes, 6)
    assert np.array_equal(sums_g[..., 0], sums_o[..., 0])
    # the wall of room r on its +axis / -axis side: makePlanesFromCuboid (Main.hs:1855-1874) gives that wall the normal +/- e_axis
    def wall(r, axis, sign):
        return int(np.argmax(sign * planes[r][:, axis]))
    conns = []
    for a, (ga, za) in enumerate(grid):
        for b, (gb, zb) in enumerate(grid):
            if (gb, zb) == (ga + 1, za):
                conns.append((X, Opposite(0.1), (a, wall(a, 0, +1)), (b, wall(b, 0, -1))))
            if (gb, zb) == (ga, za + 1):
                conns.append((Z, Opposite(0.1), (a, wall(a, 2, +1)), (b, wall(b, 2, -1))))
    conns = conns[::-1]  # connectWalls conses: newest first (Main.hs:2061)
    corner_mean = lambda r: FitCuboidBFGS.cuboidFromParams(params[r]).mean(axis=0)
    ids = list(range(len(grid)))
    res = []
    for sums in (sums_g, sums_o):
        pm = lambda r, w, s=sums: s[r, w, 3:6] / s[r, w, 0]
        moved, log = optimizeRoomPositions(ids, conns, pm, corner_mean, ctx=ctx)
        res.append(np.stack([moved[r] for r in ids]))
    assert np.allclose(res[0], res[1], rtol=1e-6, atol=1e-6)
    assert any("Aligning the X" in l for l in log) and any("Aligning the Z" in l for l in log)
    # after the move the facing walls are 0.1 m apart (to the noise of the synthetic walls)
    pm = lambda r, w: sums_g[r, w, 3:6] / sums_g[r, w, 0] + res[0][r]
    for axis, _, (r1, w1), (r2, w2) in conns:
        gap = abs(float(pm(r2, w2)[axis] - pm(r1, w1)[axis]))
        assert abs(gap - 0.1) < 5e-3, (axis, r1, r2, gap)
    assert np.abs(res[0]).max() < 3.0  # rooms close their 1-2 m gaps; nobody jumps across a neighbour (6 m grid)


def test_sharded_kth_passes_on_gpu(ctx):
    """hs_kth_shard_pass: the three radix passes of the sharded k-th on one rank, and two 'ranks' emulated by two
    clouds whose histograms are added by hand, against the sorted keys; histograms bit-exact against the oracle's"""
    import housescan_b200 as hb
    from housescan_b200 import VectorUtil

    rng = np.random.default_rng(23)
    n = 200_005
    xyz = (rng.normal(size=(n, 3)) * 2).astype(np.float32)
    xyz[rng.integers(0, n, 3000), 1] = np.float32(-0.75)
    xyz[:4, 1] = [0.0, -0.0, 2e38, -2e38]
    cl = ctx.upload(xyz)
    srt = np.sort(xyz[:, 1])
    for k in (1, n // 5, n // 2, n):
        assert VectorUtil.kthLargestBySharded(1, k, cl) == srt[::-1][k - 1] == ctx.kth_largest(cl, 1, k)
        assert VectorUtil.kthSmallestBySharded(1, k, cl) == srt[k - 1]
    h = ctx.kth_shard_pass(cl, 1, 0, 0, 0)
    assert np.array_equal(h, O.kth_shard_hist(xyz[:, 1], 0, 0, 0)) and int(h.sum()) == n
    # two shards, histograms summed by hand
    cut = 80_004
    a = ctx.upload(xyz[:cut])
    ctx_b = hb.Context(0)  # every rank has its own context (and key scratch)
    try:
        b = ctx_b.upload(xyz[cut:])
        fn = lambda p, pre, m: ctx.kth_shard_pass(a, 1, p, pre, m).astype(np.int64) + ctx_b.kth_shard_pass(b, 1, p, pre, m).astype(np.int64)
        for k in (1, 777, n // 5, n):
            assert VectorUtil.kth_sharded(fn, k, True) == srt[::-1][k - 1]
        with pytest.raises(ValueError, match="k must be >= 1"):
            VectorUtil.kth_sharded(fn, 0, True)
    finally:
        b.free()
        ctx_b.close()


def test_remove_ceiling_sharded_single_rank_equals_remove_ceiling(ctx):
    from housescan_b200 import VectorUtil

    rng = np.random.default_rng(29)
    xyz = rng.normal(size=(50_007, 3)).astype(np.float32)
    cl = ctx.upload(xyz)
    kept, _, ylim, first = VectorUtil.removeCeilingSharded(cl)
    ref, _, ylim_ref = ctx.remove_ceiling(cl)
    assert first == 0 and ylim == ylim_ref and np.array_equal(kept.download().view(np.uint32), ref.download().view(np.uint32))


def test_eval_random_room_layouts(ctx):
    """throughput kernel against the exact-products kernel over random layouts (empty rooms between rooms, rooms smaller than a
    group, unaligned offsets, clouds from one group to several tiles per block), twice per layout (tickets must come back to zero),
    and the resident session form bit-identical to the launches"""
    rng = np.random.default_rng(2024)
    for trial in range(40):
        nrooms = int(rng.integers(1, 9))
        kind = trial % 4
        hi = (30, 3_000, 60_000, 700_000)[kind]
        sizes = rng.integers(0, hi, size=nrooms)
        sizes[rng.random(nrooms) < 0.25] = 0
        lead, trail = int(rng.integers(0, 7)), int(rng.integers(0, 7))
        n = lead + int(sizes.sum()) + trail
        if n == 0:
            continue
        params = np.stack([np.concatenate([rng.normal(size=3), rng.uniform(1, 5, 3), rng.normal(size=4)]) for _ in sizes])
        xyz = (rng.normal(size=(n, 3)) * 3).astype(np.float32)
        offs = np.concatenate([[0], np.cumsum(sizes)]) + lead
        cloud = ctx.upload(xyz)
        ctx.set_mode(0, 1)
        exact = ctx.rooms_cuboid_sums(cloud, offs, params)
        ctx.set_mode(0, 0)
        for rep in range(2):
            fast = ctx.rooms_cuboid_sums(cloud, offs, params)
            assert np.array_equal(fast[:, 16:24], exact[:, 16:24]), (trial, rep, sizes.tolist())
            scale = np.abs(exact[:, :16]).max(axis=1, keepdims=True) + 1e-30
            assert np.max(np.abs(fast[:, :16] - exact[:, :16]) / scale) < 2e-5, (trial, rep)  # vs its own magnitude: cancellation-free bound is looser
        with ctx.eval_session(cloud, offs) as sess:
            last = sess.post(np.stack([params, params]))
            assert np.array_equal(sess.wait(last), fast) and np.array_equal(sess.wait(last - 1), fast), (trial, sizes.tolist())


def test_eval_side_threshold_is_the_reference_comparison(ctx):
    """Side of a wall pair (Main.hs:1371-1372 under minimumBy: strict `<`, ties keep the + wall): points packed within a few
    ulps of the mid-plane of the pair that wins the assignment, walls also the wrong way round (negative dimension): the per-wall
    counts must be the reference's, bit for bit."""
    rng = np.random.default_rng(77)
    for trial in range(15):
        j = trial % 3
        dims = np.array([5.0, 6.0, 7.0])
        dims[j] = rng.uniform(0.5, 1.5)  # axis j is the nearest pair for points around the centre
        if trial >= 12:
            dims[j] = -dims[j]  # an optimiser step may cross zero: the + wall then lies on the - side
        center = rng.normal(size=3) * (0.0 if trial < 3 else 2.0)
        quat = np.array([1.0, 0, 0, 0]) if trial < 6 else rng.normal(size=4)
        params = np.concatenate([center, dims, quat])
        R = synth.rot_rows_from_quat(params[6:])  # rows = wall normals
        n = 400_000
        u = rng.uniform(-0.2, 0.2, size=(n, 3))
        u[:, j] = rng.uniform(-3e-6, 3e-6, size=n)  # mid-plane of pair j (in room coordinates) +- a dozen ulps
        u[: n // 8, j] = 0.0
        xyz = (u @ R + center).astype(np.float32)
        rec = ctx.rooms_cuboid_sums(ctx.upload(xyz), np.array([0, n]), params[None])
        ro = O.cuboid_sums(xyz, params)
        assert np.array_equal(rec[0, 16:22], ro[16:22]), (trial, rec[0, 16:22], ro[16:22])
        assert rec[0, 16 + 2 * j] > 1000 and rec[0, 17 + 2 * j] > 1000  # both sides populated: the threshold was exercised
        assert abs(rec[0, 0] - ro[0]) <= 1e-6 * ro[0]


def test_sharded_transform_export_equals_single_export(ctx, tmp_path):
    """SURVEY.md 8e row 3 on the GPU: three point-range shards of a room are transformed (`hs_transform`) and written with
    `hs_write_ply_part` in reverse order into ONE file; byte-identical to `hs_write_ply` of the whole transformed cloud, and the
    transformed points are the oracle's bit for bit."""
    import housescan_b200 as hb
    from housescan_b200.rooms import shard_range

    rng = np.random.default_rng(12)
    n = 2_500_003  # several pinned chunks per part
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = synth.rot_rows_from_quat(synth.quat_from_axis_angle([0.3, 1.0, -0.2], 40.0)).astype(np.float32)
    m[3, :3] = [1.5, -2.0, 0.25]
    whole = ctx.transform(ctx.upload(xyz), m)
    assert np.array_equal(whole.download().view(np.uint32), O.project_cloud(xyz, m).view(np.uint32))
    for tag, colors in (("plain", None), ("rgb", rgb)):
        single, parts = str(tmp_path / f"single_{tag}.ply"), str(tmp_path / f"parts_{tag}.ply")
        ctx.write_ply(whole, single, colors)
        hb.write_ply_begin(parts, n, colors is not None)
        for rank in (2, 0, 1):
            lo, hi = shard_range(n, rank, 3)
            shard = ctx.transform(ctx.upload(xyz[lo:hi]), m)
            ctx.write_ply_part(shard, parts, lo, n, None if colors is None else colors[lo:hi])
        assert open(single, "rb").read() == open(parts, "rb").read()
