"""Room input formats: planes.txt, PCD clouds, makeInwardFacing, loadRoom.
CPU part: the host-side parsers of the C ABI against the oracle restatement.  GPU part (-m gpu): PCD -> device cloud through
hs_cloud_from_pcd / hs_load_room, bit-exact against the oracle's loader for every DATA kind."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle as O
from pcd_util import write_pcd

import housescan_b200 as hb
from housescan_b200 import RoomIO

PLANE_TEXTS = [
    "1 0 0 -2\n0 2 0 4.0\r\n0.1e1 1 1 3\n",                     # LF and CRLF, trailing newline
    "0.5 0.5 0.70710678 1.25",                                   # no trailing newline
    "-0.0123 9.99e-1 4E-2 -17.25\n3 4 0 +5\n",                   # exponents, explicit plus
    "1 0 0 -2 \n0 1 0 3\n",                                      # blank after d: endOfLine fails, parsing stops after one plane
    "1 0 0 2\n 0 1 0 3\n",                                       # next line starts with a blank: stops after one plane
    "1\n0\n0\n5\n0 0 1 7",                                       # skipSpace between a b c d spans newlines
    "1 0 0 2\n\n0 1 0 3\n",                                      # empty line: stops
    "1. 0 0 2\n",                                                # attoparsec >= 0.11: the dot is consumed, `1.` is 1 -> one plane
    "1.e1 0. 0 2.\n3 4 5 6\n",                                   # dot without digits before an exponent / at the end of a line
    ".5 0 0 2\n",                                                # needs a leading digit: error
    "",                                                          # error
    "2 0 0 1e400\n",                                             # overflow to inf
    "0 0 3 -0\n0 0 3 0\n",
    "1 2 3 4\n5 6 7 x\n",                                        # second line broken: one plane
]


@pytest.mark.parametrize("text", PLANE_TEXTS)
def test_plane_eqs_from_text_matches_oracle(built_lib, text):
    try:
        exp = O.plane_eqs_from_text(text)
    except ValueError:
        with pytest.raises(hb.HsError, match="Could not load planes"):
            RoomIO.planeEqsFromText(text)
        return
    got = RoomIO.planeEqsFromText(text)
    assert got.shape == exp.shape
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_plane_eq_is_pcl_convention(built_lib):
    """PCL writes ax + by + cz + d = 0; the reference wants n.x = d (Main.hs:1383-1386): d changes sign, both get normalised"""
    pl = RoomIO.planeEqsFromText("0 0 2 -6\n")
    assert np.allclose(pl, [[0, 0, 1, 3]])
    assert abs(float(np.dot(pl[0, :3], [5, 5, 3]) - pl[0, 3])) < 1e-6  # the point (5, 5, 3) lies on 2z - 6 = 0


def test_make_inward_facing_matches_oracle(built_lib):
    rng = np.random.default_rng(3)
    planes = np.stack([O.mk_plane_eq(rng.normal(size=3), rng.normal()) for _ in range(64)])
    means = rng.normal(size=(64, 3)).astype(np.float32) * 3
    center = rng.normal(size=3).astype(np.float32)
    means[5] = center  # inward vector zero: not > 0, so the plane flips
    got = RoomIO.makeInwardFacing(center, means, planes)
    exp = O.make_inward_facing(center, means, planes)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert (np.einsum("kc,kc->k", center[None] - means, got[:, :3]) >= 0).sum() >= 62
    assert not np.array_equal(got, planes)


def test_pcd_info_and_header_errors(built_lib, tmp_path):
    xyz = np.arange(30, dtype=np.float32).reshape(10, 3)
    for kind in ("ascii", "binary", "binary_compressed"):
        p = str(tmp_path / f"{kind}.pcd")
        write_pcd(p, xyz, rgb=np.zeros((10, 3), int) if kind != "ascii" else None, kind=kind)
        assert RoomIO.pcdInfo(p) == (10, kind != "ascii", kind)
    bad = tmp_path / "bad.pcd"
    bad.write_bytes(b"VERSION 0.7\nFIELDS a b c\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n" + b"\0" * 12)
    with pytest.raises(hb.HsError):
        RoomIO.pcdInfo(str(bad))  # no x y z
    with pytest.raises(hb.HsError):
        RoomIO.pcdInfo(str(tmp_path / "missing.pcd"))


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx():
    c = hb.Context(0)
    yield c
    c.close()


def test_pcd_width_height_overflow_is_rejected(built_lib, tmp_path):
    """WIDTH x HEIGHT that overflows int64 (or merely exceeds the file) is an error, not a 0-point cloud"""
    bad = tmp_path / "overflow.pcd"
    bad.write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4611686018427387904\nHEIGHT 4\nDATA binary\n" + b"\0" * 48)
    with pytest.raises(hb.HsError):
        RoomIO.pcdInfo(str(bad))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ascii", "binary", "binary_compressed"])
@pytest.mark.parametrize("rgb_type", [None, "F", "U"])
def test_cloud_from_pcd_bit_exact(ctx, tmp_path, kind, rgb_type):
    rng = np.random.default_rng(21)
    for n in (1, 255, 256, 257, 5000):
        xyz = (rng.normal(size=(n, 3)) * 4).astype(np.float32)
        xyz[n // 3: n // 2] = xyz[0]  # repeated records: LZF back references
        rgb = rng.integers(0, 256, (n, 3)) if rgb_type else None
        nrm = rng.normal(size=(n, 4)) if rgb_type else None  # x y z rgb normal_x normal_y normal_z curvature, as KinFu exports
        p = str(tmp_path / f"c_{n}.pcd")
        write_pcd(p, xyz, rgb, nrm, kind=kind, rgb_type=rgb_type or "F")
        cl, col = RoomIO.cloudFromFile(ctx, p)
        xo, co = O.pcd_load(p)
        assert np.array_equal(cl.download().view(np.uint32), xo.view(np.uint32)) and np.array_equal(xo, xyz)
        if rgb_type:
            assert np.array_equal(col.download().view(np.uint32), co.view(np.uint32))
        else:
            assert col is None


@pytest.mark.gpu
def test_cloud_from_pcd_unaligned_fields_and_errors(ctx, tmp_path):
    rng = np.random.default_rng(22)
    xyz = rng.normal(size=(1001, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (1001, 3))
    for kind in ("binary", "binary_compressed"):
        p = str(tmp_path / f"u_{kind}.pcd")
        write_pcd(p, xyz, rgb, kind=kind, extra_front=True)  # a 1-byte field first: 17-byte records, nothing 4-byte aligned
        cl, col = RoomIO.cloudFromFile(ctx, p)
        assert np.array_equal(cl.download(), xyz)
        assert np.array_equal(col.download(), O.pcd_load(p)[1])
    empty = str(tmp_path / "empty.pcd")
    write_pcd(empty, np.zeros((0, 3), np.float32))
    with pytest.raises(hb.HsError, match="contains no points"):  # Main.hs:1344
        RoomIO.cloudFromFile(ctx, empty)
    trunc = tmp_path / "trunc.pcd"
    trunc.write_bytes(open(str(tmp_path / "u_binary.pcd"), "rb").read()[:-100])
    with pytest.raises(hb.HsError, match="ends early"):
        RoomIO.cloudFromFile(ctx, str(trunc))


def _write_room(directory, seed, flip_some=True, n=200_000):
    from housescan_b200 import synth
    rng = np.random.default_rng(seed)
    os.makedirs(directory, exist_ok=True)
    xyz, face = synth.cuboid_room_cloud(n, synth.C1_PARAMS, sigma=0.004, seed=seed)
    write_pcd(os.path.join(directory, "cloud_downsampled.pcd"), xyz, rng.integers(0, 256, (n, 3)), rng.normal(size=(n, 4)), kind="binary")
    planes = O.planes_from_cuboid(synth.C1_PARAMS)  # inward-facing n.x = d
    lines = []
    for k in range(6):
        a, b, c, d = (float(v) for v in planes[k])
        s = -1.0 if (flip_some and k % 2) else 1.0  # PCL's sign is arbitrary: makeInwardFacing has to repair it
        scale = 1.0 + 0.5 * k
        lines.append(f"{s * a * scale!r} {s * b * scale!r} {s * c * scale!r} {-s * d * scale!r}")  # ax + by + cz + d = 0
        hull = xyz[face == k][:40]
        write_pcd(os.path.join(directory, f"cloud_plane_hull{k}.pcd"), hull, kind="ascii")
    with open(os.path.join(directory, "planes.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return xyz


@pytest.mark.gpu
def test_load_room_matches_oracle(ctx, tmp_path):
    d = str(tmp_path / "room0")
    xyz = _write_room(d, 31)
    cl, col, planes = RoomIO.loadRoom(ctx, d)
    xo, co, po = O.load_room(d)
    assert np.array_equal(cl.download().view(np.uint32), xo.view(np.uint32)) and np.array_equal(col.download(), co)
    assert planes.shape == (6, 4) and np.array_equal(planes.view(np.uint32), po.view(np.uint32))
    # every plane faces the room centre again, and the loaded room feeds the hot path: all six walls get their share of points
    c = xyz.astype(np.float64).mean(axis=0)
    assert ((planes[:, :3] @ c - planes[:, 3]) > 0).all()
    a, _ = ctx.plane_assign(cl, planes)
    assert (np.bincount(a, minlength=6) > 1000).all()
    os.remove(os.path.join(d, "cloud_plane_hull3.pcd"))
    with pytest.raises(hb.HsError):
        RoomIO.loadRoom(ctx, d)


# ------------------------------------------------------------------ transform export compatibility
def test_transform_text_round_trips_the_reference_formats(built_lib):
    """roomProjectionToXfFormat / roomProjectionToString (Main.hs:2271-2302) write the transposed matrix with `show`; reading it
    back gives roomProj bit for bit (Haskell's show prints the shortest digits that identify the Float)."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        m = np.eye(4, dtype=np.float32)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        m[:3, :3] = q.astype(np.float32)
        m[3, :3] = (rng.normal(size=3) * 7).astype(np.float32)
        for text in (hb.proj_to_xf(m), hb.proj_to_string(m)):
            back = RoomIO.transformFromText(text)
            assert np.array_equal(back.view(np.uint32), m.view(np.uint32)), text
    # the .xf layout is the LEFT-multiplicative matrix: translation in the last column
    xf = "1.0 0.0 0.0 5.0\n0.0 1.0 0.0 6.0\n0.0 0.0 1.0 7.0\n0.0 0.0 0.0 1.0\n"
    assert np.array_equal(RoomIO.transformFromText(xf)[3], [5, 6, 7, 1])
    with pytest.raises(hb.HsError):
        RoomIO.transformFromText("1 2 3")


def _write_ply_ascii(path, xyz, rgb=None, extra=False):
    with open(path, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment test\nelement vertex %d\n" % len(xyz))
        if extra:
            fh.write("property float confidence\n")
        fh.write("property float x\nproperty float y\nproperty float z\n")
        if rgb is not None:
            fh.write("property uchar red\nproperty uchar green\nproperty uchar blue\n")
        fh.write("element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        for i in range(len(xyz)):
            row = (["0.5"] if extra else []) + [repr(float(v)) for v in xyz[i]] + ([str(int(v)) for v in rgb[i]] if rgb is not None else [])
            fh.write(" ".join(row) + "\n")


@pytest.mark.gpu
def test_ply_reader_and_file_transformer(ctx, tmp_path):
    rng = np.random.default_rng(41)
    n = 20_003
    xyz = (rng.normal(size=(n, 3)) * 3).astype(np.float32)
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    m = np.eye(4, dtype=np.float32)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    m[:3, :3] = q.astype(np.float32)
    m[3, :3] = [1.5, -2.0, 0.25]
    exp = O.project_cloud(xyz, m)  # projectRoom's cloud part (Main.hs:1716-1730)
    src_bin, src_asc = str(tmp_path / "in.ply"), str(tmp_path / "in_ascii.ply")
    ctx.write_ply(ctx.upload(xyz), src_bin, rgb)  # binary_little_endian with uchar colours
    _write_ply_ascii(src_asc, xyz, rgb, extra=True)
    for src in (src_bin, src_asc):
        cl, col = RoomIO.cloudFromPly(ctx, src)
        assert np.array_equal(cl.download().view(np.uint32), xyz.view(np.uint32))
        assert np.array_equal(col.download(), rgb.astype(np.float32) / np.float32(255))
    cl, col = RoomIO.cloudFromPly(ctx, str(tmp_path / "in.ply"))
    xf = tmp_path / "room.xf"
    xf.write_text(hb.proj_to_xf(m))
    want = ctx.transform(ctx.upload(xyz), m).download()
    assert np.array_equal(want.view(np.uint32), exp.view(np.uint32))
    for dst in ("out.ply", "out.pcd"):
        npts = RoomIO.transformCloudFile(ctx, src_bin, str(xf), str(tmp_path / dst))
        assert npts == n
        back, bcol = (RoomIO.cloudFromPly if dst.endswith("ply") else RoomIO.cloudFromFile)(ctx, str(tmp_path / dst))
        assert np.array_equal(back.download().view(np.uint32), want.view(np.uint32))
        assert np.array_equal(np.rint(bcol.download() * 255).astype(np.uint8), rgb)
    # the -matrix string form and a colourless PCD source
    p = str(tmp_path / "plain.pcd")
    write_pcd(p, xyz, kind="binary")
    RoomIO.transformCloudFile(ctx, p, hb.proj_to_string(m), str(tmp_path / "plain_out.pcd"))
    back, bcol = RoomIO.cloudFromFile(ctx, str(tmp_path / "plain_out.pcd"))
    assert bcol is None and np.array_equal(back.download().view(np.uint32), want.view(np.uint32))
    xo, _ = O.pcd_load(str(tmp_path / "plain_out.pcd"))  # the files we write are readable by the independent loader
    assert np.array_equal(xo.view(np.uint32), want.view(np.uint32))


def test_pcd_header_fuzz_never_crashes(built_lib, tmp_path):
    """hostile / damaged headers: the parser answers with a status, whatever the numbers claim (no crash, no giant allocation)"""
    rng = np.random.default_rng(99)
    good = tmp_path / "good.pcd"
    write_pcd(str(good), np.arange(30, dtype=np.float32).reshape(10, 3), rgb=np.zeros((10, 3), int), kind="binary")
    raw = good.read_bytes()
    head_end = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    cases = [raw[:head_end].replace(b"POINTS 10", b"POINTS 99999999999999") + raw[head_end:],
             raw[:head_end].replace(b"COUNT 1 1 1 1", b"COUNT 1 1 1 2000000000") + raw[head_end:],
             raw[:head_end].replace(b"SIZE 4 4 4 4", b"SIZE 4 4 4 3") + raw[head_end:],
             raw[:head_end].replace(b"POINTS 10", b"POINTS -5") + raw[head_end:],
             raw[:head_end].replace(b"DATA binary", b"DATA zip") + raw[head_end:],
             raw[:40], b"", b"\n\n\n", raw[:head_end]]
    for _ in range(200):
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, head_end))] = int(rng.integers(32, 127))
        cases.append(bytes(b))
    ok = bad = 0
    for i, c in enumerate(cases):
        p = tmp_path / f"fz{i}.pcd"
        p.write_bytes(c)
        try:
            n, _, _ = RoomIO.pcdInfo(str(p))
            assert 0 <= n <= len(c)
            ok += 1
        except hb.HsError:
            bad += 1
    assert bad >= 8 and ok + bad == len(cases)


# ------------------------------------------------------------------ committed room fixture (tests/golden/room, make_room_fixture.py)
GOLDEN_ROOM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "room")


def _golden_expected():
    import json

    with open(os.path.join(GOLDEN_ROOM, "expected.json")) as fh:
        return json.load(fh)


def test_golden_room_fixture_pins_the_oracle_and_the_host_parsers(built_lib):
    exp = _golden_expected()
    cloud, cols, planes = O.load_room(GOLDEN_ROOM)
    assert len(cloud) == exp["n"] == 1203
    assert int(cloud.view(np.uint32).astype(np.uint64).sum()) == exp["cloud_u32_sum"]
    assert cloud[:3].view(np.uint32).tolist() == exp["cloud_first"] and cloud[-2:].view(np.uint32).tolist() == exp["cloud_last"]
    assert cols[:3].view(np.uint32).tolist() == exp["colors_first"]
    assert int(cols.view(np.uint32).astype(np.uint64).sum()) == exp["colors_u32_sum"]
    assert planes.view(np.uint32).tolist() == exp["planes_inward"]
    # the C ABI's host-side parsers on the same files
    raw = RoomIO.planeEqsFromFile(os.path.join(GOLDEN_ROOM, "planes.txt"))
    assert raw.view(np.uint32).tolist() == exp["planes_raw"]
    assert RoomIO.pcdInfo(os.path.join(GOLDEN_ROOM, "cloud_downsampled.pcd")) == (1203, True, "binary_compressed")
    assert RoomIO.pcdInfo(os.path.join(GOLDEN_ROOM, "cloud_plane_hull3.pcd")) == (12, False, "ascii")
    # three of the six planes were written with PCL's sign flipped: makeInwardFacing turns exactly those around
    flipped = [k for k in range(6) if np.array_equal(np.array(exp["planes_raw"][k], np.uint32).view(np.float32), -np.array(exp["planes_inward"][k], np.uint32).view(np.float32))]
    assert len(flipped) == 3


@pytest.mark.gpu
def test_golden_room_fixture_through_the_gpu_path(ctx):
    exp = _golden_expected()
    cl, col, planes = RoomIO.loadRoom(ctx, GOLDEN_ROOM)
    xyz, c = cl.download(), col.download()
    assert len(cl) == exp["n"] and int(xyz.view(np.uint32).astype(np.uint64).sum()) == exp["cloud_u32_sum"]
    assert xyz[:3].view(np.uint32).tolist() == exp["cloud_first"] and c[:3].view(np.uint32).tolist() == exp["colors_first"]
    assert int(c.view(np.uint32).astype(np.uint64).sum()) == exp["colors_u32_sum"]
    assert planes.view(np.uint32).tolist() == exp["planes_inward"]


def test_ply_header_parser_and_fuzz(built_lib, tmp_path):
    xyz = np.arange(30, dtype=np.float32).reshape(10, 3)
    p = str(tmp_path / "a.ply")
    _write_ply_ascii(p, xyz, rgb=np.zeros((10, 3), int), extra=True)
    assert RoomIO.plyInfo(p) == (10, True, "ascii")
    bin_hdr = b"ply\nformat binary_little_endian 1.0\nelement vertex 10\nproperty float x\nproperty float y\nproperty float z\nend_header\n"
    q = tmp_path / "b.ply"
    q.write_bytes(bin_hdr + xyz.tobytes())
    assert RoomIO.plyInfo(str(q)) == (10, False, "binary_little_endian")
    rng = np.random.default_rng(5)
    cases = [bin_hdr.replace(b"vertex 10", b"vertex 99999999999999") + xyz.tobytes(), bin_hdr.replace(b"float x", b"double x") + xyz.tobytes(),
             bin_hdr.replace(b"binary_little_endian", b"binary_big_endian"), bin_hdr.replace(b"element vertex 10", b"element face 3"),
             bin_hdr.replace(b"property float z", b"property list uchar int z"), b"ply\n", b"", bin_hdr[:-11]]
    for _ in range(200):
        b = bytearray(bin_hdr + xyz.tobytes())
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(bin_hdr)))] = int(rng.integers(32, 127))
        cases.append(bytes(b))
    ok = bad = 0
    for i, c in enumerate(cases):
        f = tmp_path / f"fz{i}.ply"
        f.write_bytes(c)
        try:
            n, _, _ = RoomIO.plyInfo(str(f))
            assert 0 <= n <= len(c)
            ok += 1
        except hb.HsError:
            bad += 1
    assert bad >= 8 and ok + bad == len(cases)
