"""N > 1 path on CPU: two gloo ranks shard the apartment by point range, reduce one record per room each and all-reduce
them; the summed records and the gradients must equal the unsharded evaluation.  On the CPU box the per-shard reduction is
the oracle standing in for the kernel (tests may use it as the checker); the sharding helpers, the record algebra, the
all-reduce and the host chain rule are the product's own code, i.e. exactly what runs around the kernel under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import housescan_b200 as hb
    import oracle as O
    from housescan_b200 import synth
    from housescan_b200.rooms import local_room_offsets, shard_range

    xyz, offs, params = synth.apartment(n_rooms=3, pts_per_room=20_001, seed=3)
    pe = params + 0.01
    n = len(xyz)
    lo, hi = shard_range(n, rank, world)
    loc = local_room_offsets(offs, lo, hi)
    shard = xyz[lo:hi]
    rec = np.stack([O.cuboid_sums(shard[loc[r] : loc[r + 1]], pe[r]) for r in range(3)])
    t = torch.from_numpy(rec.copy())
    dist.all_reduce(t)  # the path's only exchange: nrooms x HS_REC doubles
    total = t.numpy()
    whole = np.stack([O.cuboid_sums(xyz[offs[r] : offs[r + 1]], pe[r]) for r in range(3)])
    ok = np.array_equal(total[:, 16:22], whole[:, 16:22]) and np.allclose(total, whole, rtol=1e-12, atol=1e-9)
    for r in range(3):
        f, g, c = hb.cuboid_grad_from_sums(pe[r], total[r])
        f_o, g_o, c_o, gs = O.cuboid_residual_grad(xyz[offs[r] : offs[r + 1]], pe[r])
        ok = ok and np.array_equal(c, c_o) and abs(f - f_o) <= 1e-12 * f_o and np.max(np.abs(g - g_o) / gs) < 1e-11
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_point_range_sharding_equals_unsharded(tmp_path, built_lib, oracle_lib):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(world)] == ["1", "1"]


def _cc_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from housescan_b200 import synth
    from housescan_b200.GroupConnectedComponents import cc_label_sharded

    src, dst, n, _ = synth.voxel_building_graph(nx=32, ny=12, nz=32, seed=4)
    rng = np.random.default_rng(5)  # plus long-range edges so components straddle the shard boundary several times
    extra = rng.integers(0, n, size=(40, 2))
    src = np.concatenate([src, extra[:, 0]]).astype(np.uint32)
    dst = np.concatenate([dst, extra[:, 1]]).astype(np.uint32)
    labels, (lo, hi) = cc_label_sharded(lambda s, d, m: O.cc_label(s, d, m), src, dst, n, rank, world)
    whole = O.cc_label(src, dst, n)
    ok = np.array_equal(labels, whole[lo:hi])
    open(os.path.join(tmp, f"cc{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_vertex_range_cc_equals_unsharded(tmp_path, built_lib, oracle_lib):
    world = 2
    mp.spawn(_cc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"cc{r}").read() for r in range(world)] == ["1", "1"]


def _frames_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from housescan_b200 import synth
    from housescan_b200.rooms import depth_stream_records_sharded, shard_frames

    nf, w, h = 7, 64, 48  # 7 frames over 2 ranks: ragged ranges (4 + 3)
    frames, poses = synth.depth_stream(nf, w, h)
    planes = O.planes_from_cuboid(synth.C1_PARAMS)
    intr = (synth.KINFU_INTR * 0.1).astype(np.float32)
    fn = lambda fr, ps: O.backproject_reduce6x6(fr, w, h, planes, intr, ps)
    got = depth_stream_records_sharded(fn, frames, poses, rank, world)
    whole = fn(frames, poses)
    ranges = [shard_frames(nf, r, world) for r in range(world)]
    ok = np.array_equal(got, whole) and ranges == [(0, 4), (4, 7)]
    ok = ok and [shard_frames(10_000, r, 8) for r in (0, 7)] == [(0, 1250), (8750, 10_000)]  # BASELINE configs[4]
    ok = ok and sum(b - a for a, b in (shard_frames(5, r, 8) for r in range(8))) == 5  # more ranks than frames: empty ranges
    open(os.path.join(tmp, f"fr{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_frame_range_stream_equals_unsharded(tmp_path, built_lib, oracle_lib):
    world = 2
    mp.spawn(_frames_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"fr{r}").read() for r in range(world)] == ["1", "1"]


def _kth_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from housescan_b200.rooms import shard_range
    from housescan_b200.VectorUtil import kth_sharded

    rng = np.random.default_rng(17)
    n = 100_003
    y = (rng.normal(size=n) * 3).astype(np.float32)
    y[rng.integers(0, n, 4000)] = np.float32(1.25)  # heavy ties
    y[:6] = [0.0, -0.0, 1e-30, -1e-30, 3e38, -3e38]
    lo, hi = shard_range(n, rank, world)
    mine = y[lo:hi]
    fn = lambda p, pre, m: O.kth_shard_hist(mine, p, pre, m)  # the oracle stands in for Context.kth_shard_pass on the CPU box
    srt = np.sort(y)
    ok = True
    for k in (1, 2, n // 5, n // 2, n - 1, n):
        ok = ok and kth_sharded(fn, k, True) == srt[::-1][k - 1] and kth_sharded(fn, k, False) == srt[k - 1]
    for bad in (0, n + 1):
        try:
            kth_sharded(fn, bad, True)
            ok = False
        except ValueError:
            pass
    # removeCeiling over the two shards: global k = n `quot` 5, per-rank filter, offsets of the kept runs
    from housescan_b200.VectorUtil import remove_ceiling_sharded
    xyz = np.stack([np.zeros(n, np.float32), y, np.arange(n, dtype=np.float32)], axis=1)
    kept, ylim, first = remove_ceiling_sharded(hi - lo, fn, lambda lim: xyz[lo:hi][xyz[lo:hi, 1] <= np.float32(lim)])
    whole, _ = O.remove_ceiling(xyz, None)
    ok = ok and ylim == O.kth_largest(y, n // 5) and np.array_equal(kept, whole[first:first + len(kept)])
    ok = ok and (first == 0) == (rank == 0)
    open(os.path.join(tmp, f"kth{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_kth_equals_unsharded(tmp_path, built_lib, oracle_lib):
    world = 2
    mp.spawn(_kth_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"kth{r}").read() for r in range(world)] == ["1", "1"]


def _export_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import housescan_b200 as hb
    import oracle as O
    from housescan_b200.rooms import export_room_ply_sharded

    rng = np.random.default_rng(8)
    n = 100_003
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 1]], np.float32)
    m[3, :3] = [1.5, -2.0, 0.25]
    for tag, colors in (("plain", None), ("rgb", rgb)):
        path = os.path.join(tmp, f"room_{tag}.ply")
        export_room_ply_sharded(lambda lo, hi: (O.project_cloud(xyz[lo:hi], m), lo, hi),
                                lambda part, lo: hb.write_ply_part_host(path, part[0], lo, n, None if colors is None else colors[part[1]:part[2]]),
                                path, n, rank, world, has_rgb=colors is not None)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_transform_export_is_byte_identical(tmp_path, built_lib, oracle_lib):
    """two ranks transform their point ranges and write their parts of ONE .ply; the file equals the single-rank
    export byte for byte.  On the CPU box the per-shard transform is the oracle standing in for `hs_transform`; header, offsets,
    barrier protocol and the body writer are the product's."""
    import housescan_b200 as hb
    import oracle as O

    world = 2
    mp.spawn(_export_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(8)
    n = 100_003
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 1]], np.float32)
    m[3, :3] = [1.5, -2.0, 0.25]
    whole = O.project_cloud(xyz, m)
    for tag, colors in (("plain", None), ("rgb", rgb)):
        single = str(tmp_path / f"single_{tag}.ply")
        hb.write_ply_begin(single, n, colors is not None)
        hb.write_ply_part_host(single, whole, 0, n, colors)
        a, b = open(single, "rb").read(), open(tmp_path / f"room_{tag}.ply", "rb").read()
        assert a == b and len(a) > n * (15 if colors is not None else 12)
        assert a.startswith(b"ply\nformat binary_little_endian 1.0\n")
