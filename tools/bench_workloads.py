"""The other sharded rows of SURVEY.md section 8e on hardware (bench.py --workload c5_stream | c4_cc | c2_export).

    c5_stream  BASELINE configs[4]: replayed depth stream, 10 000 640x480 frames, frame-sharded; a step = the fused
               back-projection + point-to-plane 6x6 reduction over the rank's frames.  No collective on the data path.
    c4_cc      BASELINE configs[3]: 50 M plane-inlier vertices of a multi-storey building, vertex ranges per rank; a step = the
               local union-find labelling + one exchange of (cut-edge endpoint, local root) pairs + the contracted merge + the
               relabelling of the touched components.
    c2_export  BASELINE configs[1]: 8 M-point room, point ranges per rank; a step = rigid transform + nearest-plane residuals +
               this rank's part of ONE binary .ply.

Same launch contract as bench.py (one process per GPU, barrier + synchronize around W warm-up and K timed steps, CUDA events,
max over ranks, ONE JSON line on rank 0).  These are measurements of rows, not the headline line."""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup(args, backend="nccl"):
    import torch
    import torch.distributed as dist

    import housescan_b200 as hb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend, device_id=dev)
    ctx = hb.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    return torch, dist, hb, ctx, dev, rank, world


def _time_steps(torch, dist, world, dev, step, warmup, steps):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def _emit(rank, world, dist, line):
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------ c5
def c5_stream(args):
    torch, dist, hb, ctx, dev, rank, world = _setup(args)
    from housescan_b200 import synth
    from housescan_b200._lib import ptr
    from housescan_b200.rooms import shard_frames

    w, h, nf = 640, 480, args.frames
    lo, hi = shard_frames(nf, rank, world)
    n_loc = hi - lo
    base, _ = synth.depth_stream(8, w, h)
    reps = -(-n_loc // 8)
    frames = torch.from_numpy(base.astype(np.int32)).to(dev).to(torch.int16).repeat(reps, 1, 1)[:n_loc].contiguous()
    planes = hb.planes_from_cuboid(synth.C1_PARAMS)
    intr = np.array(synth.KINFU_INTR, np.float32)
    rec = torch.empty(max(n_loc, 1) * hb.HS_NE, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step():
        if n_loc:
            ctx._chk(ctx.lib.hs_backproject_reduce6x6_dev(ctx.h, C.c_void_p(frames.data_ptr()), n_loc, w, h, ptr(intr), None, ptr(planes), 6, C.c_void_p(rec.data_ptr())))

    ms = _time_steps(torch, dist, world, dev, step, args.warmup, args.steps)
    per = ms / args.steps
    px = nf * w * h
    achieved = 2.0 * n_loc * w * h / (per * 1e-3) / 1e9
    _emit(rank, world, dist, {
        "metric": "depth_stream_frames_per_sec", "value": nf * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u16 depth / f32 geometry / f64 accumulation",
        "data": "synthetic", "config": {"workload": "replayed OpenNI depth stream (BASELINE configs[4]): fused back-projection + point-to-plane 6x6 reduction", "frames": nf, "width": w, "height": h,
                                        "sharding": f"frame-range x{world}", "collective": "none on the data path"},
        "pixels_per_sec": px * args.steps / (ms * 1e-3), "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": _peak(), "unit": "GB/s", "frac": achieved / _peak(), "traffic": None, "bytes_per_pixel": 2.0, "kernel": "k_reduce6x6_f32"}})


# ------------------------------------------------------------------------------------------------------------------ c4
def c4_cc(args):
    torch, dist, hb, ctx, dev, rank, world = _setup(args)
    from housescan_b200.rooms import shard_range

    P, S = args.storeys, 1000
    N = P * S * S
    lo, hi = shard_range(N, rank, world, align=1)
    # the rank builds the edges whose FIRST endpoint it owns: x- and y-neighbours inside a storey, 3 % dropped (global seeded mask)
    v = torch.arange(lo, hi, device=dev, dtype=torch.int64)
    x, y = v % S, (v // S) % S
    g = torch.Generator(device=dev)
    g.manual_seed(4 + rank)
    parts = []
    for has, nb in ((x < S - 1, v + 1), (y < S - 1, v + S)):
        keep = has & (torch.rand(v.numel(), device=dev, generator=g) >= 0.03)
        parts.append(torch.stack([v[keep], nb[keep]]))
    e = torch.cat(parts, dim=1)
    inside = e[1] < hi
    src = (e[0][inside] - lo).to(torch.int32).contiguous()
    dst = (e[1][inside] - lo).to(torch.int32).contiguous()
    cut = e[:, ~inside].contiguous()  # global ids; the second endpoint belongs to a later rank
    n_loc, E = hi - lo, src.numel()
    lab = torch.empty(n_loc, dtype=torch.int32, device=dev)
    # exchange buffers: every rank's cut edges (u, v) + after labelling the local roots of its own endpoints
    ncut = torch.tensor([cut.shape[1]], device=dev)
    if world > 1:
        allc = [torch.zeros(1, dtype=ncut.dtype, device=dev) for _ in range(world)]
        dist.all_gather(allc, ncut)
        maxcut = int(max(int(c.item()) for c in allc))
    else:
        maxcut = int(ncut.item())
    pad = torch.full((2, max(maxcut, 1)), -1, dtype=torch.int64, device=dev)
    pad[:, : cut.shape[1]] = cut
    torch.cuda.synchronize()
    total_edges = torch.tensor([float(E + cut.shape[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_edges)

    def step():
        ctx._chk(ctx.lib.hs_cc_label_dev(ctx.h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), E, n_loc, C.c_void_p(lab.data_ptr())))
        if world == 1:
            return lab
        # cut-edge exchange: (u, v) of every rank, then root(u) from u's owner and root(v) from v's owner
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        allcut = torch.cat(gathered, dim=1)
        allcut = allcut[:, allcut[0] >= 0]
        roots = torch.full_like(allcut, -1)
        for k in range(2):
            mine = (allcut[k] >= lo) & (allcut[k] < hi)
            roots[k][mine] = lab[(allcut[k][mine] - lo)].long() + lo
        dist.all_reduce(roots, op=dist.ReduceOp.MAX)  # every endpoint has exactly one owner
        # contracted graph over the touched local roots: host union-find keeping the minimum (thousands of edges at most)
        r = roots.cpu().numpy()
        parent = {}

        def find(a):
            while parent.setdefault(a, a) != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        for a, b in zip(r[0].tolist(), r[1].tolist()):
            fa, fb = find(a), find(b)
            if fa != fb:
                parent[max(fa, fb)] = min(fa, fb)
        mine = sorted(a for a in parent if lo <= a < hi and find(a) != a)
        if mine:
            old = torch.tensor(mine, device=dev) - lo
            new = torch.tensor([find(a) for a in mine], device=dev)
            # relabel: every vertex whose local root is a touched root takes the merged root (ids below lo are other ranks' vertices)
            glab = lab.long()
            lut = torch.arange(n_loc, device=dev) + lo
            lut[old] = new
            return lut[glab]
        return lab

    ms = _time_steps(torch, dist, world, dev, step, args.warmup, args.steps)
    per = ms / args.steps
    # labelling traffic floor: every edge read once (8 B) + every label written once (4 B)
    achieved = (8.0 * E + 4.0 * n_loc) / (per * 1e-3) / 1e9
    _emit(rank, world, dist, {
        "metric": "cc_label_vertices_per_sec", "value": N * args.steps / (ms * 1e-3), "unit": "vertices/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 ids", "data": "synthetic",
        "config": {"workload": "GroupConnectedComponents on plane-inlier vertices of a multi-storey building (BASELINE configs[3])", "vertices": N, "edges": int(total_edges.item()),
                   "sharding": f"vertex-range x{world}", "cut_edges_max_per_rank": maxcut,
                   "collective": "none" if world == 1 else "all_gather of cut edges + all_reduce(max) of their endpoints' local roots (NCCL), contracted merge on the host"},
        "gpu_launches": 3 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": _peak(), "unit": "GB/s", "frac": achieved / _peak(), "traffic": None, "bytes": "8 B/edge + 4 B/vertex (compulsory)", "kernel": "k_cc_init / k_cc_link / k_cc_flatten"}})


# ------------------------------------------------------------------------------------------------------------------ c2
def c2_export(args):
    torch, dist, hb, ctx, dev, rank, world = _setup(args)
    import bench
    from housescan_b200 import synth
    from housescan_b200._lib import ptr
    from housescan_b200.rooms import shard_range

    n = args.points
    lo, hi = shard_range(n, rank, world)
    n_loc = hi - lo
    params = bench.room_params(1)
    buf, pts = bench.gen_points_torch(torch, dev, params, [n_loc], seed=50 + rank)
    cloud = ctx.wrap(buf.data_ptr(), n_loc, keepalive=buf)
    out = ctx.alloc(n_loc)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = synth.rot_rows_from_quat(synth.quat_from_axis_angle([0.1, 1.0, 0.05], 33.0)).astype(np.float32)
    m[3, :3] = [6.0, 0.0, -6.0]
    planes = hb.planes_from_cuboid(params[0])
    assign = torch.empty(n_loc + 16, dtype=torch.uint8, device=dev)
    resid = torch.empty(n_loc + 16, dtype=torch.float32, device=dev)
    path = os.path.join(args.out_dir or tempfile.gettempdir(), "hs_bench_room.ply")
    if rank == 0:
        hb.write_ply_begin(path, n, False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    gpu_ms = [0.0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        e0.record()
        ctx._chk(ctx.lib.hs_plane_assign_dev(ctx.h, cloud.h, ptr(planes), 6, C.c_void_p(assign.data_ptr()), C.c_void_p(resid.data_ptr())))
        ctx.transform(cloud, m, out)
        e1.record()
        ctx.write_ply_part(out, path, lo, n)  # D2H in two pinned halves overlapped with pwrite
        gpu_ms[0] += e0.elapsed_time(e1)

    ms = _time_steps(torch, dist, world, dev, step, args.warmup, args.steps)
    per = ms / args.steps
    k_ms = gpu_ms[0] / (args.steps + max(args.warmup, 3))
    achieved = (12.0 + 5.0 + 24.0) * n_loc / (k_ms * 1e-3) / 1e9  # assign: 12 in + 5 out; transform: 12 in + 12 out
    size = os.path.getsize(path) if rank == 0 else 0
    _emit(rank, world, dist, {
        "metric": "room_export_points_per_sec", "value": n * args.steps / (ms * 1e-3), "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full-resolution room export (BASELINE configs[1]): plane residuals + rigid transform + binary .ply write", "points": n, "sharding": f"point-range x{world}",
                   "file": path, "file_bytes": size, "collective": "none on the data path (rank 0 writes the header, every rank its own byte range)"},
        "file_write_gbs": 12.0 * n / (per * 1e-3) / 1e9, "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": _peak(), "unit": "GB/s", "frac": achieved / _peak(), "traffic": None, "kernel_ms": k_ms,
                     "bytes_per_point": 41.0, "kernel": "k_plane_assign<6,1> + k_affine (device part of the step; the rest is PCIe D2H + file write)"}})
    if rank == 0:
        try:
            os.unlink(path)
        except OSError:
            pass


# ------------------------------------------------------------------------------------------------------------------ a12
def a12_ceiling(args):
    """removeCeiling (Main.hs:2643-2664) over a cloud sharded by point range: the k-th largest y over all ranks (three radix passes,
    three all-reduces of 2048 counters) + the order-preserving filter of every rank's own points"""
    torch, dist, hb, ctx, dev, rank, world = _setup(args, backend="cpu:gloo,cuda:nccl")
    import bench
    from housescan_b200 import VectorUtil
    from housescan_b200.rooms import shard_range

    n = args.points
    lo, hi = shard_range(n, rank, world)
    n_loc = hi - lo
    buf, pts = bench.gen_points_torch(torch, dev, bench.room_params(1), [n_loc], seed=70 + rank)
    cloud = ctx.wrap(buf.data_ptr(), n_loc, keepalive=buf)
    torch.cuda.synchronize()
    last = {}

    def step():
        kept, _, y_limit, first = VectorUtil.removeCeilingSharded(cloud)
        last["kept"], last["y"], last["first"] = len(kept), float(y_limit), first
        kept.free()

    ms = _time_steps(torch, dist, world, dev, step, args.warmup, args.steps)
    per = ms / args.steps
    kept_total = torch.tensor([float(last["kept"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kept_total)
    achieved = (24.0 * n_loc + 24.0 * n_loc) / (per * 1e-3) / 1e9  # k-th: 24 B/point physical; filter: 12 in (x2 passes or 1) + 12 out x 0.8
    _emit(rank, world, dist, {
        "metric": "remove_ceiling_points_per_sec", "value": n * args.steps / (ms * 1e-3), "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 keys", "data": "synthetic",
        "config": {"workload": "removeCeiling: k-th largest y (k = n/5) + order-preserving filter", "points": n, "sharding": f"point-range x{world}",
                   "collective": "none" if world == 1 else "3 all-reduces of 2048 counters (gloo, host histograms) + 1 all-gather of kept counts", "y_limit": last["y"],
                   "kept_fraction": float(kept_total.item()) / n},
        "gpu_launches": None,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": _peak(), "unit": "GB/s", "frac": achieved / _peak(), "traffic": None, "bytes_per_point": 48.0,
                     "kernel": "k_sel2_first/next + k_filter_onepass (host round trips between the radix passes included)"}})


WORKLOADS = {"c5_stream": c5_stream, "c4_cc": c4_cc, "c2_export": c2_export, "a12_ceiling": a12_ceiling}
