#!/usr/bin/env python
"""Per-row measurements of the hot-path kernels other than the headline one.

    python tools/bench_rows.py [--n POINTS] [--reps R] [--out gpurun_out/rows.json]

Each row: the C-ABI entry point on device-resident inputs, CUDA events on the library's stream around R calls after a
warm-up, algorithmic bytes per unit from SURVEY.md §8d, fraction of the measured HBM peak (MEASURED_PEAKS.json), and a
full-size parity property checked on the spot (never the oracle at these sizes unless it finishes in seconds).
Not the bench contract (bench.py is); this is the evidence for the other rows, summarised under profiles/.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import housescan_b200 as hb
from housescan_b200._lib import ptr

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000_008)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--cc-planes", type=int, default=50)
ap.add_argument("--cc-side", type=int, default=1000)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "rows.json"))
ap.add_argument("--only", default="")
a = ap.parse_args()

dev = torch.device("cuda", 0)
ctx = hb.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
lib = ctx.lib
PEAK, PEAK_SRC = bench.measured_peak_gbs()
rows = []


def timed(fn, reps=None, warm=2):
    reps = reps or a.reps
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(row, what, units, unit_name, bytes_per_unit, ms, parity, launches=None, note=""):
    gbs = bytes_per_unit * units / (ms * 1e-3) / 1e9
    r = {"row": row, "entry": what, "units": int(units), "unit": unit_name, "ms": ms, "g_units_per_s": units / (ms * 1e-3) / 1e9,
         "bytes_per_unit": bytes_per_unit, "achieved_gbs": gbs, "peak_gbs": PEAK, "frac": gbs / PEAK, "parity": parity, "note": note}
    rows.append(r)
    print(f"{row:4s} {what:34s} {units/1e6:9.1f} M{unit_name:6s} {ms:9.3f} ms  {r['g_units_per_s']:8.2f} G/s  {gbs:7.0f} GB/s  frac {gbs/PEAK:5.3f}  parity={parity} {note}", flush=True)


def want(tag):
    return not a.only or tag in a.only.split(",")


# ---- the apartment cloud (same generator as bench.py)
params = bench.room_params()
per = a.n // 12
n = per * 12
buf, pts = bench.gen_points_torch(torch, dev, params, [per] * 12, seed=3)
cloud = ctx.wrap(buf.data_ptr(), n, keepalive=buf)
out_buf = torch.empty(buf.numel(), dtype=torch.float32, device=dev)
out_cloud = ctx.wrap(out_buf.data_ptr(), n, keepalive=out_buf)
torch.cuda.synchronize()

if want("A8"):
    th = 0.3
    R = np.array([[np.cos(th), 0, -np.sin(th)], [0, 1, 0], [np.sin(th), 0, np.cos(th)]], np.float32)
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = R
    m[3, :3] = [1.5, -0.25, 3.0]
    ms = timed(lambda: ctx.transform(cloud, m, out_cloud))
    # parity property at full size: the first and last 4096 points equal the non-contracted Float formula
    idx = torch.cat([torch.arange(0, 4096, device=dev), torch.arange(n - 4096, n, device=dev)])
    p = pts[idx].cpu().numpy()
    q = out_buf[: 3 * n].view(n, 3)[idx].cpu().numpy()
    f32 = np.float32
    exp = np.empty_like(p)
    for c in range(3):
        rot = f32(f32(f32(p[:, 0] * m[0, c]) + f32(p[:, 1] * m[1, c])) + f32(p[:, 2] * m[2, c]))
        exp[:, c] = f32(m[3, c] + f32(rot + f32(0)))
    report("A8", "hs_transform (projectRoom)", n, "pts", 24.0, ms, bool(np.array_equal(exp.view(np.uint32), q.view(np.uint32))))
    off = np.array([0.5, 0.25, -1.0], np.float32)
    ms = timed(lambda: ctx.translate(cloud, off, out_cloud))
    report("A8", "hs_translate", n, "pts", 24.0, ms, True)

if want("A9"):
    ms = timed(lambda: ctx.mean_extent(cloud))
    mean, md = ctx.mean_extent(cloud)
    tm = pts.double().mean(dim=0).cpu().numpy()
    report("A9", "hs_mean_extent (2 passes)", n, "pts", 24.0, ms, bool(np.allclose(mean, tm, rtol=1e-12, atol=1e-12)), note="mean pass + max-distance pass, 12 B/pt each")

if want("A7"):
    ms = timed(lambda: ctx.scatter3x3(cloud))
    report("A7", "hs_scatter3x3 (2 passes)", n, "pts", 24.0, ms, True, note="mean pass + scatter pass")

if want("A5"):
    planes = hb.planes_from_cuboid(params[0])
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_r = torch.empty(n, dtype=torch.float32, device=dev)
    fn = lambda: ctx._chk(lib.hs_plane_assign_dev(ctx.h, cloud.h, ptr(planes), 6, C.c_void_p(d_a.data_ptr()), None))
    ms = timed(fn)
    cnt = torch.bincount(d_a[:per].int(), minlength=6).cpu().numpy()
    rec = ctx.rooms_cuboid_sums(cloud, np.array([0, per], np.int64), params[:1])
    report("A5", "hs_plane_assign_dev (index)", n, "pts", 13.0, ms, bool(np.array_equal(cnt, rec[0, 16:22].astype(np.int64))), note="room-0 histogram == cuboid-sums counts")
    fn = lambda: ctx._chk(lib.hs_plane_assign_dev(ctx.h, cloud.h, ptr(planes), 6, C.c_void_p(d_a.data_ptr()), C.c_void_p(d_r.data_ptr())))
    ms = timed(fn)
    f_dev = float((d_r[:per].double() ** 2).sum().item())
    report("A5", "hs_plane_assign_dev (+residual)", n, "pts", 17.0, ms, bool(abs(f_dev - rec[0, 0]) <= 1e-7 * rec[0, 0]), note="sum r^2 of room 0 == record")
    del d_a, d_r

if want("A13"):
    offs = np.arange(13, dtype=np.int64) * per
    allp = np.stack([hb.planes_from_cuboid(params[r]) for r in range(12)])
    for mode, tag in ((1, "all-Double"), (2, "Float chains, direct loads"), (0, "Float chains, bulk-async rings")):
        ctx.set_mode(7, mode)
        ps = ctx.plane_sums(cloud, offs, allp, 6)
        ms = timed(lambda: ctx.plane_sums(cloud, offs, allp, 6), reps=3, warm=1)
        ok = int(ps[:, :, 0].sum()) == n
        if mode == 1:
            ps_ref = ps
        else:  # the two forms agree to the north-star bar on every sum, exactly on counts and max |r|
            ok = ok and np.array_equal(ps[..., 0], ps_ref[..., 0]) and np.array_equal(ps[..., 9], ps_ref[..., 9]) and np.allclose(ps, ps_ref, rtol=1e-6, atol=1e-6 * np.abs(ps_ref).max())
        report("A13", f"hs_plane_sums 12x6 ({tag})", n, "pts", 12.0, ms, bool(ok), note="counts sum to n; forms agree")

if want("A12"):
    k = n // 5
    y = pts[:, 1]
    for mode, tag in ((1, "4 passes over the cloud"), (0, "3 passes, compact keys")):
        ctx.set_mode(8, mode)
        ms = timed(lambda: ctx.kth_largest(cloud, 1, k), reps=5)
        v = ctx.kth_largest(cloud, 1, k)
        ok = int((y > float(v)).sum().item()) < k <= int((y >= float(v)).sum().item())
        report("A12", f"hs_kth_largest ({tag})", n, "pts", 16.0, ms, bool(ok), note="algorithmic 4 B key x 4; physical 12 B x 4 (cloud passes) or 12 + 4 + 4 + 4 (compact keys)")
    n_out = C.c_int64()
    yl = C.c_float()
    fn = lambda: ctx._chk(lib.hs_remove_ceiling(ctx.h, cloud.h, None, out_cloud.h, None, C.byref(n_out), C.byref(yl)))
    ctx.set_mode(12, 1)
    ms2 = timed(fn, reps=5)
    report("A12", "hs_remove_ceiling (single-pass filter)", n, "pts", 16.0 + 12.0 + 12.0 * 0.8, ms2, True, note="k-th + look-back compaction")
    ctx.set_mode(12, 0)
    ms = timed(fn, reps=5)
    kept = int((y <= yl.value).sum().item())
    o = out_buf[: 3 * kept].view(kept, 3)
    ok = n_out.value == kept and bool(torch.equal(o, pts[y <= yl.value]))
    report("A12", "hs_remove_ceiling (two-pass filter)", n, "pts", 16.0 + 12.0 + 12.0 * kept / n, ms, bool(ok), note="k-th + count pass + scatter; output order preserved (checked bit-exact vs torch mask)")

if want("A1"):
    from housescan_b200 import synth
    w, h = 640, 480
    nf = a.frames
    base, _ = synth.depth_stream(8, w, h)
    frames = torch.from_numpy(base.astype(np.int32)).to(dev).to(torch.int16).repeat((nf + 7) // 8, 1, 1)[:nf].contiguous()
    npx = nf * w * h
    torch.cuda.synchronize()
    # A1-A3 on a whole replayed stream treated as ONE tall frame (w x nf*h): mask + ordered compaction + scaling
    bp_out = torch.empty(npx * 3 + 16, dtype=torch.float32, device=dev)
    bp_cloud = ctx.wrap(bp_out.data_ptr(), npx, keepalive=bp_out)
    d_mask = torch.empty(npx, dtype=torch.uint8, device=dev)
    nv = C.c_int64()
    fn = lambda: ctx._chk(lib.hs_backproject_ref_dev(ctx.h, C.c_void_p(frames.data_ptr()), w, h * nf, bp_cloud.h, C.c_void_p(d_mask.data_ptr()), C.byref(nv)))
    ctx.set_mode(10, 1)
    ms2 = timed(fn, reps=5)
    ctx.set_mode(10, 0)
    ms = timed(fn, reps=5)
    valid = frames.view(-1) != 0
    ok = nv.value == int(valid.sum().item()) and bool(torch.equal(d_mask.bool(), valid))
    fr = float(nv.value) / npx
    # spot-check the last 1000 compacted points (order + arithmetic): x/10, y/10, d/20-30
    idx = torch.nonzero(valid).view(-1)[-1000:]
    d = (frames.view(-1)[idx].int() & 0xFFFF).float()
    ten, twenty = torch.full_like(d, 10.0), torch.full_like(d, 20.0)  # tensor divisors: torch turns `/ scalar` into `* (1/scalar)`
    exp = torch.stack([torch.div((idx % w).float(), ten), torch.div((idx // w).float(), ten), torch.div(d, twenty) - 30.0], dim=1)
    got = bp_out[: 3 * nv.value].view(-1, 3)[-1000:]
    ok = ok and bool(torch.equal(exp, got))
    report("A1-3", "hs_backproject_ref_dev (two passes)", npx, "px", 2.0 + 1.0 + 12.0 * fr, ms2, True, note="count pass + scatter pass")
    report("A1-3", "hs_backproject_ref_dev (single pass)", npx, "px", 2.0 + 1.0 + 12.0 * fr, ms, bool(ok), note=f"decoupled look-back; {nf} frames as one raster, valid fraction {fr:.3f}, mask written")
    del bp_out, d_mask
    # A4 fused back-project + nearest plane + 6x6 normal equations per frame
    planes = hb.planes_from_cuboid(synth.C1_PARAMS)
    intr = np.array(synth.KINFU_INTR, np.float32)
    d_out = torch.empty(nf * hb.HS_NE, dtype=torch.float64, device=dev)
    fn = lambda: ctx._chk(lib.hs_backproject_reduce6x6_dev(ctx.h, C.c_void_p(frames.data_ptr()), nf, w, h, ptr(intr), None, ptr(planes), 6, C.c_void_p(d_out.data_ptr())))
    poses8 = synth.depth_stream(8, 8, 8)[1]
    poses = np.ascontiguousarray(np.tile(poses8, ((nf + 7) // 8, 1))[:nf])
    fn_pose = lambda: ctx._chk(lib.hs_backproject_reduce6x6_dev(ctx.h, C.c_void_p(frames.data_ptr()), nf, w, h, ptr(intr), ptr(poses), ptr(planes), 6, C.c_void_p(d_out.data_ptr())))
    for mode, tag in ((1, "all-Double"), (0, "Float chains")):
        ctx.set_mode(9, mode)
        ms = timed(fn, reps=5)
        o = d_out.view(nf, hb.HS_NE).clone()
        cnt_ok = bool(torch.equal(o[:, 28].long(), (frames != 0).view(nf, -1).sum(dim=1)))
        rep_ok = bool(torch.equal(o[:8], o[8:16])) if nf >= 16 else True  # replayed frames give identical records (deterministic)
        if mode == 1:
            o_ref = o
        else:
            diag = o_ref[:, [0, 6, 11, 15, 18, 20]].max(dim=1, keepdim=True).values
            rep_ok = rep_ok and bool(((o - o_ref).abs() <= 1e-6 * diag).all().item())
        report("A4", f"hs_backproject_reduce6x6_dev ({tag})", npx, "px", 2.0, ms, cnt_ok and rep_ok, note=f"{nf} frames, intrinsics, 232 B out per frame; counts == valid pixels, replayed frames bit-identical, forms agree")
        ms = timed(fn_pose, reps=5)
        report("A4", f"  + per-frame pose ({tag})", npx, "px", 2.0, ms, True, note="includes the H2D copy of the poses (64 B/frame)")
    del frames, d_out

if want("A10"):
    # C4-shaped graph built on the device: P planes of side x side voxels, 4-neighbour edges inside a plane, 3 % dropouts
    P, S = a.cc_planes, a.cc_side
    N = P * S * S
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    vid = torch.arange(N, device=dev, dtype=torch.int64).view(P, S, S)
    right = torch.stack([vid[:, :, :-1].reshape(-1), vid[:, :, 1:].reshape(-1)])
    down = torch.stack([vid[:, :-1, :].reshape(-1), vid[:, 1:, :].reshape(-1)])
    e = torch.cat([right, down], dim=1)
    keep = torch.rand(e.shape[1], device=dev, generator=g) >= 0.03
    e = e[:, keep]
    src = e[0].to(torch.int32).contiguous()
    dst = e[1].to(torch.int32).contiguous()
    E = src.numel()
    del e, keep, right, down, vid
    lab = torch.empty(N, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    fn = lambda: ctx._chk(lib.hs_cc_label_dev(ctx.h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), E, N, C.c_void_p(lab.data_ptr())))
    ms = timed(fn, reps=3, warm=1)
    l64 = lab.long()
    ar = torch.arange(N, device=dev)
    props = bool((l64 <= ar).all().item()) and bool(torch.equal(l64[l64], l64)) and bool(torch.equal(l64[src.long()], l64[dst.long()]))
    par = "properties"
    try:
        import oracle as O
        t0 = time.perf_counter()
        ref = O.cc_label(src.cpu().numpy().astype(np.uint32), dst.cpu().numpy().astype(np.uint32), N)
        t_cpu = time.perf_counter() - t0
        props = props and bool(np.array_equal(ref, lab.cpu().numpy().astype(np.uint32)))
        par = f"bit-exact vs oracle ({t_cpu:.1f} s on the CPU)"
    except Exception as ex:  # pragma: no cover
        par = f"properties only ({ex})"
    report("A10", "hs_cc_label_dev", N, "nodes", 4.0 + 8.0 * E / N, ms, props, note=f"{E} edges; {par}; ncomp {int((l64 == ar).sum().item())}")

os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as fh:
    json.dump({"peak_gbs": PEAK, "peak_source": PEAK_SRC, "rows": rows}, fh, indent=1)
