"""Minimal driver for profiling the evaluation kernel under ncu (never a bench number).
usage: python tools/prof_eval.py [--n POINTS] [--mode M] [--reps R] [--bps B]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import housescan_b200 as hb

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000_008)
ap.add_argument("--mode", type=int, default=-1)
ap.add_argument("--bps", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--cons", type=int, default=0)
ap.add_argument("--var", type=int, default=0)
ap.add_argument("--tpi", type=int, default=1)
ap.add_argument("--time", action="store_true")
ap.add_argument("--psleep", type=int, default=0)
ap.add_argument("--rooms", type=int, default=12)
ap.add_argument("--blocks", action="store_true", help="print the per-block timeline of the last launch")
ap.add_argument("--blocks2", action="store_true", help="per-block timeline of the DEFAULT kernel (its timestamped instantiation, mode key 5)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
if a.mode >= 0:
    ctx.set_mode(0, a.mode)
if a.bps > 0:
    ctx.set_mode(1, a.bps)
if a.cons > 0:
    ctx.set_mode(2, a.cons)
ctx.set_mode(3, a.var)
ctx.set_mode(4, a.tpi)
ctx.set_mode(6, a.psleep)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
ctx.set_stream(s.cuda_stream)
params = bench.room_params()
pe = np.ascontiguousarray(bench.eval_params(params))
NR = a.rooms
per = a.n // NR
offs = np.arange(NR + 1, dtype=np.int64) * per
params, pe = params[:NR], np.ascontiguousarray(pe[:NR])
buf, pts = bench.gen_points_torch(torch, dev, params, [per] * NR, seed=3)
cloud = ctx.wrap(buf.data_ptr(), per * NR, keepalive=buf)
rec = torch.zeros(NR * hb.HS_REC, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(f"psleep={a.psleep} mode={a.mode} bps={a.bps} cons={a.cons} var={a.var} tpi={a.tpi} n={per*NR} rooms={NR} {ms*1e3:.1f} us/launch  {per*NR/ms/1e6:.1f} Gpts/s  {per*NR*12/ms/1e6:.0f} GB/s  frac_of_6553={per*NR*12/ms/1e6/6553.3:.3f}")

if a.blocks:
    import ctypes as C
    from housescan_b200 import _lib
    lib = C.CDLL(_lib.SO_PATH)
    ctx.set_mode(5, 1)
    for _ in range(3):
        ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
    torch.cuda.synchronize()
    nb = 148
    buf = (C.c_uint64 * (12 * nb))()
    lib.hs_dbg_block_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    rc = lib.hs_dbg_block_times(ctx.h, buf, 3 * nb)
    t = np.array(buf[: 4 * nb], dtype=np.int64).reshape(nb, 4)
    ev = np.array(buf[4 * nb :], dtype=np.int64).reshape(nb, 2, 4)
    t0 = t[:, 0].min()
    st, md, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
    print(f"rc={rc} block start us: min {st.min():.1f} max {st.max():.1f} | main-loop end us: min {md.min():.1f} median {np.median(md):.1f} p90 {np.percentile(md,90):.1f} max {md.max():.1f} | end max {en.max():.1f}")
    print("main-loop duration us: min %.1f median %.1f max %.1f" % ((md - st).min(), np.median(md - st), (md - st).max()))
    order = np.argsort(md)
    print("slowest blocks (id, main end us):", [(int(b), round(float(md[b]), 1)) for b in order[-8:]])
    print("last block:", int(np.argmax(t[:, 3])), "its end %.1f" % en[np.argmax(t[:, 3])])
    for b in [35, 36, 37, 72, 73, 110, 12, 90]:
        print("block", b, "start %.1f" % st[b], "seg events (head done, main done, rem done, reduce done):", [[round((x - t0) / 1e3, 1) if x else None for x in ev[b, k]] for k in range(2)])


if a.blocks2:
    import ctypes as C
    from housescan_b200 import _lib
    lib = C.CDLL(_lib.SO_PATH)
    ctx.set_mode(5, 1)
    for _ in range(4):
        ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    print(f"timestamped instantiation: {e0.elapsed_time(e1) / a.reps * 1e3:.1f} us/launch")
    nb = 148
    buf = (C.c_uint64 * (8 * nb))()
    lib.hs_dbg_block_times.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    rc = lib.hs_dbg_block_times(ctx.h, buf, 2 * nb)
    t = np.array(buf[: 8 * nb], dtype=np.int64).reshape(nb, 8)
    live = t[:, 0] > 0
    t0 = t[live, 0].min()
    us = lambda c: (t[live, c] - t0) / 1e3
    st, first, main, tick = us(0), us(1), us(2), us(3)
    last = int(np.flatnonzero(live)[np.argmax(t[live, 4])])
    print(f"rc={rc} blocks {int(live.sum())} | start: max {st.max():.1f} | first tile landed after start: median {np.median(first - st):.1f} max {(first - st).max():.1f}"
          f" | streaming end: min {main.min():.1f} median {np.median(main):.1f} p90 {np.percentile(main, 90):.1f} max {main.max():.1f}"
          f" | ticket taken: max {tick.max():.1f}")
    print(f"last block {last}: streaming end {(t[last, 2] - t0) / 1e3:.1f}, ticket {(t[last, 3] - t0) / 1e3:.1f}, final reduction written {(t[last, 5] - t0) / 1e3:.1f}, kernel end {(t[last, 6] - t0) / 1e3:.1f}")
    order = np.flatnonzero(live)[np.argsort(main)]
    print("latest streaming ends (block, us):", [(int(b), round(float((t[b, 2] - t0) / 1e3), 1)) for b in order[-10:]])
