"""Evaluation sessions on one GPU: records bit-identical to the one-shot launches, time per evaluation queued ahead and one by one.
usage: python tools/prof_session.py [--n POINTS] [--rooms R] [--evals K]   (never a bench number under ncu)"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import housescan_b200 as hb

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000_008)
ap.add_argument("--rooms", type=int, default=12)
ap.add_argument("--evals", type=int, default=200)
ap.add_argument("--seg-cost", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
if a.seg_cost:
    ctx.set_mode(2, a.seg_cost)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
ctx.set_stream(s.cuda_stream)
NR = a.rooms
params = bench.room_params()[:NR]
pe = np.ascontiguousarray(bench.eval_params(bench.room_params())[:NR])
per = a.n // NR
offs = np.arange(NR + 1, dtype=np.int64) * per
buf, pts = bench.gen_points_torch(torch, dev, params, [per] * NR, seed=3)
cloud = ctx.wrap(buf.data_ptr(), per * NR, keepalive=buf)
rec = torch.zeros(NR * hb.HS_REC, dtype=torch.float64, device=dev)
torch.cuda.synchronize()

# one-shot launches, queued
for _ in range(5):
    ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.evals):
    ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
e1.record()
torch.cuda.synchronize()
one_ms = e0.elapsed_time(e1) / a.evals
ref = rec.cpu().numpy().reshape(NR, hb.HS_REC).copy()
pe2 = pe * (1 + 1e-4)
ctx.rooms_cuboid_sums_async(cloud, offs, pe2, rec.data_ptr())
torch.cuda.synchronize()
ref2 = rec.cpu().numpy().reshape(NR, hb.HS_REC).copy()
print(f"one-shot: {one_ms * 1e3:.2f} us/eval  {per * NR * 12 / one_ms / 1e6:.0f} GB/s", flush=True)

# session, queued ahead
K = a.evals
batch = np.stack([pe if (i % 2 == 0) else pe2 for i in range(K)])
e0.record()
t0 = time.perf_counter()
with ctx.eval_session(cloud, offs) as sess:
    t1 = time.perf_counter()
    last = sess.post(batch)
    t2 = time.perf_counter()
    recs = [sess.wait(i) for i in range(K)] if K <= 256 else [sess.wait(last)]
    t3 = time.perf_counter()
    tq = np.array([sess.times(i) for i in range(max(K - 250, 0), K)], dtype=np.int64)
    print(f"session queued, device clock: commit-to-commit median {np.median(np.diff(tq[:, 1])) / 1e3:.2f} us, min {np.min(np.diff(tq[:, 1])) / 1e3:.2f}, max {np.max(np.diff(tq[:, 1])) / 1e3:.2f}", flush=True)
e1.record()
torch.cuda.synchronize()
t4 = time.perf_counter()
sess_ms = e0.elapsed_time(e1) / K
print(f"session queued: {sess_ms * 1e3:.2f} us/eval (events around begin..end)  {per * NR * 12 / sess_ms / 1e6:.0f} GB/s; host: begin {1e3*(t1-t0):.2f} ms, post {1e3*(t2-t1):.2f} ms, wait {1e3*(t3-t2):.2f} ms, end {1e3*(t4-t3):.2f} ms", flush=True)
if K <= 256:
    bad = [i for i in range(K) if not np.array_equal(recs[i], ref if i % 2 == 0 else ref2)]
    print("session records bit-identical to one-shot:", not bad, bad[:5], flush=True)
    if bad:
        i = bad[0]
        d = np.abs(recs[i] - (ref if i % 2 == 0 else ref2))
        print("max abs diff", d.max(), "at", np.unravel_index(d.argmax(), d.shape), recs[i][0, :4], ref[0, :4])

# session, one by one (the optimiser's sequential dependency): latency per evaluation
with ctx.eval_session(cloud, offs) as sess:
    for _ in range(5):
        sess.eval(pe)
    t0 = time.perf_counter()
    for i in range(K):
        r = sess.eval(pe if i % 2 == 0 else pe2)
    t1 = time.perf_counter()
print(f"session one-by-one: {1e6 * (t1 - t0) / K:.2f} us/eval (host clock, post+wait each)", "last ok:", np.array_equal(r, ref if (K - 1) % 2 == 0 else ref2), flush=True)

# the same through the raw C entry point with pre-converted arguments (no numpy / wrapper work per call): what a compiled host pays
import ctypes as C
lib = ctx.lib
with ctx.eval_session(cloud, offs) as sess:
    out = np.empty((NR, hb.HS_REC))
    pa, pb, po = pe.ctypes.data_as(C.c_void_p), pe2.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)
    for _ in range(5):
        lib.hs_eval_session_eval(sess.h, pa, po)
    t0 = time.perf_counter()
    for i in range(K):
        lib.hs_eval_session_eval(sess.h, pa if i % 2 == 0 else pb, po)
    t1 = time.perf_counter()
    tm = np.array([sess.times(5 + i) for i in range(max(K - 250, 0), K)], dtype=np.int64)
print(f"session one-by-one, raw C calls: {1e6 * (t1 - t0) / K:.2f} us/eval; device clock: seen->committed median {np.median(tm[:, 1] - tm[:, 0]) / 1e3:.2f} us, "
      f"committed->next seen median {np.median(tm[1:, 0] - tm[:-1, 1]) / 1e3:.2f} us", flush=True)

# one-shot one by one with a host sync after each (what the round-1 API gives an optimiser)
t0 = time.perf_counter()
for i in range(min(K, 100)):
    ctx.rooms_cuboid_sums(cloud, offs, pe)
t1 = time.perf_counter()
print(f"one-shot one-by-one (sync API): {1e6 * (t1 - t0) / min(K, 100):.2f} us/eval", flush=True)
ctx.close()
