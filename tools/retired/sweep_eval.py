"""A/B of the evaluation kernel's tuning variants (hs_ctx_set_mode key 3) on one box: interleaved rounds, one-launch form.
usage: python tools/sweep_eval.py [--n POINTS] [--reps R] [--rounds K] [--vars 0,1,2]"""
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
ap.add_argument("--reps", type=int, default=100)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--vars", default="0,1,2,3,4,5")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
ctx.set_stream(s.cuda_stream)
params = bench.room_params()
pe = np.ascontiguousarray(bench.eval_params(params))
NR = 12
per = a.n // NR
offs = np.arange(NR + 1, dtype=np.int64) * per
buf, pts = bench.gen_points_torch(torch, dev, params, [per] * NR, seed=3)
cloud = ctx.wrap(buf.data_ptr(), per * NR, keepalive=buf)
rec = torch.zeros(NR * hb.HS_REC, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
variants = [int(v) for v in a.vars.split(",")]
ref = None
for v in variants:
    ctx.set_mode(3, v)
    ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
    torch.cuda.synchronize()
    r = rec.cpu().numpy().reshape(NR, hb.HS_REC).copy()
    if ref is None:
        ref = r
    print(f"var {v}: counts equal {np.array_equal(r[:, 16:22], ref[:, 16:22])}, f max rel {np.max(np.abs(r[:, 0] - ref[:, 0]) / ref[:, 0]):.2e}", flush=True)
# burst state first: idle second, 5 warm-up launches, 20 timed ones (the driver's bench shape), variants interleaved
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
bres = {v: [] for v in variants}
for rnd in range(a.rounds):
    for v in variants:
        ctx.set_mode(3, v)
        time.sleep(1.0)
        for _ in range(5):
            ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
        e0.record()
        for _ in range(20):
            ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        bres[v].append(e0.elapsed_time(e1) / 20 * 1e3)
for v in variants:
    t = np.array(bres[v])
    print(f"burst var {v}: us/launch {' '.join(f'{x:.1f}' for x in t)}  median {np.median(t):.1f}  frac {per * NR * 12 / np.median(t) / 1e3 / 6554.9:.3f}", flush=True)
# heat the GPU so every variant is measured in the sustained state
ctx.set_mode(3, variants[0])
for _ in range(1500):
    ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {v: [] for v in variants}
for rnd in range(a.rounds):
    for v in variants:
        ctx.set_mode(3, v)
        for _ in range(5):
            ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
        e0.record()
        for _ in range(a.reps):
            ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / a.reps * 1e3)
for v in variants:
    t = np.array(res[v])
    print(f"var {v}: us/launch {' '.join(f'{x:.1f}' for x in t)}  median {np.median(t):.1f}  GB/s {per * NR * 12 / np.median(t) / 1e3:.0f}  frac {per * NR * 12 / np.median(t) / 1e3 / 6554.9:.3f}", flush=True)
ctx.close()
