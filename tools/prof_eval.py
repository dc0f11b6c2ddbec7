"""Minimal driver for profiling the evaluation kernel (one launch per evaluation) under ncu (never a bench number).
usage: python tools/prof_eval.py [--n POINTS] [--mode M] [--reps R] [--rooms NR] [--seg-cost G]"""
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
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--rooms", type=int, default=12)
ap.add_argument("--seg-cost", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
if a.mode >= 0:
    ctx.set_mode(0, a.mode)
if a.seg_cost:
    ctx.set_mode(2, a.seg_cost)
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
print(f"mode={a.mode} n={per*NR} rooms={NR} {ms*1e3:.1f} us/launch  {per*NR/ms/1e6:.1f} Gpts/s  {per*NR*12/ms/1e6:.0f} GB/s  frac_of_6555={per*NR*12/ms/1e6/6554.9:.3f}")
ctx.close()
