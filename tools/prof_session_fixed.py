"""Where the fixed cost of a short session goes: W + K posted evaluations, CUDA events around begin..stop (the bench's shape), device
stamps of every evaluation.  usage: python tools/prof_session_fixed.py [--n POINTS] [--k K]"""
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
ap.add_argument("--n", type=int, default=12_500_004)
ap.add_argument("--k", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = hb.Context(0)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
ctx.set_stream(s.cuda_stream)
NR = 12
params = bench.room_params()
pe = np.ascontiguousarray(bench.eval_params(params))
per = a.n // NR
offs = np.arange(NR + 1, dtype=np.int64) * per
buf, pts = bench.gen_points_torch(torch, dev, params, [per] * NR, seed=3)
cloud = ctx.wrap(buf.data_ptr(), per * NR, keepalive=buf)
torch.cuda.synchronize()
batch = np.ascontiguousarray(np.stack([pe] * a.k))
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    sess = ctx.eval_session(cloud, offs)
    t1 = time.perf_counter()
    sess.post(batch)
    t2 = time.perf_counter()
    sess.stop()
    e1.record()
    t3 = time.perf_counter()
    sess.wait(a.k - 1)
    tm = np.array([sess.times(i) for i in range(a.k)], dtype=np.int64)
    torch.cuda.synchronize()
    sess.close()
    ev = e0.elapsed_time(e1) * 1e3
    dev_span = (tm[-1, 1] - tm[0, 0]) / 1e3
    d = np.diff(tm[:, 1]) / 1e3
    print(f"rep {rep}: events {ev:.1f} us ({ev / a.k:.2f}/eval) | device: first seen -> last commit {dev_span:.1f} us; seen[0]->done[0] {(tm[0,1]-tm[0,0])/1e3:.1f}; "
          f"commit gaps us: {' '.join(f'{x:.1f}' for x in d[:6])} ... median {np.median(d):.1f} | unexplained by device span {ev - dev_span:.1f} us | host: begin {1e6*(t1-t0):.0f} post {1e6*(t2-t1):.0f} stop+record {1e6*(t3-t2):.0f} us", flush=True)
ctx.close()
