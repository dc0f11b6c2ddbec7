"""Per-block timeline of a session's first evaluations (HS_EVAL_TRACE): go, first tile, streaming done, final barrier.
usage: python tools/trace_session.py [--n POINTS] [--k K]"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=12_500_004)
ap.add_argument("--k", type=int, default=24)
a = ap.parse_args()
path = os.path.join(tempfile.gettempdir(), "hs_eval_trace.bin")
os.environ["HS_EVAL_TRACE"] = path
import torch

import bench
import housescan_b200 as hb

dev = torch.device("cuda", 0)
ctx = hb.Context(0)
NR = 12
params = bench.room_params()
pe = np.ascontiguousarray(bench.eval_params(params))
per = a.n // NR
offs = np.arange(NR + 1, dtype=np.int64) * per
buf, pts = bench.gen_points_torch(torch, dev, params, [per] * NR, seed=3)
cloud = ctx.wrap(buf.data_ptr(), per * NR, keepalive=buf)
torch.cuda.synchronize()
for rep in range(2):
    with ctx.eval_session(cloud, offs) as sess:
        last = sess.post(np.stack([pe] * a.k))
        sess.wait(last)
        done = np.array([sess.times(i)[1] for i in range(a.k)], dtype=np.int64)
raw = np.fromfile(path, dtype=np.uint8)
hdr = raw[:16].view(np.int32)
ne, nb = int(hdr[0]), int(hdr[1])
t = raw[16:].view(np.uint64).astype(np.int64).reshape(ne, nb, 4)[: a.k]
live = t[0, :, 0] > 0
t = t[:, live, :]
t0 = t[0, :, 0].min()
us = (t - t0) / 1e3
for e in range(8, min(a.k, 16)):
    go, first, stream, bar = us[e, :, 0], us[e, :, 1], us[e, :, 2], us[e, :, 3]
    prev_bar = us[e - 1, :, 3]
    print(f"eval {e}: go-after-prev-barrier median {np.median(go - prev_bar):.2f} max {np.max(go - prev_bar):.2f} | first tile after go median {np.median(first - go):.2f} max {np.max(first - go):.2f} | "
          f"streaming median {np.median(stream - first):.2f} min {np.min(stream - first):.2f} max {np.max(stream - first):.2f} | barrier after streaming median {np.median(bar - stream):.2f} max {np.max(bar - stream):.2f} | "
          f"block period (bar-to-bar) median {np.median(bar - prev_bar):.2f} min {np.min(bar - prev_bar):.2f} max {np.max(bar - prev_bar):.2f} | spread of barrier times {np.max(bar) - np.min(bar):.2f} | commit after last barrier {(done[e] - t0) / 1e3 - np.max(bar):.2f}")
slow = np.argsort(us[12, :, 3] - us[11, :, 3])[-6:]
print("slowest blocks of eval 12 (index among live, period):", [(int(i), round(float(us[12, i, 3] - us[11, i, 3]), 2)) for i in slow])
ctx.close()
