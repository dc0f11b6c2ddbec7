#!/usr/bin/env python
"""bench.py — plane-residual evaluation throughput on the 12-room apartment (BASELINE.json configs[2]).

One "step" = one evaluation of the cuboid objective + gradient sums of all 12 rooms over every point of the
apartment (nearest-plane assignment, Float residuals, Double reductions; hs_rooms_cuboid_sums), followed at N>1 by
the all-reduce of the 12 x 24-double records (NCCL, the path's only exchange).  Points are sharded by contiguous
point range across ranks; total work is fixed => "strong" scaling (use --scaling weak for a fixed
100 M points per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement): value = whole-job points/s with inputs resident
in HBM; e2e = same metric through the public host-buffer API with the H2D copy of every point and the D2H of the
record inside the timed region; roofline = 12 B/pt algorithmic bytes / kernel time against the measured HBM peak;
cpu_baseline = the oracle (C port of the Haskell reference, OpenMP) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROOMS = 12
PTS_PER_ROOM = 8_333_334  # 12 rooms ~ 100 M points (BASELINE.json configs[2])
METRIC = "plane_residual_eval_points_per_sec"
UNIT = "points/s"
BYTES_PER_POINT = 12.0  # residual+gradient evaluation reads 12 B/pt, writes ~0


def workload_config(n_total):
    """the `config` object: identical in both arms (the driver compares them)"""
    return {"workload": "12-room grid apartment (BASELINE configs[2]): per-room cuboid residual + gradient sums, nearest-plane assignment",
            "rooms": N_ROOMS, "points_total": int(n_total),
            "l2": "inputs larger than L2: every step streams all %.0f MB of points (L2 is 126 MB per GPU)" % (n_total * 12 / 1e6)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def room_params(n_rooms=N_ROOMS, seed=3):
    from housescan_b200 import synth

    rng = np.random.default_rng(seed)
    params = np.zeros((n_rooms, 10))
    for r, (gx, gz) in enumerate(synth.diagonal_pairs(n_rooms)):
        params[r, :3] = np.array([6.0 * gx, 0.0, 6.0 * gz]) + rng.uniform(-0.2, 0.2, size=3)
        params[r, 3:6] = np.array([5.0, 2.6, 4.0]) + rng.uniform(-0.3, 0.3, size=3)
        params[r, 6:] = synth.quat_from_axis_angle([0, 1, 0], rng.uniform(-3, 3))
    return params


def eval_params(params, seed=33):
    """the parameters the optimiser would be evaluating: a few mm / mrad away from the generating ones"""
    rng = np.random.default_rng(seed)
    p = params.copy()
    p[:, :6] += rng.normal(0, 0.004, size=(p.shape[0], 6))
    p[:, 6:] += rng.normal(0, 0.002, size=(p.shape[0], 4))
    return p


def gen_points_torch(torch, dev, params, counts, seed, sigma=0.005):
    """uniform points on the faces of each room's cuboid + normal noise, generated on the device (float32 AoS)."""
    from housescan_b200 import synth

    total = int(sum(counts))
    pad = ((total * 12 + 47) // 48) * 48 // 4 + 16
    buf = torch.empty(pad, dtype=torch.float32, device=dev)
    out = buf[: total * 3].view(total, 3)
    g = torch.Generator(device=dev)
    o = 0
    for r, n in enumerate(counts):
        if n == 0:
            continue
        g.manual_seed(seed * 1000 + r)
        p = params[r]
        dims = torch.tensor(p[3:6], dtype=torch.float32, device=dev)
        R = torch.tensor(synth.rot_rows_from_quat(p[6:]), dtype=torch.float32, device=dev)
        c = torch.tensor(p[:3], dtype=torch.float32, device=dev)
        a, b, cc = [float(v) for v in p[3:6]]
        areas = torch.tensor([b * cc, b * cc, a * cc, a * cc, a * b, a * b], dtype=torch.float32, device=dev)
        CH = 4_000_000
        for s in range(0, n, CH):
            m = min(CH, n - s)
            face = torch.multinomial(areas / areas.sum(), m, replacement=True, generator=g)
            u = (torch.rand(m, 3, device=dev, generator=g) - 0.5) * dims
            axis = face // 2
            sign = 1.0 - 2.0 * (face % 2).float()
            wall = sign * dims[axis] * 0.5 + torch.randn(m, device=dev, generator=g) * sigma
            u.scatter_(1, axis[:, None], wall[:, None])
            out[o + s : o + s + m] = u @ R + c
        o += n
    return buf, out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for line in self.f:
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                util = float(t[8])
                if util <= 0:
                    continue
                sm.append(float(t[1])); mx.append(float(t[2])); pw.append(float(t[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples under load"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle C port; no GHC in this image) on the host cores."""
    if rank != 0:
        return
    import oracle as O
    from housescan_b200 import synth

    if not os.environ.get("HS_REF_CHILD"):
        # Always in a child process with the launcher's settings removed, so that the arm is the SAME process set-up at every N:
        # torchrun exports OMP_NUM_THREADS=1 to its workers (libgomp sizes its pool and wait policy from it at load time) and its
        # workers inherit whatever affinity the launcher had.  The child uses every host CPU, unbound threads, active waiting.
        env = {k: v for k, v in os.environ.items() if not k.startswith(("TORCHELASTIC", "OMP_", "GOMP_", "KMP_", "MKL_"))}
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "GROUP_WORLD_SIZE", "ROLE_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        env["HS_REF_CHILD"] = "1"
        env["OMP_PROC_BIND"] = "false"
        env["OMP_WAIT_POLICY"] = "active"
        env["OMP_DYNAMIC"] = "false"
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--scaling", args.scaling, "--ref-pts-per-room", str(args.ref_pts_per_room)]
        out = subprocess.run(cmd, env=env, capture_output=True, text=True)
        sys.stderr.write(out.stderr)
        sys.stdout.write(out.stdout)
        sys.stdout.flush()
        if out.returncode != 0:
            raise SystemExit(out.returncode)
        return
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))  # whatever the launcher pinned us to: all host CPUs
    except Exception:
        pass
    O.build()
    per_room = args.ref_pts_per_room
    params = room_params()
    pe = eval_params(params)
    rng = np.random.default_rng(3)
    clouds = [synth.cuboid_room_cloud(per_room, params[r], sigma=0.005, rng=rng)[0] for r in range(N_ROOMS)]
    n = per_room * N_ROOMS

    def step():
        for r in range(N_ROOMS):
            O.cuboid_sums(clouds[r], pe[r])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    cores = O.num_threads()
    sample = f"{N_ROOMS} rooms x {per_room} pts per step ({n} pts = the full workload), all 12 records per step, OpenMP {cores} threads, affinity {len(os.sched_getaffinity(0))} CPUs"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 geometry / f64 accumulation",
        "data": "synthetic", "config": workload_config(n),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--pts-per-room", type=int, default=PTS_PER_ROOM)
    ap.add_argument("--ref-pts-per-room", type=int, default=PTS_PER_ROOM, help="reference arm: points per room per step (default = the full workload)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--mode", type=int, default=-1, help="evaluation kernel variant (hs_ctx_set_mode key 0)")
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5_stream", "c4_cc", "c2_export", "a12_ceiling"],
                    help="c3 (default, the headline line): 12-room apartment evaluation; the others are the remaining sharded rows of SURVEY.md 8e (tools/bench_workloads.py)")
    ap.add_argument("--frames", type=int, default=10_000, help="c5_stream: frames of the replayed stream")
    ap.add_argument("--storeys", type=int, default=50, help="c4_cc: storeys of 1000 x 1000 vertices")
    ap.add_argument("--points", type=int, default=8_000_000, help="c2_export / a12_ceiling: points of the room")
    ap.add_argument("--out-dir", default="", help="c2_export: directory of the .ply (default: the system temp dir)")
    ap.add_argument("--path", default="session", choices=["session", "launch"],
                    help="session = one resident kernel runs all K evaluations (hs_eval_session_*); launch = one kernel launch per evaluation")
    ap.add_argument("--defer-launch", action="store_true",
                    help="session path: start the resident kernel at stop() with every command in place (for profilers that block in the launch call, e.g. ncu)")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: p2p = records summed over NVLink peer memory inside the reduction kernel (product path); nccl = kernel + dist.all_reduce")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if args.workload != "c3":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_workloads

        bench_workloads.WORKLOADS[args.workload](args)
        return

    import torch
    import torch.distributed as dist

    import housescan_b200 as hb
    from housescan_b200.rooms import local_room_offsets, shard_range

    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = hb.Context(local)  # raises HS_ECUDA without an sm_100 device: no fallback
    if args.mode >= 0:
        ctx.set_mode(0, args.mode)
    if args.blocks_per_sm > 0:
        ctx.set_mode(1, args.blocks_per_sm)
    if args.defer_launch:
        ctx.set_mode(4, 1)
    # a dedicated non-default stream shared by torch (events, NCCL ordering) and the library (kernels, copies)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    # ---- workload: this rank's point range of the concatenated apartment cloud
    params = room_params()
    pe = np.ascontiguousarray(eval_params(params))
    per_room = args.pts_per_room
    if args.scaling == "weak":
        n_total = per_room * N_ROOMS * world
        offs_global = np.arange(N_ROOMS + 1, dtype=np.int64) * per_room * world
    else:
        n_total = per_room * N_ROOMS
        offs_global = np.arange(N_ROOMS + 1, dtype=np.int64) * per_room
    lo, hi = shard_range(n_total, rank, world)
    offs = local_room_offsets(offs_global, lo, hi)
    counts = np.diff(offs).tolist()
    buf, pts = gen_points_torch(torch, dev, params, counts, seed=3 + rank)
    n_local = hi - lo
    cloud = ctx.wrap(buf.data_ptr(), n_local, keepalive=buf)
    rec = torch.zeros(N_ROOMS * hb.HS_REC, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    use_p2p = world > 1 and args.collective == "p2p"
    use_session = args.path == "session" and (world == 1 or use_p2p)
    collective_check = None
    if use_p2p:
        ctx.peer_connect(rank, world)  # CUDA IPC mailboxes, handles exchanged through torch.distributed
        # cross-check once against NCCL: same records, summed by dist.all_reduce
        ctx.rooms_cuboid_sums_async(cloud, offs, pe, rec.data_ptr())
        dist.all_reduce(rec)
        ref = rec.clone()
        ctx.rooms_cuboid_sums_allreduce_async(cloud, offs, pe, rec.data_ptr())
        torch.cuda.synchronize()
        scale = ref.abs().clamp_min(1e-300)
        collective_check = float(((rec - ref).abs() / scale).max().item())
        if not collective_check < 1e-12:
            raise SystemExit(f"peer-memory all-reduce disagrees with NCCL: max rel {collective_check}")

    # The optimiser's view of the device (FitCuboidBFGS.hs:184,201,233: thousands of objective evaluations over an unchanged cloud):
    # an evaluation session = ONE resident kernel launch that runs the posted evaluations back to back, each a full pass over the
    # rank's points followed (N > 1) by the all-reduce of the records over NVLink peer memory.  One step = one evaluation.
    pe_alt = np.ascontiguousarray(pe * (1.0 + 1e-4))  # consecutive evaluations use different parameters, as an optimiser's do

    def param_batch(k):
        return np.ascontiguousarray(np.stack([pe if i % 2 == 0 else pe_alt for i in range(k)]))

    def open_steps():
        """host-side set-up of a run of steps: opens the session (no kernel yet: the resident kernel starts with the first post)"""
        return ctx.eval_session(cloud, offs, allreduce=world > 1) if use_session else None

    def run_steps(k, batch=None, sess=None):
        """enqueue k steps on the stream; returns the open session (or None) - the caller closes it after its events"""
        if use_session:
            sess = sess if sess is not None else open_steps()
            sess.post(param_batch(k) if batch is None else batch)  # parameter conversion, kernel launch, k evaluations
            sess.stop()  # non-blocking: the kernel leaves after the last posted evaluation
            return sess
        for i in range(k):
            p = pe if i % 2 == 0 else pe_alt
            if use_p2p:
                ctx.rooms_cuboid_sums_allreduce_async(cloud, offs, p, rec.data_ptr())
            else:
                ctx.rooms_cuboid_sums_async(cloud, offs, p, rec.data_ptr())
                if world > 1:
                    dist.all_reduce(rec)
        return None

    def finish(sess, want_last=False):
        out = None
        if sess is not None:
            if want_last:
                out = sess.wait(sess.posted - 1)
            sess.close()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    # ---- warm-up: W untimed steps (the contract's shape: W warm-up steps, then exactly K timed ones)
    finish(run_steps(args.warmup))
    torch.cuda.synchronize()

    # ---- value: K steps, device-timed (CUDA events on the launching stream), max over ranks
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timed_batch = param_batch(args.steps)
    align = torch.zeros(1, device=dev)

    def start_together():
        """barrier + synchronize (host side), then a stream-ordered all-reduce: the GPUs leave it within microseconds of each
        other, so the timed regions start together on the devices whatever the hosts' launch jitter is"""
        barrier()
        if world > 1:
            dist.all_reduce(align)

    sess = open_steps()
    start_together()
    torch.cuda.profiler.start()  # `ncu --profile-from-start off` lists the timed regions only (no-op without a profiler)
    l0 = ctx.launch_count
    ev0.record()
    sess = run_steps(args.steps, timed_batch, sess)
    ev1.record()
    barrier()
    launches = ctx.launch_count - l0
    timeline = None
    if sess is not None:  # the session's own device-clock stamps of the timed evaluations (per rank; the clocks of different GPUs are not comparable)
        sess.wait(sess.posted - 1, want_record=False)
        tm = np.array([sess.times(i) for i in range(args.steps)], dtype=np.int64)
        gaps = np.diff(tm[:, 1]) / 1e3 if args.steps > 1 else np.zeros(1)
        timeline = {"first_seen_to_first_commit_us": float(tm[0, 1] - tm[0, 0]) / 1e3, "commit_gap_median_us": float(np.median(gaps)), "commit_gap_max_us": float(np.max(gaps)),
                    "first_seen_to_last_commit_us": float(tm[-1, 1] - tm[0, 0]) / 1e3}
    last_rec = finish(sess, want_last=True)
    ms = ev0.elapsed_time(ev1)
    k_ms_timed = ms / args.steps  # this rank's own time per step (before the max over ranks)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if timeline is not None:
        timeline["events_minus_device_span_us"] = k_ms_timed * args.steps * 1e3 - timeline["first_seen_to_last_commit_us"]
        if world > 1:  # every rank's view, gathered: min / max over ranks
            allt = [None] * world
            dist.all_gather_object(allt, timeline)
            timeline = {k: {"min": min(t[k] for t in allt), "max": max(t[k] for t in allt)} for k in timeline}
    value = n_total * args.steps / (ms * 1e-3)
    if last_rec is None:
        last_rec = rec.cpu().numpy().reshape(N_ROOMS, hb.HS_REC).copy()
    last_params = pe if (args.steps - 1) % 2 == 0 else pe_alt

    # ---- the same K steps again after half a second of continuous load (the clock sampler needs samples under load, and this box's
    # sustained state differs from its burst state: the kernel is issue-bound and follows the SM clock under the power cap)
    n_heat = int(min(20000, max(100, round(500.0 / max(ms / args.steps, 1e-3)))))  # same value on every rank (ms is the max over ranks)
    if args.defer_launch:
        n_heat = min(n_heat, 200)  # a deferred launch must find every command in the 256-entry ring
    barrier()
    finish(run_steps(n_heat))
    sess = open_steps()
    start_together()
    ev0.record()
    sess = run_steps(args.steps, timed_batch, sess)
    ev1.record()
    barrier()
    finish(sess)
    ms_sus = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_sus], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_sus = float(t.item())

    # ---- the same evaluation as one launch per step (hs_rooms_cuboid_sums_async: what a caller without a session gets), no exchange
    barrier()
    ev0.record()
    for _ in range(args.steps):
        ctx.rooms_cuboid_sums_async(cloud, offs, last_params, rec.data_ptr())
    ev1.record()
    torch.cuda.synchronize()
    launch_ms = ev0.elapsed_time(ev1) / args.steps
    rec_host = rec.cpu().numpy().reshape(N_ROOMS, hb.HS_REC).copy()
    session_equals_launch = None
    if use_session and world == 1:  # same partition, same order of additions: bit-identical
        session_equals_launch = bool(np.array_equal(last_rec, rec_host))
        if not session_equals_launch:
            raise SystemExit("session records differ from the one-launch-per-evaluation records")
    # roofline of the dominant kernel = the kernel of the timed region: algorithmic bytes of the evaluations one launch ran / its duration
    k_ms = k_ms_timed if use_session else launch_ms
    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_POINT * n_local / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
            if int(tj.get("points_per_launch", -1)) == int(n_local):
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "profiles/traffic.json (ncu --set full capture of one evaluation launch of this workload; not measured in this run)"
    except Exception:
        pass

    # ---- e2e: host buffers in, host record out, every step (public C-ABI path with pinned host memory)
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    host_pts = torch.empty((n_local, 3), dtype=torch.float32, pin_memory=True)
    host_pts.copy_(pts)
    host_rec = torch.empty(N_ROOMS * hb.HS_REC, dtype=torch.float64, pin_memory=True)
    dcloud = ctx.alloc(n_local)
    torch.cuda.synchronize()

    def e2e_step():
        ctx.write(dcloud, host_pts.data_ptr(), n_local)  # H2D of every point of the step
        if use_p2p:
            ctx.rooms_cuboid_sums_allreduce_async(dcloud, offs, pe, rec.data_ptr())
        else:
            ctx.rooms_cuboid_sums_async(dcloud, offs, pe, rec.data_ptr())
            if world > 1:
                dist.all_reduce(rec)
        host_rec.copy_(rec, non_blocking=True)  # D2H of the result
        stream.synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    barrier()
    e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    e2e_value = n_total * e2e_steps / (e_ms * 1e-3)
    torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O

        O.build()
        hp = host_pts.numpy()
        reps, t_cpu, n_cpu = 0, 0.0, 0
        O.cuboid_sums(hp[: min(n_local, 1_000_000)], last_params[0])
        check = np.zeros((N_ROOMS, hb.HS_REC))
        while t_cpu < 10.0 and reps < 8:
            t0 = time.perf_counter()
            for r in range(N_ROOMS):
                rr = O.cuboid_sums(hp[offs[r] : offs[r + 1]], last_params[r])
                if reps == 0:
                    check[r] = rr
            t_cpu += time.perf_counter() - t0
            n_cpu += n_local
            reps += 1
        # the whole 12 x 22 record of the timed path against the oracle: counts bit-exact; sums relative to the room's largest sum
        # of the same kind (f | sum r | B-moments), the scale the optimiser sees them at
        got = last_rec
        rel = np.zeros_like(check[:, :16])
        for lo_, hi_ in ((0, 1), (1, 7), (7, 16)):
            sc = np.abs(check[:, lo_:hi_]).max(axis=1, keepdims=True)
            rel[:, lo_:hi_] = np.abs(got[:, lo_:hi_] - check[:, lo_:hi_]) / np.maximum(sc, 1e-300)
        cpu = {"value": n_cpu / t_cpu, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
               "sample": f"{reps} x the full {n_local}-pt workload (12 rooms), OpenMP over all host threads",
               "parity_counts_equal_all_rooms": bool(np.array_equal(check[:, 16:22], got[:, 16:22])),
               "parity_sums_max_rel_all_rooms": float(rel.max()),
               "parity_f_max_rel": float(rel[:, 0].max())}
        if not (cpu["parity_counts_equal_all_rooms"] and cpu["parity_f_max_rel"] < 1e-6):
            raise SystemExit(f"GPU records disagree with the oracle: {cpu}")

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32 geometry / f64 accumulation", "data": "synthetic",
            "config": workload_config(n_total),
            "detail": {"path": ("evaluation session (hs_eval_session_*): one resident kernel, one step = one posted evaluation" if use_session else "one kernel launch per evaluation"),
                       "points_per_gpu": int(n_local), "sharding": f"point-range x{world}",
                       "collective": ("none" if world == 1 else ("records summed over NVLink peer memory inside the reduction kernel (12x24 f64, CUDA IPC mailboxes)" if use_p2p
                                      else "nccl all_reduce of 12x24 f64 per step")),
                       "collective_vs_nccl_max_rel": collective_check, "session_timeline": timeline},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 12), "d2h_bytes_per_step": int(N_ROOMS * hb.HS_REC * 8),
                    "steps": e2e_steps, "ms_per_step": e_ms / e2e_steps, "h2d_gbs_per_gpu": n_local * 12 / (e_ms / e2e_steps * 1e-3) / 1e9,
                    "path": "hs_cloud_write (pinned host -> HBM) + one evaluation launch (+ exchange) + record D2H, every step"},
            "gpu_launches": int(launches),
            "evaluations_per_launch": (args.steps if use_session else 1),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": ("k_eval<session>: ONE launch ran the %d timed evaluations; kernel_ms = its duration / %d" % (args.steps, args.steps)) if use_session else "k_eval (one launch per evaluation)",
                         "kernel_ms": k_ms, "bytes_per_point": BYTES_PER_POINT,
                         "points_per_launch": int(n_local) * (args.steps if use_session else 1), "points_per_evaluation": int(n_local),
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "one_launch_per_evaluation": {"kernel_ms": launch_ms, "achieved": BYTES_PER_POINT * n_local / (launch_ms * 1e-3) / 1e9,
                                                       "frac": BYTES_PER_POINT * n_local / (launch_ms * 1e-3) / 1e9 / peak},
                         "session_records_equal_launch_records": session_equals_launch},
            "sustained": {"value": n_total * args.steps / (ms_sus * 1e-3), "ms_per_step": ms_sus / args.steps, "after_steps_of_continuous_load": n_heat,
                          "note": "same K steps re-timed after >= 0.5 s of back-to-back evaluations (power-capped clock state of this box); `value` is the contract's W warm-up + K timed steps"},
            "clocks": clocks,
            "cpu_baseline": cpu,
            "gpts_per_s": value / 1e9,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
