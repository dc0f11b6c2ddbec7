// Cuboid objective / gradient sums per room (A6 + A5; FitCuboidBFGS.hs:51-76 generalised from 8 corners to the room cloud,
// planes of Main.hs:1852-1874, distance of Main.hs:1371-1372) — the headline kernel, 12 B/point, and its resident
// multi-evaluation ("session") form.
//
// One CTA per SM.  Warps 0..NW-1 are consumers; then come the producer, the reducer, the finaliser and (session form, block 0
// only) the dispatcher warp.
//   producer   one lane streams the block's point range global -> shared in 72 KB tiles with 1-D bulk async copies (the TMA
//              engine) through a STAGES-deep mbarrier ring.  The cloud does not change between evaluations, so in the session
//              form it simply keeps streaming: the first tiles of evaluation e+1 are in shared memory before e has finished.
//   consumers  each thread takes GPT groups (4 points = 48 B = 3 x LDS.128, conflict-free) per tile, evaluates them with the
//              predicated point block of k_eval_point.cuh into 21 Float chains, and every FLUSH_TILES tiles the warp transposes
//              its chains with shuffles (lane l ends up with the warp total of chain l: 104 instructions instead of the 22
//              butterflies of round 1) and lane l adds its total into ONE Double register.  At the end of a room segment the
//              lanes leave their Doubles in shared memory, sync, and go straight on to the next segment / evaluation: they never
//              touch global memory and never wait for a reduction (at most EV_D evaluations in flight).  In the session form
//              what the next evaluation needs (control words, plane constants) was fetched with cp.async while this one streamed,
//              and the go decision rides on the segment's own barrier: no bubble between evaluations.  A segment's ragged head /
//              tail points (room offsets are arbitrary) are staged with cp.async too and evaluated after the tiles.
//   reducer    adds the warps' Doubles in warp order into the block's partial record of (evaluation, room), publishes it and
//              takes the room's ticket; the block that delivers the LAST partial of a room adds the room's partials in block
//              order (deterministic) and converts the raw sums into the HS_REC record.  Never waits for anything remote.
//   finaliser  the block that completed the last room of an evaluation finalises it: with a peer group the records are
//              exchanged over NVLink peer memory (k_peer.cuh) and summed in rank order; evaluations commit in order.
//   dispatcher polls the host-mapped command ring, copies new plane tables into device memory and publishes them; enforces
//              the idle watchdog (a forgotten session cannot hang the GPU).
//
// The static partition (EvalPlan, computed on the host) gives blocks that hold a room boundary fewer groups, so that all
// blocks finish an evaluation together although a second segment costs a second flush / block sum.
#pragma once
#include "k_common.cuh"
#include "k_ring.cuh"
#include "k_eval_point.cuh"
#include "k_peer.cuh"

namespace hsk {

constexpr int EV_MAXB = 256;   // blocks of a plan (>= SM count)
constexpr int EV_D = 8;        // evaluations in flight per GPU: partial-record ring, tickets (PEER_SLOTS >= 2 * EV_D)
constexpr int EV_NRAW = 22;    // raw sums per (block, room): f, T[3], M[3], B[9], C1, C2, Cm[3], N
constexpr int EV_QCAP = 256;   // command / result ring entries of a session
constexpr int EV_PARK = 16;    // entries of the consumers -> reducer queue
constexpr int EV_WS = 4;       // segments the reducer may lag behind the consumers (buffers of warp sums)
constexpr int EV_FQ = 16;      // finaliser queue entries (>= EV_D)
constexpr int EV_TRACE_EVALS = 32;  // evaluations covered by a session trace
constexpr int EV_PF = 4;       // segments of a block whose constants are prefetched for the next evaluation (more: loaded at use)
static_assert(PEER_SLOTS >= 2 * EV_D, "mailbox slots must cover two windows of in-flight evaluations");

struct EvalPlan {  // device memory; fixed for a (cloud, room offsets, grid) triple
  int32_t nblocks, nrooms, nrooms_nonempty;
  uint32_t empty_mask;  // bit r: room r holds no point (its record is zero)
  int64_t n;
  int64_t off[HS_MAX_ROOMS + 1];
  int64_t blk_g0[EV_MAXB + 1];  // block b owns the 4-point groups [blk_g0[b], blk_g0[b+1])
  int32_t blk_rfirst[EV_MAXB], blk_rlast[EV_MAXB];      // rooms with points in the block's range (rlast < rfirst: none)
  int32_t room_blo[HS_MAX_ROOMS], room_nb[HS_MAX_ROOMS];  // blocks with points of the room: room_blo .. room_blo + room_nb - 1
};

constexpr int EV_CMD_F = 16;  // floats per room: one 64-byte line
struct EvalCmd {  // one evaluation's plane constants: per room n[3][3] (normals of the + walls), dp[3], dm[3], [15] = bits of (sequence number + 1)
  float c[HS_MAX_ROOMS][EV_CMD_F];
};

struct EvalCtl {  // device memory, zero between launches / at session begin
  uint32_t posted;    // session: commands available in d_cmds
  uint32_t stop;      // session: no more commands will come
  uint32_t done_seq;  // evaluations finalised (in order)
  uint32_t error;     // 1: peer timeout, 2: idle watchdog
  uint32_t room_ticket[EV_D][HS_MAX_ROOMS];
  uint32_t rooms_done[EV_D];
};

struct alignas(16) EvalHostCtl {  // mapped pinned host memory (session)
  uint32_t posted;  // host -> device: commands written to h_cmds  } read together by the dispatcher
  uint32_t stop;    // host -> device                               }
  uint32_t done;    // device -> host: evaluations whose records are in h_results
  uint32_t error;   // device -> host
};

struct EvalArgs {
  const float* xyz;
  const EvalPlan* plan;
  EvalCtl* ctl;
  double* partials;   // [EV_D][nrooms][nblocks][EV_NRAW]
  double* local_rec;  // [EV_D][HS_MAX_ROOMS * HS_REC]: this GPU's records before the exchange
  double* out;        // one-shot: nrooms x HS_REC; session: ring [EV_QCAP][nrooms * HS_REC]
  EvalCmd* d_cmds;              // session: device command ring [EV_QCAP]
  const EvalCmd* h_cmds;        // session: host-mapped command ring [EV_QCAP]
  EvalHostCtl* h_ctl;           // session
  double* h_results;            // session: host-mapped result ring [EV_QCAP][nrooms * HS_REC]
  unsigned long long* h_times;  // session: host-mapped [EV_QCAP][2] %globaltimer stamps: command seen by the device, records committed
  unsigned long long* trace;    // session tracing (HS_EVAL_TRACE=<file>): [EV_TRACE_EVALS][nblocks][4] %globaltimer stamps, else nullptr
  uint32_t* h_status;           // mapped word of the ctx: set to HS_ENCCL when a peer timed out
  unsigned long long idle_timeout_ns;
  uint32_t epoch0;              // peer epoch of evaluation 0 (epochs are > 0)
  uint32_t pad;
  PeerExchange px;              // world <= 1: no exchange
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_sys_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ float4 ld_sys_v4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

template <int NCONS>
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// Transposing warp reduction of the 22 chain values: on return lane l (< 22) holds the warp total of value l.
// Step S exchanges halves between lanes l and l ^ S: lanes with bit S keep the upper S slots.  Slots 22..31 are zero, so the
// first step needs selects only for the 6 slot pairs (i, i + 16) that are both live.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float keep = up ? v[i + 16] : v[i], send = up ? v[i] : v[i + 16];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 6; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);  // upper lanes end up with unused copies
  }
#pragma unroll
  for (int S = 8; S >= 1; S >>= 1) {
    const bool up = lane & S;
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const float keep = up ? v[i + S] : v[i], send = up ? v[i] : v[i + S];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
    }
  }
  return v[0];
}

// chains of the whole warp -> one Double per lane (lane l accumulates raw sum l).  npts = points this lane added since the last flush.
__device__ __forceinline__ void flush_chains(ChainsP& c, int& npts, double& dacc, int lane) {
  float v[32];
  v[0] = c.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    v[1 + j] = c.T[j]; v[4 + j] = c.M[j]; v[18 + j] = c.Cm[j];
#pragma unroll
    for (int q = 0; q < 3; ++q) v[7 + 3 * j + q] = c.B[j][q];
  }
  v[16] = c.C1; v[17] = c.C2; v[21] = static_cast<float>(npts);
#pragma unroll
  for (int i = 22; i < 32; ++i) v[i] = 0.f;
  dacc += static_cast<double>(warp_transpose_sum(v, lane));
  c.clear();
  npts = 0;
}

__device__ __forceinline__ void load_room_consts(RoomK& R, const float* c16) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    R.n[j][0] = c16[3 * j]; R.n[j][1] = c16[3 * j + 1]; R.n[j][2] = c16[3 * j + 2];
    R.dp[j] = c16[9 + j]; R.dm[j] = c16[12 + j];
  }
}

// raw room sums -> record component L (= lane); raw value k lives on lane k of the calling warp, lanes >= EV_NRAW hold 0.
//   rec[0] = f; rec[1+2j] = T_j - M_j (sum r over the + wall), rec[2+2j] = -M_j (sum r over the - wall, r = -s);
//   rec[7..15] = B; rec[16+2j] = count of the + wall = axis count - Cm_j, rec[17+2j] = Cm_j; axis counts N - C1 - C2, C1, C2
__device__ __forceinline__ double raw_to_record(double raw, int L) {
  int ia = 31, ib = 31, ic = 31, id = 31;  // rec = raw[ia] - raw[ib] - raw[ic] - raw[id]; 31 = a lane that holds zero
  bool neg = false;
  if (L == 0) ia = 0;
  else if (L <= 6) { const int j = (L - 1) >> 1; if (L & 1) { ia = 1 + j; ib = 4 + j; } else { ia = 4 + j; neg = true; } }
  else if (L <= 15) ia = L;
  else if (L <= 21) {
    const int j = (L - 16) >> 1;
    if (L & 1) ia = 18 + j;
    else { id = 18 + j; if (j == 0) { ia = 21; ib = 16; ic = 17; } else ia = 15 + j; }
  }
  const double va = __shfl_sync(0xffffffffu, raw, ia), vb = __shfl_sync(0xffffffffu, raw, ib);
  const double vc = __shfl_sync(0xffffffffu, raw, ic), vd = __shfl_sync(0xffffffffu, raw, id);
  const double r = ((va - vb) - vc) - vd;
  return neg ? -r : r;
}

__device__ __forceinline__ void add_group(ChainsP& ch, const RoomK& R, const float4& q0, const float4& q1, const float4& q2) {
  add_point(ch, R, q0.x, q0.y, q0.z); add_point(ch, R, q0.w, q1.x, q1.y);
  add_point(ch, R, q1.z, q1.w, q2.x); add_point(ch, R, q2.y, q2.z, q2.w);
}

template <int NCONS, int STAGES, int GPT, int FLUSH_TILES, bool SESSION>
__global__ void __launch_bounds__(NCONS + (SESSION ? 128 : 96), 1)
k_eval(const __grid_constant__ EvalArgs a, const __grid_constant__ EvalCmd cmd0) {
  constexpr int NW = NCONS / 32;
  constexpr int W_PRODUCER = NW, W_REDUCER = NW + 1, W_FINAL = NW + 2, W_DISPATCH = NW + 3;
  constexpr int TILE_GROUPS = GPT * NCONS;
  constexpr uint32_t TILE_BYTES = TILE_GROUPS * 48;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STAGES) * TILE_BYTES);
  uint64_t* empty = full + STAGES;
  double* wsum = reinterpret_cast<double*>(empty + STAGES);  // [EV_WS][NW][EV_NRAW]: the warps' Doubles of the last EV_WS segments, summed by the reducer
  __shared__ int64_t s_lo[HS_MAX_ROOMS], s_hi[HS_MAX_ROOMS];  // point range of each room segment of this block
  __shared__ uint32_t s_rq[EV_PARK], s_fq[EV_FQ];            // consumers -> reducer (e << 8 | room), reducer -> finaliser (e)
  __shared__ volatile uint32_t s_rq_tail, s_rq_head, s_fq_tail, s_exit, s_exit2, s_consumed_lo, s_consumed_hi, s_go;
  __shared__ volatile uint32_t s_go_fast[2];  // [parity of the evaluation]: decided before the previous evaluation's final barrier
  // session: what the NEXT evaluation needs, fetched with cp.async while the current one streams (no bubble between evaluations):
  // a snapshot of the control words {posted, stop, done_seq, error} and the plane constants of the block's first EV_PF segments
  __shared__ float s_rag[6][3];  // a segment's ragged head (<= 3) and tail (<= 3) points, staged with cp.async while the tiles stream
  __shared__ __align__(16) uint32_t s_ctl_snap[4];
  __shared__ volatile uint32_t s_pf_ok[2];  // [parity of the evaluation]: its constants were prefetched from a command that had been published before
  __shared__ __align__(16) float s_cmd_pf[2][EV_PF][EV_CMD_F];

  const EvalPlan* __restrict__ plan = a.plan;
  const int b = static_cast<int>(blockIdx.x);
  const int nrooms = plan->nrooms, nblocks = plan->nblocks;
  const int rfirst = plan->blk_rfirst[b], rlast = plan->blk_rlast[b];
  const int nseg = rlast - rfirst + 1;
  const int warp = static_cast<int>(threadIdx.x >> 5), lane = static_cast<int>(threadIdx.x & 31);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_rq_tail = 0; s_rq_head = 0; s_fq_tail = 0; s_exit = 0; s_exit2 = 0; s_consumed_lo = 0; s_consumed_hi = 0; s_go = 0; s_go_fast[0] = 0; s_go_fast[1] = 0; s_pf_ok[0] = 0; s_pf_ok[1] = 0;
  }
  if (static_cast<int>(threadIdx.x) < nseg) {
    const int64_t p0 = plan->blk_g0[b] * 4, p1 = min(plan->blk_g0[b + 1] * 4, plan->n);
    const int r = rfirst + static_cast<int>(threadIdx.x);
    s_lo[threadIdx.x] = max(plan->off[r], p0);
    s_hi[threadIdx.x] = min(plan->off[r + 1], p1);
  }
  __syncthreads();
  if (nseg <= 0 && !(SESSION && b == 0 && warp == W_DISPATCH)) return;  // nothing to stream (tiny clouds): no ticket counts on this block

  // =================================================================================================== dispatcher (session)
  if (SESSION && warp == W_DISPATCH) {
    if (b != 0) return;
    uint32_t seq = 0, spins = 0;
    unsigned long long t_last = peer_now_ns();
    for (;;) {
      uint32_t hp, hstop;  // {posted, stop} sit side by side: ONE read over PCIe per poll
      asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(hp), "=r"(hstop) : "l"(&a.h_ctl->posted) : "memory");
      if (hp != seq) {
        // copy the new commands host -> device ring (one 64-byte line per room) and publish them one by one: the consumers start on
        // the first while the rest still crosses PCIe.  All loads of a command are in flight together (<= 4 per lane for 32 rooms).
        const unsigned long long t_seen = peer_now_ns();
        while (seq != hp) {
          if (lane == 0) a.h_times[2 * (seq % EV_QCAP)] = t_seen;
          const float4* src = reinterpret_cast<const float4*>(a.h_cmds[seq % EV_QCAP].c);
          float4* dst = reinterpret_cast<float4*>(a.d_cmds[seq % EV_QCAP].c);
          const int nv = nrooms * (EV_CMD_F / 4);
          float4 v[HS_MAX_ROOMS * (EV_CMD_F / 4) / 32];
#pragma unroll
          for (int k = 0; k < HS_MAX_ROOMS * (EV_CMD_F / 4) / 32; ++k)
            if (lane + 32 * k < nv) v[k] = ld_sys_v4(src + lane + 32 * k);
#pragma unroll
          for (int k = 0; k < HS_MAX_ROOMS * (EV_CMD_F / 4) / 32; ++k)
            if (lane + 32 * k < nv) dst[lane + 32 * k] = v[k];
          ++seq;
          __threadfence();
          __syncwarp();
          if (lane == 0) st_release_u32(&a.ctl->posted, seq);
        }
        t_last = peer_now_ns();
        spins = 0;
        continue;
      }
      if (hstop) {  // the host stores `posted` before `stop` (release order), and this read saw both: nothing is left behind
        if (lane == 0) st_release_u32(&a.ctl->stop, 1u);
        return;
      }
      if ((++spins & 15u) == 0u) {  // off the fast path: errors raised by a finaliser, the idle watchdog
        if (ld_acquire_u32(&a.ctl->error)) { if (lane == 0) st_release_u32(&a.ctl->stop, 1u); return; }
        if (ld_acquire_u32(&a.ctl->done_seq) != seq) t_last = peer_now_ns();  // evaluations still running: not idle
        else if (peer_now_ns() - t_last > a.idle_timeout_ns) {  // watchdog: nobody posts, nobody stops
          if (lane == 0) { a.ctl->error = 2u; st_sys_u32(&a.h_ctl->error, 2u); __threadfence_system(); st_release_u32(&a.ctl->stop, 1u); }
          return;
        }
      }
      __nanosleep(100);
    }
  }

  // =================================================================================================== producer
  if (warp == W_PRODUCER) {
    if (lane != 0) return;
    uint64_t tt = 0;  // tiles issued so far (continuous across rooms and evaluations)
    bool stop = false;
    for (uint32_t e = 0; (SESSION || e == 0) && !stop; ++e) {
      for (int si = 0; si < nseg && !stop; ++si) {
        const int64_t gl = (s_lo[si] + 3) >> 2, gh = s_hi[si] >> 2;
        for (int64_t tg = gl; tg < gh; tg += TILE_GROUPS, ++tt) {
          const int s = static_cast<int>(tt % STAGES);
          if (tt >= STAGES) {
            const uint32_t par = static_cast<uint32_t>(((tt / STAGES) - 1) & 1);
            uint32_t ok;
            for (;;) {  // poll, then sleep: a bare spin takes issue slots from the consumer warps on this scheduler
              asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                           : "=r"(ok) : "r"(smem_u32(empty + s)), "r"(par) : "memory");
              if (ok) break;
              if (SESSION && s_exit) { stop = true; break; }
              __nanosleep(100);
            }
            if (stop) break;
          }
          const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(TILE_GROUPS), gh - tg) * 48);
          mbar_expect_tx(full + s, bytes);
          bulk_g2s(tiles + static_cast<size_t>(s) * TILE_GROUPS * 3, reinterpret_cast<const float4*>(a.xyz) + 3 * tg, bytes, full + s);
        }
      }
      if (SESSION && tt == 0) break;  // a block whose segments hold no whole group never streams
    }
    if (SESSION) {  // tiles in flight that nobody will consume: the copies must have landed before the block's shared memory goes away
      while (!s_exit) __nanosleep(100);
      const uint64_t consumed = (static_cast<uint64_t>(s_consumed_hi) << 32) | s_consumed_lo;
      for (uint64_t t = consumed; t < tt; ++t) mbar_wait(full + (t % STAGES), static_cast<uint32_t>((t / STAGES) & 1));
    }
    return;
  }

  // =================================================================================================== reducer
  if (warp == W_REDUCER) {
    uint32_t head = 0, ftail = 0;
    // the plan entries this warp needs, read once: lane i keeps (first block, block count) of room rfirst + i
    const int nonempty = plan->nrooms_nonempty;
    const uint32_t empty_mask = plan->empty_mask;
    const int my_blo = (lane < nseg) ? plan->room_blo[rfirst + lane] : 0, my_nb = (lane < nseg) ? plan->room_nb[rfirst + lane] : 0;
    for (;;) {
      bool leave = false;
      while (s_rq_tail == head) {  // long naps: every poll takes issue slots from the consumer warps of this scheduler
        if (s_exit && s_rq_tail == head) { leave = true; break; }
        __nanosleep(200);
      }
      if (leave) break;
      __threadfence_block();
      const uint32_t item = s_rq[head % EV_PARK];
      const uint32_t e = item >> 8;
      const int r = static_cast<int>(item & 255u);
      const uint32_t slot = e % EV_D;
      // ---- publish the block's partial record of (evaluation, room)
      const int blo = __shfl_sync(0xffffffffu, my_blo, r - rfirst), nb = __shfl_sync(0xffffffffu, my_nb, r - rfirst);
      double* part = a.partials + ((static_cast<size_t>(slot) * nrooms + r) * nblocks) * EV_NRAW;
      if (lane < EV_NRAW) {  // the warps' Doubles added in warp order (deterministic)
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) sum += wsum[((head % EV_WS) * NW + w) * EV_NRAW + lane];
        part[static_cast<size_t>(b - blo) * EV_NRAW + lane] = sum;
        __threadfence();
      }
      __syncwarp();
      ++head;
      if (lane == 0) s_rq_head = head;  // this segment's warp sums may be overwritten
      uint32_t last = 0;
      if (lane == 0) {
        const uint32_t t = atomicAdd(&a.ctl->room_ticket[slot][r], 1u);
        last = (t == static_cast<uint32_t>(nb) - 1u);
        if (last) { a.ctl->room_ticket[slot][r] = 0u; __threadfence(); }
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (!last) continue;
      // ---- last partial of the room: add the room's partials in block order (all loads of a batch in flight together)
      double raw = 0.0;
      if (lane < EV_NRAW) {
        constexpr int U = 8;
        for (int k0 = 0; k0 < nb; k0 += U) {
          double v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = (k0 + u < nb) ? __ldcg(part + static_cast<size_t>(k0 + u) * EV_NRAW + lane) : 0.0;
#pragma unroll
          for (int u = 0; u < U; ++u) raw += v[u];
        }
      }
      const double rec = raw_to_record(raw, lane);
      const bool direct = !SESSION && a.px.world <= 1;  // one-shot on one GPU: the record goes straight to its destination
      double* lrec = direct ? a.out : a.local_rec + static_cast<size_t>(slot) * (HS_MAX_ROOMS * HS_REC);
      if (lane < HS_REC) { lrec[r * HS_REC + lane] = rec; __threadfence(); }
      __syncwarp();
      last = 0;
      if (lane == 0) {
        const uint32_t t = atomicAdd(&a.ctl->rooms_done[slot], 1u);
        last = (t == static_cast<uint32_t>(nonempty) - 1u);
        if (last) { a.ctl->rooms_done[slot] = 0u; __threadfence(); }
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (!last) continue;
      // ---- last room of the evaluation
      for (uint32_t m = empty_mask; m; m &= m - 1u)  // rooms without points: zero records
        if (lane < HS_REC) lrec[(__ffs(m) - 1) * HS_REC + lane] = 0.0;
      if (direct) continue;  // the end of the kernel publishes a.out
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        s_fq[ftail % EV_FQ] = e;
        __threadfence_block();
        s_fq_tail = ++ftail;
      }
    }
    if (lane == 0) { __threadfence_block(); s_exit2 = 1u; }
    return;
  }

  // =================================================================================================== finaliser
  if (warp == W_FINAL) {
    uint32_t head = 0;
    const int count = nrooms * HS_REC;
    for (;;) {
      bool leave = false;
      while (s_fq_tail == head) {
        if (s_exit2 && s_fq_tail == head) { leave = true; break; }
        __nanosleep(400);
      }
      if (leave) break;
      __threadfence_block();
      const uint32_t e = s_fq[head % EV_FQ];
      ++head;
      __threadfence();
      const double* lrec = a.local_rec + static_cast<size_t>(e % EV_D) * (HS_MAX_ROOMS * HS_REC);
      double* dst = SESSION ? a.out + static_cast<size_t>(e % EV_QCAP) * count : a.out;
      bool ok = true;
      // the exchange of evaluation e depends on the peers only: it runs while earlier evaluations are still being finalised by
      // other blocks; only the publication below is in order
      if (a.px.world > 1) {
        // (pushing from the reducer warp instead - so that a push never queues behind a collect - was measured at 8 GPUs: 39 us per
        // evaluation against 33, profiles/r02f_bench_n8_reducer_push_rejected.json: the reducer is the busier warp of the block)
        peer_push_warp(a.px, a.epoch0 + e, lrec, count);
        ok = peer_collect_warp(a.px, a.epoch0 + e, dst, count);
      } else {
        for (int i = lane; i < count; i += 32) dst[i] = __ldcg(lrec + i);
      }
      if (SESSION) {
        double* hdst = a.h_results + static_cast<size_t>(e % EV_QCAP) * count;
        for (int i = lane; i < count; i += 32) hdst[i] = dst[i];  // each lane re-reads what it wrote itself; posted PCIe writes
        // in-order commit: the done counters advance evaluation by evaluation (the writes above travel meanwhile)
        const unsigned long long t0 = peer_now_ns();
        while (ld_acquire_u32(&a.ctl->done_seq) != e) {
          if (peer_now_ns() - t0 > 2 * PEER_TIMEOUT_NS) { ok = false; break; }
          __nanosleep(250);
        }
      }
      if (!ok && lane == 0) {
        if (a.h_status) st_sys_u32(a.h_status, static_cast<uint32_t>(HS_ENCCL));
        if (SESSION) { a.ctl->error = 1u; st_sys_u32(&a.h_ctl->error, 1u); }  // (the one-launch form leaves the control block untouched)
      }
      if (SESSION) {
        if (lane == 0) a.h_times[2 * (e % EV_QCAP) + 1] = peer_now_ns();
        __threadfence_system();  // every lane's record stores (device ring, host ring) before the counters
        __syncwarp();
        if (lane == 0) {
          st_sys_u32(&a.h_ctl->done, e + 1u);
          st_release_u32(&a.ctl->done_seq, e + 1u);
        }
      }
    }
    return;
  }

  // =================================================================================================== consumers
  uint64_t tt = 0;  // tiles consumed so far
  uint32_t stage = 0, parity = 0;
  int last_si = -1;  // last segment of the block that holds points (the evaluation's final barrier happens there)
  for (int si = 0; si < nseg; ++si)
    if (s_lo[si] < s_hi[si]) last_si = si;
  uint32_t qtail = 0;  // segments handed to the reducer so far
  const uint32_t tiles_s = smem_u32(tiles) + threadIdx.x * 48, empty_s = smem_u32(empty);
  for (uint32_t e = 0; SESSION || e == 0; ++e) {
    if (SESSION && !(e > 0 && s_go_fast[e & 1])) {
      // wait for command e (or the end of the session); never run more than EV_D evaluations ahead of the finalised ones.
      // One thread decides for the block: a split decision would leave warps behind at the named barrier.  (Fast path: thread 0
      // took the decision from the control snapshot before the previous evaluation's final barrier, see below - no second barrier.)
      if (threadIdx.x == 0) {
        uint32_t go = 0, known_posted = 0;
        while (!go) {
          const uint32_t posted = ld_acquire_u32(&a.ctl->posted), done = ld_acquire_u32(&a.ctl->done_seq);
          known_posted = posted;
          if (static_cast<int32_t>(posted - e) > 0) {
            if (e - done < static_cast<uint32_t>(EV_D)) { go = 1; break; }
          } else if (ld_acquire_u32(&a.ctl->stop) && static_cast<int32_t>(ld_acquire_u32(&a.ctl->posted) - e) <= 0) break;
          if (ld_acquire_u32(&a.ctl->error)) break;
          __nanosleep(40);
        }
        s_go = go;
        // command e + 1 may be prefetched only if it is known to be published already (the dispatcher writes a command's lines
        // unordered and publishes `posted` after a fence: an unpublished command can be half written)
        s_pf_ok[(e + 1) & 1] = static_cast<int32_t>(known_posted - (e + 1)) > 0 ? 1u : 0u;
      }
      consumers_sync<NCONS>();
      if (!s_go) break;
    }
    if (SESSION) {
      // prefetch for evaluation e + 1 (asynchronous, lands long before this evaluation's first segment ends)
      if (threadIdx.x == 0) cp_async16(smem_u32(s_ctl_snap), a.ctl);
      else if (s_pf_ok[(e + 1) & 1] && static_cast<int>(threadIdx.x) - 32 >= 0 && static_cast<int>(threadIdx.x) - 32 < 4 * min(nseg, EV_PF)) {
        const int q = static_cast<int>(threadIdx.x) - 32, sj = q >> 2, part = q & 3;
        cp_async16(smem_u32(&s_cmd_pf[(e + 1) & 1][sj][4 * part]), a.d_cmds[(e + 1) % EV_QCAP].c[rfirst + sj] + 4 * part);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    bool pf_waited = !SESSION;
    const bool tracing = SESSION && a.trace != nullptr && e < static_cast<uint32_t>(EV_TRACE_EVALS) && threadIdx.x == 64;
    unsigned long long* tr = tracing ? a.trace + (static_cast<size_t>(e) * nblocks + b) * 4 : nullptr;
    if (tracing) tr[0] = peer_now_ns();
    for (int si = 0; si < nseg; ++si) {
      const int r = rfirst + si;
      const int64_t lo = s_lo[si], hi = s_hi[si];
      if (lo >= hi) continue;  // a room without points between two rooms of this block: no segment, no partial, no ticket (room_nb counts none)
      RoomK R;
      if (SESSION) {
        // constants of (evaluation e, room r): from the prefetch buffer when it holds exactly this command (stamp), else from L2
        const float* pf = s_cmd_pf[e & 1][si < EV_PF ? si : 0];
        if (si < EV_PF && e > 0 && s_pf_ok[e & 1] && __float_as_uint(pf[15]) == e + 1u) {
          load_room_consts(R, pf);
        } else {
          const float* c16 = a.d_cmds[e % EV_QCAP].c[r];
          float t[15];
#pragma unroll
          for (int i = 0; i < 15; ++i) t[i] = __ldcg(c16 + i);
          load_room_consts(R, t);
        }
      } else {
        load_room_consts(R, cmd0.c[r]);
      }
      ChainsP ch;
      ch.clear();
      int npts = 0;
      double dacc = 0.0;
      bool has_rag = false;
      int ragslot = 0;
      const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
      if (gl <= gh) {
        // ragged head / tail points (at most 3 each): the same per-point block, one point per thread
        const int64_t head_end = gl * 4, tail_begin = gh * 4;
        const int64_t nh = head_end - lo, ntail = hi - tail_begin;
        // staged into shared memory asynchronously, evaluated after the tiles: no warp waits for a global load at the segment's start
        const bool rag_head = static_cast<int64_t>(threadIdx.x) < nh, rag_tail = threadIdx.x >= 32 && static_cast<int64_t>(threadIdx.x) - 32 < ntail;
        has_rag = rag_head || rag_tail;
        ragslot = rag_head ? static_cast<int>(threadIdx.x) : 3 + static_cast<int>(threadIdx.x) - 32;
        if (has_rag) {
          const int64_t i = rag_head ? lo + threadIdx.x : tail_begin + threadIdx.x - 32;
#pragma unroll
          for (int c = 0; c < 3; ++c) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_rag[ragslot][c])), "l"(a.xyz + 3 * i + c) : "memory");
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const int64_t ngroups = gh - gl;
        const int nfull = static_cast<int>(ngroups / TILE_GROUPS);
        const int rem_groups = static_cast<int>(ngroups - static_cast<int64_t>(nfull) * TILE_GROUPS);
        int since_flush = 0;
        for (int t = 0; t < nfull; ++t) {
          mbar_wait(full + stage, parity);
          if (tracing && t == 0 && si == 0) tr[1] = peer_now_ns();
          const uint32_t base = tiles_s + stage * TILE_BYTES;
          float4 q[GPT][3];
#pragma unroll
          for (int g = 0; g < GPT; ++g) { q[g][0] = lds_v4(base + g * NCONS * 48); q[g][1] = lds_v4(base + g * NCONS * 48 + 16); q[g][2] = lds_v4(base + g * NCONS * 48 + 32); }
          mbar_arrive_s(empty_s + 8 * stage);  // the values are in registers: hand the slot back before the math
#pragma unroll
          for (int g = 0; g < GPT; ++g) add_group(ch, R, q[g][0], q[g][1], q[g][2]);
          npts += 4 * GPT;
          if (++stage == STAGES) { stage = 0; parity ^= 1u; }
          if (++since_flush == FLUSH_TILES) { flush_chains(ch, npts, dacc, lane); since_flush = 0; }
        }
        tt += nfull;
        if (rem_groups) {  // partial last tile of the segment: the threads whose group is in range
          mbar_wait(full + stage, parity);
          const uint32_t base = tiles_s + stage * TILE_BYTES;
          float4 q[GPT][3];
#pragma unroll
          for (int g = 0; g < GPT; ++g) {  // out-of-range slots: stale but valid shared memory, skipped below
            q[g][0] = lds_v4(base + g * NCONS * 48); q[g][1] = lds_v4(base + g * NCONS * 48 + 16); q[g][2] = lds_v4(base + g * NCONS * 48 + 32);
          }
          mbar_arrive_s(empty_s + 8 * stage);
#pragma unroll
          for (int g = 0; g < GPT; ++g)
            if (g * NCONS + static_cast<int>(threadIdx.x) < rem_groups) {
              add_group(ch, R, q[g][0], q[g][1], q[g][2]);
              npts += 4;
            }
          if (++stage == STAGES) { stage = 0; parity ^= 1u; }
          ++tt;
        }
      } else {
        const int64_t i = lo + threadIdx.x;
        if (i < hi) { add_point(ch, R, a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2]); ++npts; }
      }
      if (tracing && si == last_si) tr[2] = peer_now_ns();
      if (has_rag) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        add_point(ch, R, s_rag[ragslot][0], s_rag[ragslot][1], s_rag[ragslot][2]);
        ++npts;
      }
      flush_chains(ch, npts, dacc, lane);
      // ---- the block's sums of this segment: every warp leaves its Doubles in buffer qtail % EV_WS and moves on at once; the
      // reducer warp adds them in warp order.  The buffer is free once the reducer has taken segment qtail - EV_WS.
      if (qtail >= static_cast<uint32_t>(EV_WS))
        while (static_cast<int32_t>(s_rq_head - (qtail - (EV_WS - 1))) < 0) __nanosleep(32);  // only with many tiny rooms per block
      if (lane < EV_NRAW) wsum[((qtail % EV_WS) * NW + warp) * EV_NRAW + lane] = dacc;
      if (!pf_waited) { asm volatile("cp.async.wait_all;" ::: "memory"); pf_waited = true; }  // own copies done; the barrier publishes them
      if (SESSION && si == last_si && threadIdx.x == 0) {
        // decision for evaluation e + 1 from the snapshot this thread fetched at the start of e: posted, not throttled, no error
        // => everybody goes straight on after this barrier.  Anything else takes the polling path at the top of the loop, AFTER
        // this evaluation's sums have been handed over (a one-by-one caller posts e + 1 only when it has seen e's records).
        const uint32_t ps = s_ctl_snap[0], ds = s_ctl_snap[2];
        const bool fast = static_cast<int32_t>(ps - (e + 1)) > 0 && (e + 1) - ds < static_cast<uint32_t>(EV_D) && s_ctl_snap[3] == 0u;
        s_go_fast[(e + 1) & 1] = fast ? 1u : 0u;  // (a raised stop flag does not matter while posted commands are left)
        if (fast) s_pf_ok[(e + 2) & 1] = static_cast<int32_t>(ps - (e + 2)) > 0 ? 1u : 0u;
      }
      consumers_sync<NCONS>();
      if (tracing && si == last_si) tr[3] = peer_now_ns();
      if (threadIdx.x == 0) {
        s_rq[qtail % EV_PARK] = (e << 8) | static_cast<uint32_t>(r);
        __threadfence_block();
        s_rq_tail = qtail + 1;
      }
      ++qtail;
    }
  }
  // ---- leaving: tell the producer how far the ring was consumed and let producer / reducer / finaliser drain
  consumers_sync<NCONS>();
  if (threadIdx.x == 0) {
    s_consumed_lo = static_cast<uint32_t>(tt);
    s_consumed_hi = static_cast<uint32_t>(tt >> 32);
    __threadfence_block();
    s_exit = 1u;
  }
}

}  // namespace hsk
