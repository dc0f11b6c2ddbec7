// Cuboid objective / gradient sums per room — the throughput kernel (A6, 12 B/point, HBM-bound target), scalar-predicated form.
//
// Same record and the same bit-exact plane assignment as k_rooms_cuboid_sums<AccExact> (k_planes.cu).  What differs is the
// instruction mix, chosen from the measured sm_100 dispatch model (profiles/r01_ubench3_dispatch.log): one warp instruction
// issues per cycle per SMSP; scalar FP32 (FMUL/FADD/FFMA, also predicated) costs 1 FMA-pipe cycle, every compare/select
// (FSETP/FSEL/FSET) 2 ALU-pipe cycles, and the two pipes overlap.  So the nearest-wall selection is done with PREDICATES:
//   * per axis j: t = ((nx*x + ny*y) + nz*z), sp = t - d+, sm = t + d- (Float, no contraction, exactly signedDistanceToPlaneEq
//     Main.hs:1371-1372 for the + wall and minus that of the - wall), P = |sm| < |sp|, s_j = P ? sm : sp, pf_j = P ? 1 : 0;
//   * axis predicates E0/E1/E2 from three more compares (strict '<': ties keep the lower wall index, as minimumBy does);
//   * every accumulation is ONE predicated FADD/FFMA on a Float chain (@E_j B[j][c] += s_j * p_c ...): no one-hot selects.
// ~62 issue slots per point (44 FP, 14 ALU, 0.75 LDS.128) against ~90 dispatch cycles of the packed f32x2 form it replaces.
// Point tiles are streamed global -> shared by a producer warp with 1-D bulk async copies (cp.async.bulk, the TMA engine)
// through an mbarrier ring; per-thread Float chains (64 points) are flushed into per-thread Double accumulators in shared
// memory; block partials are summed by the last block in block order (deterministic for a fixed grid).
#include <cstring>

#include "k_common.cuh"
#include "k_ring.cuh"
#include "k_eval_point.cuh"
#include "k_eval_group_gen.cuh"
#include "k_peer.cuh"

namespace hsk {

constexpr int EP_FLUSH_POINTS = 128;  // Float chain length (points per thread between flushes into the Double accumulators)

struct PredTable {
  int32_t nrooms;
  int32_t pad;
  int64_t off[HS_MAX_ROOMS + 1];
  float n[HS_MAX_ROOMS][3][3];  // normal of the + wall of each axis
  float dp[HS_MAX_ROOMS][3];    // d of the + wall
  float dm[HS_MAX_ROOMS][3];    // d of the - wall (whose normal is -n)
};

// Per-thread Double accumulators live in shared memory (slot c of thread t at acc[c * NCONS + t]: conflict-free, private,
// no synchronisation) so the register file is left to the Float chains; they are touched once per 64 points.
template <int NCONS>
__device__ __forceinline__ void flush_chains_p(ChainsP& c, double* acc, int npoints) {
  double* a = acc + threadIdx.x;
  a[0] += static_cast<double>(c.f);
  const double C1 = c.C1, C2 = c.C2;
  const double Caxis[3] = {static_cast<double>(npoints) - C1 - C2, C1, C2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double T = c.T[j], M = c.M[j], Cm = c.Cm[j];
    a[(1 + 2 * j) * NCONS] += T - M;  // sum of r over the + wall (r = s)
    a[(2 + 2 * j) * NCONS] -= M;      // sum of r over the - wall (r = -s)
    a[(16 + 2 * j) * NCONS] += Caxis[j] - Cm;
    a[(17 + 2 * j) * NCONS] += Cm;
#pragma unroll
    for (int q = 0; q < 3; ++q) a[(7 + 3 * j + q) * NCONS] += static_cast<double>(c.B[j][q]);
  }
  c.clear();
}

// scalar edge path (ragged points at room / chunk borders, partial tiles): exact Double products, same record
template <int NCONS>
__device__ __noinline__ void add_point_exact_p(double* acc, const float* pl /* 6 x 4 */, float x, float y, float z) {
  float rb = plane_dist(pl[0], pl[1], pl[2], pl[3], x, y, z);
  float ab = fabsf(rb);
  int kb = 0;
#pragma unroll
  for (int k = 1; k < 6; ++k) {
    const float rk = plane_dist(pl[4 * k], pl[4 * k + 1], pl[4 * k + 2], pl[4 * k + 3], x, y, z);
    const float ak = fabsf(rk);
    const bool lt = ak < ab;
    ab = lt ? ak : ab; rb = lt ? rk : rb; kb = lt ? k : kb;
  }
  const double rd = rb, s = (kb & 1) ? -rd : rd;
  const int j = kb >> 1;
  double* a = acc + threadIdx.x;
  a[0] = fma(rd, rd, a[0]);
  a[(1 + kb) * NCONS] += rd;
  a[(16 + kb) * NCONS] += 1.0;
  a[(7 + 3 * j + 0) * NCONS] = fma(s, static_cast<double>(x), a[(7 + 3 * j + 0) * NCONS]);
  a[(7 + 3 * j + 1) * NCONS] = fma(s, static_cast<double>(y), a[(7 + 3 * j + 1) * NCONS]);
  a[(7 + 3 * j + 2) * NCONS] = fma(s, static_cast<double>(z), a[(7 + 3 * j + 2) * NCONS]);
}

__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

#define DBG_EVT(k) do { if (dbg && threadIdx.x == 0 && (r - rfirst) < 2) dbg[4 * gridDim.x + 8 * blockIdx.x + 4 * (r - rfirst) + (k)] = globaltimer_ns(); } while (0)

template <int NCONS>
__device__ __forceinline__ void consumers_sync_p() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// block b owns groups [b*gpb, (b+1)*gpb); per overlapping room: stream whole-group tiles through the ring.
// GPT = 48-byte groups per thread per tile (tile = GPT * NCONS groups); FORM 0 = scalar products, 1 = packed products
template <int NCONS, int STAGES, int BPS, int GPT, int FORM>
__global__ void __launch_bounds__(NCONS + 32, BPS)
k_rooms_cuboid_sums_pred(const float* __restrict__ xyz, int64_t n, const __grid_constant__ PredTable tbl, int64_t gpb,
                         double* __restrict__ partials, int* __restrict__ meta, unsigned int* ticket, double* __restrict__ out,
                         unsigned long long* __restrict__ dbg) {
  constexpr int TILE_GROUPS = GPT * NCONS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STAGES) * TILE_GROUPS * 48);
  uint64_t* empty = full + STAGES;
  double* acc = reinterpret_cast<double*>(empty + STAGES);           // [HS_NACC][NCONS] per-thread Double accumulators
  double* red = acc + static_cast<size_t>(HS_NACC) * NCONS;           // [NCONS/32][HS_NACC]
  float* spl = reinterpret_cast<float*>(red + (NCONS / 32) * HS_NACC);  // 6 x 4 planes of the current room (edge path)
  __shared__ bool is_last;

  const int nrooms = tbl.nrooms;
  const int64_t G = (n + 3) >> 2;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * gpb;
  const int64_t g1 = min(g0 + gpb, G);
  const int64_t p0 = g0 * 4, p1 = min(g1 * 4, n);
  int rfirst = -1, rlast = -2;
  for (int r = 0; r < nrooms; ++r)
    if (tbl.off[r] < p1 && tbl.off[r + 1] > p0) { if (rfirst < 0) rfirst = r; rlast = r; }

  if (threadIdx.x == 0) {
    if (dbg) dbg[4 * blockIdx.x] = globaltimer_ns();
    meta[blockIdx.x] = rfirst;
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x >= NCONS) {
    // ---------------- producer warp: one lane issues the bulk copies, same tile order as the consumers
    if (threadIdx.x == NCONS) {
      int64_t tt = 0;
      for (int r = rfirst; r <= rlast; ++r) {
        const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
        const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
        for (int64_t tg = gl; tg < gh; tg += TILE_GROUPS, ++tt) {
          const int s = static_cast<int>(tt % STAGES);
          if (tt >= STAGES) mbar_wait(empty + s, static_cast<uint32_t>(((tt / STAGES) - 1) & 1));
          const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(TILE_GROUPS), gh - tg) * 48);
          mbar_expect_tx(full + s, bytes);
          bulk_g2s(tiles + static_cast<size_t>(s) * TILE_GROUPS * 3, reinterpret_cast<const float4*>(xyz) + 3 * tg, bytes, full + s);
        }
      }
    }
    return;
  }

  // ---------------- consumers
  int64_t tt = 0;
  for (int r = rfirst; r <= rlast; ++r) {
    const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
    RoomK R;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      R.n[j][0] = tbl.n[r][j][0]; R.n[j][1] = tbl.n[r][j][1]; R.n[j][2] = tbl.n[r][j][2];
      R.dp[j] = tbl.dp[r][j]; R.dm[j] = tbl.dm[r][j];
    }
    if (threadIdx.x < 6) {
      const int j = threadIdx.x >> 1;
      const bool minus = threadIdx.x & 1;
      spl[4 * threadIdx.x + 0] = minus ? -tbl.n[r][j][0] : tbl.n[r][j][0];
      spl[4 * threadIdx.x + 1] = minus ? -tbl.n[r][j][1] : tbl.n[r][j][1];
      spl[4 * threadIdx.x + 2] = minus ? -tbl.n[r][j][2] : tbl.n[r][j][2];
      spl[4 * threadIdx.x + 3] = minus ? tbl.dm[r][j] : tbl.dp[r][j];
    }
#pragma unroll
    for (int i = 0; i < HS_NACC; ++i) acc[i * NCONS + threadIdx.x] = 0.0;
    consumers_sync_p<NCONS>();
    ChainsP ch;
    ch.clear();
    const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
    if (gl <= gh) {
      const int64_t head_end = gl * 4, tail_begin = gh * 4;
      const int64_t nh = head_end - lo, ntail = hi - tail_begin;
      if (threadIdx.x < nh) { const int64_t i = lo + threadIdx.x; add_point_exact_p<NCONS>(acc, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      else if (threadIdx.x >= 32 && threadIdx.x - 32 < ntail) { const int64_t i = tail_begin + threadIdx.x - 32; add_point_exact_p<NCONS>(acc, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      // full tiles: the hot loop.  Ring position is carried as (stage, parity); `tt` only keeps producer and consumers on
      // the same global tile count across rooms.
      DBG_EVT(0);
      const int64_t ngroups = gh - gl;
      const int nfull = static_cast<int>(ngroups / TILE_GROUPS);
      const int rem_groups = static_cast<int>(ngroups - static_cast<int64_t>(nfull) * TILE_GROUPS);
      constexpr uint32_t TILE_BYTES = TILE_GROUPS * 48;
      constexpr int FLUSH_TILES = EP_FLUSH_POINTS / (4 * GPT);
      uint32_t stage = static_cast<uint32_t>(tt % STAGES);
      uint32_t parity = static_cast<uint32_t>((tt / STAGES) & 1);
      const uint32_t tiles_s = smem_u32(tiles) + threadIdx.x * 48, full_s = smem_u32(full), empty_s = smem_u32(empty);
      const RoomK2 R2 = make_room_k2(R);
      int since_flush = 0;
      for (int t = 0; t < nfull; ++t) {
        mbar_wait_s(full_s + 8 * stage, parity);
        // this thread's GPT groups of the tile (group g*NCONS + tid): 4 consecutive points = 3 x LDS.128 each
        // (48 B lane stride: conflict-free quarter-warps)
        const uint32_t base = tiles_s + stage * TILE_BYTES;
        if (FORM == 0) {
          float4 q[GPT][3];
#pragma unroll
          for (int g = 0; g < GPT; ++g) {
            q[g][0] = lds_v4(base + g * NCONS * 48); q[g][1] = lds_v4(base + g * NCONS * 48 + 16); q[g][2] = lds_v4(base + g * NCONS * 48 + 32);
          }
          mbar_arrive_s(empty_s + 8 * stage);  // the values are in registers: hand the slot back before the math
#pragma unroll
          for (int g = 0; g < GPT; ++g) {
            add_point_pred(ch, R, q[g][0].x, q[g][0].y, q[g][0].z);
            add_point_pred(ch, R, q[g][0].w, q[g][1].x, q[g][1].y);
            add_point_pred(ch, R, q[g][1].z, q[g][1].w, q[g][2].x);
            add_point_pred(ch, R, q[g][2].y, q[g][2].z, q[g][2].w);
          }
        } else {
          unsigned long long w[GPT][6];
#pragma unroll
          for (int g = 0; g < GPT; ++g) {
            lds_v2b64(base + g * NCONS * 48, w[g][0], w[g][1]); lds_v2b64(base + g * NCONS * 48 + 16, w[g][2], w[g][3]);
            lds_v2b64(base + g * NCONS * 48 + 32, w[g][4], w[g][5]);
          }
          mbar_arrive_s(empty_s + 8 * stage);
#pragma unroll
          for (int g = 0; g < GPT; ++g) add_group_v1(ch, R2, w[g][0], w[g][1], w[g][2], w[g][3], w[g][4], w[g][5]);
        }
        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
        if (++since_flush == FLUSH_TILES) { flush_chains_p<NCONS>(ch, acc, 4 * GPT * FLUSH_TILES); since_flush = 0; }
      }
      flush_chains_p<NCONS>(ch, acc, 4 * GPT * since_flush);
      tt += nfull;
      DBG_EVT(1);
      if (rem_groups) {  // partial last tile of the room segment: exact scalar path for the in-range groups
        mbar_wait_s(full_s + 8 * stage, parity);
        const float* tile = reinterpret_cast<const float*>(tiles) + static_cast<size_t>(stage) * TILE_GROUPS * 12;
        float v[GPT][12];
#pragma unroll
        for (int g = 0; g < GPT; ++g)
#pragma unroll
          for (int e = 0; e < 12; ++e) v[g][e] = tile[12 * (g * NCONS + threadIdx.x) + e];  // out-of-range slots: stale but valid shared memory
        mbar_arrive_s(empty_s + 8 * stage);
#pragma unroll
        for (int g = 0; g < GPT; ++g)
          if (g * NCONS + static_cast<int>(threadIdx.x) < rem_groups)
            for (int e = 0; e < 4; ++e) add_point_exact_p<NCONS>(acc, spl, v[g][3 * e], v[g][3 * e + 1], v[g][3 * e + 2]);
        ++tt;
      }
    } else {
      const int64_t i = lo + threadIdx.x;
      if (i < hi) add_point_exact_p<NCONS>(acc, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
    DBG_EVT(2);
    // consumer-only deterministic block reduction
    {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
      for (int c = 0; c < HS_NACC; ++c) {
        const double sum = warp_sum(acc[c * NCONS + threadIdx.x]);
        if (lane == 0) red[warp * HS_NACC + c] = sum;
      }
      consumers_sync_p<NCONS>();
      if (threadIdx.x < HS_NACC) {
        double sum = 0;
#pragma unroll
        for (int w = 0; w < NCONS / 32; ++w) sum += red[w * HS_NACC + threadIdx.x];
        partials[(static_cast<int64_t>(blockIdx.x) * nrooms + (r - rfirst)) * HS_NACC + threadIdx.x] = sum;
      }
      consumers_sync_p<NCONS>();
    }
    DBG_EVT(3);
  }

  // ---------------- last block sums the partials per room in block order
  if (dbg && threadIdx.x == 0) dbg[4 * blockIdx.x + 1] = globaltimer_ns();
  __threadfence();
  consumers_sync_p<NCONS>();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  consumers_sync_p<NCONS>();
  if (dbg && threadIdx.x == 0) { dbg[4 * blockIdx.x + 2] = globaltimer_ns(); dbg[4 * blockIdx.x + 3] = is_last; }
  if (!is_last) return;
  __threadfence();
  // The other SMs are idle from here on, so this must be short: all partials are fetched with independent loads into shared
  // memory first (one round trip instead of one per block), then every output is summed from shared memory in block order.
  const int64_t ppb = gpb * 4;
  int* smeta = reinterpret_cast<int*>(red);
  const int nblocks = static_cast<int>(gridDim.x);
  constexpr int SMETA_CAP = (NCONS / 32) * HS_NACC * 2, STAGE_CAP = HS_NACC * NCONS;
  // per room: first overlapping block, number of overlapping blocks, prefix of those counts (one 64-bit division pair per room)
  __shared__ int s_blo[HS_MAX_ROOMS], s_nbr[HS_MAX_ROOMS], s_base[HS_MAX_ROOMS + 1];
  if (threadIdx.x < nrooms) {
    const int r = threadIdx.x;
    const bool nonempty = tbl.off[r + 1] > tbl.off[r];
    const int64_t b_lo = nonempty ? tbl.off[r] / ppb : 0;
    s_blo[r] = static_cast<int>(b_lo);
    s_nbr[r] = nonempty ? static_cast<int>((tbl.off[r + 1] - 1) / ppb - b_lo) + 1 : 0;
  }
  consumers_sync_p<NCONS>();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int r = 0; r < nrooms; ++r) { s_base[r] = run; run += s_nbr[r]; }
    s_base[nrooms] = run;
  }
  consumers_sync_p<NCONS>();
  const int total_slots = s_base[nrooms];  // sum over rooms of the number of blocks overlapping the room
  if (nblocks <= SMETA_CAP && total_slots * HS_NACC <= STAGE_CAP) {
    for (int b = threadIdx.x; b < nblocks; b += NCONS) smeta[b] = __ldcg(meta + b);
    consumers_sync_p<NCONS>();
    const int total = total_slots * HS_NACC;
    constexpr int U = 4;
    for (int i0 = 0; i0 < total; i0 += NCONS * U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NCONS + threadIdx.x;
        v[u] = 0.0;
        if (i < total) {
          const int q = i / HS_NACC, c = i - q * HS_NACC;
          int r = 0;
          while (q >= s_base[r + 1]) ++r;  // room of slot q (empty rooms have equal prefixes and are skipped)
          const int bb = s_blo[r] + (q - s_base[r]);
          v[u] = __ldcg(partials + (static_cast<int64_t>(bb) * nrooms + (r - smeta[bb])) * HS_NACC + c);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NCONS + threadIdx.x;
        if (i < total) acc[i] = v[u];
      }
    }
    consumers_sync_p<NCONS>();
    for (int o = threadIdx.x; o < nrooms * HS_REC; o += NCONS) {
      const int r = o / HS_REC, c = o % HS_REC;
      double sum = 0.0;
      if (c < HS_NACC)
        for (int k = 0; k < s_nbr[r]; ++k) sum += acc[(s_base[r] + k) * HS_NACC + c];
      out[o] = sum;
    }
    if (dbg && threadIdx.x == 0) dbg[4 * blockIdx.x + 2] = globaltimer_ns();
    return;
  }
  for (int o = threadIdx.x; o < nrooms * HS_REC; o += NCONS) {  // general fallback: same sums, one dependent load per block
    const int r = o / HS_REC, c = o % HS_REC;
    double sum = 0.0;
    if (c < HS_NACC && tbl.off[r + 1] > tbl.off[r]) {
      const int64_t b_lo = tbl.off[r] / ppb, b_hi = (tbl.off[r + 1] - 1) / ppb;
      for (int64_t b = b_lo; b <= b_hi; ++b) {
        const int slot = r - __ldcg(meta + b);
        sum += __ldcg(partials + (b * nrooms + slot) * HS_NACC + c);
      }
    }
    out[o] = sum;
  }
}


// ======================================================================================================================
// Warp-accumulator form (mode key 3 = 5).  Same per-point PTX block (add_point_pred) and the same ring, but the per-thread
// Double accumulators in shared memory (176 B per thread) are gone: Float chains are summed across the warp with shuffles
// (five more roundings, like five more chain steps) and only the 22 warp totals are added in Double into a [warps][22]
// table.  That frees ~100 KB of shared memory and, with scalar chains (21 registers instead of 42 for packed pairs), the
// block holds 24 consumer warps (6 per scheduler instead of 4) at 80 registers — the kernel is bound by instruction
// issue, and issue efficiency is what more resident warps buy.  Ragged points go through the same per-point block.
// ======================================================================================================================
constexpr int ES_FLUSH_TILES = 32;  // 128 points per thread between flushes (Float counts stay exact: <= 4096 per warp)

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// chains of the whole warp -> the warp's 22 Double accumulators.  npts = points this lane added since the last flush.
__device__ __forceinline__ void flush_warp(ChainsP& c, int npts, double* wacc /* [HS_NACC] of this warp */, float* wtmp /* [24] of this warp */) {
  const int lane = threadIdx.x & 31;
  float v[23];
  v[0] = c.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    v[1 + j] = c.T[j]; v[4 + j] = c.M[j]; v[16 + j] = c.Cm[j];
#pragma unroll
    for (int q = 0; q < 3; ++q) v[7 + 3 * j + q] = c.B[j][q];
  }
  v[19] = c.C1; v[20] = c.C2; v[21] = static_cast<float>(npts); v[22] = 0.f;
#pragma unroll
  for (int i = 0; i < 22; ++i) v[i] = warp_sum_f(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 22; ++i) wtmp[i] = v[i];
  }
  __syncwarp();
  if (lane < HS_NACC) {
    // record component `lane` from the warp totals (same algebra as flush_chains_p)
    double add;
    if (lane == 0) add = wtmp[0];
    else if (lane <= 6) { const int j = (lane - 1) >> 1; const double T = wtmp[1 + j], M = wtmp[4 + j]; add = (lane & 1) ? T - M : -M; }
    else if (lane <= 15) add = wtmp[lane];  // B[j][q] sits at 7 + 3j + q in both layouts
    else {
      const int j = (lane - 16) >> 1;
      const double C1 = wtmp[19], C2 = wtmp[20], N = wtmp[21], Cm = wtmp[16 + j];
      const double Cax = j == 0 ? N - C1 - C2 : (j == 1 ? C1 : C2);
      add = (lane & 1) ? Cm : Cax - Cm;
    }
    wacc[lane] += add;
  }
  __syncwarp();
  c.clear();
}

template <int PT>
__device__ __forceinline__ void point(ChainsP& c, const RoomK& R, float x, float y, float z) {
  if (PT == 2) add_point_pred2(c, R, x, y, z);
  else add_point_pred(c, R, x, y, z);
}

// tools only (mode key 5): per-block timestamps of the DBG instantiation land here; the product instantiation (DBG = false) is
// compiled without a single extra instruction
__device__ unsigned long long* g_warp_dbg = nullptr;
#define WDBG(k) do { if (DBG && threadIdx.x == 0) g_warp_dbg[8 * blockIdx.x + (k)] = globaltimer_ns(); } while (0)

template <int NCONS, int STAGES, int GPT, int PT, int WH, bool DBG = false>
__global__ void __launch_bounds__(NCONS + 32, 1)
k_rooms_cuboid_sums_warp(const float* __restrict__ xyz, int64_t n, const __grid_constant__ PredTable tbl, int64_t gpb,
                         double* __restrict__ partials, int* __restrict__ meta, unsigned int* ticket, double* __restrict__ out,
                         const __grid_constant__ PeerExchange px) {
  constexpr int NW = NCONS / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STAGES) * GPT * NCONS * 48);
  uint64_t* empty = full + STAGES;
  double* wacc = reinterpret_cast<double*>(empty + STAGES);       // [NW][HS_NACC]
  float* wtmp = reinterpret_cast<float*>(wacc + NW * HS_NACC);    // [NW][24]
  double* stage_d = reinterpret_cast<double*>(tiles);             // final reduction staging: the ring is idle by then
  __shared__ bool is_last;

  const int nrooms = tbl.nrooms;
  const int64_t G = (n + 3) >> 2;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * gpb;
  const int64_t g1 = min(g0 + gpb, G);
  const int64_t p0 = g0 * 4, p1 = min(g1 * 4, n);
  int rfirst = -1, rlast = -2;
  for (int r = 0; r < nrooms; ++r)
    if (tbl.off[r] < p1 && tbl.off[r + 1] > p0) { if (rfirst < 0) rfirst = r; rlast = r; }

  if (threadIdx.x == 0) {
    meta[blockIdx.x] = rfirst;
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x >= NCONS) {
    // ---------------- producer warp: one lane issues the bulk copies, same tile order as the consumers
    if (threadIdx.x == NCONS) {
      int64_t tt = 0;
      for (int r = rfirst; r <= rlast; ++r) {
        const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
        const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
        for (int64_t tg = gl; tg < gh; tg += GPT * NCONS, ++tt) {
          const int s = static_cast<int>(tt % STAGES);
          if (tt >= STAGES) {
            const uint32_t par = static_cast<uint32_t>(((tt / STAGES) - 1) & 1);
            uint32_t ok;
            for (;;) {  // poll, then sleep: a bare spin takes issue slots from the consumer warps on this scheduler
              asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                           : "=r"(ok) : "r"(smem_u32(empty + s)), "r"(par) : "memory");
              if (ok) break;
              __nanosleep(100);
            }
          }
          const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(GPT * NCONS), gh - tg) * 48);
          mbar_expect_tx(full + s, bytes);
          bulk_g2s(tiles + static_cast<size_t>(s) * GPT * NCONS * 3, reinterpret_cast<const float4*>(xyz) + 3 * tg, bytes, full + s);
        }
      }
    }
    return;
  }

  // ---------------- consumers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* my_wacc = wacc + warp * HS_NACC;
  float* my_wtmp = wtmp + warp * 24;
  int64_t tt = 0;
  WDBG(0);  // block start
  for (int r = rfirst; r <= rlast; ++r) {
    const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
    RoomK R;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      R.n[j][0] = tbl.n[r][j][0]; R.n[j][1] = tbl.n[r][j][1]; R.n[j][2] = tbl.n[r][j][2];
      R.dp[j] = tbl.dp[r][j]; R.dm[j] = tbl.dm[r][j];
    }
    if (lane < HS_NACC) my_wacc[lane] = 0.0;
    __syncwarp();
    ChainsP ch;
    ch.clear();
    int npts = 0;
    const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
    if (gl <= gh) {
      // ragged head / tail points (at most 3 each): the same per-point block, one point per thread
      const int64_t head_end = gl * 4, tail_begin = gh * 4;
      const int64_t nh = head_end - lo, ntail = hi - tail_begin;
      if (threadIdx.x < nh) { const int64_t i = lo + threadIdx.x; point<PT>(ch, R, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); ++npts; }
      else if (threadIdx.x >= 32 && threadIdx.x - 32 < ntail) { const int64_t i = tail_begin + threadIdx.x - 32; point<PT>(ch, R, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); ++npts; }
      const int64_t ngroups = gh - gl;
      const int nfull = static_cast<int>(ngroups / (GPT * NCONS));
      const int rem_groups = static_cast<int>(ngroups - static_cast<int64_t>(nfull) * (GPT * NCONS));
      constexpr uint32_t TILE_BYTES = GPT * NCONS * 48;
      uint32_t stage = static_cast<uint32_t>(tt % STAGES);
      uint32_t parity = static_cast<uint32_t>((tt / STAGES) & 1);
      const uint32_t tiles_s = smem_u32(tiles) + threadIdx.x * 48, full_s = smem_u32(full), empty_s = smem_u32(empty);
      int since_flush = 0;
      for (int t = 0; t < nfull; ++t) {
        if (WH == 2) mbar_wait_s_test(full_s + 8 * stage, parity); else if (WH == 1) mbar_wait_s(full_s + 8 * stage, parity); else mbar_wait(full + stage, parity);
        if (DBG && t == 0 && r == rfirst) WDBG(1);  // first tile has landed
        // this thread's GPT groups of the tile (group g * NCONS + tid): 4 consecutive points = 3 x LDS.128 each (48 B lane stride:
        // conflict-free quarter-warps)
        const uint32_t base = tiles_s + stage * TILE_BYTES;
        float4 q[GPT][3];
#pragma unroll
        for (int g = 0; g < GPT; ++g) { q[g][0] = lds_v4(base + g * NCONS * 48); q[g][1] = lds_v4(base + g * NCONS * 48 + 16); q[g][2] = lds_v4(base + g * NCONS * 48 + 32); }
        mbar_arrive_s(empty_s + 8 * stage);  // the values are in registers: hand the slot back before the math
#pragma unroll
        for (int g = 0; g < GPT; ++g) {
          point<PT>(ch, R, q[g][0].x, q[g][0].y, q[g][0].z);
          point<PT>(ch, R, q[g][0].w, q[g][1].x, q[g][1].y);
          point<PT>(ch, R, q[g][1].z, q[g][1].w, q[g][2].x);
          point<PT>(ch, R, q[g][2].y, q[g][2].z, q[g][2].w);
        }
        npts += 4 * GPT;
        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
        if (++since_flush == ES_FLUSH_TILES / GPT) { flush_warp(ch, npts, my_wacc, my_wtmp); npts = 0; since_flush = 0; }
      }
      tt += nfull;
      if (rem_groups) {  // partial last tile of the room segment: the threads whose group is in range
        if (WH == 2) mbar_wait_s_test(full_s + 8 * stage, parity); else if (WH == 1) mbar_wait_s(full_s + 8 * stage, parity); else mbar_wait(full + stage, parity);
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
            point<PT>(ch, R, q[g][0].x, q[g][0].y, q[g][0].z);
            point<PT>(ch, R, q[g][0].w, q[g][1].x, q[g][1].y);
            point<PT>(ch, R, q[g][1].z, q[g][1].w, q[g][2].x);
            point<PT>(ch, R, q[g][2].y, q[g][2].z, q[g][2].w);
            npts += 4;
          }
        ++tt;
      }
    } else {
      const int64_t i = lo + threadIdx.x;
      if (i < hi) { point<PT>(ch, R, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); ++npts; }
    }
    flush_warp(ch, npts, my_wacc, my_wtmp);
    // the block's record of this room: warp tables summed in warp order
    consumers_sync_p<NCONS>();
    if (threadIdx.x < HS_NACC) {
      double sum = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) sum += wacc[w * HS_NACC + threadIdx.x];
      partials[(static_cast<int64_t>(blockIdx.x) * nrooms + (r - rfirst)) * HS_NACC + threadIdx.x] = sum;
    }
    consumers_sync_p<NCONS>();
  }

  // ---------------- last block sums the partials per room in block order (parallel fetch, then fixed-order sums)
  WDBG(2);  // streaming and per-room block sums done
  __threadfence();
  consumers_sync_p<NCONS>();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  consumers_sync_p<NCONS>();
  WDBG(3);  // ticket taken
  if (DBG && threadIdx.x == 0) g_warp_dbg[8 * blockIdx.x + 4] = is_last;
  if (!is_last) return;
  __threadfence();
  const int64_t ppb = gpb * 4;
  __shared__ int s_blo[HS_MAX_ROOMS], s_nbr[HS_MAX_ROOMS], s_base[HS_MAX_ROOMS + 1];
  __shared__ int smeta[1024];
  const int nblocks = static_cast<int>(gridDim.x);
  constexpr int STAGE_CAP = STAGES * GPT * NCONS * 48 / 8;
  if (threadIdx.x < nrooms) {
    const int r = threadIdx.x;
    const bool nonempty = tbl.off[r + 1] > tbl.off[r];
    const int64_t b_lo = nonempty ? tbl.off[r] / ppb : 0;
    s_blo[r] = static_cast<int>(b_lo);
    s_nbr[r] = nonempty ? static_cast<int>((tbl.off[r + 1] - 1) / ppb - b_lo) + 1 : 0;
  }
  consumers_sync_p<NCONS>();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int r = 0; r < nrooms; ++r) { s_base[r] = run; run += s_nbr[r]; }
    s_base[nrooms] = run;
  }
  consumers_sync_p<NCONS>();
  const int total_slots = s_base[nrooms];
  if (nblocks <= 1024 && total_slots * HS_NACC <= STAGE_CAP) {
    for (int b = threadIdx.x; b < nblocks; b += NCONS) smeta[b] = __ldcg(meta + b);
    consumers_sync_p<NCONS>();
    const int total = total_slots * HS_NACC;
    // every thread issues ALL its loads of a batch before it stores the first one: a load followed by its own store per loop
    // iteration serialises the L2 round trips (~0.7 us each, nine of them per thread at 148 blocks x 12 rooms: the per-block
    // timeline showed this loop as 6 of the ~9 us the last block spent here)
    constexpr int U = 10;
    for (int i0 = 0; i0 < total; i0 += NCONS * U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NCONS + static_cast<int>(threadIdx.x);
        v[u] = 0.0;
        if (i < total) {
          const int q = i / HS_NACC, c = i - q * HS_NACC;
          int r = 0;
          while (q >= s_base[r + 1]) ++r;  // room of slot q (empty rooms have equal prefixes and are skipped)
          const int bb = s_blo[r] + (q - s_base[r]);
          v[u] = __ldcg(partials + (static_cast<int64_t>(bb) * nrooms + (r - smeta[bb])) * HS_NACC + c);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NCONS + static_cast<int>(threadIdx.x);
        if (i < total) stage_d[i] = v[u];
      }
    }
    consumers_sync_p<NCONS>();
    for (int o = threadIdx.x; o < nrooms * HS_REC; o += NCONS) {
      const int r = o / HS_REC, c = o % HS_REC;
      double sum = 0.0;
      if (c < HS_NACC)
        for (int k = 0; k < s_nbr[r]; ++k) sum += stage_d[(s_base[r] + k) * HS_NACC + c];
      out[o] = sum;
    }
  } else {
    for (int o = threadIdx.x; o < nrooms * HS_REC; o += NCONS) {  // general fallback: one dependent load per block
      const int r = o / HS_REC, c = o % HS_REC;
      double sum = 0.0;
      if (c < HS_NACC && tbl.off[r + 1] > tbl.off[r]) {
        const int64_t b_lo = tbl.off[r] / ppb, b_hi = (tbl.off[r + 1] - 1) / ppb;
        for (int64_t b = b_lo; b <= b_hi; ++b) {
          const int slot = r - __ldcg(meta + b);
          sum += __ldcg(partials + (b * nrooms + slot) * HS_NACC + c);
        }
      }
      out[o] = sum;
    }
  }
  WDBG(5);  // final reduction written
  // ---------------- multi-GPU: the records of all ranks are summed over peer memory before the kernel ends (k_peer.cuh)
  if (px.world > 1) {
    consumers_sync_p<NCONS>();
    peer_allreduce(px, out, nrooms * HS_REC, static_cast<int>(threadIdx.x), NCONS, [] { consumers_sync_p<NCONS>(); });
  }
  WDBG(6);
}


// ======================================================================================================================
// Warp-private rings (mode key 3 = 6).  Every consumer warp streams ITS OWN tiles (GPT x 32 groups, 48 B each) through its own
// D-slot mbarrier ring: lane 0 issues the 1-D bulk copy of the tile D-1 steps ahead right after the warp has read the slot
// it reuses.  No producer warp, no "empty" barriers, no block-wide coupling: warps never wait for each other inside a room
// segment, so nobody spins while somebody else computes (in the shared ring the fast warps' polling takes issue slots from
// the slow ones, and the kernel is issue-bound).  Tile k of a room segment belongs to warp k % NW.
// ======================================================================================================================
template <int NW, int GPT, int D>
__global__ void __launch_bounds__(NW * 32, 1)
k_rooms_cuboid_sums_wring(const float* __restrict__ xyz, int64_t n, const __grid_constant__ PredTable tbl, int64_t gpb,
                          double* __restrict__ partials, int* __restrict__ meta, unsigned int* ticket, double* __restrict__ out) {
  constexpr int NT = NW * 32;
  constexpr int TG = 32 * GPT;                 // groups per tile
  constexpr uint32_t TILE_BYTES = TG * 48;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* rings = smem_raw;                                                           // [NW][D][TILE_BYTES]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(NW) * D * TILE_BYTES);  // [NW][D]
  double* wacc = reinterpret_cast<double*>(full + NW * D);                                   // [NW][HS_NACC]
  float* wtmp = reinterpret_cast<float*>(wacc + NW * HS_NACC);                               // [NW][24]
  double* stage_d = reinterpret_cast<double*>(rings);                                        // final reduction staging (rings idle)
  __shared__ bool is_last;

  const int nrooms = tbl.nrooms;
  const int64_t G = (n + 3) >> 2;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * gpb;
  const int64_t g1 = min(g0 + gpb, G);
  const int64_t p0 = g0 * 4, p1 = min(g1 * 4, n);
  int rfirst = -1, rlast = -2;
  for (int r = 0; r < nrooms; ++r)
    if (tbl.off[r] < p1 && tbl.off[r + 1] > p0) { if (rfirst < 0) rfirst = r; rlast = r; }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* my_full = full + warp * D;
  if (threadIdx.x == 0) meta[blockIdx.x] = rfirst;
  if (lane == 0) {
    for (int s = 0; s < D; ++s) mbar_init(my_full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double* my_wacc = wacc + warp * HS_NACC;
  float* my_wtmp = wtmp + warp * 24;
  const uint32_t ring_s = smem_u32(rings) + warp * (D * TILE_BYTES), full_s = smem_u32(my_full);
  uint32_t uses = 0;  // tiles this warp has pushed through its ring so far (slot = uses % D, parity = (uses / D) & 1)
  for (int r = rfirst; r <= rlast; ++r) {
    const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
    RoomK R;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      R.n[j][0] = tbl.n[r][j][0]; R.n[j][1] = tbl.n[r][j][1]; R.n[j][2] = tbl.n[r][j][2];
      R.dp[j] = tbl.dp[r][j]; R.dm[j] = tbl.dm[r][j];
    }
    if (lane < HS_NACC) my_wacc[lane] = 0.0;
    __syncwarp();
    ChainsP ch;
    ch.clear();
    int npts = 0;
    const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
    if (gl <= gh) {
      // ragged head / tail points (at most 3 each): the same per-point block, one point per thread
      const int64_t head_end = gl * 4, tail_begin = gh * 4;
      const int64_t nh = head_end - lo, ntail = hi - tail_begin;
      if (threadIdx.x < nh) { const int64_t i = lo + threadIdx.x; add_point_pred(ch, R, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); ++npts; }
      else if (threadIdx.x >= 32 && threadIdx.x - 32 < ntail) { const int64_t i = tail_begin + threadIdx.x - 32; add_point_pred(ch, R, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); ++npts; }
      const int64_t ngroups = gh - gl;
      const int64_t ntiles = (ngroups + TG - 1) / TG;
      const int64_t mine = ntiles > warp ? (ntiles - warp + NW - 1) / NW : 0;  // tiles warp, warp + NW, ...
      const float4* src = reinterpret_cast<const float4*>(xyz) + 3 * gl;
      auto issue = [&](int64_t i) {  // lane 0: bulk copy of my i-th tile of this segment into slot (uses0 + i) % D
        const int64_t tg = (warp + i * NW) * TG;
        const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(TG), ngroups - tg) * 48);
        const uint32_t slot = (uses + static_cast<uint32_t>(i)) % D;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_s + 8 * slot), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ring_s + slot * TILE_BYTES), "l"(src + 3 * tg), "r"(bytes), "r"(full_s + 8 * slot) : "memory");
      };
      // `uses` is only advanced after the loop: issue() addresses slots relative to it
      if (lane == 0)
        for (int64_t i = 0; i < D - 1 && i < mine; ++i) issue(i);
      int since_flush = 0;
      for (int64_t i = 0; i < mine; ++i) {
        __syncwarp();  // every lane has read the slot that tile i + D - 1 overwrites (it held tile i - 1)
        if (lane == 0 && i + D - 1 < mine) issue(i + D - 1);
        const uint32_t u = uses + static_cast<uint32_t>(i);
        const uint32_t slot = u % D;
        mbar_wait_s_spin(full_s + 8 * slot, (u / D) & 1u);
        const int64_t tg = (warp + i * NW) * TG;
        const int64_t left = ngroups - tg;  // groups in this tile (>= 1)
        const uint32_t base = ring_s + slot * TILE_BYTES + lane * 48;
        float4 q[GPT][3];
#pragma unroll
        for (int g = 0; g < GPT; ++g) {  // lanes past the end of a partial tile read stale (valid) shared memory and skip the math
          q[g][0] = lds_v4(base + g * 32 * 48); q[g][1] = lds_v4(base + g * 32 * 48 + 16); q[g][2] = lds_v4(base + g * 32 * 48 + 32);
        }
#pragma unroll
        for (int g = 0; g < GPT; ++g) {
          if (left >= TG || g * 32 + lane < left) {
            add_point_pred(ch, R, q[g][0].x, q[g][0].y, q[g][0].z);
            add_point_pred(ch, R, q[g][0].w, q[g][1].x, q[g][1].y);
            add_point_pred(ch, R, q[g][1].z, q[g][1].w, q[g][2].x);
            add_point_pred(ch, R, q[g][2].y, q[g][2].z, q[g][2].w);
            npts += 4;
          }
        }
        if (++since_flush == ES_FLUSH_TILES / GPT) { flush_warp(ch, npts, my_wacc, my_wtmp); npts = 0; since_flush = 0; }
      }
      uses += static_cast<uint32_t>(mine);
    } else {
      const int64_t i = lo + threadIdx.x;
      if (i < hi) { add_point_pred(ch, R, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); ++npts; }
    }
    flush_warp(ch, npts, my_wacc, my_wtmp);
    // the block's record of this room: warp tables summed in warp order
    __syncthreads();
    if (threadIdx.x < HS_NACC) {
      double sum = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) sum += wacc[w * HS_NACC + threadIdx.x];
      partials[(static_cast<int64_t>(blockIdx.x) * nrooms + (r - rfirst)) * HS_NACC + threadIdx.x] = sum;
    }
    __syncthreads();
  }

  // ---------------- last block sums the partials per room in block order (parallel fetch, then fixed-order sums)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int64_t ppb = gpb * 4;
  __shared__ int s_blo[HS_MAX_ROOMS], s_nbr[HS_MAX_ROOMS], s_base[HS_MAX_ROOMS + 1];
  __shared__ int smeta[1024];
  const int nblocks = static_cast<int>(gridDim.x);
  constexpr int STAGE_CAP = NW * D * static_cast<int>(TILE_BYTES) / 8;
  if (threadIdx.x < nrooms) {
    const int r = threadIdx.x;
    const bool nonempty = tbl.off[r + 1] > tbl.off[r];
    const int64_t b_lo = nonempty ? tbl.off[r] / ppb : 0;
    s_blo[r] = static_cast<int>(b_lo);
    s_nbr[r] = nonempty ? static_cast<int>((tbl.off[r + 1] - 1) / ppb - b_lo) + 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int r = 0; r < nrooms; ++r) { s_base[r] = run; run += s_nbr[r]; }
    s_base[nrooms] = run;
  }
  __syncthreads();
  const int total_slots = s_base[nrooms];
  if (nblocks <= 1024 && total_slots * HS_NACC <= STAGE_CAP) {
    for (int b = threadIdx.x; b < nblocks; b += NT) smeta[b] = __ldcg(meta + b);
    __syncthreads();
    const int total = total_slots * HS_NACC;
    for (int i = threadIdx.x; i < total; i += NT) {
      const int q = i / HS_NACC, c = i - q * HS_NACC;
      int r = 0;
      while (q >= s_base[r + 1]) ++r;
      const int bb = s_blo[r] + (q - s_base[r]);
      stage_d[i] = __ldcg(partials + (static_cast<int64_t>(bb) * nrooms + (r - smeta[bb])) * HS_NACC + c);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < nrooms * HS_REC; o += NT) {
      const int r = o / HS_REC, c = o % HS_REC;
      double sum = 0.0;
      if (c < HS_NACC)
        for (int k = 0; k < s_nbr[r]; ++k) sum += stage_d[(s_base[r] + k) * HS_NACC + c];
      out[o] = sum;
    }
    return;
  }
  for (int o = threadIdx.x; o < nrooms * HS_REC; o += NT) {  // general fallback: one dependent load per block
    const int r = o / HS_REC, c = o % HS_REC;
    double sum = 0.0;
    if (c < HS_NACC && tbl.off[r + 1] > tbl.off[r]) {
      const int64_t b_lo = tbl.off[r] / ppb, b_hi = (tbl.off[r + 1] - 1) / ppb;
      for (int64_t b = b_lo; b <= b_hi; ++b) {
        const int slot = r - __ldcg(meta + b);
        sum += __ldcg(partials + (b * nrooms + slot) * HS_NACC + c);
      }
    }
    out[o] = sum;
  }
}

// standalone exchange for the kernels that do not carry it in their tail (one block)
__global__ void __launch_bounds__(HS_TPB) k_peer_allreduce(double* buf, int count, const __grid_constant__ PeerExchange px) {
  peer_allreduce(px, buf, count, static_cast<int>(threadIdx.x), HS_TPB, [] { __syncthreads(); });
}

}  // namespace hsk

using namespace hsk;

int32_t launch_peer_allreduce(hs_ctx* ctx, double* d_buf, int count, const PeerExchange& px) {
  k_peer_allreduce<<<1, HS_TPB, 0, ctx->stream>>>(d_buf, count, px);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

template <int NCONS, int STAGES, int BPS, int GPT, int FORM>
static int32_t launch_pred_t(hs_ctx* ctx, const float* xyz, int64_t n, const PredTable& tbl, double* d_rec_out) {
  const int64_t G = (n + 3) >> 2;
  int64_t nb = static_cast<int64_t>(ctx->sm_count) * BPS;
  const int64_t cap = (G + NCONS - 1) / NCONS;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  const int64_t gpb = (G + nb - 1) / nb > 0 ? (G + nb - 1) / nb : 1;
  const size_t need = static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double) + static_cast<size_t>(nb) * sizeof(int) + 64;
  if (int32_t rc = hs_ensure_scratch(ctx, need)) return rc;
  double* partials = reinterpret_cast<double*>(ctx->d_scratch);
  int* meta = reinterpret_cast<int*>(ctx->d_scratch + static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double));
  const size_t smem = static_cast<size_t>(STAGES) * GPT * NCONS * 48 + 2 * STAGES * 8 + static_cast<size_t>(HS_NACC) * NCONS * 8 +
                      static_cast<size_t>(NCONS / 32) * HS_NACC * 8 + 6 * 4 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_rooms_cuboid_sums_pred<NCONS, STAGES, BPS, GPT, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  k_rooms_cuboid_sums_pred<NCONS, STAGES, BPS, GPT, FORM><<<static_cast<int>(nb), NCONS + 32, smem, ctx->stream>>>(xyz, n, tbl, gpb, partials, meta, ctx->d_ticket, d_rec_out, ctx->d_dbg);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

template <int NCONS, int STAGES, int GPT = 1, int PT = 1, int WH = 0, bool DBG = false>
static int32_t launch_warp_t(hs_ctx* ctx, const float* xyz, int64_t n, const PredTable& tbl, double* d_rec_out) {
  const int64_t G = (n + 3) >> 2;
  int64_t nb = ctx->sm_count;
  const int64_t cap = (G + NCONS - 1) / NCONS;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  const int64_t gpb = (G + nb - 1) / nb > 0 ? (G + nb - 1) / nb : 1;
  const size_t need = static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double) + static_cast<size_t>(nb) * sizeof(int) + 64;
  if (int32_t rc = hs_ensure_scratch(ctx, need)) return rc;
  double* partials = reinterpret_cast<double*>(ctx->d_scratch);
  int* meta = reinterpret_cast<int*>(ctx->d_scratch + static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double));
  const size_t smem = static_cast<size_t>(STAGES) * GPT * NCONS * 48 + 2 * STAGES * 8 + static_cast<size_t>(NCONS / 32) * HS_NACC * 8 +
                      static_cast<size_t>(NCONS / 32) * 24 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_rooms_cuboid_sums_warp<NCONS, STAGES, GPT, PT, WH, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    if (DBG) HS_CUDA_TRY(ctx, cudaMemcpyToSymbol(g_warp_dbg, &ctx->d_dbg, sizeof(ctx->d_dbg)));
    attr_set = true;
  }
  PeerExchange px = {};
  if (ctx->px_next) { px = ctx->px; px.epoch = ++ctx->px.epoch; ctx->px_next = false; }  // fused exchange: consumed by this launch
  k_rooms_cuboid_sums_warp<NCONS, STAGES, GPT, PT, WH, DBG><<<static_cast<int>(nb), NCONS + 32, smem, ctx->stream>>>(xyz, n, tbl, gpb, partials, meta, ctx->d_ticket, d_rec_out, px);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

template <int NW, int GPT, int D>
static int32_t launch_wring_t(hs_ctx* ctx, const float* xyz, int64_t n, const PredTable& tbl, double* d_rec_out) {
  const int64_t G = (n + 3) >> 2;
  int64_t nb = ctx->sm_count;
  const int64_t cap = (G + NW * 32 - 1) / (NW * 32);
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  const int64_t gpb = (G + nb - 1) / nb > 0 ? (G + nb - 1) / nb : 1;
  const size_t need = static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double) + static_cast<size_t>(nb) * sizeof(int) + 64;
  if (int32_t rc = hs_ensure_scratch(ctx, need)) return rc;
  double* partials = reinterpret_cast<double*>(ctx->d_scratch);
  int* meta = reinterpret_cast<int*>(ctx->d_scratch + static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double));
  const size_t smem = static_cast<size_t>(NW) * D * (32 * GPT * 48) + static_cast<size_t>(NW) * D * 8 + static_cast<size_t>(NW) * HS_NACC * 8 +
                      static_cast<size_t>(NW) * 24 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_rooms_cuboid_sums_wring<NW, GPT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  k_rooms_cuboid_sums_wring<NW, GPT, D><<<static_cast<int>(nb), NW * 32, smem, ctx->stream>>>(xyz, n, tbl, gpb, partials, meta, ctx->d_ticket, d_rec_out);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

// caller guarantees the planes are paired (cuboid rooms)
int32_t launch_rooms_cuboid_sums_pred(hs_ctx* ctx, const float* xyz, int64_t n, const RoomTable& rt, double* d_rec_out) {
  PredTable t;
  memset(&t, 0, sizeof t);
  t.nrooms = rt.nrooms;
  for (int r = 0; r <= rt.nrooms; ++r) t.off[r] = rt.off[r];
  for (int r = 0; r < rt.nrooms; ++r)
    for (int j = 0; j < 3; ++j) {
      for (int c = 0; c < 3; ++c) t.n[r][j][c] = rt.pl[r][2 * j][c];
      t.dp[r][j] = rt.pl[r][2 * j][3];
      t.dm[r][j] = rt.pl[r][2 * j + 1][3];
    }
  // product default: 12 consumer warps + the producer warp, 3 ring stages of 72 KB, 16 points per thread per tile, FMNMX point
  // block (profiles/r01_sweep8_tile_size.log: 220.6 us per 100 M points = 0.83 of the measured HBM peak)
  if (ctx->modes[HS_MODE_EVAL_VARIANT] == 0) {
    if (ctx->modes[HS_MODE_DEBUG_TIMES] && ctx->d_dbg) return launch_warp_t<384, 3, 4, 2, 0, true>(ctx, xyz, n, t, d_rec_out);  // tools: timestamps
    return launch_warp_t<384, 3, 4, 2>(ctx, xyz, n, t, d_rec_out);
  }
  if (ctx->modes[HS_MODE_EVAL_VARIANT] == 6) {  // warp-private rings: consumers = warps * 100 + groups per thread * 10 + ring depth
    switch (ctx->modes[HS_MODE_EVAL_CONSUMERS]) {
      case 1614: return launch_wring_t<16, 1, 4>(ctx, xyz, n, t, d_rec_out);
      case 1618: return launch_wring_t<16, 1, 8>(ctx, xyz, n, t, d_rec_out);
      case 1623: return launch_wring_t<16, 2, 3>(ctx, xyz, n, t, d_rec_out);
      case 2414: return launch_wring_t<24, 1, 4>(ctx, xyz, n, t, d_rec_out);
      case 2416: return launch_wring_t<24, 1, 6>(ctx, xyz, n, t, d_rec_out);
      case 3214: return launch_wring_t<32, 1, 4>(ctx, xyz, n, t, d_rec_out);
      default: return launch_wring_t<16, 2, 4>(ctx, xyz, n, t, d_rec_out);
    }
  }
  if (ctx->modes[HS_MODE_EVAL_VARIANT] == 5) {  // warp-accumulator form
    switch (ctx->modes[HS_MODE_EVAL_CONSUMERS]) {
      case 512: return launch_warp_t<512, 4>(ctx, xyz, n, t, d_rec_out);
      case 7682: return launch_warp_t<768, 4, 1, 2>(ctx, xyz, n, t, d_rec_out);   // FMNMX point block
      case 5122: return launch_warp_t<512, 4, 2, 1>(ctx, xyz, n, t, d_rec_out);   // two groups per thread per tile
      case 5123: return launch_warp_t<512, 4, 2, 2>(ctx, xyz, n, t, d_rec_out);   // both
      case 7683: return launch_warp_t<768, 3, 2, 2>(ctx, xyz, n, t, d_rec_out);
      case 5124: return launch_warp_t<512, 4, 2, 2, 1>(ctx, xyz, n, t, d_rec_out);  // parked consumer waits
      case 38435: return launch_warp_t<384, 3, 4, 2, 2>(ctx, xyz, n, t, d_rec_out);  // test_wait polling
      case 5133: return launch_warp_t<512, 3, 3, 2>(ctx, xyz, n, t, d_rec_out);     // three groups per thread per tile
      case 3843: return launch_warp_t<384, 4, 3, 2>(ctx, xyz, n, t, d_rec_out);
      case 6402: return launch_warp_t<640, 3, 2, 2>(ctx, xyz, n, t, d_rec_out);
      case 38434: return launch_warp_t<384, 3, 4, 2>(ctx, xyz, n, t, d_rec_out);
      case 51224: return launch_warp_t<512, 2, 4, 2>(ctx, xyz, n, t, d_rec_out);
      case 48033: return launch_warp_t<480, 3, 3, 2>(ctx, xyz, n, t, d_rec_out);
      case 25636: return launch_warp_t<256, 3, 6, 2>(ctx, xyz, n, t, d_rec_out);
      case 640: return launch_warp_t<640, 4>(ctx, xyz, n, t, d_rec_out);
      case 896: return launch_warp_t<896, 4>(ctx, xyz, n, t, d_rec_out);
      case 992: return launch_warp_t<992, 4>(ctx, xyz, n, t, d_rec_out);
      default: return launch_warp_t<768, 4>(ctx, xyz, n, t, d_rec_out);
    }
  }
  switch (ctx->modes[HS_MODE_EVAL_CONSUMERS]) {  // tuning variants (tools/prof_eval.py --cons N); 0 is the product default
    case 1: return launch_pred_t<480, 4, 1, 1, 0>(ctx, xyz, n, t, d_rec_out);
    case 2: return launch_pred_t<480, 4, 1, 1, 1>(ctx, xyz, n, t, d_rec_out);
    case 3: return launch_pred_t<480, 3, 1, 2, 0>(ctx, xyz, n, t, d_rec_out);
    case 4: return launch_pred_t<480, 3, 1, 2, 1>(ctx, xyz, n, t, d_rec_out);
    case 5: return launch_pred_t<384, 4, 1, 2, 0>(ctx, xyz, n, t, d_rec_out);
    case 6: return launch_pred_t<384, 4, 1, 2, 1>(ctx, xyz, n, t, d_rec_out);
    case 7: return launch_pred_t<608, 2, 1, 2, 0>(ctx, xyz, n, t, d_rec_out);
    case 8: return launch_pred_t<608, 2, 1, 2, 1>(ctx, xyz, n, t, d_rec_out);
    default: return launch_pred_t<480, 3, 1, 2, 0>(ctx, xyz, n, t, d_rec_out);
  }
}
