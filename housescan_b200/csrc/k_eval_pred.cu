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

}  // namespace hsk

using namespace hsk;

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
