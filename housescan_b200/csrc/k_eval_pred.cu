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

namespace hsk {

constexpr int EP_FLUSH_TILES = 16;  // Float chain length: 16 tiles x 4 points per thread

struct PredTable {
  int32_t nrooms;
  int32_t pad;
  int64_t off[HS_MAX_ROOMS + 1];
  float n[HS_MAX_ROOMS][3][3];  // normal of the + wall of each axis
  float dp[HS_MAX_ROOMS][3];    // d of the + wall
  float dm[HS_MAX_ROOMS][3];    // d of the - wall (whose normal is -n)
};

struct RoomK {  // one room's constants in registers (uniform across the block)
  float n[3][3], dp[3], dm[3];
};

// Float chains of one thread since the last flush
struct ChainsP {
  float f, T[3], M[3], B[3][3], C1, C2, Cm[3];
  __device__ __forceinline__ void clear() {
    f = C1 = C2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T[j] = M[j] = Cm[j] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) B[j][c] = 0.f;
    }
  }
};

// One point.  Written as one PTX block so that every accumulation is a predicated FP instruction (the C++ front end is free
// to turn `if (E) acc += v` into selects, which cost an ALU-pipe slot each).  ptxas schedules across consecutive blocks.
__device__ __forceinline__ void add_point_pred(ChainsP& c, const RoomK& R, float x, float y, float z) {
  asm("{\n"
      ".reg .pred P, Q1, E0, E1, E2;\n"
      ".reg .f32 a, b, t, sp, sm, asp, asm_, s0, s1, s2, p0, p1, p2, a0, a1, a2, a01;\n"
      // axis 0
      "mul.rn.f32 a, %24, %21;\n mul.rn.f32 b, %25, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %26, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 sp, t, %33;\n add.rn.f32 sm, t, %36;\n abs.f32 asp, sp;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n selp.f32 s0, sm, sp, P;\n selp.f32 p0, 0f3F800000, 0f00000000, P;\n"
      // axis 1
      "mul.rn.f32 a, %27, %21;\n mul.rn.f32 b, %28, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %29, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 sp, t, %34;\n add.rn.f32 sm, t, %37;\n abs.f32 asp, sp;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n selp.f32 s1, sm, sp, P;\n selp.f32 p1, 0f3F800000, 0f00000000, P;\n"
      // axis 2
      "mul.rn.f32 a, %30, %21;\n mul.rn.f32 b, %31, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %32, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 sp, t, %35;\n add.rn.f32 sm, t, %38;\n abs.f32 asp, sp;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n selp.f32 s2, sm, sp, P;\n selp.f32 p2, 0f3F800000, 0f00000000, P;\n"
      // nearest axis, sequential first-minimum semantics (NaN compares false and keeps the earlier wall)
      "abs.f32 a0, s0;\n abs.f32 a1, s1;\n abs.f32 a2, s2;\n"
      "setp.lt.f32 Q1, a1, a0;\n selp.f32 a01, a1, a0, Q1;\n setp.lt.f32 E2, a2, a01;\n"
      "setp.lt.and.f32 E1, a1, a0, !E2;\n setp.geu.and.f32 E0, a1, a0, !E2;\n"
      // predicated accumulation
      "@E0 fma.rn.f32 %0, s0, s0, %0;\n @E0 add.rn.f32 %1, %1, s0;\n @E0 fma.rn.f32 %4, s0, p0, %4;\n @E0 add.rn.f32 %18, %18, p0;\n"
      "@E0 fma.rn.f32 %7, s0, %21, %7;\n @E0 fma.rn.f32 %8, s0, %22, %8;\n @E0 fma.rn.f32 %9, s0, %23, %9;\n"
      "@E1 fma.rn.f32 %0, s1, s1, %0;\n @E1 add.rn.f32 %2, %2, s1;\n @E1 fma.rn.f32 %5, s1, p1, %5;\n @E1 add.rn.f32 %19, %19, p1;\n"
      "@E1 fma.rn.f32 %10, s1, %21, %10;\n @E1 fma.rn.f32 %11, s1, %22, %11;\n @E1 fma.rn.f32 %12, s1, %23, %12;\n @E1 add.rn.f32 %16, %16, 0f3F800000;\n"
      "@E2 fma.rn.f32 %0, s2, s2, %0;\n @E2 add.rn.f32 %3, %3, s2;\n @E2 fma.rn.f32 %6, s2, p2, %6;\n @E2 add.rn.f32 %20, %20, p2;\n"
      "@E2 fma.rn.f32 %13, s2, %21, %13;\n @E2 fma.rn.f32 %14, s2, %22, %14;\n @E2 fma.rn.f32 %15, s2, %23, %15;\n @E2 add.rn.f32 %17, %17, 0f3F800000;\n"
      "}\n"
      : "+f"(c.f), "+f"(c.T[0]), "+f"(c.T[1]), "+f"(c.T[2]), "+f"(c.M[0]), "+f"(c.M[1]), "+f"(c.M[2]),              // 0..6
        "+f"(c.B[0][0]), "+f"(c.B[0][1]), "+f"(c.B[0][2]), "+f"(c.B[1][0]), "+f"(c.B[1][1]), "+f"(c.B[1][2]),       // 7..12
        "+f"(c.B[2][0]), "+f"(c.B[2][1]), "+f"(c.B[2][2]), "+f"(c.C1), "+f"(c.C2),                                  // 13..17
        "+f"(c.Cm[0]), "+f"(c.Cm[1]), "+f"(c.Cm[2])                                                                 // 18..20
      : "f"(x), "f"(y), "f"(z),                                                                                      // 21..23
        "f"(R.n[0][0]), "f"(R.n[0][1]), "f"(R.n[0][2]), "f"(R.n[1][0]), "f"(R.n[1][1]), "f"(R.n[1][2]),             // 24..29
        "f"(R.n[2][0]), "f"(R.n[2][1]), "f"(R.n[2][2]),                                                              // 30..32
        "f"(R.dp[0]), "f"(R.dp[1]), "f"(R.dp[2]), "f"(R.dm[0]), "f"(R.dm[1]), "f"(R.dm[2]));                        // 33..38
}

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

template <int NCONS>
__device__ __forceinline__ void consumers_sync_p() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// block b owns groups [b*gpb, (b+1)*gpb); per overlapping room: stream whole-group tiles through the ring.
template <int NCONS, int STAGES, int BPS>
__global__ void __launch_bounds__(NCONS + 32, BPS)
k_rooms_cuboid_sums_pred(const float* __restrict__ xyz, int64_t n, const __grid_constant__ PredTable tbl, int64_t gpb,
                         double* __restrict__ partials, int* __restrict__ meta, unsigned int* ticket, double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STAGES) * NCONS * 48);
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
        for (int64_t tg = gl; tg < gh; tg += NCONS, ++tt) {
          const int s = static_cast<int>(tt % STAGES);
          if (tt >= STAGES) mbar_wait(empty + s, static_cast<uint32_t>(((tt / STAGES) - 1) & 1));
          const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(NCONS), gh - tg) * 48);
          mbar_expect_tx(full + s, bytes);
          bulk_g2s(tiles + static_cast<size_t>(s) * NCONS * 3, reinterpret_cast<const float4*>(xyz) + 3 * tg, bytes, full + s);
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
      const int64_t ngroups = gh - gl;
      const int nfull = static_cast<int>(ngroups / NCONS);
      const int rem_groups = static_cast<int>(ngroups - static_cast<int64_t>(nfull) * NCONS);
      constexpr uint32_t TILE_BYTES = NCONS * 48;
      uint32_t stage = static_cast<uint32_t>(tt % STAGES);
      uint32_t parity = static_cast<uint32_t>((tt / STAGES) & 1);
      const uint32_t tiles_s = smem_u32(tiles) + threadIdx.x * 48, full_s = smem_u32(full), empty_s = smem_u32(empty);
      int since_flush = 0;
      for (int t = 0; t < nfull; ++t) {
        mbar_wait_s(full_s + 8 * stage, parity);
        // this thread's group of the tile: 4 consecutive points = 3 x LDS.128 (48 B lane stride: conflict-free quarter-warps)
        const uint32_t base = tiles_s + stage * TILE_BYTES;
        const float4 q0 = lds_v4(base), q1 = lds_v4(base + 16), q2 = lds_v4(base + 32);
        mbar_arrive_s(empty_s + 8 * stage);  // the values are in registers: hand the slot back before the math
        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
        add_point_pred(ch, R, q0.x, q0.y, q0.z);
        add_point_pred(ch, R, q0.w, q1.x, q1.y);
        add_point_pred(ch, R, q1.z, q1.w, q2.x);
        add_point_pred(ch, R, q2.y, q2.z, q2.w);
        if (++since_flush == EP_FLUSH_TILES) { flush_chains_p<NCONS>(ch, acc, 4 * EP_FLUSH_TILES); since_flush = 0; }
      }
      flush_chains_p<NCONS>(ch, acc, 4 * since_flush);
      tt += nfull;
      if (rem_groups) {  // partial last tile of the room segment: exact scalar path for the in-range groups
        mbar_wait_s(full_s + 8 * stage, parity);
        const float* tile = reinterpret_cast<const float*>(tiles) + static_cast<size_t>(stage) * NCONS * 12;
        float v[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) v[e] = tile[12 * threadIdx.x + e];  // out-of-range slots read stale but valid shared memory
        mbar_arrive_s(empty_s + 8 * stage);
        if (static_cast<int>(threadIdx.x) < rem_groups)
          for (int e = 0; e < 4; ++e) add_point_exact_p<NCONS>(acc, spl, v[3 * e], v[3 * e + 1], v[3 * e + 2]);
        ++tt;
      }
    } else {
      const int64_t i = lo + threadIdx.x;
      if (i < hi) add_point_exact_p<NCONS>(acc, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
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
  }

  // ---------------- last block sums the partials per room in block order
  __threadfence();
  consumers_sync_p<NCONS>();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  consumers_sync_p<NCONS>();
  if (!is_last) return;
  __threadfence();
  const int64_t ppb = gpb * 4;
  for (int o = threadIdx.x; o < nrooms * HS_REC; o += NCONS) {
    const int r = o / HS_REC, c = o % HS_REC;
    double s = 0.0;
    if (c < HS_NACC && tbl.off[r + 1] > tbl.off[r]) {
      const int64_t b_lo = tbl.off[r] / ppb, b_hi = (tbl.off[r + 1] - 1) / ppb;
      for (int64_t b = b_lo; b <= b_hi; ++b) {
        const int slot = r - __ldcg(meta + b);
        s += __ldcg(partials + (b * nrooms + slot) * HS_NACC + c);
      }
    }
    out[o] = s;
  }
}

}  // namespace hsk

using namespace hsk;

template <int NCONS, int STAGES, int BPS>
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
  const size_t smem = static_cast<size_t>(STAGES) * NCONS * 48 + 2 * STAGES * 8 + static_cast<size_t>(HS_NACC) * NCONS * 8 +
                      static_cast<size_t>(NCONS / 32) * HS_NACC * 8 + 6 * 4 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_rooms_cuboid_sums_pred<NCONS, STAGES, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  k_rooms_cuboid_sums_pred<NCONS, STAGES, BPS><<<static_cast<int>(nb), NCONS + 32, smem, ctx->stream>>>(xyz, n, tbl, gpb, partials, meta, ctx->d_ticket, d_rec_out);
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
  switch (ctx->modes[HS_MODE_EVAL_CONSUMERS]) {
    case 1: return launch_pred_t<224, 4, 2>(ctx, xyz, n, t, d_rec_out);   // 2 CTAs/SM x (7 consumer warps + producer)
    case 2: return launch_pred_t<320, 3, 2>(ctx, xyz, n, t, d_rec_out);   // 2 CTAs/SM x (10 + 1)
    case 3: return launch_pred_t<608, 3, 1>(ctx, xyz, n, t, d_rec_out);   // 19 + 1 warps
    case 4: return launch_pred_t<736, 2, 1>(ctx, xyz, n, t, d_rec_out);   // 23 + 1 warps
    case 5: return launch_pred_t<352, 4, 1>(ctx, xyz, n, t, d_rec_out);   // 11 + 1 warps
    default: return launch_pred_t<480, 4, 1>(ctx, xyz, n, t, d_rec_out);  // 15 consumer warps + producer
  }
}
