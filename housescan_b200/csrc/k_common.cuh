// Device helpers shared by the streaming kernels.
#pragma once
#include "hs_internal.cuh"

namespace hsk {

// signedDistanceToPlaneEq (Main.hs:1371-1372) in Float with GHC's evaluation order and no FMA contraction:
//   ((nx*px + ny*py) + nz*pz) - d
__device__ __forceinline__ float plane_dist(float nx, float ny, float nz, float d, float x, float y, float z) {
  return __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, x), __fmul_rn(ny, y)), __fmul_rn(nz, z)), d);
}

// First minimum of |distance| over the planes of a table (strict <: ties keep the lower index = minimumBy semantics); returns
// the index and leaves min |distance| in ab.  Only the index is carried through the chain (FSETP + SEL) and the running minimum
// is an FMNMX; callers that need the signed residual recompute it from the winner's plane.
//   KT > 0: plane count known at compile time (unrolled, constants straight from the constant bank); KT == 0: tbl.K at run time.
//   PAIRED (KT == 6 only): planes 2j and 2j+1 have exactly negated normals (a cuboid room's walls, Main.hs:1855-1874), so one dot
//   product t_j serves both: fl(-a x) == -fl(a x) and round-to-nearest is symmetric, hence the dot product of the negated normal
//   is exactly -t_j and its distance -t_j - d is the same Float the generic path computes.  15 of the 36 operations go away.
template <int KT, bool PAIRED>
__device__ __forceinline__ int nearest_plane(const PlaneTable& t, float x, float y, float z, float& ab) {
  int kb = 0;
  if (KT == 6 && PAIRED) {
    float a[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float tj = __fadd_rn(__fadd_rn(__fmul_rn(t.pl[2 * j][0], x), __fmul_rn(t.pl[2 * j][1], y)), __fmul_rn(t.pl[2 * j][2], z));
      a[2 * j] = fabsf(__fsub_rn(tj, t.pl[2 * j][3]));
      a[2 * j + 1] = fabsf(__fsub_rn(-tj, t.pl[2 * j + 1][3]));
    }
    ab = a[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) { kb = (a[k] < ab) ? k : kb; ab = fminf(ab, a[k]); }
  } else if (KT > 0) {
    ab = fabsf(plane_dist(t.pl[0][0], t.pl[0][1], t.pl[0][2], t.pl[0][3], x, y, z));
#pragma unroll
    for (int k = 1; k < KT; ++k) {
      const float ak = fabsf(plane_dist(t.pl[k][0], t.pl[k][1], t.pl[k][2], t.pl[k][3], x, y, z));
      kb = (ak < ab) ? k : kb;
      ab = fminf(ab, ak);
    }
  } else {
    ab = fabsf(plane_dist(t.pl[0][0], t.pl[0][1], t.pl[0][2], t.pl[0][3], x, y, z));
    for (int k = 1; k < t.K; ++k) {
      const float ak = fabsf(plane_dist(t.pl[k][0], t.pl[k][1], t.pl[k][2], t.pl[k][3], x, y, z));
      kb = (ak < ab) ? k : kb;
      ab = fminf(ab, ak);
    }
  }
  return kb;
}

// Correctly rounded Float division by a divisor that is the same for every point, without the MUFU.RCP + range-check sequence
// that `/` compiles to (~12 instructions and a slow-path branch per quotient).  y must be RN(1/b) (computed once on the host with
// an IEEE division).  q0 = RN(a y) is within 2 ulp of a/b; one residual step makes it faithful; by Markstein's theorem
// (y = RN(1/b), q faithful, r = a - b q exact in an FMA  =>  RN(q + r y) = RN(a/b)) the second step is the IEEE quotient.
// Valid while nothing under/overflows, which holds for pixel coordinates and depths.
__device__ __forceinline__ float div_rn_by(float a, float b, float y) {
  const float q0 = __fmul_rn(a, y);
  const float q1 = __fmaf_rn(__fmaf_rn(-b, q0, a), y, q0);
  return __fmaf_rn(__fmaf_rn(-b, q1, a), y, q1);
}
// Same for integer-valued a in [0, 2^24) and b in {10, 20}: one residual step already gives the IEEE quotient for every such a
// (checked exhaustively: tests/test_host_logic.py::test_constant_division_sequence).
__device__ __forceinline__ float div_rn_small(float a, float b, float y) {
  const float q0 = __fmul_rn(a, y);
  return __fmaf_rn(__fmaf_rn(-b, q0, a), y, q0);
}

// one thread's 4 consecutive AoS points = 3 x float4 (48 B, 16 B aligned)
struct Pts4 {
  float x[4], y[4], z[4];
};
__device__ __forceinline__ Pts4 load_group(const float* __restrict__ xyz, int64_t g) {
  const float4* q = reinterpret_cast<const float4*>(xyz) + 3 * g;
  const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  Pts4 p;
  p.x[0] = a.x; p.y[0] = a.y; p.z[0] = a.z;
  p.x[1] = a.w; p.y[1] = b.x; p.z[1] = b.y;
  p.x[2] = b.z; p.y[2] = b.w; p.z[2] = c.x;
  p.x[3] = c.y; p.y[3] = c.z; p.z[3] = c.w;
  return p;
}
__device__ __forceinline__ void store_group(float* __restrict__ xyz, int64_t g, const Pts4& p) {
  float4* q = reinterpret_cast<float4*>(xyz) + 3 * g;
  q[0] = make_float4(p.x[0], p.y[0], p.z[0], p.x[1]);
  q[1] = make_float4(p.y[1], p.z[1], p.x[2], p.y[2]);
  q[2] = make_float4(p.z[2], p.x[3], p.y[3], p.z[3]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction of N per-thread doubles; thread c < N writes the block total of component c to dst[c].
template <int N>
__device__ __forceinline__ void block_sum_store(double (&v)[N], double* dst, double* smem /* [HS_TPB/32][N] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    const double s = warp_sum(v[c]);
    if (lane == 0) smem[warp * N + c] = s;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < HS_TPB / 32; ++w) s += smem[w * N + threadIdx.x];
    dst[threadIdx.x] = s;
  }
  __syncthreads();
}

// "last block done" ticket: returns true in every thread of the block that arrives last.
__device__ __forceinline__ bool last_block_arrives(unsigned int* ticket, unsigned int nblocks) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == nblocks - 1);
    if (is_last) *ticket = 0u;  // re-arm for the next launch on this stream
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// Exclusive scan of ntiles unsigned counts in place by ONE block; returns the grand total in every thread.  Used by the
// order-preserving compactions (V.filter semantics).  Every thread owns one contiguous chunk of tiles: it sums the chunk,
// the 256 chunk sums are scanned across the block, and the chunk is rewritten with its running prefix — two sweeps of
// independent loads per thread instead of ntiles / 256 block-wide synchronised rounds.
__device__ __forceinline__ unsigned int block_scan_tiles_exclusive(unsigned int* tile_off, int64_t ntiles) {
  __shared__ unsigned int wsum[HS_TPB / 32];
  const int64_t chunk = (ntiles + HS_TPB - 1) / HS_TPB;
  const int64_t t0 = min(static_cast<int64_t>(threadIdx.x) * chunk, ntiles), t1 = min(t0 + chunk, ntiles);
  unsigned int mine = 0;
  for (int64_t t = t0; t < t1; ++t) mine += __ldcg(tile_off + t);
  unsigned int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += u;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  unsigned int run = incl - mine, total = 0;
#pragma unroll
  for (int w = 0; w < HS_TPB / 32; ++w) {
    if (w < (threadIdx.x >> 5)) run += wsum[w];
    total += wsum[w];
  }
  for (int64_t t = t0; t < t1; ++t) {
    const unsigned int v = __ldcg(tile_off + t);
    tile_off[t] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// Sum of per-block partial records by the last block: partials[b * N + c] for b < nblocks -> dst[c].  All HS_TPB threads load
// (slice s = tid / N takes blocks s, s + S, ...), then the slices are added in slice order: a fixed tree for a fixed grid.
// MAXC >= 0 marks one component that is reduced with max instead of + (non-negative values).
template <int N, int MAXC = -1>
__device__ __forceinline__ void last_block_sum(const double* __restrict__ partials, unsigned int nblocks, double* __restrict__ dst, double* smem /* [HS_TPB] */) {
  constexpr int S = HS_TPB / N;  // slices
  const int c = threadIdx.x % N, sl = threadIdx.x / N;
  double acc = 0.0;
  if (sl < S)
    for (unsigned int b = sl; b < nblocks; b += S) {
      const double v = __ldcg(partials + static_cast<int64_t>(b) * N + c);
      acc = (c == MAXC) ? fmax(acc, v) : acc + v;
    }
  smem[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < N) {
    double t = 0.0;
    for (int q = 0; q < S; ++q) { const double v = smem[q * N + threadIdx.x]; t = (static_cast<int>(threadIdx.x) == MAXC) ? fmax(t, v) : t + v; }
    dst[threadIdx.x] = t;
  }
  __syncthreads();
}

// Decoupled look-back for single-pass order-preserving compactions.  Claims (tiles) are handed out in order through a counter;
// claim t publishes its element count as soon as it knows it (status AGGREGATE), then one warp looks back over the predecessors
// 32 at a time, adding aggregates until it meets one that already published its INCLUSIVE prefix, and publishes its own.
// A predecessor was claimed earlier, is therefore running, and publishes its aggregate before it waits for anything: no
// deadlock for any grid size.  Status and value share one 64-bit word (one store, one load: nothing else to order); the state
// array must be zero before the launch.  Call with all 32 lanes of ONE warp; returns the exclusive prefix of claim t.
#define HS_LB_AGG (1ull << 62)
#define HS_LB_INC (2ull << 62)
#define HS_LB_MASK (3ull << 62)
__device__ __forceinline__ unsigned long long tile_lookback(volatile unsigned long long* vstate, int64_t t, unsigned long long count) {
  const int lane = threadIdx.x & 31;
  unsigned long long prefix = 0;
  if (t > 0) {
    if (lane == 0) vstate[t] = HS_LB_AGG | count;
    int64_t look = t - 1;
    for (;;) {
      const int64_t idx = look - lane;
      unsigned long long sv = HS_LB_INC;  // before claim 0: "inclusive prefix 0"
      if (idx >= 0) { do { sv = vstate[idx]; } while ((sv & HS_LB_MASK) == 0); }
      const unsigned int inc = __ballot_sync(0xffffffffu, (sv & HS_LB_MASK) == HS_LB_INC);
      const int first = inc ? __ffs(inc) - 1 : 32;  // nearest predecessor that already knows its inclusive prefix
      unsigned long long pv = (lane <= first) ? (sv & ~HS_LB_MASK) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
      prefix += pv;
      if (inc) break;
      look -= 32;
    }
  }
  if (lane == 0) vstate[t] = HS_LB_INC | (prefix + count);
  return prefix;
}

// exclusive prefix of a per-thread count inside the block (raster order of threads); smem wsum[HS_TPB/32]
__device__ __forceinline__ unsigned int block_exclusive_prefix(unsigned int c, unsigned int* wsum) {
  unsigned int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += u;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  unsigned int pos = incl - c;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) pos += wsum[w];
  return pos;
}

}  // namespace hsk
