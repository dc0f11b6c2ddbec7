// Order statistics and order-preserving filters.
//   kthSmallestBy / kthLargestBy (VectorUtil.hs:11-19): the k-th order statistic by a Float key (1-based k).
//   removeCeiling (Main.hs:2643-2664): yLimit = k-th largest y with k = n `quot` 5; keep y <= yLimit in input order.
// The order statistic is found by a 4-pass MSB radix select over order-preserving uint images of the Float keys
// (no sort, no copy of the cloud); the filter is a two-pass block-scan compaction.
#include <algorithm>

#include "k_common.cuh"

namespace hsk {

struct SelState {
  unsigned int prefix;
  unsigned int mask;
  unsigned long long k_rem;
  unsigned int hist[256];
};

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}

__global__ void k_sel_init(SelState* st, unsigned long long k) {
  if (threadIdx.x == 0) { st->prefix = 0; st->mask = 0; st->k_rem = k; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) st->hist[i] = 0;
}

template <bool LARGEST>
__global__ void __launch_bounds__(HS_TPB)
k_sel_pass(const float* __restrict__ xyz, int64_t n, int axis, int shift, SelState* st, unsigned int* ticket, float* __restrict__ out) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;  // HS_TPB == 256
  __syncthreads();
  const unsigned int prefix = st->prefix, mask = st->mask;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; i < n; i += stride) {
    const unsigned int u = f2ord(__ldg(xyz + 3 * i + axis));
    if ((u & mask) == prefix) atomicAdd(&sh[(u >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
  if (!last_block_arrives(ticket, gridDim.x)) return;
  if (threadIdx.x == 0) {
    unsigned long long k = st->k_rem, cum = 0;
    int digit = 0;
    for (int t = 0; t < 256; ++t) {
      const int b = LARGEST ? 255 - t : t;
      const unsigned long long c = __ldcg(&st->hist[b]);
      if (cum + c >= k) { digit = b; break; }
      cum += c;
    }
    st->k_rem = k - cum;
    st->prefix = prefix | (static_cast<unsigned int>(digit) << shift);
    st->mask = mask | (255u << shift);
    if (shift == 0) *out = ord2f(st->prefix);
  }
  __syncthreads();
  st->hist[threadIdx.x] = 0;
}

// ---- k-th over compacted keys (default) ------------------------------------------------------------------------------
// The cloud is AoS, so a pass over one coordinate still moves all 12 B/point.  The first pass therefore also writes the
// order-preserving uint image of the key as a dense 4 B/point array; the remaining passes read only that.  Three passes of
// 11 + 11 + 10 bits: 12 + 4 (write) + 4 + 4 = 24 B/point of traffic instead of 4 x 12 = 48.
#define SEL_BINS 2048
struct SelState2 {
  unsigned int prefix;
  unsigned int mask;
  unsigned long long k_rem;
  unsigned int hist[SEL_BINS];
};

__global__ void k_sel2_init(SelState2* st, unsigned long long k) {
  if (threadIdx.x == 0) { st->prefix = 0; st->mask = 0; st->k_rem = k; }
  for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) st->hist[i] = 0;
}

// merge the block histogram, and in the last block pick the digit that holds the k-th key (all 256 threads: 8 bins each)
template <bool LARGEST>
__device__ __forceinline__ void sel2_finish(unsigned int* sh, int shift, int bits, SelState2* st, unsigned int* ticket, float* __restrict__ out) {
  __shared__ unsigned int wsum[HS_TPB / 32];
  const int nbins = 1 << bits;
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += HS_TPB)
    if (sh[b]) atomicAdd(&st->hist[b], sh[b]);
  if (!out) return;  // sharded mode: the digit is picked on the host after the histograms of all ranks are summed
  if (!last_block_arrives(ticket, gridDim.x)) return;
  // scan order: descending bins for the k-th largest, ascending for the k-th smallest; thread t owns 8 consecutive bins of it
  unsigned int c[SEL_BINS / HS_TPB], mine = 0;
#pragma unroll
  for (int j = 0; j < SEL_BINS / HS_TPB; ++j) {
    const int pos = threadIdx.x * (SEL_BINS / HS_TPB) + j;
    const int b = LARGEST ? nbins - 1 - pos : pos;
    c[j] = (pos < nbins) ? __ldcg(&st->hist[b]) : 0u;
    mine += c[j];
  }
  const unsigned long long before = block_exclusive_prefix(mine, wsum);  // counts fit 32 bits per pass only if n < 2^32: checked by the launcher
  const unsigned long long k = st->k_rem;
  const unsigned int prefix = st->prefix, mask = st->mask;
  __syncthreads();
  if (before < k && k <= before + mine) {  // exactly one thread
    unsigned long long cum = before;
#pragma unroll
    for (int j = 0; j < SEL_BINS / HS_TPB; ++j) {
      if (cum + c[j] >= k) {
        const int pos = threadIdx.x * (SEL_BINS / HS_TPB) + j;
        const unsigned int digit = static_cast<unsigned int>(LARGEST ? nbins - 1 - pos : pos);
        st->k_rem = k - cum;
        st->prefix = prefix | (digit << shift);
        st->mask = mask | (static_cast<unsigned int>(nbins - 1) << shift);
        if (shift == 0) *out = ord2f(prefix | digit);
        break;
      }
      cum += c[j];
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < SEL_BINS; b += HS_TPB) st->hist[b] = 0;
}

// pass 1: key image + histogram of the top 11 bits
template <bool LARGEST>
__global__ void __launch_bounds__(HS_TPB)
k_sel2_first(const float* __restrict__ xyz, int64_t n, int axis, unsigned int* __restrict__ keys, SelState2* st, unsigned int* ticket, float* __restrict__ out) {
  __shared__ unsigned int sh[SEL_BINS];
  for (int b = threadIdx.x; b < SEL_BINS; b += HS_TPB) sh[b] = 0;
  __syncthreads();
  const int64_t gfull = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; g < gfull; g += stride) {
    const Pts4 p = load_group(xyz, g);
    uint4 u;
    u.x = f2ord(axis == 0 ? p.x[0] : (axis == 1 ? p.y[0] : p.z[0]));
    u.y = f2ord(axis == 0 ? p.x[1] : (axis == 1 ? p.y[1] : p.z[1]));
    u.z = f2ord(axis == 0 ? p.x[2] : (axis == 1 ? p.y[2] : p.z[2]));
    u.w = f2ord(axis == 0 ? p.x[3] : (axis == 1 ? p.y[3] : p.z[3]));
    reinterpret_cast<uint4*>(keys)[g] = u;
    atomicAdd(&sh[u.x >> 21], 1u); atomicAdd(&sh[u.y >> 21], 1u); atomicAdd(&sh[u.z >> 21], 1u); atomicAdd(&sh[u.w >> 21], 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = gfull * 4 + threadIdx.x;
    const unsigned int u = f2ord(xyz[3 * i + axis]);
    keys[i] = u;
    atomicAdd(&sh[u >> 21], 1u);
  }
  sel2_finish<LARGEST>(sh, 21, 11, st, ticket, out);
}

// passes 2 and 3: histogram of the next digit over the keys that match the prefix so far
template <bool LARGEST>
__global__ void __launch_bounds__(HS_TPB)
k_sel2_next(const unsigned int* __restrict__ keys, int64_t n, int shift, int bits, SelState2* st, unsigned int* ticket, float* __restrict__ out) {
  __shared__ unsigned int sh[SEL_BINS];
  for (int b = threadIdx.x; b < SEL_BINS; b += HS_TPB) sh[b] = 0;
  __syncthreads();
  const unsigned int prefix = st->prefix, mask = st->mask, dm = (1u << bits) - 1u;
  const int64_t gfull = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; g < gfull; g += stride) {
    const uint4 u = __ldcs(reinterpret_cast<const uint4*>(keys) + g);
    if ((u.x & mask) == prefix) atomicAdd(&sh[(u.x >> shift) & dm], 1u);
    if ((u.y & mask) == prefix) atomicAdd(&sh[(u.y >> shift) & dm], 1u);
    if ((u.z & mask) == prefix) atomicAdd(&sh[(u.z >> shift) & dm], 1u);
    if ((u.w & mask) == prefix) atomicAdd(&sh[(u.w >> shift) & dm], 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const unsigned int u = keys[gfull * 4 + threadIdx.x];
    if ((u & mask) == prefix) atomicAdd(&sh[(u >> shift) & dm], 1u);
  }
  sel2_finish<LARGEST>(sh, shift, bits, st, ticket, out);
}

// ---- order-preserving filter  (comp axis) <= limit -----------------------------------------------------------------
#define FL_TILE 1024  // points per tile: one group of 4 per thread

// this thread's group of 4 points of tile t (whole group: 3 x LDG.128; ragged end: scalar loads, missing points flagged)
__device__ __forceinline__ int load_tile_group(const float* __restrict__ xyz, int64_t n, int64_t i0, Pts4& p) {
  if (i0 + 4 <= n) { p = load_group(xyz, i0 >> 2); return 4; }
  int m = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    p.x[e] = p.y[e] = p.z[e] = 0.f;
    if (i0 + e < n) { p.x[e] = __ldg(xyz + 3 * (i0 + e)); p.y[e] = __ldg(xyz + 3 * (i0 + e) + 1); p.z[e] = __ldg(xyz + 3 * (i0 + e) + 2); m = e + 1; }
  }
  return m;
}

// A block takes FL_CT consecutive tiles per round: their loads are all in flight before the first count is reduced, and the
// counts of two tiles share one shuffle tree (each fits 16 bits), so a round costs one barrier pair for FL_CT * 1024 points.
#define FL_CT 4
__global__ void __launch_bounds__(HS_TPB)
k_filter_count(const float* __restrict__ xyz, int64_t n, int axis, float limit, unsigned int* __restrict__ tile_off,
               unsigned int* ticket, int64_t* __restrict__ n_out) {
  __shared__ unsigned int wsum[FL_CT / 2][HS_TPB / 32];
  const int64_t ntiles = (n + FL_TILE - 1) / FL_TILE;
  const int64_t nrounds = (ntiles + FL_CT - 1) / FL_CT;
  for (int64_t rd = blockIdx.x; rd < nrounds; rd += gridDim.x) {
    const int64_t t0 = rd * FL_CT;
    unsigned int c[FL_CT];
#pragma unroll
    for (int q = 0; q < FL_CT; ++q) {
      const int64_t i0 = (t0 + q) * FL_TILE + 4 * threadIdx.x;
      Pts4 p;
      const int m = load_tile_group(xyz, n, i0, p);
      c[q] = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) c[q] += (e < m) && ((axis == 0 ? p.x[e] : (axis == 1 ? p.y[e] : p.z[e])) <= limit);
    }
#pragma unroll
    for (int q = 0; q < FL_CT / 2; ++q) {
      unsigned int cc = c[2 * q] | (c[2 * q + 1] << 16);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cc += __shfl_xor_sync(0xffffffffu, cc, o);
      if ((threadIdx.x & 31) == 0) wsum[q][threadIdx.x >> 5] = cc;
    }
    __syncthreads();
    if (threadIdx.x < FL_CT / 2) {
      unsigned int sum = 0;
      for (int w = 0; w < HS_TPB / 32; ++w) sum += wsum[threadIdx.x][w];
      const int64_t t = t0 + 2 * threadIdx.x;
      if (t < ntiles) tile_off[t] = sum & 0xffffu;
      if (t + 1 < ntiles) tile_off[t + 1] = sum >> 16;
    }
    __syncthreads();
  }
  if (!last_block_arrives(ticket, gridDim.x)) return;
  const unsigned int total = block_scan_tiles_exclusive(tile_off, ntiles);
  if (threadIdx.x == 0) *n_out = total;
}

// the tile's kept points are compacted in shared memory, shifted so that shared and global float indices agree modulo 4,
// and leave as 16-byte stores (same scheme as k_bp_scatter); the colour cloud rides along with the same positions
__device__ __forceinline__ void store_run(const float* stage, int a, int nfloats, float* __restrict__ out, int64_t dst0) {
  const int lo = a, hi = a + nfloats;
  float* gbase = out + (dst0 - a);
  const int lo4 = (lo + 3) & ~3, hi4 = hi & ~3;
  if (lo4 < hi4) {
    if (static_cast<int>(threadIdx.x) < lo4 - lo) gbase[lo + threadIdx.x] = stage[lo + threadIdx.x];
    const float4* s4 = reinterpret_cast<const float4*>(stage);
    float4* g4 = reinterpret_cast<float4*>(gbase);
    for (int v = (lo4 >> 2) + threadIdx.x; v < (hi4 >> 2); v += HS_TPB) __stcs(g4 + v, s4[v]);
    if (static_cast<int>(threadIdx.x) < hi - hi4) gbase[hi4 + threadIdx.x] = stage[hi4 + threadIdx.x];
  } else {
    for (int q = lo + threadIdx.x; q < hi; q += HS_TPB) gbase[q] = stage[q];
  }
}

// warp-private flavour of store_run: the 32 lanes of one warp write the run staged in the warp's own slice
__device__ __forceinline__ void store_run_warp(const float* ws, int a, int nfloats, float* __restrict__ out, int64_t dst0, int lane) {
  const int lo = a, hi = a + nfloats;
  float* gbase = out + (dst0 - a);
  const int lo4 = (lo + 3) & ~3, hi4 = hi & ~3;
  if (lo4 < hi4) {
    if (lane < lo4 - lo) gbase[lo + lane] = ws[lo + lane];
    const float4* s4 = reinterpret_cast<const float4*>(ws);
    float4* g4 = reinterpret_cast<float4*>(gbase);
    for (int v = (lo4 >> 2) + lane; v < (hi4 >> 2); v += 32) __stcs(g4 + v, s4[v]);
    if (lane < hi - hi4) gbase[hi4 + lane] = ws[hi4 + lane];
  } else {
    for (int q = lo + lane; q < hi; q += 32) gbase[q] = ws[q];
  }
}

// Every warp stages the kept points of ITS 128 points in its own slice of shared memory and writes its own run: one block-wide
// barrier per tile (for the warp offsets) instead of three — the compaction was barrier-bound with block-wide staging.
__global__ void __launch_bounds__(HS_TPB)
k_filter_scatter(const float* __restrict__ xyz, int64_t n, int axis, float limit, const unsigned int* __restrict__ tile_off,
                 const float* __restrict__ extra_in, float* __restrict__ out, float* __restrict__ extra_out) {
  constexpr int NW = HS_TPB / 32, SLICE = 128 * 3 + 4;  // 388 floats = 97 x 16 B: slices stay 16-byte aligned
  __shared__ unsigned int wsum[2][NW];
  __shared__ __align__(16) float stage[NW * SLICE];
  __shared__ __align__(16) float stage2[NW * SLICE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ws = stage + warp * SLICE;
  float* ws2 = stage2 + warp * SLICE;
  const int64_t ntiles = (n + FL_TILE - 1) / FL_TILE;
  int par = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, par ^= 1) {
    const int64_t i0 = t * FL_TILE + 4 * threadIdx.x;
    Pts4 p, q;
    const int m = load_tile_group(xyz, n, i0, p);
    if (extra_in) load_tile_group(extra_in, n, i0, q);
    bool keep[4];
    unsigned int c = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { keep[e] = (e < m) && ((axis == 0 ? p.x[e] : (axis == 1 ? p.y[e] : p.z[e])) <= limit); c += keep[e]; }
    unsigned int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const unsigned int cw = __shfl_sync(0xffffffffu, incl, 31);  // kept points of this warp
    if (lane == 0) wsum[par][warp] = cw;
    __syncthreads();  // the only block-wide barrier of the tile (wsum alternates, so the next tile cannot overwrite it early)
    unsigned int wbase = 0;
#pragma unroll
    for (int u = 0; u < NW; ++u) wbase += (u < warp) ? wsum[par][u] : 0u;
    const int64_t dst0 = 3 * (static_cast<int64_t>(tile_off[t]) + wbase);
    const int a = static_cast<int>(dst0 & 3);
    int sp = a + 3 * static_cast<int>(incl - c);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (keep[e]) {
        ws[sp] = p.x[e]; ws[sp + 1] = p.y[e]; ws[sp + 2] = p.z[e];
        if (extra_in) { ws2[sp] = q.x[e]; ws2[sp + 1] = q.y[e]; ws2[sp + 2] = q.z[e]; }
        sp += 3;
      }
    __syncwarp();
    store_run_warp(ws, a, 3 * static_cast<int>(cw), out, dst0, lane);
    if (extra_in) store_run_warp(ws2, a, 3 * static_cast<int>(cw), extra_out, dst0, lane);
    __syncwarp();  // the slice is reused by the warp's next tile
  }
}

// Single-pass filter (mode key 12 = 1; measured no faster than the two passes at 100 M points, so not the default): the same compaction with the offsets from a decoupled look-back over claims of FL_SUPER tiles
// (tile_lookback, k_common.cuh).  A claim's points are read once from HBM; the second read, for the scatter, comes from L2
// (48 KB per claim).  12 + 12 kept bytes per point instead of 24 + 12 kept, and one launch less.
#define FL_SUPER 4
__global__ void __launch_bounds__(HS_TPB)
k_filter_onepass(const float* __restrict__ xyz, int64_t n, int axis, float limit, unsigned long long* state /* [nsuper], zeroed */,
                 unsigned int* counters /* [0] next claim, [1] blocks done */, const float* __restrict__ extra_in, float* __restrict__ out,
                 float* __restrict__ extra_out, int64_t* __restrict__ n_out) {
  __shared__ unsigned int wsum[FL_SUPER][HS_TPB / 32];
  __shared__ __align__(16) float stage[FL_TILE * 3 + 4];
  __shared__ __align__(16) float stage2[FL_TILE * 3 + 4];
  __shared__ unsigned int s_claim;
  __shared__ unsigned long long s_prefix;
  const int64_t nsuper = (n + FL_SUPER * FL_TILE - 1) / (FL_SUPER * FL_TILE);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  volatile unsigned long long* vstate = state;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_claim = atomicAdd(counters, 1u);
    __syncthreads();
    const int64_t t = s_claim;
    if (t >= nsuper) break;
    unsigned int keepbits = 0;  // bit 4 q + e: point e of this thread's group in tile q is kept
#pragma unroll
    for (int q = 0; q < FL_SUPER; ++q) {
      const int64_t i0 = (t * FL_SUPER + q) * FL_TILE + 4 * threadIdx.x;
      Pts4 p;
      const int m = load_tile_group(xyz, n, i0, p);
      unsigned int c = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool k = (e < m) && ((axis == 0 ? p.x[e] : (axis == 1 ? p.y[e] : p.z[e])) <= limit);
        keepbits |= static_cast<unsigned int>(k) << (4 * q + e);
        c += k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) wsum[q][warp] = c;
    }
    __syncthreads();
    unsigned int all = 0;
#pragma unroll
    for (int q = 0; q < FL_SUPER; ++q)
#pragma unroll
      for (int u = 0; u < HS_TPB / 32; ++u) all += wsum[q][u];
    if (warp == 0) {
      const unsigned long long prefix = tile_lookback(vstate, t, all);
      if (lane == 0) {
        s_prefix = prefix;
        if (t == nsuper - 1) *n_out = static_cast<int64_t>(prefix + all);
      }
    }
    __syncthreads();
    int64_t pts0 = static_cast<int64_t>(s_prefix);
#pragma unroll 1
    for (int q = 0; q < FL_SUPER; ++q) {
      const int64_t i0 = (t * FL_SUPER + q) * FL_TILE + 4 * threadIdx.x;
      const unsigned int kb = (keepbits >> (4 * q)) & 15u;
      // position of this thread's first kept point inside the tile: kept points of the lower lanes + of the lower warps
      unsigned int incl = __popc(kb);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      unsigned int pos = incl - __popc(kb), tq = 0;
      for (int u = 0; u < HS_TPB / 32; ++u) { pos += (u < warp) ? wsum[q][u] : 0u; tq += wsum[q][u]; }
      const int64_t dst0 = 3 * pts0;
      const int a = static_cast<int>(dst0 & 3);
      if (kb) {
        Pts4 p, c;
        load_tile_group(xyz, n, i0, p);
        if (extra_in) load_tile_group(extra_in, n, i0, c);
        int sp = a + 3 * static_cast<int>(pos);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if ((kb >> e) & 1u) {
            stage[sp] = p.x[e]; stage[sp + 1] = p.y[e]; stage[sp + 2] = p.z[e];
            if (extra_in) { stage2[sp] = c.x[e]; stage2[sp + 1] = c.y[e]; stage2[sp + 2] = c.z[e]; }
            sp += 3;
          }
      }
      __syncthreads();
      store_run(stage, a, 3 * static_cast<int>(tq), out, dst0);
      if (extra_in) store_run(stage2, a, 3 * static_cast<int>(tq), extra_out, dst0);
      pts0 += tq;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counters + 1, 1u) == gridDim.x - 1) { counters[0] = 0u; counters[1] = 0u; }
  }
}

}  // namespace hsk

using namespace hsk;

// One radix pass of the sharded k-th: the local histogram of the pass's digit over this rank's points, restricted to
// the keys that match (prefix, mask).  Pass 0 also builds the dense key array, which stays in the ctx scratch for passes 1 and 2.
int32_t launch_kth_shard_pass(hs_ctx* ctx, const float* xyz, int64_t n, int axis, int pass, uint32_t prefix, uint32_t mask, uint32_t* d_hist_out) {
  const size_t keys_off = (sizeof(SelState2) + 255) & ~static_cast<size_t>(255);
  if (int32_t rc = hs_ensure_scratch(ctx, keys_off + static_cast<size_t>(n) * 4 + 16)) return rc;
  SelState2* st = reinterpret_cast<SelState2*>(ctx->d_scratch);
  unsigned int* keys = reinterpret_cast<unsigned int*>(ctx->d_scratch + keys_off);
  SelState2 head;  // prefix, mask, k_rem: only the first 16 bytes travel; the histogram is zeroed on the device
  head.prefix = prefix; head.mask = mask; head.k_rem = 0;
  HS_CUDA_TRY(ctx, cudaMemsetAsync(st, 0, sizeof(SelState2), ctx->stream));
  HS_CUDA_TRY(ctx, cudaMemcpyAsync(st, &head, 16, cudaMemcpyHostToDevice, ctx->stream));
  HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // `head` lives on this stack frame
  int64_t nb = ((n >> 2) + HS_TPB - 1) / HS_TPB;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  const int g = static_cast<int>(nb);
  if (pass == 0) k_sel2_first<true><<<g, HS_TPB, 0, ctx->stream>>>(xyz, n, axis, keys, st, ctx->d_ticket, nullptr);
  else if (pass == 1) k_sel2_next<true><<<g, HS_TPB, 0, ctx->stream>>>(keys, n, 10, 11, st, ctx->d_ticket, nullptr);
  else k_sel2_next<true><<<g, HS_TPB, 0, ctx->stream>>>(keys, n, 0, 10, st, ctx->d_ticket, nullptr);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  HS_CUDA_TRY(ctx, cudaMemcpyAsync(d_hist_out, st->hist, sizeof(unsigned int) * SEL_BINS, cudaMemcpyDeviceToDevice, ctx->stream));
  return HS_OK;
}

int32_t launch_kth(hs_ctx* ctx, const float* xyz, int64_t n, int axis, int64_t k, bool largest, float* d_out) {
  if (ctx->modes[HS_MODE_SEL_KERNEL] != 1 && n < (1ll << 32) && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0) {
    const size_t keys_off = (sizeof(SelState2) + 255) & ~static_cast<size_t>(255);
    if (int32_t rc = hs_ensure_scratch(ctx, keys_off + static_cast<size_t>(n) * 4 + 16)) return rc;
    SelState2* st = reinterpret_cast<SelState2*>(ctx->d_scratch);
    unsigned int* keys = reinterpret_cast<unsigned int*>(ctx->d_scratch + keys_off);
    k_sel2_init<<<1, 256, 0, ctx->stream>>>(st, static_cast<unsigned long long>(k));
    int64_t nb = ((n >> 2) + HS_TPB - 1) / HS_TPB;
    const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    const int g = static_cast<int>(nb);
    if (largest) {
      k_sel2_first<true><<<g, HS_TPB, 0, ctx->stream>>>(xyz, n, axis, keys, st, ctx->d_ticket, d_out);
      k_sel2_next<true><<<g, HS_TPB, 0, ctx->stream>>>(keys, n, 10, 11, st, ctx->d_ticket, d_out);
      k_sel2_next<true><<<g, HS_TPB, 0, ctx->stream>>>(keys, n, 0, 10, st, ctx->d_ticket, d_out);
    } else {
      k_sel2_first<false><<<g, HS_TPB, 0, ctx->stream>>>(xyz, n, axis, keys, st, ctx->d_ticket, d_out);
      k_sel2_next<false><<<g, HS_TPB, 0, ctx->stream>>>(keys, n, 10, 11, st, ctx->d_ticket, d_out);
      k_sel2_next<false><<<g, HS_TPB, 0, ctx->stream>>>(keys, n, 0, 10, st, ctx->d_ticket, d_out);
    }
    ctx->launches += 4;
    HS_CUDA_TRY(ctx, cudaGetLastError());
    return HS_OK;
  }
  if (int32_t rc = hs_ensure_scratch(ctx, sizeof(SelState))) return rc;
  SelState* st = reinterpret_cast<SelState*>(ctx->d_scratch);
  k_sel_init<<<1, 256, 0, ctx->stream>>>(st, static_cast<unsigned long long>(k));
  ctx->launches++;
  int64_t nb = (n + HS_TPB - 1) / HS_TPB;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (largest) k_sel_pass<true><<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(xyz, n, axis, shift, st, ctx->d_ticket, d_out);
    else k_sel_pass<false><<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(xyz, n, axis, shift, st, ctx->d_ticket, d_out);
    ctx->launches++;
  }
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

int32_t launch_filter_le(hs_ctx* ctx, const float* xyz, int64_t n, int axis, float limit, const float* extra_in, float* out,
                         float* extra_out, int64_t* d_nout) {
  const int64_t ntiles = (n + FL_TILE - 1) / FL_TILE;
  int64_t nb = ntiles < static_cast<int64_t>(ctx->sm_count) * 8 ? ntiles : static_cast<int64_t>(ctx->sm_count) * 8;
  if (nb < 1) nb = 1;
  if (int32_t rc = hs_ensure_scratch(ctx, static_cast<size_t>(ntiles + 1) * sizeof(unsigned long long))) return rc;
  if (ctx->modes[HS_MODE_FILTER_KERNEL] == 1 && n > 0 && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0) {  // single pass (opt-in: measured no faster)
    unsigned long long* state = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
    const int64_t nsuper = (ntiles + FL_SUPER - 1) / FL_SUPER;
    HS_CUDA_TRY(ctx, cudaMemsetAsync(state, 0, static_cast<size_t>(nsuper) * sizeof(unsigned long long), ctx->stream));
    const int64_t nb1 = std::max<int64_t>(1, std::min<int64_t>(nsuper, static_cast<int64_t>(ctx->sm_count) * 6));
    k_filter_onepass<<<static_cast<int>(nb1), HS_TPB, 0, ctx->stream>>>(xyz, n, axis, limit, state, ctx->d_ticket + 24, extra_in, out, extra_out, d_nout);
    ctx->launches++;
    HS_CUDA_TRY(ctx, cudaGetLastError());
    return HS_OK;
  }
  unsigned int* tile_off = reinterpret_cast<unsigned int*>(ctx->d_scratch);
  const int64_t nbc = std::max<int64_t>(1, std::min<int64_t>((ntiles + FL_CT - 1) / FL_CT, static_cast<int64_t>(ctx->sm_count) * 8));
  k_filter_count<<<static_cast<int>(nbc), HS_TPB, 0, ctx->stream>>>(xyz, n, axis, limit, tile_off, ctx->d_ticket, d_nout);
  ctx->launches++;
  k_filter_scatter<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(xyz, n, axis, limit, tile_off, extra_in, out, extra_out);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
