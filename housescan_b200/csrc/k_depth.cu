// Depth frames -> points (north-star piece 1).
//   frame format: row-major host-endian uint16 (HoniHelper.hs:34-36, 45-46)
//   addDevicePointCloud (Main.hs:1296-1313): (y,x) = i `quotRem` width; drop depth == 0 keeping raster order;
//   scalePoints = (x/10.0, y/10.0, d/20.0 - 30.0) in Float with true IEEE division.
// and the fused per-frame point-to-plane 6x6 normal equations (north-star addition, no reference code).
#include <algorithm>

#include "k_common.cuh"

namespace hsk {

#define BP_TILE 2048  // pixels per block-tile: 8 consecutive pixels (one 16-byte load) per thread

// x, y < 2^24 (checked by the API) and d is a uint16, so the quotients come from div_rn_small: bit-identical to `/`.
// RN(1/10) and RN(1/20) as literals: 1/10 and 0.1 are the same real number, so the decimal literal rounds to the same Float
// (an __fdiv_rn(1.0f, 10.0f) here is NOT folded by the compiler and costs a full division per use)
#define HS_RCP10 0.1f
#define HS_RCP20 0.05f
__device__ __forceinline__ void scale_point(int x, int y, unsigned int d, float& X, float& Y, float& Z) {
  X = div_rn_small(static_cast<float>(x), 10.0f, HS_RCP10);
  Y = div_rn_small(static_cast<float>(y), 10.0f, HS_RCP10);
  Z = __fsub_rn(div_rn_small(static_cast<float>(d), 20.0f, HS_RCP20), 30.0f);
}

// this thread's 8 pixels of tile t (zeros past the end of the frame)
__device__ __forceinline__ void load_px8(const uint16_t* __restrict__ depth, int64_t npx, int64_t i0, bool aligned, unsigned int (&d)[8]) {
  if (aligned && i0 + 8 <= npx) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(depth + i0));
    d[0] = v.x & 0xffffu; d[1] = v.x >> 16; d[2] = v.y & 0xffffu; d[3] = v.y >> 16;
    d[4] = v.z & 0xffffu; d[5] = v.z >> 16; d[6] = v.w & 0xffffu; d[7] = v.w >> 16;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = (i0 + e < npx) ? depth[i0 + e] : 0u;
  }
}

// bit 15 / bit 31 set iff the low / high uint16 of w is non-zero (carry-free SWAR test: 3 logic ops for two pixels)
__device__ __forceinline__ unsigned int nz16x2(unsigned int w) {
  return (((w & 0x7fff7fffu) + 0x7fff7fffu) | w) & 0x80008000u;
}
// uint16 -> Float without I2F (quarter-rate XU pipe): 2^23 + d has d in its low mantissa bits
__device__ __forceinline__ float u16_lo_f(unsigned int w) { return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), 8388608.0f); }
__device__ __forceinline__ float u16_hi_f(unsigned int w) { return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)), 8388608.0f); }

// pass 1: valid-pixel count per tile (+ optional byte mask); the last block turns counts into exclusive offsets.
// A block takes BP_CT consecutive tiles per iteration: their loads are all in flight before the first count is reduced, and the
// counts of two tiles share one shuffle tree (each fits 16 bits), so a round costs one barrier pair for BP_CT * 2048 pixels.
#define BP_CT 4
__device__ __forceinline__ unsigned int count_group(const uint16_t* __restrict__ depth, int64_t npx, uint8_t* __restrict__ mask, int64_t i0,
                                                    bool aligned, bool swar) {
  if (i0 >= npx) return 0u;
  if (swar && i0 + 8 <= npx) {  // whole 16-byte group: two pixels per logic op, no per-pixel compare
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(depth + i0));
    const unsigned int n0 = nz16x2(v.x), n1 = nz16x2(v.y), n2 = nz16x2(v.z), n3 = nz16x2(v.w);
    if (mask) {
      uint2 m;  // byte e = (pixel e != 0)
      m.x = __byte_perm(n0 >> 15, n1 >> 15, 0x6420);
      m.y = __byte_perm(n2 >> 15, n3 >> 15, 0x6420);
      __stcs(reinterpret_cast<uint2*>(mask + i0), m);
    }
    return __popc(n0 | (n1 >> 1) | (n2 >> 2) | (n3 >> 3));
  }
  unsigned int d[8], c = 0;
  load_px8(depth, npx, i0, aligned, d);
#pragma unroll
  for (int e = 0; e < 8; ++e) c += d[e] != 0;
  if (mask) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (i0 + e < npx) mask[i0 + e] = static_cast<uint8_t>(d[e] != 0);
  }
  return c;
}

__global__ void __launch_bounds__(HS_TPB)
k_bp_count(const uint16_t* __restrict__ depth, int64_t npx, uint8_t* __restrict__ mask, unsigned int* __restrict__ tile_off,
           unsigned int* ticket, int64_t* __restrict__ n_valid) {
  __shared__ unsigned int wsum[BP_CT / 2][HS_TPB / 32];
  const int64_t ntiles = (npx + BP_TILE - 1) / BP_TILE;
  const int64_t nrounds = (ntiles + BP_CT - 1) / BP_CT;
  const bool aligned = (reinterpret_cast<uintptr_t>(depth) & 15) == 0;
  const bool mask_aligned = mask && (reinterpret_cast<uintptr_t>(mask) & 7) == 0;
  const bool swar = aligned && (!mask || mask_aligned);
  for (int64_t rd = blockIdx.x; rd < nrounds; rd += gridDim.x) {
    const int64_t t0 = rd * BP_CT;
    unsigned int c[BP_CT];
#pragma unroll
    for (int q = 0; q < BP_CT; ++q) c[q] = count_group(depth, npx, mask, (t0 + q) * BP_TILE + 8 * threadIdx.x, aligned, swar);
#pragma unroll
    for (int q = 0; q < BP_CT / 2; ++q) {
      unsigned int cc = c[2 * q] | (c[2 * q + 1] << 16);  // a tile holds 2048 pixels: both counts fit 16 bits all the way up
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cc += __shfl_xor_sync(0xffffffffu, cc, o);
      if ((threadIdx.x & 31) == 0) wsum[q][threadIdx.x >> 5] = cc;
    }
    __syncthreads();
    if (threadIdx.x < BP_CT / 2) {
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
  if (threadIdx.x == 0) *n_valid = total;
}

// pass 2: order-preserving scatter of the scaled points.
//   * a warp owns 256 consecutive pixels of the tile.  Each lane loads 16 bytes (8 pixels) and parks them in the warp's slice
//     of shared memory; the warp then walks its pixels in 8 steps of 32 CONSECUTIVE pixels (lane l takes pixel 32 e + l), so
//     the position of a valid pixel is a ballot + popcount and neighbouring lanes write neighbouring points of the staging
//     buffer (stride 3 words: conflict-free; 8 pixels per lane would put the lanes 24 words apart, an 8-way bank conflict);
//   * the tile's points are staged shifted so that a shared-memory float index and its global float index agree modulo 4, and
//     the run leaves as 16-byte stores (coalesced STG.128) with at most three scalar floats at either end.
__global__ void __launch_bounds__(HS_TPB)
k_bp_scatter(const uint16_t* __restrict__ depth, int64_t npx, int w, const unsigned int* __restrict__ tile_off, float* __restrict__ xyz) {
  __shared__ unsigned int wsum[HS_TPB / 32];
  __shared__ __align__(16) uint16_t sraw[BP_TILE];
  __shared__ __align__(16) float stage[BP_TILE * 3 + 4];
  const int64_t ntiles = (npx + BP_TILE - 1) / BP_TILE;
  const bool aligned = (reinterpret_cast<uintptr_t>(depth) & 15) == 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t i0 = t * BP_TILE + 8 * threadIdx.x;
    if (aligned && i0 + 8 <= npx) {
      reinterpret_cast<uint4*>(sraw)[threadIdx.x] = __ldg(reinterpret_cast<const uint4*>(depth + i0));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) sraw[8 * threadIdx.x + e] = (i0 + e < npx) ? depth[i0 + e] : static_cast<uint16_t>(0);
    }
    const int64_t dst0 = 3 * static_cast<int64_t>(tile_off[t]);  // first float of the tile's run in xyz
    const int a = static_cast<int>(dst0 & 3);
    __syncwarp();
    unsigned int d[8], bal[8], cw = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d[e] = sraw[256 * warp + 32 * e + lane];
      bal[e] = __ballot_sync(0xffffffffu, d[e] != 0);
      cw += __popc(bal[e]);
    }
    if (lane == 0) wsum[warp] = cw;
    __syncthreads();
    unsigned int run = 0, total = 0;
#pragma unroll
    for (int q = 0; q < HS_TPB / 32; ++q) { run += (q < warp) ? wsum[q] : 0u; total += wsum[q]; }
    // row / column of this lane's first pixel (a 64-bit division costs more than the rest of the iteration)
    const int64_t p0 = t * BP_TILE + 256 * warp + lane;
    int y, x;
    if (npx <= 0xffffffffll) { const unsigned int q = static_cast<unsigned int>(p0) / static_cast<unsigned int>(w); y = static_cast<int>(q); x = static_cast<int>(static_cast<unsigned int>(p0) - q * static_cast<unsigned int>(w)); }
    else { y = static_cast<int>(p0 / w); x = static_cast<int>(p0 - static_cast<int64_t>(y) * w); }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (d[e] != 0) {
        float* sp = stage + a + 3 * (run + __popc(bal[e] & lt));
        sp[0] = div_rn_small(static_cast<float>(x), 10.0f, HS_RCP10);
        sp[1] = div_rn_small(static_cast<float>(y), 10.0f, HS_RCP10);
        sp[2] = __fsub_rn(div_rn_small(static_cast<float>(d[e]), 20.0f, HS_RCP20), 30.0f);
      }
      run += __popc(bal[e]);
      x += 32;
      while (x >= w) { x -= w; ++y; }
    }
    __syncthreads();
    const int lo = a, hi = a + 3 * static_cast<int>(total);  // the run in shared-memory float indices; global index = dst0 - a + s
    float* gbase = xyz + (dst0 - a);                          // 16-byte aligned
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
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Single-pass form (default when points are wanted): count, mask, order-preserving offsets and scatter in ONE sweep over the
// depth data, the offsets coming from a decoupled look-back (tile_lookback, k_common.cuh) over claims of BP_SUPER tiles.
// The depth data is read once (2 B/px instead of 4) and one launch goes away; the scatter half is the code of k_bp_scatter.
// ------------------------------------------------------------------------------------------------------------------
#define BP_SUPER 4  // tiles per claim: one counter increment and one look-back per 8192 pixels

__global__ void __launch_bounds__(HS_TPB)
k_bp_onepass(const uint16_t* __restrict__ depth, int64_t npx, int w, uint8_t* __restrict__ mask, unsigned long long* state /* [nsuper], zeroed */,
             unsigned int* counters /* [0] next super-tile, [1] blocks done; zero between launches */, float* __restrict__ xyz, int64_t* __restrict__ n_valid) {
  __shared__ unsigned int wsum[BP_SUPER][HS_TPB / 32];
  __shared__ __align__(16) uint16_t sraw[BP_SUPER * BP_TILE];
  __shared__ __align__(16) float stage[(HS_TPB / 32) * (256 * 3 + 4)];
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_prefix;
  const int64_t nsuper = (npx + BP_SUPER * BP_TILE - 1) / (BP_SUPER * BP_TILE);
  const bool aligned = (reinterpret_cast<uintptr_t>(depth) & 15) == 0;
  const bool mask_aligned = mask && (reinterpret_cast<uintptr_t>(mask) & 7) == 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  volatile unsigned long long* vstate = state;
  for (;;) {
    __syncthreads();  // the previous claim's buffers and s_tile / s_prefix are free again
    if (threadIdx.x == 0) s_tile = atomicAdd(counters, 1u);
    __syncthreads();
    const int64_t t = s_tile;
    if (t >= nsuper) break;
    // ---- all BP_SUPER tiles of the claim: loads first (independent, all in flight), then mask + parking in shared memory
    uint4 v[BP_SUPER];
    bool fast[BP_SUPER];
#pragma unroll
    for (int q = 0; q < BP_SUPER; ++q) {
      const int64_t i0 = (t * BP_SUPER + q) * BP_TILE + 8 * threadIdx.x;
      fast[q] = aligned && i0 + 8 <= npx;
      v[q] = make_uint4(0u, 0u, 0u, 0u);
      if (fast[q]) v[q] = __ldg(reinterpret_cast<const uint4*>(depth + i0));
    }
#pragma unroll
    for (int q = 0; q < BP_SUPER; ++q) {
      const int64_t i0 = (t * BP_SUPER + q) * BP_TILE + 8 * threadIdx.x;
      if (!fast[q]) {  // ragged end of the raster or an unaligned frame
        unsigned int dd[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dd[e] = (i0 + e < npx) ? depth[i0 + e] : 0u;
        v[q] = make_uint4(dd[0] | (dd[1] << 16), dd[2] | (dd[3] << 16), dd[4] | (dd[5] << 16), dd[6] | (dd[7] << 16));
      }
      reinterpret_cast<uint4*>(sraw)[q * HS_TPB + threadIdx.x] = v[q];
      if (mask) {
        uint2 m;  // byte e = (pixel e != 0)
        m.x = __byte_perm(nz16x2(v[q].x) >> 15, nz16x2(v[q].y) >> 15, 0x6420);
        m.y = __byte_perm(nz16x2(v[q].z) >> 15, nz16x2(v[q].w) >> 15, 0x6420);
        if (mask_aligned && i0 + 8 <= npx) __stcs(reinterpret_cast<uint2*>(mask + i0), m);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (i0 + e < npx) mask[i0 + e] = static_cast<uint8_t>(((e < 4 ? m.x : m.y) >> (8 * (e & 3))) & 1u);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < BP_SUPER; ++q) {
      unsigned int cw = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) cw += __popc(__ballot_sync(0xffffffffu, sraw[q * BP_TILE + 256 * warp + 32 * e + lane] != 0));
      if (lane == 0) wsum[q][warp] = cw;
    }
    __syncthreads();
    unsigned int total[BP_SUPER], all = 0;
#pragma unroll
    for (int q = 0; q < BP_SUPER; ++q) {
      total[q] = 0;
#pragma unroll
      for (int u = 0; u < HS_TPB / 32; ++u) total[q] += wsum[q][u];
      all += total[q];
    }
    if (warp == 0) {
      const unsigned long long prefix = tile_lookback(vstate, t, all);
      if (lane == 0) {
        s_prefix = prefix;
        if (t == nsuper - 1) *n_valid = static_cast<int64_t>(prefix + all);
      }
    }
    __syncthreads();
    // ---- scatter: every warp stages ITS 256 pixels of a tile in its own slice of shared memory and writes its own run, so the
    // tiles of the claim need no block-wide barrier at all (the compaction was barrier-bound with block-wide staging)
    int64_t pts0 = static_cast<int64_t>(s_prefix);  // points before the current tile
    float* ws = stage + warp * (256 * 3 + 4);
#pragma unroll 1
    for (int q = 0; q < BP_SUPER; ++q) {
      unsigned int wbase = 0, tq = 0;
#pragma unroll
      for (int u = 0; u < HS_TPB / 32; ++u) { wbase += (u < warp) ? wsum[q][u] : 0u; tq += wsum[q][u]; }
      const unsigned int cwq = wsum[q][warp];
      const int64_t dst0 = 3 * (pts0 + wbase);  // first float of this warp's run in xyz
      const int a = static_cast<int>(dst0 & 3);
      const int64_t p0 = (t * BP_SUPER + q) * BP_TILE + 256 * warp + lane;
      int y, x;
      if (npx <= 0xffffffffll) { const unsigned int qq = static_cast<unsigned int>(p0) / static_cast<unsigned int>(w); y = static_cast<int>(qq); x = static_cast<int>(static_cast<unsigned int>(p0) - qq * static_cast<unsigned int>(w)); }
      else { y = static_cast<int>(p0 / w); x = static_cast<int>(p0 - static_cast<int64_t>(y) * w); }
      unsigned int run = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned int d = sraw[q * BP_TILE + 256 * warp + 32 * e + lane];
        const unsigned int bal = __ballot_sync(0xffffffffu, d != 0);
        if (d != 0) {
          float* sp = ws + a + 3 * (run + __popc(bal & lt));
          sp[0] = div_rn_small(static_cast<float>(x), 10.0f, HS_RCP10);
          sp[1] = div_rn_small(static_cast<float>(y), 10.0f, HS_RCP10);
          sp[2] = __fsub_rn(div_rn_small(static_cast<float>(d), 20.0f, HS_RCP20), 30.0f);
        }
        run += __popc(bal);
        x += 32;
        while (x >= w) { x -= w; ++y; }
      }
      __syncwarp();
      const int lo = a, hi = a + 3 * static_cast<int>(cwq);  // the run in this warp's staging indices; global index = dst0 - a + s
      float* gbase = xyz + (dst0 - a);                        // 16-byte aligned
      const int lo4 = (lo + 3) & ~3, hi4 = hi & ~3;
      if (lo4 < hi4) {
        if (lane < lo4 - lo) gbase[lo + lane] = ws[lo + lane];
        const float4* s4 = reinterpret_cast<const float4*>(ws);
        float4* g4 = reinterpret_cast<float4*>(gbase);
        for (int vv = (lo4 >> 2) + lane; vv < (hi4 >> 2); vv += 32) __stcs(g4 + vv, s4[vv]);
        if (lane < hi - hi4) gbase[hi4 + lane] = ws[hi4 + lane];
      } else {
        for (int qq = lo + lane; qq < hi; qq += 32) gbase[qq] = ws[qq];
      }
      pts0 += tq;
      __syncwarp();  // the slice is reused by the warp's next tile
    }
  }
  if (threadIdx.x == 0) {  // the last block to run out of tiles re-arms the counters for the next launch on this stream
    __threadfence();
    if (atomicAdd(counters + 1, 1u) == gridDim.x - 1) { counters[0] = 0u; counters[1] = 0u; }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// fused per-frame: back-project -> (pose) -> nearest plane -> J = [p x n, n], accumulate J^T J (21), J^T r (6), r^2, count.
// One block per frame (grid-stride over frames); Float geometry, Double accumulation, deterministic per-frame order.
// ------------------------------------------------------------------------------------------------------------------
struct FrameGeom {
  int use_intr;
  float fx, fy, cx, cy;
};

// One pixel's geometry: back-project, (pose), first-minimum plane, J = [p x n, n] and r, all in Float.  Branch-free so that
// the pixels a thread holds are evaluated with instruction-level parallelism; invalid pixels (d == 0) compute garbage that
// the caller never accumulates.  KT > 0: number of planes known at compile time (full unroll); KT == 0: tbl.K at run time.
struct PixelJ { float j[6], r; };
template <bool INTR, bool POSE, int KT>
__device__ __forceinline__ PixelJ ne_geometry(const FrameGeom& geo, const float (&M)[12], const PlaneTable& tbl, const float4* __restrict__ spl,
                                              int x, int y, unsigned int d) {
  float X, Y, Z;
  if (INTR) {
    Z = __fmul_rn(static_cast<float>(d), 0.001f);
    X = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(x), geo.cx), Z), geo.fx);
    Y = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(y), geo.cy), Z), geo.fy);
  } else {
    scale_point(x, y, d, X, Y, Z);
  }
  float px = X, py = Y, pz = Z;
  if (POSE) {
    px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[0]), __fmul_rn(Y, M[3])), __fmul_rn(Z, M[6])), M[9]);
    py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[1]), __fmul_rn(Y, M[4])), __fmul_rn(Z, M[7])), M[10]);
    pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[2]), __fmul_rn(Y, M[5])), __fmul_rn(Z, M[8])), M[11]);
  }
  float rb = plane_dist(tbl.pl[0][0], tbl.pl[0][1], tbl.pl[0][2], tbl.pl[0][3], px, py, pz);
  float ab = fabsf(rb);
  int kb = 0;
  if (KT > 0) {
#pragma unroll
    for (int k = 1; k < KT; ++k) {
      const float rk = plane_dist(tbl.pl[k][0], tbl.pl[k][1], tbl.pl[k][2], tbl.pl[k][3], px, py, pz);
      const float ak = fabsf(rk);
      const bool lt = ak < ab;
      ab = lt ? ak : ab; rb = lt ? rk : rb; kb = lt ? k : kb;
    }
  } else {
    for (int k = 1; k < tbl.K; ++k) {  // planes come from the constant bank with a uniform index
      const float rk = plane_dist(tbl.pl[k][0], tbl.pl[k][1], tbl.pl[k][2], tbl.pl[k][3], px, py, pz);
      const float ak = fabsf(rk);
      const bool lt = ak < ab;
      ab = lt ? ak : ab; rb = lt ? rk : rb; kb = lt ? k : kb;
    }
  }
  const float4 nn = spl[kb];  // the winner's normal: a per-lane index, so it is read from shared memory (one LDS.128), not selected
  PixelJ o;
  // crossprod p n in Float (y*c - z*b, z*a - x*c, x*b - y*a)
  o.j[0] = __fsub_rn(__fmul_rn(py, nn.z), __fmul_rn(pz, nn.y));
  o.j[1] = __fsub_rn(__fmul_rn(pz, nn.x), __fmul_rn(px, nn.z));
  o.j[2] = __fsub_rn(__fmul_rn(px, nn.y), __fmul_rn(py, nn.x));
  o.j[3] = nn.x; o.j[4] = nn.y; o.j[5] = nn.z;
  o.r = rb;
  return o;
}
// every product of the normal equations is formed in Double (exact for Float operands)
__device__ __forceinline__ void ne_accumulate(double (&acc)[HS_NE], const PixelJ& p) {
  const double J[6] = {p.j[0], p.j[1], p.j[2], p.j[3], p.j[4], p.j[5]};
  const double rd = p.r;
  int t = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) { acc[t] = fma(J[a], J[b], acc[t]); ++t; }
#pragma unroll
  for (int a = 0; a < 6; ++a) acc[21 + a] = fma(J[a], rd, acc[21 + a]);
  acc[27] = fma(rd, rd, acc[27]);
  acc[28] += 1.0;
}

// Frames are handed out to the resident blocks through a counter (any block may take any frame: a frame's record is computed
// by one block with a fixed thread -> pixel mapping and a fixed reduction tree, so the result does not depend on who took it).
// Each thread reads 8 consecutive pixels with one 16-byte load and evaluates them four at a time.
template <bool INTR, bool POSE, int KT>
__global__ void __launch_bounds__(HS_TPB, 2)
k_reduce6x6(const uint16_t* __restrict__ frames, int64_t nframes, int w, int h, const FrameGeom geo, const float* __restrict__ poses,
            const __grid_constant__ PlaneTable tbl, double* __restrict__ out, unsigned int* counters /* [0] next frame, [1] blocks done */) {
  __shared__ double smem[(HS_TPB / 32) * HS_NE];
  __shared__ float4 spl[16];
  __shared__ unsigned int s_frame;
  if (threadIdx.x < tbl.K) spl[threadIdx.x] = make_float4(tbl.pl[threadIdx.x][0], tbl.pl[threadIdx.x][1], tbl.pl[threadIdx.x][2], tbl.pl[threadIdx.x][3]);
  const int npx = w * h;
  const bool vec_ok = (npx & 7) == 0 && (reinterpret_cast<uintptr_t>(frames) & 15) == 0;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_frame = atomicAdd(counters, 1u);
    __syncthreads();
    const int64_t fr = s_frame;
    if (fr >= nframes) break;
    const uint16_t* dep = frames + fr * npx;
    float M[12];
    if (POSE) {
      const float* pm = poses + 16 * fr;
      M[0] = pm[0]; M[1] = pm[1]; M[2] = pm[2]; M[3] = pm[4]; M[4] = pm[5]; M[5] = pm[6];
      M[6] = pm[8]; M[7] = pm[9]; M[8] = pm[10]; M[9] = pm[12]; M[10] = pm[13]; M[11] = pm[14];
    }
    double acc[HS_NE];
#pragma unroll
    for (int i = 0; i < HS_NE; ++i) acc[i] = 0.0;
    if (vec_ok) {
      const uint4* dep8 = reinterpret_cast<const uint4*>(dep);
      for (int g = threadIdx.x; g < (npx >> 3); g += HS_TPB) {
        const uint4 v = __ldcs(dep8 + g);
        const unsigned int wv[4] = {v.x, v.y, v.z, v.w};
        const int i0 = g << 3;
        int y = i0 / w, x = i0 - y * w;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          PixelJ pj[4];
          unsigned int dd[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            dd[e] = (wv[2 * half + (e >> 1)] >> (16 * (e & 1))) & 0xffffu;
            pj[e] = ne_geometry<INTR, POSE, KT>(geo, M, tbl, spl, x, y, dd[e]);
            if (++x == w) { x = 0; ++y; }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (dd[e] != 0) ne_accumulate(acc, pj[e]);
        }
      }
    } else {
      for (int i = threadIdx.x; i < npx; i += HS_TPB) {
        const unsigned int d = dep[i];
        if (d == 0) continue;
        const int y = i / w, x = i - y * w;
        ne_accumulate(acc, ne_geometry<INTR, POSE, KT>(geo, M, tbl, spl, x, y, d));
      }
    }
    block_sum_store<HS_NE>(acc, out + fr * HS_NE, smem);
  }
  if (threadIdx.x == 0) {  // the last block to run out of frames re-arms the counters for the next launch on this stream
    __threadfence();
    if (atomicAdd(counters + 1, 1u) == gridDim.x - 1) { counters[0] = 0u; counters[1] = 0u; }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Throughput form of the same record (default).  Identical per-pixel Float geometry (so the plane choice, J and r are
// bit-identical to k_reduce6x6 and to the oracle); what changes is everything around it:
//   * the two back-projection quotients per pixel use the exact constant-divisor sequences of k_common.cuh instead of `/`;
//   * the winner's residual is recomputed from the winner's plane (one LDS.128 + 6 ops) instead of being carried through the
//     selection chain, and the running minimum is an FMNMX: 5 issue cycles per plane instead of 8;
//   * the 29 sums run as Float FMA chains over 16 pixels per thread and are then added into per-thread Doubles (the product
//     inside an FMA is exact; only the 16-term partial sums are rounded to Float, ~1e-8 of the record after the Double sums) —
//     29 FFMA + 7 F2F/DADD per pixel-row instead of 29 DFMA at half rate;
//   * a frame is cut into `parts` row bands handed out through the item counter, so 1000 frames fill 296 resident blocks
//     evenly; the band that finishes a frame last adds the bands in band order (deterministic).
// Requires w % 8 == 0 (a thread's 8 pixels share a row) and 16-byte aligned frames; the launcher falls back otherwise.
// ------------------------------------------------------------------------------------------------------------------
template <bool INTR, bool POSE, int KT, bool PAIRED>
__device__ __forceinline__ PixelJ ne_geometry_fast(const FrameGeom& geo, float rfx, float rfy, const float (&M)[12], const PlaneTable& tbl,
                                                   const float4* __restrict__ spl, float xf, float yc /* INTR: y - cy; else y/10 */, unsigned int d) {
  float X, Y, Z;
  const float df = static_cast<float>(d);
  if (INTR) {
    Z = __fmul_rn(df, 0.001f);
    X = div_rn_by(__fmul_rn(__fsub_rn(xf, geo.cx), Z), geo.fx, rfx);
    Y = div_rn_by(__fmul_rn(yc, Z), geo.fy, rfy);
  } else {
    X = div_rn_small(xf, 10.0f, HS_RCP10);
    Y = yc;
    Z = __fsub_rn(div_rn_small(df, 20.0f, HS_RCP20), 30.0f);
  }
  float px = X, py = Y, pz = Z;
  if (POSE) {
    px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[0]), __fmul_rn(Y, M[3])), __fmul_rn(Z, M[6])), M[9]);
    py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[1]), __fmul_rn(Y, M[4])), __fmul_rn(Z, M[7])), M[10]);
    pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[2]), __fmul_rn(Y, M[5])), __fmul_rn(Z, M[8])), M[11]);
  }
  float ab;
  const int kb = nearest_plane<KT, PAIRED>(tbl, px, py, pz, ab);
  const float4 nn = spl[kb];
  PixelJ o;
  o.j[0] = __fsub_rn(__fmul_rn(py, nn.z), __fmul_rn(pz, nn.y));
  o.j[1] = __fsub_rn(__fmul_rn(pz, nn.x), __fmul_rn(px, nn.z));
  o.j[2] = __fsub_rn(__fmul_rn(px, nn.y), __fmul_rn(py, nn.x));
  o.j[3] = nn.x; o.j[4] = nn.y; o.j[5] = nn.z;
  o.r = plane_dist(nn.x, nn.y, nn.z, nn.w, px, py, pz);  // same operations on the same operands as in the loop: same bits
  return o;
}
// every accumulation is a predicated FP instruction (no branch around the block, no selects): the four pixels a thread has in
// flight stay independent instruction streams for the scheduler
__device__ __forceinline__ void ne_accumulate_f32(float (&acc)[HS_NE], const PixelJ& p, unsigned int d) {
  asm("{\n .reg .pred p;\n setp.ne.u32 p, %36, 0;\n"
      "@p fma.rn.f32 %0, %29, %29, %0;\n"
      "@p fma.rn.f32 %1, %29, %30, %1;\n"
      "@p fma.rn.f32 %2, %29, %31, %2;\n"
      "@p fma.rn.f32 %3, %29, %32, %3;\n"
      "@p fma.rn.f32 %4, %29, %33, %4;\n"
      "@p fma.rn.f32 %5, %29, %34, %5;\n"
      "@p fma.rn.f32 %6, %30, %30, %6;\n"
      "@p fma.rn.f32 %7, %30, %31, %7;\n"
      "@p fma.rn.f32 %8, %30, %32, %8;\n"
      "@p fma.rn.f32 %9, %30, %33, %9;\n"
      "@p fma.rn.f32 %10, %30, %34, %10;\n"
      "@p fma.rn.f32 %11, %31, %31, %11;\n"
      "@p fma.rn.f32 %12, %31, %32, %12;\n"
      "@p fma.rn.f32 %13, %31, %33, %13;\n"
      "@p fma.rn.f32 %14, %31, %34, %14;\n"
      "@p fma.rn.f32 %15, %32, %32, %15;\n"
      "@p fma.rn.f32 %16, %32, %33, %16;\n"
      "@p fma.rn.f32 %17, %32, %34, %17;\n"
      "@p fma.rn.f32 %18, %33, %33, %18;\n"
      "@p fma.rn.f32 %19, %33, %34, %19;\n"
      "@p fma.rn.f32 %20, %34, %34, %20;\n"
      "@p fma.rn.f32 %21, %29, %35, %21;\n"
      "@p fma.rn.f32 %22, %30, %35, %22;\n"
      "@p fma.rn.f32 %23, %31, %35, %23;\n"
      "@p fma.rn.f32 %24, %32, %35, %24;\n"
      "@p fma.rn.f32 %25, %33, %35, %25;\n"
      "@p fma.rn.f32 %26, %34, %35, %26;\n"
      "@p fma.rn.f32 %27, %35, %35, %27;\n"
      "@p add.rn.f32 %28, %28, 0f3F800000;\n"
      "}\n"
      : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7]), "+f"(acc[8]), "+f"(acc[9]), "+f"(acc[10]), "+f"(acc[11]), "+f"(acc[12]), "+f"(acc[13]), "+f"(acc[14]), "+f"(acc[15]), "+f"(acc[16]), "+f"(acc[17]), "+f"(acc[18]), "+f"(acc[19]), "+f"(acc[20]), "+f"(acc[21]), "+f"(acc[22]), "+f"(acc[23]), "+f"(acc[24]), "+f"(acc[25]), "+f"(acc[26]), "+f"(acc[27]), "+f"(acc[28])
      : "f"(p.j[0]), "f"(p.j[1]), "f"(p.j[2]), "f"(p.j[3]), "f"(p.j[4]), "f"(p.j[5]), "f"(p.r), "r"(d));
}

#ifndef NE_ILP
#define NE_ILP 4
#endif
template <bool INTR, bool POSE, int KT, bool PAIRED>
__global__ void __launch_bounds__(HS_TPB, 2)
k_reduce6x6_f32(const uint16_t* __restrict__ frames, int64_t nframes, int w, int h, const FrameGeom geo, float rfx, float rfy,
                const float* __restrict__ poses, const __grid_constant__ PlaneTable tbl, int parts, double* __restrict__ partials,
                unsigned int* __restrict__ part_done, double* __restrict__ out, unsigned int* counters /* [0] next item, [1] blocks done */) {
  extern __shared__ double sdacc[];  // [HS_NE][HS_TPB]: every thread's Double sums (registers are for the Float chains)
  __shared__ double smem[(HS_TPB / 32) * HS_NE];
  __shared__ float4 spl[16];
  __shared__ unsigned int s_item;
  __shared__ bool s_last;
  if (threadIdx.x < tbl.K) spl[threadIdx.x] = make_float4(tbl.pl[threadIdx.x][0], tbl.pl[threadIdx.x][1], tbl.pl[threadIdx.x][2], tbl.pl[threadIdx.x][3]);
  const int npx = w * h;
  const int ng = npx >> 3;                    // 8-pixel groups per frame
  const int gpp = (ng + parts - 1) / parts;   // groups per band
  const int64_t nitems = nframes * parts;
  const int dy = (8 * HS_TPB) / w, dx = (8 * HS_TPB) - dy * w;  // one loop step = 2048 pixels further on
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(counters, 1u);
    __syncthreads();
    const int64_t item = s_item;
    if (item >= nitems) break;
    const int64_t fr = item / parts;
    const int part = static_cast<int>(item - fr * parts);
    const uint4* dep8 = reinterpret_cast<const uint4*>(frames + fr * npx);
    float M[12];
    if (POSE) {
      const float* pm = poses + 16 * fr;
      M[0] = pm[0]; M[1] = pm[1]; M[2] = pm[2]; M[3] = pm[4]; M[4] = pm[5]; M[5] = pm[6];
      M[6] = pm[8]; M[7] = pm[9]; M[8] = pm[10]; M[9] = pm[12]; M[10] = pm[13]; M[11] = pm[14];
    }
    float acc[HS_NE];
#pragma unroll
    for (int i = 0; i < HS_NE; ++i) { sdacc[i * HS_TPB + threadIdx.x] = 0.0; acc[i] = 0.0f; }
    const int g_lo = part * gpp, g_hi = min(g_lo + gpp, ng);
    int g = g_lo + threadIdx.x;
    int y = (g << 3) / w, x = (g << 3) - y * w;
    int it = 0;
    uint4 vn = make_uint4(0u, 0u, 0u, 0u);
    if (g < g_hi) vn = __ldcs(dep8 + g);
    for (; g < g_hi; g += HS_TPB, ++it) {
      const uint4 v = vn;
      if (g + HS_TPB < g_hi) vn = __ldcs(dep8 + g + HS_TPB);  // the next 8 pixels are in flight while these are evaluated
      const unsigned int wv[4] = {v.x, v.y, v.z, v.w};
      const float x0f = static_cast<float>(x), yf = static_cast<float>(y);
      const float yc = INTR ? __fsub_rn(yf, geo.cy) : div_rn_small(yf, 10.0f, HS_RCP10);
#pragma unroll
      for (int q = 0; q < 8 / NE_ILP; ++q) {  // NE_ILP pixels' geometry in flight, then their accumulation
        PixelJ pj[NE_ILP];
        unsigned int dd[NE_ILP];
#pragma unroll
        for (int e = 0; e < NE_ILP; ++e) {
          const int px = NE_ILP * q + e;
          dd[e] = (wv[px >> 1] >> (16 * (px & 1))) & 0xffffu;
          pj[e] = ne_geometry_fast<INTR, POSE, KT, PAIRED>(geo, rfx, rfy, M, tbl, spl, __fadd_rn(x0f, static_cast<float>(px)), yc, dd[e]);
        }
#pragma unroll
        for (int e = 0; e < NE_ILP; ++e) ne_accumulate_f32(acc, pj[e], dd[e]);
      }
      if (it & 1) {  // 16 pixels per chain
#pragma unroll
        for (int i = 0; i < HS_NE; ++i) { sdacc[i * HS_TPB + threadIdx.x] += static_cast<double>(acc[i]); acc[i] = 0.0f; }
      }
      x += dx; y += dy;
      if (x >= w) { x -= w; ++y; }
    }
    double dacc[HS_NE];
#pragma unroll
    for (int i = 0; i < HS_NE; ++i) dacc[i] = sdacc[i * HS_TPB + threadIdx.x] + static_cast<double>(acc[i]);
    if (parts == 1) {
      block_sum_store<HS_NE>(dacc, out + fr * HS_NE, smem);
      continue;
    }
    block_sum_store<HS_NE>(dacc, partials + item * HS_NE, smem);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int t = atomicAdd(part_done + fr, 1u);
      s_last = (t == static_cast<unsigned int>(parts) - 1u);
      if (s_last) part_done[fr] = 0u;  // re-armed for the next launch
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (threadIdx.x < HS_NE) {
        double s = 0.0;
        for (int p = 0; p < parts; ++p) s += __ldcg(partials + (fr * parts + p) * HS_NE + threadIdx.x);
        out[fr * HS_NE + threadIdx.x] = s;
      }
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counters + 1, 1u) == gridDim.x - 1) { counters[0] = 0u; counters[1] = 0u; }
  }
}

}  // namespace hsk

using namespace hsk;

int32_t launch_backproject(hs_ctx* ctx, const uint16_t* d_depth, int32_t w, int32_t h, float* d_xyz, uint8_t* d_mask, int64_t* d_nvalid) {
  const int64_t npx = static_cast<int64_t>(w) * h;
  const int64_t ntiles = (npx + BP_TILE - 1) / BP_TILE;
  int64_t nb = std::min<int64_t>(ntiles, static_cast<int64_t>(ctx->sm_count) * 8);
  if (nb < 1) nb = 1;
  const int64_t nbc = std::max<int64_t>(1, std::min<int64_t>((ntiles + BP_CT - 1) / BP_CT, static_cast<int64_t>(ctx->sm_count) * 8));
  if (int32_t rc = hs_ensure_scratch(ctx, static_cast<size_t>(ntiles + 1) * sizeof(unsigned long long))) return rc;
  if (d_xyz && ctx->modes[HS_MODE_BP_KERNEL] != 1 && npx > 0) {  // single pass with decoupled look-back
    unsigned long long* state = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
    const int64_t nsuper = (ntiles + BP_SUPER - 1) / BP_SUPER;
    HS_CUDA_TRY(ctx, cudaMemsetAsync(state, 0, static_cast<size_t>(nsuper) * sizeof(unsigned long long), ctx->stream));
    nb = std::max<int64_t>(1, std::min<int64_t>(nsuper, static_cast<int64_t>(ctx->sm_count) * 5));
    k_bp_onepass<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(d_depth, npx, w, d_mask, state, ctx->d_ticket + 16, d_xyz, d_nvalid);
    ctx->launches++;
    HS_CUDA_TRY(ctx, cudaGetLastError());
    return HS_OK;
  }
  unsigned int* tile_off = reinterpret_cast<unsigned int*>(ctx->d_scratch);
  k_bp_count<<<static_cast<int>(nbc), HS_TPB, 0, ctx->stream>>>(d_depth, npx, d_mask, tile_off, ctx->d_ticket, d_nvalid);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  if (d_xyz) {
    k_bp_scatter<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(d_depth, npx, w, tile_off, d_xyz);
    ctx->launches++;
    HS_CUDA_TRY(ctx, cudaGetLastError());
  }
  return HS_OK;
}

// bands per frame so that the item count is >= 16 x the resident blocks (even tail), 1 when there are plenty of frames
static int ne_parts(const hs_ctx* ctx, int64_t nframes, int32_t w, int32_t h) {
  const int64_t want = static_cast<int64_t>(ctx->sm_count) * 2 * 16;
  int parts = 1;
  while (parts < 8 && nframes * parts < want && (static_cast<int64_t>(w) * h >> 3) / (parts * 2) >= 4 * HS_TPB) parts *= 2;
  return parts;
}
size_t reduce6x6_work_bytes(const hs_ctx* ctx, int64_t nframes, int32_t w, int32_t h) {
  const int parts = ne_parts(ctx, nframes, w, h);
  return static_cast<size_t>(nframes) * parts * HS_NE * sizeof(double) + static_cast<size_t>(nframes) * sizeof(unsigned int) + 64;
}

int32_t launch_reduce6x6(hs_ctx* ctx, const uint16_t* d_frames, int64_t nframes, int32_t w, int32_t h, const float* intr,
                         const float* d_poses, const PlaneTable& tbl, double* d_out, char* d_work) {
  FrameGeom geo{};
  geo.use_intr = intr != nullptr;
  if (intr) { geo.fx = intr[0]; geo.fy = intr[1]; geo.cx = intr[2]; geo.cy = intr[3]; }
  unsigned int* counters = ctx->d_ticket + 8;  // zero between launches (re-armed by the kernel)
  const bool fast_ok = (w % 8) == 0 && (reinterpret_cast<uintptr_t>(d_frames) & 15) == 0 && d_work != nullptr;
  if (ctx->modes[HS_MODE_NE_KERNEL] != 1 && fast_ok) {
    const int parts = ne_parts(ctx, nframes, w, h);
    double* partials = reinterpret_cast<double*>(d_work);
    unsigned int* part_done = reinterpret_cast<unsigned int*>(d_work + static_cast<size_t>(nframes) * parts * HS_NE * sizeof(double));
    if (parts > 1) HS_CUDA_TRY(ctx, cudaMemsetAsync(part_done, 0, static_cast<size_t>(nframes) * sizeof(unsigned int), ctx->stream));
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nframes * parts, static_cast<int64_t>(ctx->sm_count) * 2)));
    const float rfx = intr ? 1.0f / geo.fx : 0.f, rfy = intr ? 1.0f / geo.fy : 0.f;  // RN(1/f): IEEE division on the host
    const int dsm = HS_NE * HS_TPB * static_cast<int>(sizeof(double));
#define HS_NEF_LAUNCH(I, P, KT_, PR_)                                                                                      \
  cudaFuncSetAttribute(k_reduce6x6_f32<I, P, KT_, PR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm);                 \
  k_reduce6x6_f32<I, P, KT_, PR_><<<grid, HS_TPB, dsm, ctx->stream>>>(d_frames, nframes, w, h, geo, rfx, rfy, d_poses, tbl, parts, partials, part_done, d_out, counters)
#define HS_NEF_PICK(KT_, PR_)                                   \
  do {                                                          \
    if (intr && d_poses) { HS_NEF_LAUNCH(true, true, KT_, PR_); }  \
    else if (intr) { HS_NEF_LAUNCH(true, false, KT_, PR_); }       \
    else if (d_poses) { HS_NEF_LAUNCH(false, true, KT_, PR_); }    \
    else { HS_NEF_LAUNCH(false, false, KT_, PR_); }                \
  } while (0)
    if (tbl.K == 6 && tbl.paired) HS_NEF_PICK(6, true);  // a cuboid room's walls
    else if (tbl.K == 6) HS_NEF_PICK(6, false);
    else HS_NEF_PICK(0, false);
#undef HS_NEF_PICK
#undef HS_NEF_LAUNCH
    ctx->launches++;
    HS_CUDA_TRY(ctx, cudaGetLastError());
    return HS_OK;
  }
  int64_t nb = std::min<int64_t>(nframes, static_cast<int64_t>(ctx->sm_count) * 2);
  if (nb < 1) nb = 1;
  const int grid = static_cast<int>(nb);
#define HS_NE_LAUNCH(I, P, KT_) k_reduce6x6<I, P, KT_><<<grid, HS_TPB, 0, ctx->stream>>>(d_frames, nframes, w, h, geo, d_poses, tbl, d_out, counters)
#define HS_NE_PICK(KT_)                                  \
  do {                                                   \
    if (intr && d_poses) HS_NE_LAUNCH(true, true, KT_);  \
    else if (intr) HS_NE_LAUNCH(true, false, KT_);       \
    else if (d_poses) HS_NE_LAUNCH(false, true, KT_);    \
    else HS_NE_LAUNCH(false, false, KT_);                \
  } while (0)
  if (tbl.K == 6) HS_NE_PICK(6);  // a cuboid room's six walls: fully unrolled plane loop
  else HS_NE_PICK(0);
#undef HS_NE_PICK
#undef HS_NE_LAUNCH
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
