// Depth frames -> points (north-star piece 1).
//   frame format: row-major host-endian uint16 (HoniHelper.hs:34-36, 45-46)
//   addDevicePointCloud (Main.hs:1296-1313): (y,x) = i `quotRem` width; drop depth == 0 keeping raster order;
//   scalePoints = (x/10.0, y/10.0, d/20.0 - 30.0) in Float with true IEEE division.
// and the fused per-frame point-to-plane 6x6 normal equations (north-star addition, no reference code).
#include <algorithm>

#include "k_common.cuh"

namespace hsk {

#define BP_TILE 1024  // pixels per block-tile (4 per thread)

__device__ __forceinline__ void scale_point(int x, int y, unsigned int d, float& X, float& Y, float& Z) {
  X = __fdiv_rn(static_cast<float>(x), 10.0f);
  Y = __fdiv_rn(static_cast<float>(y), 10.0f);
  Z = __fsub_rn(__fdiv_rn(static_cast<float>(d), 20.0f), 30.0f);
}

// pass 1: valid-pixel count per tile (+ optional byte mask); the last block turns counts into exclusive offsets
__global__ void __launch_bounds__(HS_TPB)
k_bp_count(const uint16_t* __restrict__ depth, int64_t npx, uint8_t* __restrict__ mask, unsigned int* __restrict__ tile_off,
           unsigned int* ticket, int64_t* __restrict__ n_valid) {
  __shared__ unsigned int wsum[HS_TPB / 32];
  const int64_t ntiles = (npx + BP_TILE - 1) / BP_TILE;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t i0 = t * BP_TILE + 4 * threadIdx.x;
    unsigned int c = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t i = i0 + e;
      if (i < npx) {
        const unsigned int v = depth[i] != 0;
        if (mask) mask[i] = static_cast<uint8_t>(v);
        c += v;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int s = 0;
      for (int w = 0; w < HS_TPB / 32; ++w) s += wsum[w];
      tile_off[t] = s;
    }
    __syncthreads();
  }
  if (!last_block_arrives(ticket, gridDim.x)) return;
  const unsigned int total = block_scan_tiles_exclusive(tile_off, ntiles);
  if (threadIdx.x == 0) *n_valid = total;
}

// pass 2: order-preserving scatter of the scaled points
__global__ void __launch_bounds__(HS_TPB)
k_bp_scatter(const uint16_t* __restrict__ depth, int64_t npx, int w, const unsigned int* __restrict__ tile_off, float* __restrict__ xyz) {
  __shared__ unsigned int wsum[HS_TPB / 32];
  const int64_t ntiles = (npx + BP_TILE - 1) / BP_TILE;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t i0 = t * BP_TILE + 4 * threadIdx.x;
    unsigned int dv[4], c = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { const int64_t i = i0 + e; dv[e] = i < npx ? depth[i] : 0u; c += dv[e] != 0; }
    unsigned int pos = tile_off[t] + block_exclusive_prefix(c, wsum);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (dv[e] != 0) {
        const int64_t i = i0 + e;
        const int y = static_cast<int>(i / w), x = static_cast<int>(i - static_cast<int64_t>(y) * w);
        float X, Y, Z;
        scale_point(x, y, dv[e], X, Y, Z);
        xyz[3 * static_cast<int64_t>(pos)] = X; xyz[3 * static_cast<int64_t>(pos) + 1] = Y; xyz[3 * static_cast<int64_t>(pos) + 2] = Z;
        ++pos;
      }
    __syncthreads();
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

__global__ void __launch_bounds__(HS_TPB)
k_reduce6x6(const uint16_t* __restrict__ frames, int64_t nframes, int w, int h, const FrameGeom geo, const float* __restrict__ poses,
            const __grid_constant__ PlaneTable tbl, double* __restrict__ out) {
  __shared__ double smem[(HS_TPB / 32) * HS_NE];
  const int npx = w * h;
  for (int64_t fr = blockIdx.x; fr < nframes; fr += gridDim.x) {
    const uint16_t* dep = frames + fr * npx;
    float M[12];
    const bool has_pose = poses != nullptr;
    if (has_pose) {
      const float* pm = poses + 16 * fr;
      M[0] = pm[0]; M[1] = pm[1]; M[2] = pm[2]; M[3] = pm[4]; M[4] = pm[5]; M[5] = pm[6];
      M[6] = pm[8]; M[7] = pm[9]; M[8] = pm[10]; M[9] = pm[12]; M[10] = pm[13]; M[11] = pm[14];
    }
    double acc[HS_NE];
#pragma unroll
    for (int i = 0; i < HS_NE; ++i) acc[i] = 0.0;
    for (int i = threadIdx.x; i < npx; i += HS_TPB) {
      const unsigned int d = dep[i];
      if (d == 0) continue;
      const int y = i / w, x = i - y * w;
      float X, Y, Z;
      if (geo.use_intr) {
        Z = __fmul_rn(static_cast<float>(d), 0.001f);
        X = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(x), geo.cx), Z), geo.fx);
        Y = __fdiv_rn(__fmul_rn(__fsub_rn(static_cast<float>(y), geo.cy), Z), geo.fy);
      } else {
        scale_point(x, y, d, X, Y, Z);
      }
      float px = X, py = Y, pz = Z;
      if (has_pose) {
        px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[0]), __fmul_rn(Y, M[3])), __fmul_rn(Z, M[6])), M[9]);
        py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[1]), __fmul_rn(Y, M[4])), __fmul_rn(Z, M[7])), M[10]);
        pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, M[2]), __fmul_rn(Y, M[5])), __fmul_rn(Z, M[8])), M[11]);
      }
      float rb = plane_dist(tbl.pl[0][0], tbl.pl[0][1], tbl.pl[0][2], tbl.pl[0][3], px, py, pz);
      float ab = fabsf(rb);
      int kb = 0;
      for (int k = 1; k < tbl.K; ++k) {
        const float rk = plane_dist(tbl.pl[k][0], tbl.pl[k][1], tbl.pl[k][2], tbl.pl[k][3], px, py, pz);
        const float ak = fabsf(rk);
        const bool lt = ak < ab;
        ab = lt ? ak : ab; rb = lt ? rk : rb; kb = lt ? k : kb;
      }
      const float nx = tbl.pl[kb][0], ny = tbl.pl[kb][1], nz = tbl.pl[kb][2];
      // crossprod p n in Float (y*c - z*b, z*a - x*c, x*b - y*a)
      const double J[6] = {static_cast<double>(__fsub_rn(__fmul_rn(py, nz), __fmul_rn(pz, ny))),
                           static_cast<double>(__fsub_rn(__fmul_rn(pz, nx), __fmul_rn(px, nz))),
                           static_cast<double>(__fsub_rn(__fmul_rn(px, ny), __fmul_rn(py, nx))),
                           static_cast<double>(nx), static_cast<double>(ny), static_cast<double>(nz)};
      const double rd = rb;
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
    block_sum_store<HS_NE>(acc, out + fr * HS_NE, smem);
  }
}

}  // namespace hsk

using namespace hsk;

int32_t launch_backproject(hs_ctx* ctx, const uint16_t* d_depth, int32_t w, int32_t h, float* d_xyz, uint8_t* d_mask, int64_t* d_nvalid) {
  const int64_t npx = static_cast<int64_t>(w) * h;
  const int64_t ntiles = (npx + BP_TILE - 1) / BP_TILE;
  int64_t nb = std::min<int64_t>(ntiles, static_cast<int64_t>(ctx->sm_count) * 8);
  if (nb < 1) nb = 1;
  if (int32_t rc = hs_ensure_scratch(ctx, static_cast<size_t>(ntiles + 1) * sizeof(unsigned int))) return rc;
  unsigned int* tile_off = reinterpret_cast<unsigned int*>(ctx->d_scratch);
  k_bp_count<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(d_depth, npx, d_mask, tile_off, ctx->d_ticket, d_nvalid);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  if (d_xyz) {
    k_bp_scatter<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(d_depth, npx, w, tile_off, d_xyz);
    ctx->launches++;
    HS_CUDA_TRY(ctx, cudaGetLastError());
  }
  return HS_OK;
}

int32_t launch_reduce6x6(hs_ctx* ctx, const uint16_t* d_frames, int64_t nframes, int32_t w, int32_t h, const float* intr,
                         const float* d_poses, const PlaneTable& tbl, double* d_out) {
  FrameGeom geo{};
  geo.use_intr = intr != nullptr;
  if (intr) { geo.fx = intr[0]; geo.fy = intr[1]; geo.cx = intr[2]; geo.cy = intr[3]; }
  int64_t nb = std::min<int64_t>(nframes, static_cast<int64_t>(ctx->sm_count) * 4);
  if (nb < 1) nb = 1;
  k_reduce6x6<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(d_frames, nframes, w, h, geo, d_poses, tbl, d_out);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
