// Rigid transforms and whole-cloud reductions.
//   rotateAround c R p = ((p &- c) .* R) &+ c          Main.hs:1582-1583 (rotateCloudAround Main.hs:1657-1659)
//   translateCloud off = V.map (off &+)                Main.hs:1697-1699
//   projectRoom cloud  = translateCloud off . rotateCloudAround zero R   Main.hs:1716
//   pointMean          = foldl' (&+) zero / n          Main.hs:1596-1601  (here: Double accumulation)
//   max extent         = V.maximum . V.map (distance m)  Main.hs:1527
//   fitPlane scatter   = sum (p-m)(p-m)^T in Double of Float-centred points   Main.hs:1441-1445
// Float arithmetic is written with explicit round-to-nearest intrinsics so nvcc cannot contract it into FMAs.
#include "k_common.cuh"

namespace hsk {

struct Affine {
  float R[9];
  float pre[3];   // subtracted before the rotation (rotation centre)
  float post[3];  // added after
};

// v .* R component c = ((vx*R0c + vy*R1c) + vz*R2c)
__device__ __forceinline__ float rowdot(float vx, float vy, float vz, float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(vx, a), __fmul_rn(vy, b)), __fmul_rn(vz, c));
}

// KIND 0: rotate_around (pre = post = centre)   ((p - c) .* R) + c
// KIND 1: project      (pre = 0)                 off + (((p - 0) .* R) + 0)
// KIND 2: translate                               off + p
template <int KIND>
__device__ __forceinline__ void xform(const Affine& A, float& x, float& y, float& z) {
  if (KIND == 2) {
    x = __fadd_rn(A.post[0], x); y = __fadd_rn(A.post[1], y); z = __fadd_rn(A.post[2], z);
    return;
  }
  const float vx = __fsub_rn(x, A.pre[0]), vy = __fsub_rn(y, A.pre[1]), vz = __fsub_rn(z, A.pre[2]);
  float rx = rowdot(vx, vy, vz, A.R[0], A.R[3], A.R[6]);
  float ry = rowdot(vx, vy, vz, A.R[1], A.R[4], A.R[7]);
  float rz = rowdot(vx, vy, vz, A.R[2], A.R[5], A.R[8]);
  if (KIND == 0) {
    x = __fadd_rn(rx, A.pre[0]); y = __fadd_rn(ry, A.pre[1]); z = __fadd_rn(rz, A.pre[2]);
  } else {
    rx = __fadd_rn(rx, 0.0f); ry = __fadd_rn(ry, 0.0f); rz = __fadd_rn(rz, 0.0f);  // (&+ zero): only turns -0 into +0
    x = __fadd_rn(A.post[0], rx); y = __fadd_rn(A.post[1], ry); z = __fadd_rn(A.post[2], rz);
  }
}

// A tile = HS_TPB groups of 4 points (12 KB).  Global traffic is fully coalesced in both directions (lane i moves the i-th
// 16 bytes of a contiguous 512-byte run per warp instruction); the 48-byte-per-thread AoS view exists only in shared memory,
// where 3 x LDS.128 / STS.128 at a 48 B lane stride are conflict free.  Direct 48 B-strided LDG/STG.128 touch half of every
// 32 B sector per instruction and cost ~25 % of the copy rate (profiles/r01_rows.txt).
template <int KIND>
__global__ void __launch_bounds__(HS_TPB)
k_affine(const float* __restrict__ in, float* __restrict__ out, int64_t n, const __grid_constant__ Affine A) {
  __shared__ float4 tile[3 * HS_TPB];
  const int64_t gfull = n >> 2;
  const int64_t ntiles = gfull / HS_TPB;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t base = t * (3 * HS_TPB);
    const float4 a = __ldcs(in4 + base + threadIdx.x), b = __ldcs(in4 + base + HS_TPB + threadIdx.x), c = __ldcs(in4 + base + 2 * HS_TPB + threadIdx.x);
    tile[threadIdx.x] = a; tile[HS_TPB + threadIdx.x] = b; tile[2 * HS_TPB + threadIdx.x] = c;
    __syncthreads();
    const float4 u = tile[3 * threadIdx.x], v = tile[3 * threadIdx.x + 1], w = tile[3 * threadIdx.x + 2];
    Pts4 p;
    p.x[0] = u.x; p.y[0] = u.y; p.z[0] = u.z;
    p.x[1] = u.w; p.y[1] = v.x; p.z[1] = v.y;
    p.x[2] = v.z; p.y[2] = v.w; p.z[2] = w.x;
    p.x[3] = w.y; p.y[3] = w.z; p.z[3] = w.w;
#pragma unroll
    for (int e = 0; e < 4; ++e) xform<KIND>(A, p.x[e], p.y[e], p.z[e]);
    tile[3 * threadIdx.x] = make_float4(p.x[0], p.y[0], p.z[0], p.x[1]);
    tile[3 * threadIdx.x + 1] = make_float4(p.y[1], p.z[1], p.x[2], p.y[2]);
    tile[3 * threadIdx.x + 2] = make_float4(p.z[2], p.x[3], p.y[3], p.z[3]);
    __syncthreads();
    __stcs(out4 + base + threadIdx.x, tile[threadIdx.x]);
    __stcs(out4 + base + HS_TPB + threadIdx.x, tile[HS_TPB + threadIdx.x]);
    __stcs(out4 + base + 2 * HS_TPB + threadIdx.x, tile[2 * HS_TPB + threadIdx.x]);
    __syncthreads();
  }
  // groups past the last whole tile (< HS_TPB of them) and the ragged points (< 4)
  if (blockIdx.x == 0) {
    const int64_t g = ntiles * HS_TPB + threadIdx.x;
    if (g < gfull) {
      Pts4 p = load_group(in, g);
#pragma unroll
      for (int e = 0; e < 4; ++e) xform<KIND>(A, p.x[e], p.y[e], p.z[e]);
      store_group(out, g, p);
    }
    if (threadIdx.x < (n & 3)) {
      const int64_t i = gfull * 4 + threadIdx.x;
      float x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
      xform<KIND>(A, x, y, z);
      out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
  }
}

// sum of points in Double -> out[0..2]
__global__ void __launch_bounds__(HS_TPB)
k_sum3(const float* __restrict__ xyz, int64_t n, double* __restrict__ partials, unsigned int* ticket, double* __restrict__ out) {
  __shared__ double smem[(HS_TPB / 32) * 3];
  double acc[3] = {0.0, 0.0, 0.0};
  const int64_t gfull = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; g < gfull; g += stride) {
    const Pts4 p = load_group(xyz, g);
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[0] += p.x[e]; acc[1] += p.y[e]; acc[2] += p.z[e]; }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = gfull * 4 + threadIdx.x;
    acc[0] += xyz[3 * i]; acc[1] += xyz[3 * i + 1]; acc[2] += xyz[3 * i + 2];
  }
  block_sum_store<3>(acc, partials + 3 * static_cast<int64_t>(blockIdx.x), smem);
  if (!last_block_arrives(ticket, gridDim.x)) return;
  __shared__ double fin[HS_TPB];
  last_block_sum<3>(partials, gridDim.x, out, fin);
}

// max over points of normsqr (m - p) in Float (exact order of `distance`), as ordered uint bits
__global__ void __launch_bounds__(HS_TPB)
k_max_nsq(const float* __restrict__ xyz, int64_t n, float mx, float my, float mz, unsigned int* __restrict__ out_bits) {
  float best = 0.0f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  auto upd = [&](float x, float y, float z) {
    const float dx = __fsub_rn(mx, x), dy = __fsub_rn(my, y), dz = __fsub_rn(mz, z);
    const float q = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    best = fmaxf(best, q);
  };
  const int64_t gfull = n >> 2;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; g < gfull; g += stride) {
    const Pts4 p = load_group(xyz, g);
#pragma unroll
    for (int e = 0; e < 4; ++e) upd(p.x[e], p.y[e], p.z[e]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = gfull * 4 + threadIdx.x;
    upd(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(best));  // non-negative floats order like uints
}

// scatter of Float-centred points, products and sums in Double: xx,xy,xz,yy,yz,zz
__global__ void __launch_bounds__(HS_TPB)
k_scatter(const float* __restrict__ xyz, int64_t n, float mx, float my, float mz, double* __restrict__ partials,
          unsigned int* ticket, double* __restrict__ out) {
  __shared__ double smem[(HS_TPB / 32) * 6];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  auto upd = [&](float x, float y, float z) {
    const double dx = __fsub_rn(x, mx), dy = __fsub_rn(y, my), dz = __fsub_rn(z, mz);
    acc[0] = fma(dx, dx, acc[0]); acc[1] = fma(dx, dy, acc[1]); acc[2] = fma(dx, dz, acc[2]);
    acc[3] = fma(dy, dy, acc[3]); acc[4] = fma(dy, dz, acc[4]); acc[5] = fma(dz, dz, acc[5]);
  };
  const int64_t gfull = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; g < gfull; g += stride) {
    const Pts4 p = load_group(xyz, g);
#pragma unroll
    for (int e = 0; e < 4; ++e) upd(p.x[e], p.y[e], p.z[e]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = gfull * 4 + threadIdx.x;
    upd(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }
  block_sum_store<6>(acc, partials + 6 * static_cast<int64_t>(blockIdx.x), smem);
  if (!last_block_arrives(ticket, gridDim.x)) return;
  __shared__ double fin[HS_TPB];
  last_block_sum<6>(partials, gridDim.x, out, fin);
}

}  // namespace hsk

using namespace hsk;

static int stream_blocks(const hs_ctx* ctx, int64_t n, int per_sm) {
  int64_t nb = static_cast<int64_t>(ctx->sm_count) * per_sm;
  const int64_t cap = ((n + 3) / 4 + HS_TPB - 1) / HS_TPB;
  if (nb > cap) nb = cap;
  return static_cast<int>(nb < 1 ? 1 : nb);
}

int32_t launch_affine(hs_ctx* ctx, const float* in, float* out, int64_t n, const float R[9], const float pre[3], const float post[3], int kind) {
  Affine A;
  for (int i = 0; i < 9; ++i) A.R[i] = R ? R[i] : (i % 4 == 0 ? 1.f : 0.f);
  for (int i = 0; i < 3; ++i) { A.pre[i] = pre ? pre[i] : 0.f; A.post[i] = post ? post[i] : 0.f; }
  const int nb = stream_blocks(ctx, n, 8);
  if (kind == 0) k_affine<0><<<nb, HS_TPB, 0, ctx->stream>>>(in, out, n, A);
  else if (kind == 1) k_affine<1><<<nb, HS_TPB, 0, ctx->stream>>>(in, out, n, A);
  else k_affine<2><<<nb, HS_TPB, 0, ctx->stream>>>(in, out, n, A);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

int32_t launch_mean(hs_ctx* ctx, const float* xyz, int64_t n, double* d_sum3) {
  const int nb = stream_blocks(ctx, n, 4);
  if (int32_t rc = hs_ensure_scratch(ctx, static_cast<size_t>(nb) * 3 * sizeof(double))) return rc;
  k_sum3<<<nb, HS_TPB, 0, ctx->stream>>>(xyz, n, reinterpret_cast<double*>(ctx->d_scratch), ctx->d_ticket, d_sum3);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

int32_t launch_max_nsq(hs_ctx* ctx, const float* xyz, int64_t n, const float m[3], unsigned int* d_maxbits) {
  HS_CUDA_TRY(ctx, cudaMemsetAsync(d_maxbits, 0, sizeof(unsigned int), ctx->stream));
  const int nb = stream_blocks(ctx, n, 4);
  k_max_nsq<<<nb, HS_TPB, 0, ctx->stream>>>(xyz, n, m[0], m[1], m[2], d_maxbits);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

int32_t launch_scatter(hs_ctx* ctx, const float* xyz, int64_t n, const float m[3], double* d_sc6) {
  const int nb = stream_blocks(ctx, n, 4);
  if (int32_t rc = hs_ensure_scratch(ctx, static_cast<size_t>(nb) * 6 * sizeof(double))) return rc;
  k_scatter<<<nb, HS_TPB, 0, ctx->stream>>>(xyz, n, m[0], m[1], m[2], reinterpret_cast<double*>(ctx->d_scratch), ctx->d_ticket, d_sc6);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
