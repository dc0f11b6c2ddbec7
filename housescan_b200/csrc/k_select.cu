// Order statistics and order-preserving filters.
//   kthSmallestBy / kthLargestBy (VectorUtil.hs:11-19): the k-th order statistic by a Float key (1-based k).
//   removeCeiling (Main.hs:2643-2664): yLimit = k-th largest y with k = n `quot` 5; keep y <= yLimit in input order.
// The order statistic is found by a 4-pass MSB radix select over order-preserving uint images of the Float keys
// (no sort, no copy of the cloud); the filter is a two-pass block-scan compaction.
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

// ---- order-preserving filter  (comp axis) <= limit -----------------------------------------------------------------
#define FL_TILE 1024  // points per tile: one group of 4 per thread

__global__ void __launch_bounds__(HS_TPB)
k_filter_count(const float* __restrict__ xyz, int64_t n, int axis, float limit, unsigned int* __restrict__ tile_off,
               unsigned int* ticket, int64_t* __restrict__ n_out) {
  __shared__ unsigned int wsum[HS_TPB / 32];
  const int64_t ntiles = (n + FL_TILE - 1) / FL_TILE;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t i0 = t * FL_TILE + 4 * threadIdx.x;
    unsigned int c = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { const int64_t i = i0 + e; if (i < n) c += __ldg(xyz + 3 * i + axis) <= limit; }
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
  if (threadIdx.x == 0) *n_out = total;
}

__global__ void __launch_bounds__(HS_TPB)
k_filter_scatter(const float* __restrict__ xyz, int64_t n, int axis, float limit, const unsigned int* __restrict__ tile_off,
                 const float* __restrict__ extra_in, float* __restrict__ out, float* __restrict__ extra_out) {
  __shared__ unsigned int wsum[HS_TPB / 32];
  const int64_t ntiles = (n + FL_TILE - 1) / FL_TILE;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t i0 = t * FL_TILE + 4 * threadIdx.x;
    float p[4][3];
    bool keep[4];
    unsigned int c = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t i = i0 + e;
      keep[e] = false;
      if (i < n) {
        p[e][0] = __ldg(xyz + 3 * i); p[e][1] = __ldg(xyz + 3 * i + 1); p[e][2] = __ldg(xyz + 3 * i + 2);
        keep[e] = (axis == 0 ? p[e][0] : (axis == 1 ? p[e][1] : p[e][2])) <= limit;
      }
      c += keep[e];
    }
    int64_t pos = static_cast<int64_t>(tile_off[t]) + block_exclusive_prefix(c, wsum);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (keep[e]) {
        out[3 * pos] = p[e][0]; out[3 * pos + 1] = p[e][1]; out[3 * pos + 2] = p[e][2];
        if (extra_in) {
          const int64_t i = i0 + e;
          extra_out[3 * pos] = __ldg(extra_in + 3 * i); extra_out[3 * pos + 1] = __ldg(extra_in + 3 * i + 1); extra_out[3 * pos + 2] = __ldg(extra_in + 3 * i + 2);
        }
        ++pos;
      }
    __syncthreads();
  }
}

}  // namespace hsk

using namespace hsk;

int32_t launch_kth(hs_ctx* ctx, const float* xyz, int64_t n, int axis, int64_t k, bool largest, float* d_out) {
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
  if (int32_t rc = hs_ensure_scratch(ctx, static_cast<size_t>(ntiles + 1) * sizeof(unsigned int))) return rc;
  unsigned int* tile_off = reinterpret_cast<unsigned int*>(ctx->d_scratch);
  k_filter_count<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(xyz, n, axis, limit, tile_off, ctx->d_ticket, d_nout);
  ctx->launches++;
  k_filter_scatter<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(xyz, n, axis, limit, tile_off, extra_in, out, extra_out);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
