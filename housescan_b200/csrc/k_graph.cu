// Connected components on dense vertex ids (north-star piece 4).
// Reference semantics: GroupConnectedComponents.groupCCContiguous (GroupConnectedComponents.hs:39-54) =
// Data.Graph.components of the undirected graph; the canonical label of a vertex is the minimum vertex id of its
// component (components are listed in ascending order of that minimum).  Implemented as a lock-free union-find:
// every link hangs the larger root under the smaller one (atomicCAS), so each root is its set's minimum.
#include <algorithm>

#include "k_common.cuh"
#define HS_CC_DEFAULT_CHUNKS 1

namespace hsk {

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  uint32_t p = parent[x];
  while (p != x) {
    const uint32_t gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving; racy but monotone (values only move toward the root)
    x = p;
    p = gp;
  }
  return x;
}

__global__ void __launch_bounds__(HS_TPB) k_cc_init(uint32_t* __restrict__ parent, uint32_t N) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; v < N; v += stride) parent[v] = static_cast<uint32_t>(v);
}

__global__ void __launch_bounds__(HS_TPB)
k_cc_link(const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst, int64_t E, uint32_t* parent) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; e < E; e += stride) {
    uint32_t u = uf_find(parent, src[e]), v = uf_find(parent, dst[e]);
    while (u != v) {
      if (u < v) { const uint32_t t = u; u = v; v = t; }  // u is the larger root
      const uint32_t old = atomicCAS(parent + u, u, v);
      if (old == u) break;
      u = uf_find(parent, old);
      v = uf_find(parent, v);
    }
  }
}

__global__ void __launch_bounds__(HS_TPB) k_cc_flatten(uint32_t* parent, uint32_t N) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; v < N; v += stride) {
    uint32_t r = static_cast<uint32_t>(v), p = parent[r];
    while (p != r) { r = p; p = parent[r]; }
    parent[v] = r;
  }
}

}  // namespace hsk

using namespace hsk;

int32_t launch_cc(hs_ctx* ctx, const uint32_t* d_src, const uint32_t* d_dst, int64_t E, uint32_t N, uint32_t* d_label) {
  auto blocks = [&](int64_t items) {
    int64_t nb = (items + HS_TPB - 1) / HS_TPB;
    const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 16;
    return static_cast<int>(nb < 1 ? 1 : (nb > cap ? cap : nb));
  };
  if (N == 0) return HS_OK;
  k_cc_init<<<blocks(N), HS_TPB, 0, ctx->stream>>>(d_label, N);
  ctx->launches++;
  if (E > 0) {
    // The edge list is linked in `chunks` slices with a flatten in between: a flatten costs one sweep over the vertices (8 B each)
    // and leaves every tree one level deep, so the finds of the next slice are one or two loads instead of pointer chases
    // through whatever chains the previous slice built (mode key 11; 0 = default).
    int chunks = ctx->modes[HS_MODE_CC_CHUNKS] > 0 ? ctx->modes[HS_MODE_CC_CHUNKS] : HS_CC_DEFAULT_CHUNKS;
    if (E < (1 << 20)) chunks = 1;
    const int64_t per = (E + chunks - 1) / chunks;
    for (int c = 0; c < chunks; ++c) {
      const int64_t e0 = c * per, e1 = std::min<int64_t>(E, e0 + per);
      if (e0 >= e1) break;
      if (c > 0) { k_cc_flatten<<<blocks(N), HS_TPB, 0, ctx->stream>>>(d_label, N); ctx->launches++; }
      k_cc_link<<<blocks(e1 - e0), HS_TPB, 0, ctx->stream>>>(d_src + e0, d_dst + e0, e1 - e0, d_label);
      ctx->launches++;
    }
  }
  k_cc_flatten<<<blocks(N), HS_TPB, 0, ctx->stream>>>(d_label, N);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
