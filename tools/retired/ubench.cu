// Micro-benchmarks that answer design questions for the evaluation kernel (not part of the product):
//  - streaming read bandwidth of the access patterns under consideration
//  - FMA-pipe throughput scalar vs packed f32x2, ALU-pipe throughput of FSEL/FSETP
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// ---- bandwidth: each thread reads 3 x float4 of one 48-byte group (the product's current pattern)
__global__ void __launch_bounds__(256) bw_group48(const float4* __restrict__ p, int64_t groups, float* out) {
  float acc = 0;
  int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < groups; g += stride) {
    float4 a = __ldg(p + 3 * g), b = __ldg(p + 3 * g + 1), c = __ldg(p + 3 * g + 2);
    acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w;
  }
  if (acc == 12345.678f) out[0] = acc;
}
// contiguous block chunk per block, 2 groups in flight
__global__ void __launch_bounds__(256) bw_group48_chunk(const float4* __restrict__ p, int64_t groups, int64_t gpb, float* out) {
  float acc = 0;
  int64_t g0 = (int64_t)blockIdx.x * gpb, g1 = min(g0 + gpb, groups);
  int64_t g = g0 + threadIdx.x;
  for (; g + 256 < g1; g += 512) {
    float4 a = __ldg(p + 3 * g), b = __ldg(p + 3 * g + 1), c = __ldg(p + 3 * g + 2);
    float4 d = __ldg(p + 3 * (g + 256)), e = __ldg(p + 3 * (g + 256) + 1), f = __ldg(p + 3 * (g + 256) + 2);
    acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w;
    acc += d.x + d.y + d.z + d.w + e.x + e.y + e.z + e.w + f.x + f.y + f.z + f.w;
  }
  for (; g < g1; g += 256) {
    float4 a = __ldg(p + 3 * g), b = __ldg(p + 3 * g + 1), c = __ldg(p + 3 * g + 2);
    acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w;
  }
  if (acc == 12345.678f) out[0] = acc;
}
// perfectly coalesced float4 grid-stride, UNROLL loads in flight
template <int U>
__global__ void __launch_bounds__(256) bw_coalesced(const float4* __restrict__ p, int64_t n4, float* out) {
  float acc = 0;
  int64_t stride = (int64_t)gridDim.x * 256;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  for (; i < n4; i += stride) { float4 v = __ldg(p + i); acc += v.x + v.y + v.z + v.w; }
  if (acc == 12345.678f) out[0] = acc;
}

// 1-D bulk async copy (TMA engine, no tensor map) global -> shared with an mbarrier ring; consumers read LDS.128 with 48 B stride
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
  asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int STAGES, int TILE_GROUPS>  // TILE_GROUPS groups of 48 B per stage
__global__ void __launch_bounds__(256) bw_bulk(const float4* __restrict__ p, int64_t groups, int64_t gpb, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* tiles = reinterpret_cast<float4*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * TILE_GROUPS * 48);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int64_t g0 = (int64_t)blockIdx.x * gpb, g1 = min(g0 + gpb, groups);
  int64_t ntiles = (g1 - g0 + TILE_GROUPS - 1) / TILE_GROUPS;
  float acc = 0;
  // prologue: producer thread fills the ring
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES && s < ntiles; ++s) {
      int64_t tg0 = g0 + (int64_t)s * TILE_GROUPS;
      uint32_t bytes = (uint32_t)(min((int64_t)TILE_GROUPS, g1 - tg0) * 48);
      mbar_expect_tx(full + s, bytes);
      bulk_g2s(tiles + (size_t)s * TILE_GROUPS * 3, p + 3 * tg0, bytes, full + s);
    }
  }
  for (int64_t t = 0; t < ntiles; ++t) {
    int s = (int)(t % STAGES);
    uint32_t ph = (uint32_t)((t / STAGES) & 1);
    mbar_wait(full + s, ph);
    int64_t tg0 = g0 + t * TILE_GROUPS;
    int ng = (int)min((int64_t)TILE_GROUPS, g1 - tg0);
    const float4* tile = tiles + (size_t)s * TILE_GROUPS * 3;
    for (int g = threadIdx.x; g < ng; g += 256) {
      float4 a = tile[3 * g], b = tile[3 * g + 1], c = tile[3 * g + 2];
      acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w;
    }
    mbar_arrive(empty + s);
    if (threadIdx.x == 0 && t + STAGES < ntiles) {
      mbar_wait(empty + s, ph);
      int64_t ng0 = g0 + (t + STAGES) * TILE_GROUPS;
      uint32_t bytes = (uint32_t)(min((int64_t)TILE_GROUPS, g1 - ng0) * 48);
      mbar_expect_tx(full + s, bytes);
      bulk_g2s(tiles + (size_t)s * TILE_GROUPS * 3, p + 3 * ng0, bytes, full + s);
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

// ---- pipe throughput: N dependent-chain-free ops per thread
template <int MODE>
__global__ void __launch_bounds__(256) pipe_tput(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  float2 y0 = make_float2(x0, x1), y1 = make_float2(x2, x3), y2 = make_float2(x4, x5), y3 = make_float2(x6, x7);
  float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // scalar FMUL + FADD, 8 independent chains: 16 FMA-pipe instr
      x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b); x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b);
      x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b); x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b);
    } else if (MODE == 1) {  // packed FMUL2 + FADD2: same 16 flop-pairs in 8 instr
      y0 = __fadd2_rn(__fmul2_rn(y0, aa), bb); y1 = __fadd2_rn(__fmul2_rn(y1, aa), bb); y2 = __fadd2_rn(__fmul2_rn(y2, aa), bb); y3 = __fadd2_rn(__fmul2_rn(y3, aa), bb);
    } else if (MODE == 2) {  // FSETP+FSEL pairs: 16 ALU instr
      x0 = (fabsf(x1) < fabsf(x0)) ? x1 + 0.f : x0; x1 = (fabsf(x2) < fabsf(x1)) ? x2 : x1; x2 = (fabsf(x3) < fabsf(x2)) ? x3 : x2; x3 = (fabsf(x4) < fabsf(x3)) ? x4 : x3;
      x4 = (fabsf(x5) < fabsf(x4)) ? x5 : x4; x5 = (fabsf(x6) < fabsf(x5)) ? x6 : x5; x6 = (fabsf(x7) < fabsf(x6)) ? x7 : x6; x7 = (fabsf(x0) < fabsf(x7)) ? x0 : x7;
      x0 = __fmul_rn(x0, a); x4 = __fmul_rn(x4, b);
    } else if (MODE == 3) {  // mixed: 8 FMUL2/FADD2 (16 pairs) + 8 ALU
      y0 = __fadd2_rn(__fmul2_rn(y0, aa), bb); y1 = __fadd2_rn(__fmul2_rn(y1, aa), bb); y2 = __fadd2_rn(__fmul2_rn(y2, aa), bb); y3 = __fadd2_rn(__fmul2_rn(y3, aa), bb);
      x0 = (fabsf(y1.x) < fabsf(y0.x)) ? y1.x : x0; x1 = (fabsf(y2.x) < fabsf(y1.y)) ? y2.x : x1; x2 = (fabsf(y3.x) < fabsf(y2.y)) ? y3.x : x2; x3 = (fabsf(y0.y) < fabsf(y3.y)) ? y0.y : x3;
    } else if (MODE == 4) {  // double: 8 DFMA
      double d0 = x0, d1 = x1;
      d0 = fma(d0, (double)a, (double)b); d1 = fma(d1, (double)a, (double)b);
      x0 = (float)d0; x1 = (float)d1;
    }
  }
  out[blockIdx.x * 256 + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + y0.x + y0.y + y1.x + y1.y + y2.x + y2.y + y3.x + y3.y;
}

template <class F>
static float time_ms(F f, int reps = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const int64_t npts = 100000008;
  const int64_t groups = npts / 4;
  const size_t bytes = (size_t)groups * 48;
  float4* d; float* out;
  CK(cudaMalloc(&d, bytes + 256)); CK(cudaMalloc(&out, 1 << 22));
  CK(cudaMemset(d, 0, bytes));
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  for (int bps : {2, 4, 8}) {
    int nb = sms * bps;
    float ms = time_ms([&] { bw_group48<<<nb, 256>>>(d, groups, out); });
    printf("bw_group48 gridstride bps=%d: %.1f us  %.0f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
    int64_t gpb = (groups + nb - 1) / nb;
    ms = time_ms([&] { bw_group48_chunk<<<nb, 256>>>(d, groups, gpb, out); });
    printf("bw_group48 chunk      bps=%d: %.1f us  %.0f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
    ms = time_ms([&] { bw_coalesced<4><<<nb, 256>>>(d, groups * 3, out); });
    printf("bw_coalesced U4       bps=%d: %.1f us  %.0f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
    ms = time_ms([&] { bw_coalesced<8><<<nb, 256>>>(d, groups * 3, out); });
    printf("bw_coalesced U8       bps=%d: %.1f us  %.0f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
  }
  {
    constexpr int ST = 4, TG = 256;  // 12 KB per stage
    size_t sm = (size_t)ST * TG * 48 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(bw_bulk<ST, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int bps : {1, 2, 3, 4}) {
      int nb = sms * bps; int64_t gpb = (groups + nb - 1) / nb;
      float ms = time_ms([&] { bw_bulk<ST, TG><<<nb, 256, sm>>>(d, groups, gpb, out); });
      printf("bw_bulk 4x12KB        bps=%d: %.1f us  %.0f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
    }
  }
  {
    constexpr int ST = 4, TG = 512;  // 24 KB per stage
    size_t sm = (size_t)ST * TG * 48 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(bw_bulk<ST, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int bps : {1, 2}) {
      int nb = sms * bps; int64_t gpb = (groups + nb - 1) / nb;
      float ms = time_ms([&] { bw_bulk<ST, TG><<<nb, 256, sm>>>(d, groups, gpb, out); });
      printf("bw_bulk 4x24KB        bps=%d: %.1f us  %.0f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
    }
  }
  // pipe throughput
  const int iters = 4096;
  int nb = sms * 8;
  double clk = prop.clockRate * 1e3;
  auto rep = [&](const char* name, float ms, double ops_per_iter) {
    double per_sm_clk = ops_per_iter * iters * 256.0 * nb / (ms * 1e-3) / sms / clk;
    printf("%-28s %.1f us  %.1f thread-ops/clk/SM (at %.0f MHz nominal)\n", name, ms * 1e3, per_sm_clk, clk / 1e6);
  };
  rep("FMUL+FADD scalar (16/iter)", time_ms([&] { pipe_tput<0><<<nb, 256>>>(out, iters, 1.0001f, 0.5f); }), 16);
  rep("FMUL2+FADD2 (16 flops/iter)", time_ms([&] { pipe_tput<1><<<nb, 256>>>(out, iters, 1.0001f, 0.5f); }), 16);
  rep("FSETP+FSEL (16 ALU/iter)", time_ms([&] { pipe_tput<2><<<nb, 256>>>(out, iters, 1.0001f, 0.5f); }), 18);
  rep("mix 16 f32x2-flops + 8 ALU", time_ms([&] { pipe_tput<3><<<nb, 256>>>(out, iters, 1.0001f, 0.5f); }), 24);
  return 0;
}
