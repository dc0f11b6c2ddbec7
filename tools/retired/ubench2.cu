// Can 4 warps per SMSP keep the packed-FP32 pipe busy?  FFMA2 chains with a given ILP at 16 warps/SM, plus an
// FFMA2 + FSET/FSEL mix that mimics the evaluation kernel's alternating FMA/ALU dependency chain.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 pack2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int ILP>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
  f2 x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = pack2(threadIdx.x * 1e-3f + i, 1.f + i);
  f2 aa = pack2(a, a), bb = pack2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma2(x[i], aa, bb);
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) { float lo, hi; unpack2(x[i], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// chain: FFMA2 -> FSET(lanes) -> FFMA2 -> ... (alternating pipes), ILP independent chains
template <int ILP>
__global__ void k_mix(float* out, int iters, float a, float b) {
  f2 x[ILP], y[ILP];
  for (int i = 0; i < ILP; ++i) { x[i] = pack2(threadIdx.x * 1e-3f + i, 1.f + i); y[i] = pack2(0.5f + i, 0.25f); }
  f2 aa = pack2(a, a), bb = pack2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        f2 t = fma2(x[i], aa, bb);
        float tl, th, yl, yh;
        unpack2(t, tl, th); unpack2(y[i], yl, yh);
        f2 p = pack2(fabsf(tl) < fabsf(yl) ? 1.f : 0.f, fabsf(th) < fabsf(yh) ? 1.f : 0.f);
        x[i] = fma2(p, t, y[i]);
        y[i] = fma2(p, bb, t);
      }
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) { float lo, hi; unpack2(x[i], lo, hi); s += lo + hi; unpack2(y[i], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount; double clk = prop.clockRate * 1e3;
  float* out; cudaMalloc(&out, 1 << 24);
  const int iters = 2048;
  for (int threads : {512, 1024}) {
    printf("threads/SM = %d (%d warps per SMSP)\n", threads, threads / 128);
#define RUN(K, ILP, OPS) { float ms = time_ms([&] { K<ILP><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); }); \
    printf("  %-8s ILP=%d: %.3f packed-instr/clk/SMSP\n", #K, ILP, (double)OPS * iters * (threads / 32.0) / 4 / (ms * 1e-3 * clk)); }
    RUN(k_ffma2, 1, 8) RUN(k_ffma2, 2, 16) RUN(k_ffma2, 4, 32) RUN(k_ffma2, 8, 64)
    RUN(k_mix, 1, 12) RUN(k_mix, 2, 24) RUN(k_mix, 4, 48)
  }
  printf("(k_mix counts 3 FFMA2 per step; each step also has 2 FSET)\n");
  return 0;
}
