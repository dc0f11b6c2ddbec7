// Do the packed-FP32 pipe (FFMA2) and the ALU pipe (FSETP/FSEL) overlap when one warp's instruction stream holds LONG RUNS of
// each (the shape of the cuboid-sums hot loop: dots -> selection network -> accumulation), or only when the two kinds alternate
// finely?  Same work per iteration in both kernels:
//   k_phased    : per pair  F(8 FFMA2) -> A(26 FSETP/FSEL) -> F2(8 FFMA2), each phase depending on the one before
//   k_pipelined : software pipeline over iterations: F(pair t+1) | A(pair t) | F2(pair t-1) are independent inside the body
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 pack2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

struct K { f2 c[8]; };
struct Q { f2 sp[3], sm[3]; f2 x, y; };
struct Z { f2 z[3], sf; f2 x, y; };

__device__ __forceinline__ Q phaseF(f2 x, f2 y, const K& k) {  // 8 FFMA2
  Q q;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const f2 t = fma2(x, k.c[j], fma2(y, k.c[j + 3], k.c[6]));
    q.sp[j] = t;  // stands for t - d+
    q.sm[j] = fma2(t, k.c[7], k.c[j]);
    if (j == 2) break;
  }
  q.sp[2] = fma2(x, k.c[2], k.c[5]);
  q.sm[2] = fma2(y, k.c[1], k.c[4]);
  q.x = x; q.y = y;
  return q;
}
__device__ __forceinline__ void lane(const float (&sp)[3], const float (&sm)[3], float& sf, float (&z)[3]) {  // 13 ALU
  float s[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { const bool P = fabsf(sm[j]) < fabsf(sp[j]); s[j] = P ? sm[j] : sp[j]; }
  const bool Q1 = fabsf(s[1]) < fabsf(s[0]);
  const float s01 = Q1 ? s[1] : s[0];
  const bool Q2 = fabsf(s[2]) < fabsf(s01);
  sf = Q2 ? s[2] : s01;
  const float z01 = Q2 ? 0.0f : s01;
  z[2] = Q2 ? s[2] : 0.0f;
  z[1] = Q1 ? z01 : 0.0f;
  z[0] = Q1 ? 0.0f : z01;
}
__device__ __forceinline__ Z phaseA(const Q& q) {  // 26 ALU
  float spl[3], sph[3], sml[3], smh[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { unpack2(q.sp[j], spl[j], sph[j]); unpack2(q.sm[j], sml[j], smh[j]); }
  float sfl, sfh, zl[3], zh[3];
  lane(spl, sml, sfl, zl);
  lane(sph, smh, sfh, zh);
  Z o;
  o.sf = pack2(sfl, sfh);
#pragma unroll
  for (int j = 0; j < 3; ++j) o.z[j] = pack2(zl[j], zh[j]);
  o.x = q.x; o.y = q.y;
  return o;
}
struct Acc { f2 a[8]; };
__device__ __forceinline__ void phaseF2(Acc& A, const Z& z) {  // 8 FFMA2
  A.a[0] = fma2(z.sf, z.sf, A.a[0]);
  A.a[1] = fma2(z.z[0], z.x, A.a[1]);
  A.a[2] = fma2(z.z[0], z.y, A.a[2]);
  A.a[3] = fma2(z.z[1], z.x, A.a[3]);
  A.a[4] = fma2(z.z[1], z.y, A.a[4]);
  A.a[5] = fma2(z.z[2], z.x, A.a[5]);
  A.a[6] = fma2(z.z[2], z.y, A.a[6]);
  A.a[7] = fma2(z.z[2], z.sf, A.a[7]);
}
__device__ __forceinline__ K make_k(float a, float b) {
  K k;
  for (int i = 0; i < 8; ++i) k.c[i] = pack2(a + 0.01f * i, b - 0.02f * i);
  return k;
}
__device__ __forceinline__ void finish(float* out, const Acc& A) {
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; unpack2(A.a[i], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NP>  // pairs per iteration (the real kernel has 2)
__global__ void k_phased(float* out, int iters, float a, float b) {
  const K k = make_k(a, b);
  Acc A;
  for (int i = 0; i < 8; ++i) A.a[i] = pack2(0.f, 0.f);
  f2 x = pack2(threadIdx.x * 1e-3f, 1.f), y = pack2(0.5f, threadIdx.x * 2e-3f);
  const f2 dx = pack2(1e-3f, 2e-3f), one = pack2(1.f, 1.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const Q q = phaseF(x, y, k);
      const Z z = phaseA(q);
      phaseF2(A, z);
      x = fma2(x, one, dx); y = fma2(y, one, dx);
    }
  }
  finish(out, A);
}
template <int NP>
__global__ void k_pipelined(float* out, int iters, float a, float b) {
  const K k = make_k(a, b);
  Acc A;
  for (int i = 0; i < 8; ++i) A.a[i] = pack2(0.f, 0.f);
  f2 x = pack2(threadIdx.x * 1e-3f, 1.f), y = pack2(0.5f, threadIdx.x * 2e-3f);
  const f2 dx = pack2(1e-3f, 2e-3f), one = pack2(1.f, 1.f);
  Q q[NP];
  Z z[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) { q[p] = phaseF(x, y, k); z[p] = phaseA(q[p]); x = fma2(x, one, dx); y = fma2(y, one, dx); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      phaseF2(A, z[p]);           // pair t-1
      z[p] = phaseA(q[p]);        // pair t
      q[p] = phaseF(x, y, k);     // pair t+1
      x = fma2(x, one, dx); y = fma2(y, one, dx);
    }
  }
  finish(out, A);
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
  const int iters = 4096;
  printf("per pair: 18 FFMA2 (2 of them the x/y update) + 26 FSETP/FSEL; cycles are per pair per warp on one SMSP, clock %.0f MHz nominal\n", clk / 1e6);
  for (int threads : {384, 512, 768, 1024}) {
#define RUN(KN, NP) { float ms = time_ms([&] { KN<NP><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); }); \
    printf("  threads/SM %4d  %-12s NP=%d: %.1f cycles/pair/warp (sum model 88, max model 52)\n", threads, #KN, NP, ms * 1e-3 * clk / (double(iters) * NP * (threads / 128.0))); }
    RUN(k_phased, 1) RUN(k_phased, 2) RUN(k_phased, 4) RUN(k_pipelined, 1) RUN(k_pipelined, 2) RUN(k_pipelined, 4)
  }
  return 0;
}
