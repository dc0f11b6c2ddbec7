// Compute-only timing of the per-point formulations of the cuboid-sums kernel (no HBM traffic: tiles are read from shared
// memory over and over).  Reports issue cycles per point per SMSP-warp, and checks that all formulations produce bit-identical
// chains.  Not part of the product.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -Ihousescan_b200/csrc -o tools/ubench4 tools/ubench4.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "k_eval_point.cuh"
#include "k_eval_group_gen.cuh"
#include "ubench4_ablations.cuh"
using namespace hsk;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int VAR>
__global__ void __launch_bounds__(1024, 1) k_eval(const float* __restrict__ pts, float* out, int iters, RoomK R) {
  extern __shared__ __align__(16) float tile[];
  const int nthr = blockDim.x;
  for (int i = threadIdx.x; i < nthr * 12; i += nthr) tile[i] = pts[i];
  __syncthreads();
  ChainsP ch;
  ch.clear();
  const RoomK2 R2 = make_room_k2(R);
  int g = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint32_t base = s_u32(tile) + g * 48;
    g += 37; if (g >= nthr) g -= nthr;
    if (VAR == 0) {
      float4 q0, q1, q2;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(base));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "r"(base + 16));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w) : "r"(base + 32));
      add_point_pred(ch, R, q0.x, q0.y, q0.z);
      add_point_pred(ch, R, q0.w, q1.x, q1.y);
      add_point_pred(ch, R, q1.z, q1.w, q2.x);
      add_point_pred(ch, R, q2.y, q2.z, q2.w);
    } else {
      unsigned long long w0, w1, w2, w3, w4, w5;
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(base));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "r"(base + 16));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w4), "=l"(w5) : "r"(base + 32));
      if (VAR == 1) add_group_v1(ch, R2, w0, w1, w2, w3, w4, w5);
      if (VAR == 2) add_group_nopred(ch, R2, w0, w1, w2, w3, w4, w5);
      if (VAR == 3) add_group_front(ch, R2, w0, w1, w2, w3, w4, w5);
    }
  }
  float* o = out + (static_cast<size_t>(blockIdx.x) * nthr + threadIdx.x) * 21;
  o[0] = ch.f; o[16] = ch.C1; o[17] = ch.C2;
  for (int j = 0; j < 3; ++j) { o[1 + j] = ch.T[j]; o[4 + j] = ch.M[j]; o[18 + j] = ch.Cm[j]; for (int q = 0; q < 3; ++q) o[7 + 3 * j + q] = ch.B[j][q]; }
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); const double clk = khz * 1e3;
  // a 5 x 2.6 x 4 m room rotated a little; points near its walls
  RoomK R;
  const float c = 0.9396926f, s = 0.3420201f;
  const float n[3][3] = {{c, 0, s}, {0, 1, 0}, {-s, 0, c}};
  const float half[3] = {2.5f, 1.3f, 2.0f}, ctr[3] = {0.3f, -0.2f, 4.0f};
  for (int j = 0; j < 3; ++j) { float cj = 0; for (int q = 0; q < 3; ++q) { R.n[j][q] = n[j][q]; cj += n[j][q] * ctr[q]; } R.dp[j] = cj + half[j]; R.dm[j] = -(cj - half[j]); }
  const int maxthr = 1024;
  std::vector<float> h(maxthr * 12);
  srand(7);
  for (int i = 0; i < maxthr * 4; ++i) {
    float u[3]; for (int q = 0; q < 3; ++q) u[q] = (rand() / (float)RAND_MAX - 0.5f) * 2 * half[q];
    const int face = rand() % 6; u[face / 2] = (face & 1 ? -1.f : 1.f) * half[face / 2] + (rand() / (float)RAND_MAX - 0.5f) * 0.02f;
    for (int q = 0; q < 3; ++q) h[3 * i + q] = ctr[q] + u[0] * n[0][q] + u[1] * n[1][q] + u[2] * n[2][q];
  }
  float *d_pts, *d_out[3];
  CK(cudaMalloc(&d_pts, h.size() * 4)); CK(cudaMemcpy(d_pts, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  const size_t out_n = static_cast<size_t>(sms) * maxthr * 21;
  for (int v = 0; v < 3; ++v) CK(cudaMalloc(&d_out[v], out_n * 4));
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int nvar = 4;
  printf("device %s, %d SMs, nominal %.0f MHz; cycles are per point per SMSP-warp (4 SMSPs share the block's warps)\n", prop.name, sms, clk / 1e6);
  for (int thr : {512, 768}) {
    for (int v = 0; v < nvar; ++v) {
      auto launch = [&](int it) {
        const size_t sm = static_cast<size_t>(thr) * 48;
        if (v == 0) k_eval<0><<<sms, thr, sm>>>(d_pts, d_out[0], it, R);
        if (v == 1) k_eval<1><<<sms, thr, sm>>>(d_pts, d_out[1], it, R);
        if (v == 2) k_eval<2><<<sms, thr, sm>>>(d_pts, d_out[2], it, R);
        if (v == 3) k_eval<3><<<sms, thr, sm>>>(d_pts, d_out[2], it, R);
      };
      launch(64); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double cyc = ms * 1e-3 * clk / (static_cast<double>(iters) * 4 * (thr / 128.0));
      printf("threads %4d variant %d: %8.3f ms  %6.2f cycles/point/SMSP  -> %.0f Gpts/s chip-wide at this clock\n", thr, v, ms, cyc, sms * 4 * 32 / cyc * clk / 1e9);
    }
    // bit-exact agreement of the chains (same iteration count => same sums)
    std::vector<float> a(out_n), b(out_n);
    CK(cudaMemcpy(a.data(), d_out[0], out_n * 4, cudaMemcpyDeviceToHost));
    for (int v = 1; v < 2; ++v) {
      CK(cudaMemcpy(b.data(), d_out[v], out_n * 4, cudaMemcpyDeviceToHost));
      size_t bad = 0; const size_t lim = static_cast<size_t>(sms) * thr * 21;
      for (size_t i = 0; i < lim; ++i) bad += memcmp(&a[i], &b[i], 4) != 0;
      printf("threads %4d variant %d vs 0: %zu of %zu chain values differ\n", thr, v, bad, lim);
    }
  }
  return 0;
}
