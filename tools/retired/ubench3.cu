// Dispatch-cost model of the sm_100 SMSP for the instruction mixes the evaluation kernel can be written in.
// Every kernel runs 16 warps per SM on all SMs and reports cycles per warp-instruction-group per SMSP
// (one "group" = the instruction mix named in the label).  Not part of the product.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ubench3 tools/ubench3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// A body is one asm block holding R repetitions of a mix over 8 independent float chains v0..v7 (in/out) with
// constants a,b (in) and predicates p0..p3 prepared once per block from the integer k.
#define REGS "+f"(v0), "+f"(v1), "+f"(v2), "+f"(v3), "+f"(v4), "+f"(v5), "+f"(v6), "+f"(v7)
#define INS "f"(a), "f"(b), "r"(k)
#define PRE "{\n.reg .pred p0, p1, p2, p3, q0, q1, q2, q3;\n.reg .f32 t0, t1, t2, t3;\n.reg .b32 i0, i1, i2, i3;\n.reg .b64 w0, w1, w2, w3, wa, wb;\n" \
            "setp.ne.s32 p0, %10, 0;\nsetp.ne.s32 p1, %10, 1;\nsetp.ne.s32 p2, %10, 2;\nsetp.ne.s32 p3, %10, 3;\n" \
            "mov.b64 w0, {%0, %1};\nmov.b64 w1, {%2, %3};\nmov.b64 w2, {%1, %0};\nmov.b64 w3, {%3, %2};\nmov.b64 wa, {%8, %8};\nmov.b64 wb, {%9, %9};\n"
#define POST "mov.b64 {%0, %1}, w0;\nmov.b64 {%2, %3}, w1;\nmov.b64 {t0, t1}, w2;\nmov.b64 {t2, t3}, w3;\nadd.rn.f32 %0, %0, t0;\nadd.rn.f32 %1, %1, t1;\nadd.rn.f32 %2, %2, t2;\nadd.rn.f32 %3, %3, t3;\n}\n"
#define POSTS "}\n"

#define X4(s) s s s s
#define X8(s) X4(s) X4(s)

// scalar FP
#define M_FFMA "fma.rn.f32 %0, %0, %8, %9;\nfma.rn.f32 %1, %1, %8, %9;\nfma.rn.f32 %2, %2, %8, %9;\nfma.rn.f32 %3, %3, %8, %9;\n"
#define M_FFMA_B "fma.rn.f32 %4, %4, %8, %9;\nfma.rn.f32 %5, %5, %8, %9;\nfma.rn.f32 %6, %6, %8, %9;\nfma.rn.f32 %7, %7, %8, %9;\n"
#define M_FADD "add.rn.f32 %0, %0, %8;\nadd.rn.f32 %1, %1, %8;\nadd.rn.f32 %2, %2, %8;\nadd.rn.f32 %3, %3, %8;\n"
#define M_PFADD "@p0 add.rn.f32 %0, %0, %8;\n@p1 add.rn.f32 %1, %1, %8;\n@p2 add.rn.f32 %2, %2, %8;\n@p3 add.rn.f32 %3, %3, %8;\n"
#define M_PFFMA "@p0 fma.rn.f32 %0, %4, %8, %0;\n@p1 fma.rn.f32 %1, %5, %8, %1;\n@p2 fma.rn.f32 %2, %6, %8, %2;\n@p3 fma.rn.f32 %3, %7, %8, %3;\n"
// packed FP (on w0..w3)
#define M_FFMA2 "fma.rn.f32x2 w0, w0, wa, wb;\nfma.rn.f32x2 w1, w1, wa, wb;\nfma.rn.f32x2 w2, w2, wa, wb;\nfma.rn.f32x2 w3, w3, wa, wb;\n"
// ALU class
#define M_FSEL "selp.f32 %4, %5, %4, p0;\nselp.f32 %5, %6, %5, p1;\nselp.f32 %6, %7, %6, p2;\nselp.f32 %7, %4, %7, p3;\n"
#define M_FSETP "setp.lt.f32 q0, %0, %4;\nsetp.lt.f32 q1, %1, %5;\nsetp.lt.f32 q2, %2, %6;\nsetp.lt.f32 q3, %3, %7;\n"
#define M_FSETP_ABS "{\n.reg .f32 aa, bb;\nabs.f32 aa, %0;\nabs.f32 bb, %4;\nsetp.lt.f32 q0, aa, bb;\nabs.f32 aa, %1;\nabs.f32 bb, %5;\nsetp.lt.f32 q1, aa, bb;\nabs.f32 aa, %2;\nabs.f32 bb, %6;\nsetp.lt.f32 q2, aa, bb;\nabs.f32 aa, %3;\nabs.f32 bb, %7;\nsetp.lt.f32 q3, aa, bb;\n}\n"
#define M_FSETP_AND "setp.lt.and.f32 q0, %0, %4, p1;\nsetp.lt.and.f32 q1, %1, %5, p2;\nsetp.lt.and.f32 q2, %2, %6, p3;\nsetp.lt.and.f32 q3, %3, %7, p0;\n"
#define M_QFADD "@q0 add.rn.f32 %4, %4, %8;\n@q1 add.rn.f32 %5, %5, %8;\n@q2 add.rn.f32 %6, %6, %8;\n@q3 add.rn.f32 %7, %7, %8;\n"
#define M_QFSEL "selp.f32 %4, %5, %4, q0;\nselp.f32 %5, %6, %5, q1;\nselp.f32 %6, %7, %6, q2;\nselp.f32 %7, %4, %7, q3;\n"
#define M_FSET "set.lt.f32.f32 t0, %0, %4;\nset.lt.f32.f32 t1, %1, %5;\nset.lt.f32.f32 t2, %2, %6;\nset.lt.f32.f32 t3, %3, %7;\nadd.rn.f32 %4, %4, t0;\nadd.rn.f32 %5, %5, t1;\nadd.rn.f32 %6, %6, t2;\nadd.rn.f32 %7, %7, t3;\n"
#define M_FMNMX "min.f32 %4, %4, %0;\nmin.f32 %5, %5, %1;\nmin.f32 %6, %6, %2;\nmin.f32 %7, %7, %3;\n"
#define M_FMNMX3 "min.f32 %4, %4, %0, %1;\nmin.f32 %5, %5, %1, %2;\nmin.f32 %6, %6, %2, %3;\nmin.f32 %7, %7, %3, %0;\n"
#define M_PLOP "and.pred q0, q1, p0;\nor.pred q1, q2, p1;\nand.pred q2, q3, p2;\nor.pred q3, q0, p3;\n"
#define M_IADD "mov.b32 i0, %4;\nadd.s32 i0, i0, i1;\nmov.b32 %4, i0;\nmov.b32 i1, %5;\nadd.s32 i1, i1, i2;\nmov.b32 %5, i1;\nmov.b32 i2, %6;\nadd.s32 i2, i2, i3;\nmov.b32 %6, i2;\nmov.b32 i3, %7;\nadd.s32 i3, i3, i0;\nmov.b32 %7, i3;\n"
#define M_IMAD "mov.b32 i0, %4;\nmad.lo.s32 i0, i0, %10, %10;\nmov.b32 %4, i0;\nmov.b32 i1, %5;\nmad.lo.s32 i1, i1, %10, %10;\nmov.b32 %5, i1;\nmov.b32 i2, %6;\nmad.lo.s32 i2, i2, %10, %10;\nmov.b32 %6, i2;\nmov.b32 i3, %7;\nmad.lo.s32 i3, i3, %10, %10;\nmov.b32 %7, i3;\n"
#define M_LOP "mov.b32 i0, %4;\nmov.b32 i1, %5;\nmov.b32 i2, %6;\nmov.b32 i3, %7;\nlop3.b32 i0, i0, i1, i2, 0xe8;\nlop3.b32 i1, i1, i2, i3, 0xe8;\nlop3.b32 i2, i2, i3, i0, 0xe8;\nlop3.b32 i3, i3, i0, i1, 0xe8;\nmov.b32 %4, i0;\nmov.b32 %5, i1;\nmov.b32 %6, i2;\nmov.b32 %7, i3;\n"

#define KERNEL(name, body, post)                                                                     \
  __global__ void __launch_bounds__(512, 1) name(float* out, int iters, float a, float b, int k) {  \
    float v0 = threadIdx.x * 1e-3f, v1 = v0 + 1.f, v2 = v0 + 2.f, v3 = v0 + 3.f, v4 = v0 + 4.f, v5 = v0 + 5.f, v6 = v0 + 6.f, v7 = v0 + 7.f; \
    _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { asm volatile(PRE body post : REGS : INS); }                                       \
    out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;                                                        \
  }

// each kernel body = 8 x (mix); "n" in the table below = instructions per mix
KERNEL(k_ffma, X8(M_FFMA M_FFMA_B), POSTS)                       // 8 FFMA
KERNEL(k_fadd, X8(M_FADD), POSTS)                                // 4 FADD
KERNEL(k_pfadd, X8(M_PFADD), POSTS)                              // 4 predicated FADD
KERNEL(k_pffma, X8(M_PFFMA), POSTS)                              // 4 predicated FFMA
KERNEL(k_ffma2, X8(M_FFMA2), POST)                               // 4 FFMA2
KERNEL(k_fsel, X8(M_FSEL), POSTS)                                // 4 FSEL
KERNEL(k_fsetp_fadd, X8(M_FSETP M_QFADD), POSTS)                 // 4 FSETP + 4 @FADD
KERNEL(k_fsetpabs_fadd, X8(M_FSETP_ABS M_QFADD), POSTS)          // 4 FSETP(|a|,|b|) + 4 @FADD
KERNEL(k_fsetpand_fadd, X8(M_FSETP_AND M_QFADD), POSTS)          // 4 FSETP.AND + 4 @FADD
KERNEL(k_fsetp_fsel, X8(M_FSETP M_QFSEL), POSTS)                 // 4 FSETP + 4 FSEL
KERNEL(k_fset_fadd, X8(M_FSET), POSTS)                           // 4 FSET + 4 FADD
KERNEL(k_fmnmx, X8(M_FMNMX), POSTS)                              // 4 FMNMX
KERNEL(k_fmnmx3, X8(M_FMNMX3), POSTS)                            // 4 FMNMX3
KERNEL(k_plop_fadd, X8(M_FSETP M_PLOP M_QFADD), POSTS)           // 4 FSETP + 4 PLOP3 + 4 @FADD
KERNEL(k_iadd, X8(M_IADD), POSTS)                                // 4 IADD
KERNEL(k_imad, X8(M_IMAD), POSTS)                                // 4 IMAD
KERNEL(k_lop, X8(M_LOP), POSTS)                                  // 4 LOP3
KERNEL(k_mix_1f1s, X8(M_FFMA M_FSEL), POSTS)                     // 4 FFMA + 4 FSEL
KERNEL(k_mix_2f1s, X8(M_FFMA M_FSEL M_FADD), POSTS)              // 8 FP + 4 FSEL
KERNEL(k_mix_3f1s, X8(M_FFMA M_FSEL M_FADD M_PFFMA), POSTS)      // 12 FP + 4 FSEL
KERNEL(k_mix_4f1s, X4(M_FFMA M_FSEL M_FADD M_PFFMA M_PFADD), POSTS)  // 16 FP + 4 FSEL (4 reps)
KERNEL(k_mix_1p1s, X8(M_FFMA2 M_FSEL), POST)                     // 4 FFMA2 + 4 FSEL
KERNEL(k_mix_2p1s, X8(M_FFMA2 M_FSEL M_FFMA2), POST)             // 8 FFMA2 + 4 FSEL
KERNEL(k_mix_1p1f, X8(M_FFMA2 M_FFMA_B), POST)                   // 4 FFMA2 + 4 FFMA
KERNEL(k_mix_1i1f, X8(M_IMAD M_FFMA), POSTS)                     // 4 IMAD + 4 FFMA
KERNEL(k_mix_1i1s, X8(M_IMAD M_FSEL), POSTS)                     // 4 IMAD + 4 FSEL (FSEL touches the same regs: see SASS)

// shared-memory loads mixed with FP: 3 LDS.32 per 16 / 32 FFMA
__global__ void __launch_bounds__(512, 1) k_lds_mix(float* out, int iters, float a, float b, int k, int nf) {
  __shared__ float sm[512 * 3 * 4];
  for (int i = threadIdx.x; i < 512 * 12; i += 512) sm[i] = i * 1e-6f;
  __syncthreads();
  float v0 = threadIdx.x * 1e-3f, v1 = v0 + 1.f, v2 = v0 + 2.f, v3 = v0 + 3.f;
  uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 12;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float x, y, z;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(base + r * 6144));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y) : "r"(base + r * 6144 + 4));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(z) : "r"(base + r * 6144 + 8));
      if (nf == 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { v0 = fmaf(v0, x, b); v1 = fmaf(v1, y, b); v2 = fmaf(v2, z, b); v3 = fmaf(v3, a, x); }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) { v0 = fmaf(v0, x, b); v1 = fmaf(v1, y, b); v2 = fmaf(v2, z, b); v3 = fmaf(v3, a, x); }
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3;
}

template <class F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int i = 0; i < 3; ++i) f(); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3;
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); double clk = khz * 1e3;
  float* out; CK(cudaMalloc(&out, 1 << 24));
  const int iters = 4096;
  printf("device %s, %d SMs, clock %.0f MHz (nominal; cycles below assume it)\n", prop.name, sms, clk / 1e6);
  printf("%-18s %-34s %10s %12s\n", "kernel", "mix (per group)", "cyc/group", "cyc/instr");
  for (int threads : {512, 1024}) {
    printf("-- %d threads per SM (%d warps per SMSP)\n", threads, threads / 128);
#define RUN(K, reps, ninstr, label) { float ms = time_ms([&] { K<<<sms, threads>>>(out, iters, 1.0001f, 0.5f, 5); }); \
      double cyc = ms * 1e-3 * clk / ((double)iters * reps * (threads / 128)); printf("%-18s %-34s %10.2f %12.3f\n", #K, label, cyc, cyc / ninstr); }
    RUN(k_ffma, 8, 8, "8 FFMA") RUN(k_fadd, 8, 4, "4 FADD") RUN(k_pfadd, 8, 4, "4 @p FADD") RUN(k_pffma, 8, 4, "4 @p FFMA")
    RUN(k_ffma2, 8, 4, "4 FFMA2") RUN(k_fsel, 8, 4, "4 FSEL") RUN(k_fsetp_fadd, 8, 8, "4 FSETP + 4 @q FADD")
    RUN(k_fsetpabs_fadd, 8, 8, "4 FSETP|.| + 4 @q FADD") RUN(k_fsetpand_fadd, 8, 8, "4 FSETP.AND + 4 @q FADD")
    RUN(k_fsetp_fsel, 8, 8, "4 FSETP + 4 FSEL") RUN(k_fset_fadd, 8, 8, "4 FSET + 4 FADD") RUN(k_fmnmx, 8, 4, "4 FMNMX")
    RUN(k_fmnmx3, 8, 4, "4 FMNMX3") RUN(k_plop_fadd, 8, 12, "4 FSETP + 4 PLOP3 + 4 @q FADD")
    RUN(k_iadd, 8, 4, "4 IADD") RUN(k_imad, 8, 4, "4 IMAD") RUN(k_lop, 8, 4, "4 LOP3")
    RUN(k_mix_1f1s, 8, 8, "4 FFMA + 4 FSEL") RUN(k_mix_2f1s, 8, 12, "8 FP + 4 FSEL") RUN(k_mix_3f1s, 8, 16, "12 FP + 4 FSEL")
    RUN(k_mix_4f1s, 4, 20, "16 FP + 4 FSEL") RUN(k_mix_1p1s, 8, 8, "4 FFMA2 + 4 FSEL") RUN(k_mix_2p1s, 8, 12, "8 FFMA2 + 4 FSEL")
    RUN(k_mix_1p1f, 8, 8, "4 FFMA2 + 4 FFMA") RUN(k_mix_1i1f, 8, 8, "4 IMAD + 4 FFMA") RUN(k_mix_1i1s, 8, 8, "4 IMAD + 4 FSEL")
    for (int nf : {16, 32}) {
      float ms = time_ms([&] { k_lds_mix<<<sms, threads>>>(out, iters, 1.0001f, 0.5f, 5, nf); });
      double cyc = ms * 1e-3 * clk / ((double)iters * 4 * (threads / 128));
      printf("%-18s 3 LDS.32 + %d FFMA %*s %10.2f %12.3f\n", "k_lds_mix", nf, 14, "", cyc, cyc / (3 + nf));
    }
  }
  return 0;
}
