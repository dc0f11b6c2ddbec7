// Per-point arithmetic of the cuboid-sums throughput kernel (k_eval.cuh), kept in a header so that the micro-benchmarks under
// tools/ time exactly the code the product runs.
#pragma once

namespace hsk {

struct RoomK {  // one room's constants in registers (uniform across the block)
  float n[3][3], dp[3], dm[3];
};

// Float chains of one thread since the last flush
struct ChainsP {
  float f, T[3], M[3], B[3][3], C1, C2, Cm[3];
  __device__ __forceinline__ void clear() {
    f = C1 = C2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T[j] = M[j] = Cm[j] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) B[j][c] = 0.f;
    }
  }
};

// One point.  Written as one PTX block so that every accumulation is a predicated FP instruction (the C++ front end is free to
// turn `if (E) acc += v` into selects).  The axis minimum comes from FMNMX, and sum r^2 takes |r| of the nearest wall straight
// from the minimum (|r|^2 == r^2 exactly).  Forms measured and rejected on B200: tools/retired/ (packed f32x2, FP-pipe selects).
__device__ __forceinline__ void add_point(ChainsP& c, const RoomK& R, float x, float y, float z) {
  asm("{\n"
      ".reg .pred P, Q1, E0, E1, E2;\n"
      ".reg .f32 a, b, t, sp, sm, asp, asm_, s0, s1, s2, p0, p1, p2, a0, a1, a2, a01;\n"
      // axis 0
      "mul.rn.f32 a, %24, %21;\n mul.rn.f32 b, %25, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %26, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 sp, t, %33;\n add.rn.f32 sm, t, %36;\n abs.f32 asp, sp;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n selp.f32 s0, sm, sp, P;\n selp.f32 p0, 0f3F800000, 0f00000000, P;\n"
      // axis 1
      "mul.rn.f32 a, %27, %21;\n mul.rn.f32 b, %28, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %29, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 sp, t, %34;\n add.rn.f32 sm, t, %37;\n abs.f32 asp, sp;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n selp.f32 s1, sm, sp, P;\n selp.f32 p1, 0f3F800000, 0f00000000, P;\n"
      // axis 2
      "mul.rn.f32 a, %30, %21;\n mul.rn.f32 b, %31, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %32, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 sp, t, %35;\n add.rn.f32 sm, t, %38;\n abs.f32 asp, sp;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n selp.f32 s2, sm, sp, P;\n selp.f32 p2, 0f3F800000, 0f00000000, P;\n"
      // nearest axis, sequential first-minimum semantics (NaN compares false and keeps the earlier wall)
      "abs.f32 a0, s0;\n abs.f32 a1, s1;\n abs.f32 a2, s2;\n"
      "min.f32 a01, a0, a1;\n setp.lt.f32 E2, a2, a01;\n min.f32 a01, a01, a2;\n fma.rn.f32 %0, a01, a01, %0;\n"
      "setp.lt.and.f32 E1, a1, a0, !E2;\n setp.geu.and.f32 E0, a1, a0, !E2;\n"
      // predicated accumulation
      "@E0 add.rn.f32 %1, %1, s0;\n @E0 fma.rn.f32 %4, s0, p0, %4;\n @E0 add.rn.f32 %18, %18, p0;\n"
      "@E0 fma.rn.f32 %7, s0, %21, %7;\n @E0 fma.rn.f32 %8, s0, %22, %8;\n @E0 fma.rn.f32 %9, s0, %23, %9;\n"
      "@E1 add.rn.f32 %2, %2, s1;\n @E1 fma.rn.f32 %5, s1, p1, %5;\n @E1 add.rn.f32 %19, %19, p1;\n"
      "@E1 fma.rn.f32 %10, s1, %21, %10;\n @E1 fma.rn.f32 %11, s1, %22, %11;\n @E1 fma.rn.f32 %12, s1, %23, %12;\n @E1 add.rn.f32 %16, %16, 0f3F800000;\n"
      "@E2 add.rn.f32 %3, %3, s2;\n @E2 fma.rn.f32 %6, s2, p2, %6;\n @E2 add.rn.f32 %20, %20, p2;\n"
      "@E2 fma.rn.f32 %13, s2, %21, %13;\n @E2 fma.rn.f32 %14, s2, %22, %14;\n @E2 fma.rn.f32 %15, s2, %23, %15;\n @E2 add.rn.f32 %17, %17, 0f3F800000;\n"
      "}\n"
      : "+f"(c.f), "+f"(c.T[0]), "+f"(c.T[1]), "+f"(c.T[2]), "+f"(c.M[0]), "+f"(c.M[1]), "+f"(c.M[2]),              // 0..6
        "+f"(c.B[0][0]), "+f"(c.B[0][1]), "+f"(c.B[0][2]), "+f"(c.B[1][0]), "+f"(c.B[1][1]), "+f"(c.B[1][2]),       // 7..12
        "+f"(c.B[2][0]), "+f"(c.B[2][1]), "+f"(c.B[2][2]), "+f"(c.C1), "+f"(c.C2),                                  // 13..17
        "+f"(c.Cm[0]), "+f"(c.Cm[1]), "+f"(c.Cm[2])                                                                 // 18..20
      : "f"(x), "f"(y), "f"(z),                                                                                      // 21..23
        "f"(R.n[0][0]), "f"(R.n[0][1]), "f"(R.n[0][2]), "f"(R.n[1][0]), "f"(R.n[1][1]), "f"(R.n[1][2]),             // 24..29
        "f"(R.n[2][0]), "f"(R.n[2][1]), "f"(R.n[2][2]),                                                              // 30..32
        "f"(R.dp[0]), "f"(R.dp[1]), "f"(R.dp[2]), "f"(R.dm[0]), "f"(R.dm[1]), "f"(R.dm[2]));                        // 33..38
}

}  // namespace hsk
