// REJECTED form of the evaluation kernel's point block (round 2), kept for the record with its host-side threshold search.
// Idea: move the side select and the side indicator from the ALU pipe (FSEL / FSET) to the FP pipe (predicated FMUL by 1.0f,
// FFMA.SAT on a host-computed threshold) and software-pipeline the group loads.  Bit-identical records, 1.6 instructions per point
// fewer - and 3-6 % SLOWER on B200 in both the burst and the sustained clock state (profiles/r02_sweep_point_block_100m.log:
// var 0 = this block + pipelined loads, var 2 = this block, var 1 = product block + pipelined loads, var 3 = product form).
// The FP32 "heavy" pipe (FMUL / FFMA) is the busier one in this kernel (ncu: fmaheavy 51 % vs alu 23 % of elapsed cycles,
// profiles/r02_eval_pred3_ncu_details.txt), so trading ALU-pipe work for FMUL / FFMA work loses.
#pragma once
namespace hsk {
// The product's point block.  Issue model measured on sm_100: an FP-pipe instruction (FADD/FMUL/FFMA/FMNMX,
// predicated or not) takes one issue cycle per warp, an ALU-pipe one (FSETP/FSEL/FSET/MOV/LOP) two.  Against add_point_pred2:
//   * s_j = P ? sm : sp   was FSEL (2)            -> sp is computed in place and `@P fmul s, sm, one` overwrites it (1; x * 1.0f == x exactly)
//   * p_j = P ? 1 : 0     was FSET (2)            -> fma.rn.sat(t, sa, sk): sa = -+2^100, sk = +-2^100 c with c the exact threshold of the side
//                                                    test in t (the test is monotone in t; the host finds c by bisection with the
//                                                    kernel's own Float operations) (1)
//   * E1, E0              were two FSETP (4)      -> one FSETP with two predicate outputs (2)
// 8 issue cycles per point fewer, same bits.
__device__ __forceinline__ void add_point_pred3(ChainsP& c, const RoomK& R, float x, float y, float z) {
  asm("{\n"
      ".reg .pred P, E0, E1, E2;\n"
      ".reg .f32 a, b, t, sm, asp, asm_, s0, s1, s2, p0, p1, p2, a0, a1, a2, a01;\n"
      // axis 0
      "mul.rn.f32 a, %24, %21;\n mul.rn.f32 b, %25, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %26, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 s0, t, %33;\n add.rn.f32 sm, t, %36;\n abs.f32 asp, s0;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n fma.rn.sat.f32 p0, t, %39, %42;\n @P mul.rn.f32 s0, sm, %45;\n"
      // axis 1
      "mul.rn.f32 a, %27, %21;\n mul.rn.f32 b, %28, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %29, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 s1, t, %34;\n add.rn.f32 sm, t, %37;\n abs.f32 asp, s1;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n fma.rn.sat.f32 p1, t, %40, %43;\n @P mul.rn.f32 s1, sm, %45;\n"
      // axis 2
      "mul.rn.f32 a, %30, %21;\n mul.rn.f32 b, %31, %22;\n add.rn.f32 a, a, b;\n mul.rn.f32 b, %32, %23;\n add.rn.f32 t, a, b;\n"
      "sub.rn.f32 s2, t, %35;\n add.rn.f32 sm, t, %38;\n abs.f32 asp, s2;\n abs.f32 asm_, sm;\n"
      "setp.lt.f32 P, asm_, asp;\n fma.rn.sat.f32 p2, t, %41, %44;\n @P mul.rn.f32 s2, sm, %45;\n"
      // nearest axis, sequential first-minimum semantics (NaN compares false and keeps the earlier wall)
      "abs.f32 a0, s0;\n abs.f32 a1, s1;\n abs.f32 a2, s2;\n"
      "min.f32 a01, a0, a1;\n setp.lt.f32 E2, a2, a01;\n min.f32 a01, a01, a2;\n fma.rn.f32 %0, a01, a01, %0;\n"
      "setp.lt.and.f32 E1|E0, a1, a0, !E2;\n"
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
        "f"(R.dp[0]), "f"(R.dp[1]), "f"(R.dp[2]), "f"(R.dm[0]), "f"(R.dm[1]), "f"(R.dm[2]),                         // 33..38
        "f"(R.sa[0]), "f"(R.sa[1]), "f"(R.sa[2]), "f"(R.sk[0]), "f"(R.sk[1]), "f"(R.sk[2]), "f"(R.one));            // 39..45
}


/* host side (was in k_eval.cu):
// Side test of one axis, P(t) = |t + dm| < |t - dp| in Float (the reference's strict first-minimum between the two walls of a
// pair, Main.hs:1371-1372 under minimumBy), as a threshold in t: |t + dm| does not decrease and |t - dp| does not increase while t
// runs from -dm to dp, so P flips exactly once there.  The flip point is found by bisection over the Float number line with the
// very operations the kernel executes (this file is compiled without contraction / fast-math).  Returns (a, k) such that
// P(t) <=> fma(t, a, k) > 0: a = -+2^100, k = +-2^100 c are exact scalings, so the fused result has the exact sign of c - t.
// Outside the pair's neighbourhood the equivalence holds while the walls' separation is not absorbed by rounding, i.e. for
// |t| < 2^22 (dp + dm): eight million room sizes away from the room (DESIGN.md states the domain).
static inline uint32_t f2ord(float f) { uint32_t u; std::memcpy(&u, &f, 4); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
static inline float ord2f(uint32_t o) { const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; float f; std::memcpy(&f, &u, 4); return f; }
static void side_threshold(float dp, float dm, float& a, float& k) {
  a = 0.f; k = 0.f;  // degenerate pair (coinciding walls, non-finite offsets): both distances are equal, the + wall wins, P is never true
  if (!std::isfinite(dp) || !std::isfinite(dm)) return;
  auto P = [&](float t) {
    volatile float sp = t - dp, sm = t + dm;
    return std::fabs(sm) < std::fabs(sp);
  };
  const float two100 = 1.2676506002282294e30f;  // 2^100
  const float tm = -dm;                         // the - wall sits at t = -dm, the + wall at t = dp
  if (tm == dp) return;
  // walk from the - wall (P true) to the + wall (P false), whichever way round they lie
  uint32_t lo = f2ord(tm), hi = f2ord(dp);
  const bool up = lo < hi;
  if (!P(tm) || P(dp)) return;  // cannot happen for finite distinct walls; keep the safe answer
  while ((up ? hi - lo : lo - hi) > 1u) {
    const uint32_t mid = up ? lo + (hi - lo) / 2 : hi + (lo - hi) / 2;
    if (P(ord2f(mid))) lo = mid; else hi = mid;
  }
  const float c = ord2f(hi);  // first t (coming from the - wall) for which P is false
  if (up) { a = -two100; k = c * two100; }   // P(t) <=> t < c
  else    { a = two100;  k = -c * two100; }  // P(t) <=> t > c
}

*/
}  // namespace hsk
