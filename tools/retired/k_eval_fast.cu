// Cuboid objective / gradient sums per room — the throughput kernel (A6, 12 B/point, HBM-bound target).
//
// Same result record as k_rooms_cuboid_sums<AccExact> (k_planes.cu) and the same bit-exact plane assignment, but laid
// out for the B200 pipes (measured in profiles/r01_ubench.log: ALU 64, FP32 128 scalar / 256 packed thread-ops/clk/SM):
//   * point tiles (NCONS groups x 48 B) are streamed global -> shared by a producer warp with 1-D bulk async copies
//     (cp.async.bulk / UBLKCP, the TMA engine) through a 4-stage mbarrier ring; consumers read 3 x LDS.128 (conflict free);
//   * the 6 cuboid planes come in antiparallel pairs (n, d+), (-n, d-): one Float dot product per axis serves both walls,
//     ((nx*x + ny*y) + nz*z) exactly as signedDistanceToPlaneEq (Main.hs:1371-1372), -(t) - d- == -(t + d-) exactly;
//   * two points ride in every packed f32x2 instruction (FFMA2: a*b+(-0) and a*1+b are the correctly rounded product / sum,
//     so nothing is contracted); comparisons produce 0/1 floats (FSET) and every select / one-hot is an exact
//     multiply-add with those indicators, keeping the scarce ALU pipe at ~5 ops per point;
//   * per-thread Float chains (64 points) are flushed into per-thread Double accumulators; block partials are summed
//     by the last block in block order (deterministic for a fixed grid).
#include <cstring>

#include "k_common.cuh"

namespace hsk {

constexpr int EV_STAGES = 4;
constexpr int EV_FLUSH_TILES = 15;  // Float chain length: 15 tiles x 4 points per thread (6-bit count fields hold 60)

struct PairedTable {
  int32_t nrooms;
  float one;  // 1.0f, deliberately a run-time value (see add_pair)
  int64_t off[HS_MAX_ROOMS + 1];
  float n[HS_MAX_ROOMS][3][3];  // normal of the + wall of each axis
  float dp[HS_MAX_ROOMS][3];    // d of the + wall
  float dm[HS_MAX_ROOMS][3];    // d of the - wall (whose normal is -n)
};

// Packed pairs live in 64-bit registers from birth to death (inline PTX on .b64) so ptxas keeps them in aligned
// register pairs instead of re-assembling float2 halves with MOVs around every packed instruction.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// volatile: exactly one register-pair assembly per use site (the optimiser otherwise clones the cheap-looking pack next to
// every consumer, which costs ~30 MOVs per point)
__device__ __forceinline__ f2 pack2v(float lo, float hi) { f2 r; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 bc2(float a) { return pack2(a, a); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// c - p*b: p in {0,1} -> exactly c or c - b.  The negation is written per half so ptxas folds it into the operand modifier.
__device__ __forceinline__ f2 fnma2(f2 p, f2 b, f2 c) {
  f2 d;
  asm("{\n.reg .f32 lo, hi;\n.reg .b64 np;\nmov.b64 {lo, hi}, %1;\nneg.f32 lo, lo;\nneg.f32 hi, hi;\nmov.b64 np, {lo, hi};\n"
      "fma.rn.f32x2 %0, np, %2, %3;\n}" : "=l"(d) : "l"(p), "l"(b), "l"(c));
  return d;
}
// 1.0 where |a| < |b| (strict: ties keep the lower plane index), else 0.0
__device__ __forceinline__ f2 lt_abs2(f2 a, f2 b) {
  float ax, ay, bx, by;
  unpack2(a, ax, ay);
  unpack2(b, bx, by);
  return pack2(fabsf(ax) < fabsf(bx) ? 1.0f : 0.0f, fabsf(ay) < fabsf(by) ? 1.0f : 0.0f);
}
// p ? a : b for p in {0,1}: b - p*b is exactly b or 0, p*a + that is exactly a or b
__device__ __forceinline__ f2 sel2(f2 p, f2 a, f2 b) { return fma2(p, a, fnma2(p, b, b)); }

struct RoomRegs {
  f2 nx[3], ny[3], nz[3], ndp[3], dm[3], one;
};

struct Chains {  // Float partial sums, lane lo / hi = the two points of a pair
  f2 f, T[3], M[3], B[3][3];
  f2 E[2], C[3];     // VAR_ARITH only: points on axis 1 / axis 2, and on the - wall of each axis, as exact Float counts
  unsigned int cnt;  // VAR_PRED only: 5 x 6-bit fields, points assigned to walls 0..4 since the last flush (wall 5 = the rest)
  __device__ __forceinline__ void clear() {
    f = bc2(0.f);
    cnt = 0u;
    E[0] = E[1] = bc2(0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T[j] = M[j] = C[j] = bc2(0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) B[j][c] = bc2(0.f);
    }
  }
};

// One lane (= one point) of the selection network.  Inputs: sp[j] = t_j - d+_j (distance to the + wall), sm[j] = t_j + d-_j
// (minus the distance to the - wall).  First minimum of |distance| over walls 0..5 = +x -x +y -y +z -z with strict
// comparisons (ties keep the lower index).  Outputs: sf = s of the nearest wall, z[j] = sf on the nearest wall's axis else 0,
// pf[j] = 1.0 if the nearer wall of axis j is its - wall (used as an exact 0/1 multiplier), and the packed count increment.
// Every predicate dies right after it is produced: only 7 predicate registers exist and 4 lanes are in flight.
__device__ __forceinline__ void select_lane(const float (&sp)[3], const float (&sm)[3], float& sf, float (&z)[3], float (&pf)[3], unsigned int& cnt) {
  float s[3];
  unsigned int inc[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const bool P = fabsf(sm[j]) < fabsf(sp[j]);
    s[j] = P ? sm[j] : sp[j];
    inc[j] = P ? (j < 2 ? 1u << (12 * j + 6) : 0u) : (1u << (12 * j));
    pf[j] = fabsf(sm[j]) < fabsf(sp[j]) ? 1.0f : 0.0f;  // FSET: a second compare is cheaper than keeping P alive
  }
  const bool Q1 = fabsf(s[1]) < fabsf(s[0]);
  const float s01 = Q1 ? s[1] : s[0];
  const unsigned int inc01 = Q1 ? inc[1] : inc[0];
  const bool Q2 = fabsf(s[2]) < fabsf(s01);
  sf = Q2 ? s[2] : s01;
  cnt += Q2 ? inc[2] : inc01;
  const float z01 = Q2 ? 0.0f : s01;
  z[2] = Q2 ? s[2] : 0.0f;
  z[1] = Q1 ? z01 : 0.0f;
  z[0] = Q1 ? 0.0f : z01;
}

// Two points per call.  Distances and every accumulation are packed f32x2 (FMA pipe, 2 lanes per issue slot); the
// selection network runs per lane on the ALU pipe (FSETP/FSEL/SEL/FSET) — the split keeps both pipes near 2:1, their width ratio.
__device__ __forceinline__ void add_pair_pred(Chains& c, const RoomRegs& R, f2 x, f2 y, f2 z) {
  float spl[3], sph[3], sml[3], smh[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    // ((nx*x + ny*y) + nz*z) with three roundings per sum as in the reference.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2
    // into one FFMA2 (checked in SASS, even with --fmad=false), so each sum is written p*ONE + q with ONE = 1.0f read from
    // the kernel parameters: fl(p*1 + q) == fl(p + q), and an FMA with two genuine products cannot be contracted further.
    const f2 t = fma2(fma2(mul2(R.nx[j], x), R.one, mul2(R.ny[j], y)), R.one, mul2(R.nz[j], z));
    unpack2(add2(t, R.ndp[j]), spl[j], sph[j]);  // t - d+
    unpack2(add2(t, R.dm[j]), sml[j], smh[j]);   // t + d-  == -((-t) - d-)
  }
  float sfl, sfh, zl[3], zh[3], pfl[3], pfh[3];
  select_lane(spl, sml, sfl, zl, pfl, c.cnt);
  select_lane(sph, smh, sfh, zh, pfh, c.cnt);
  const f2 sf = pack2(sfl, sfh);
  c.f = fma2(sf, sf, c.f);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const f2 zj = pack2(zl[j], zh[j]);
    c.T[j] = add2(c.T[j], zj);
    c.M[j] = fma2(zj, pack2(pfl[j], pfh[j]), c.M[j]);  // z * {0,1}: exact
    c.B[j][0] = fma2(zj, x, c.B[j][0]);
    c.B[j][1] = fma2(zj, y, c.B[j][1]);
    c.B[j][2] = fma2(zj, z, c.B[j][2]);
  }
}

// The same two points with the selection network written as exact arithmetic on 0/1 indicator floats: only the five
// comparisons per point touch the ALU pipe (FSET), everything else is packed FFMA2 work.  An ALU instruction costs two
// issue cycles per 32 points, a packed one half a cycle, so this form is bound by the FP32 pipe (~57 lane-ops per point).
__device__ __forceinline__ void add_pair_arith(Chains& c, const RoomRegs& R, f2 x, f2 y, f2 z) {
  f2 s[3], pf[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const f2 t = fma2(fma2(mul2(R.nx[j], x), R.one, mul2(R.ny[j], y)), R.one, mul2(R.nz[j], z));  // see add_pair_pred
    const f2 sp = add2(t, R.ndp[j]);
    const f2 sm = add2(t, R.dm[j]);
    pf[j] = lt_abs2(sm, sp);
    s[j] = sel2(pf[j], sm, sp);
  }
  const f2 q1 = lt_abs2(s[1], s[0]);
  const f2 s01 = sel2(q1, s[1], s[0]);
  const f2 q2 = lt_abs2(s[2], s01);
  f2 zz[3];
  zz[2] = mul2(q2, s[2]);
  const f2 z01 = fnma2(q2, s01, s01);  // s01 unless axis 2 wins
  zz[1] = mul2(q1, z01);
  zz[0] = sub2(z01, zz[1]);            // exact: one of the two is zero
  const f2 e1 = fnma2(q1, q2, q1);     // q1 * (1 - q2)
  c.E[0] = add2(c.E[0], e1);
  c.E[1] = add2(c.E[1], q2);
  const f2 e0 = sub2(sub2(bc2(1.0f), q2), e1);
  c.C[0] = fma2(e0, pf[0], c.C[0]);
  c.C[1] = fma2(e1, pf[1], c.C[1]);
  c.C[2] = fma2(q2, pf[2], c.C[2]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    c.f = fma2(zz[j], zz[j], c.f);
    c.T[j] = add2(c.T[j], zz[j]);
    c.M[j] = fma2(zz[j], pf[j], c.M[j]);
    c.B[j][0] = fma2(zz[j], x, c.B[j][0]);
    c.B[j][1] = fma2(zz[j], y, c.B[j][1]);
    c.B[j][2] = fma2(zz[j], z, c.B[j][2]);
  }
}

// Selection by minimum instead of compare+select chains: FMNMX (|.| operand modifiers, full rate) produces the per-axis and
// cross-axis minima of |distance|, five FSET turn them into exact 0/1 indicators (e0: axis 0 wins ties, e2: axis 2 only on a
// strict minimum, pf_j: - wall only when strictly nearer), and everything else is packed FP32 arithmetic on those indicators.
__device__ __forceinline__ void mnmx_lane(const float (&sp)[3], const float (&sm)[3], float& e0, float& e2, float (&pf)[3], float& m) {
  float a[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    a[j] = fminf(fabsf(sp[j]), fabsf(sm[j]));
    pf[j] = fabsf(sm[j]) < fabsf(sp[j]) ? 1.0f : 0.0f;
  }
  const float m12 = fminf(a[1], a[2]), m01 = fminf(a[0], a[1]);
  e0 = a[0] <= m12 ? 1.0f : 0.0f;
  e2 = a[2] < m01 ? 1.0f : 0.0f;
  m = fminf(m01, a[2]);
}
__device__ __forceinline__ void add_pair_mnmx(Chains& c, const RoomRegs& R, f2 x, f2 y, f2 z) {
  f2 sp[3], sm[3];
  float spl[3], sph[3], sml[3], smh[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const f2 t = fma2(fma2(mul2(R.nx[j], x), R.one, mul2(R.ny[j], y)), R.one, mul2(R.nz[j], z));  // see add_pair_pred
    sp[j] = add2(t, R.ndp[j]);
    sm[j] = add2(t, R.dm[j]);
    unpack2(sp[j], spl[j], sph[j]);
    unpack2(sm[j], sml[j], smh[j]);
  }
  float e0l, e0h, e2l, e2h, pfl[3], pfh[3], ml, mh;
  mnmx_lane(spl, sml, e0l, e2l, pfl, ml);
  mnmx_lane(sph, smh, e0h, e2h, pfh, mh);
  f2 e[3], pf[3];
  e[0] = pack2(e0l, e0h);
  e[2] = pack2(e2l, e2h);
  e[1] = sub2(sub2(R.one, e[0]), e[2]);
  const f2 m = pack2(ml, mh);
  c.f = fma2(m, m, c.f);  // |r|^2 of the nearest wall
  c.E[0] = add2(c.E[0], e[1]);
  c.E[1] = add2(c.E[1], e[2]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    pf[j] = pack2(pfl[j], pfh[j]);
    const f2 zj = mul2(e[j], sel2(pf[j], sm[j], sp[j]));
    c.C[j] = fma2(e[j], pf[j], c.C[j]);
    c.T[j] = add2(c.T[j], zj);
    c.M[j] = fma2(zj, pf[j], c.M[j]);
    c.B[j][0] = fma2(zj, x, c.B[j][0]);
    c.B[j][1] = fma2(zj, y, c.B[j][1]);
    c.B[j][2] = fma2(zj, z, c.B[j][2]);
  }
}

enum { VAR_PRED = 0, VAR_ARITH = 1, VAR_MNMX = 2 };
template <int VAR>
__device__ __forceinline__ void add_pair(Chains& c, const RoomRegs& R, f2 x, f2 y, f2 z) {
  if (VAR == VAR_ARITH) add_pair_arith(c, R, x, y, z);
  else if (VAR == VAR_MNMX) add_pair_mnmx(c, R, x, y, z);
  else add_pair_pred(c, R, x, y, z);
}

// Per-thread Double accumulators live in shared memory (slot c of thread t at acc[c * NCONS + t]: conflict-free, private,
// no synchronisation) so the register file is left to the packed Float chains; they are touched once per 60 points.
template <int NCONS, int VAR>
__device__ __forceinline__ void flush_chains(Chains& c, double* acc, int npoints) {
  auto d = [](f2 a) { float lo, hi; unpack2(a, lo, hi); return static_cast<double>(lo) + static_cast<double>(hi); };
  double* a = acc + threadIdx.x;
  a[0] += d(c.f);
  if (VAR == VAR_PRED) {
    int rest = npoints;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int ck = static_cast<int>((c.cnt >> (6 * k)) & 63u);
      rest -= ck;
      a[(16 + k) * NCONS] += static_cast<double>(ck);
    }
    a[21 * NCONS] += static_cast<double>(rest);
  } else {
    const double E1 = d(c.E[0]), E2 = d(c.E[1]), C0 = d(c.C[0]), C1 = d(c.C[1]), C2 = d(c.C[2]);
    const double E0 = static_cast<double>(npoints) - E1 - E2;
    a[16 * NCONS] += E0 - C0; a[17 * NCONS] += C0;
    a[18 * NCONS] += E1 - C1; a[19 * NCONS] += C1;
    a[20 * NCONS] += E2 - C2; a[21 * NCONS] += C2;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double T = d(c.T[j]), M = d(c.M[j]);
    a[(1 + 2 * j) * NCONS] += T - M;  // sum of r over the + wall (r = s)
    a[(2 + 2 * j) * NCONS] -= M;      // sum of r over the - wall (r = -s)
#pragma unroll
    for (int q = 0; q < 3; ++q) a[(7 + 3 * j + q) * NCONS] += d(c.B[j][q]);
  }
  c.clear();
}

// scalar edge path (ragged points at room / chunk borders, partial tiles): exact Double products, same record
template <int NCONS>
__device__ __noinline__ void add_point_exact(double* acc, const float* pl /* 6 x 4 */, float x, float y, float z) {
  float rb = plane_dist(pl[0], pl[1], pl[2], pl[3], x, y, z);
  float ab = fabsf(rb);
  int kb = 0;
#pragma unroll
  for (int k = 1; k < 6; ++k) {
    const float rk = plane_dist(pl[4 * k], pl[4 * k + 1], pl[4 * k + 2], pl[4 * k + 3], x, y, z);
    const float ak = fabsf(rk);
    const bool lt = ak < ab;
    ab = lt ? ak : ab; rb = lt ? rk : rb; kb = lt ? k : kb;
  }
  const double rd = rb, s = (kb & 1) ? -rd : rd;
  const int j = kb >> 1;
  double* a = acc + threadIdx.x;
  a[0] = fma(rd, rd, a[0]);
  a[(1 + kb) * NCONS] += rd;
  a[(16 + kb) * NCONS] += 1.0;
  a[(7 + 3 * j + 0) * NCONS] = fma(s, static_cast<double>(x), a[(7 + 3 * j + 0) * NCONS]);
  a[(7 + 3 * j + 1) * NCONS] = fma(s, static_cast<double>(y), a[(7 + 3 * j + 1) * NCONS]);
  a[(7 + 3 * j + 2) * NCONS] = fma(s, static_cast<double>(z), a[(7 + 3 * j + 2) * NCONS]);
}

// ---- mbarrier / bulk-copy primitives --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
// Producer-side wait: poll, then sleep.  The producer lane waits almost all the time; a bare try_wait loop re-issues every few
// cycles and takes issue slots from the four consumer warps that share its scheduler, and the ring makes every other warp
// of the block wait for those four (ncu: the producer's SMSP executed 7.5 % more instructions than the others).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* b, uint32_t parity, unsigned int ns) {
  uint32_t ok;
  for (;;) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// shared-window (32-bit address) flavours for the hot loop
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }

template <int NCONS>
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// block b owns groups [b*gpb, (b+1)*gpb); per overlapping room: stream whole-group tiles through the ring.
template <int NCONS, int VAR, int TPI, int STG>
__global__ void __launch_bounds__(NCONS + 32, (NCONS <= 160 ? 3 : NCONS <= 256 ? 2 : 1))
k_rooms_cuboid_sums_fast(const float* __restrict__ xyz, int64_t n, const __grid_constant__ PairedTable tbl, int64_t gpb,
                         double* __restrict__ partials, int* __restrict__ meta, unsigned int* ticket, double* __restrict__ out, unsigned int psleep) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* tiles = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STG) * NCONS * 48);
  uint64_t* empty = full + STG;
  double* acc = reinterpret_cast<double*>(empty + STG);     // [HS_NACC][NCONS] per-thread Double accumulators
  double* red = acc + static_cast<size_t>(HS_NACC) * NCONS;       // [NCONS/32][HS_NACC]
  float* spl = reinterpret_cast<float*>(red + (NCONS / 32) * HS_NACC);  // 6 x 4 planes of the current room (edge path)
  __shared__ bool is_last;

  const int nrooms = tbl.nrooms;
  const int64_t G = (n + 3) >> 2;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * gpb;
  const int64_t g1 = min(g0 + gpb, G);
  const int64_t p0 = g0 * 4, p1 = min(g1 * 4, n);
  int rfirst = -1, rlast = -2;
  for (int r = 0; r < nrooms; ++r)
    if (tbl.off[r] < p1 && tbl.off[r + 1] > p0) { if (rfirst < 0) rfirst = r; rlast = r; }

  if (threadIdx.x == 0) {
    meta[blockIdx.x] = rfirst;
    for (int s = 0; s < STG; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const bool producer = threadIdx.x >= NCONS;
  if (producer) {
    // ---------------- producer warp: one lane issues the bulk copies, same tile order as the consumers
    if (threadIdx.x == NCONS) {
      int64_t tt = 0;
      for (int r = rfirst; r <= rlast; ++r) {
        const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
        const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
        for (int64_t tg = gl; tg < gh; tg += NCONS, ++tt) {
          const int s = static_cast<int>(tt % STG);
          if (tt >= STG) { if (psleep) mbar_wait_sleep(empty + s, static_cast<uint32_t>(((tt / STG) - 1) & 1), psleep); else mbar_wait(empty + s, static_cast<uint32_t>(((tt / STG) - 1) & 1)); }
          const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(NCONS), gh - tg) * 48);
          mbar_expect_tx(full + s, bytes);
          bulk_g2s(tiles + static_cast<size_t>(s) * NCONS * 3, reinterpret_cast<const float4*>(xyz) + 3 * tg, bytes, full + s);
        }
      }
    }
    return;
  }

  // ---------------- consumers
  int64_t tt = 0;
  for (int r = rfirst; r <= rlast; ++r) {
    const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
    RoomRegs R;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      R.nx[j] = bc2(tbl.n[r][j][0]); R.ny[j] = bc2(tbl.n[r][j][1]); R.nz[j] = bc2(tbl.n[r][j][2]);
      R.ndp[j] = bc2(-tbl.dp[r][j]); R.dm[j] = bc2(tbl.dm[r][j]); R.one = bc2(tbl.one);
    }
    if (threadIdx.x < 6) {
      const int j = threadIdx.x >> 1;
      const bool minus = threadIdx.x & 1;
      spl[4 * threadIdx.x + 0] = minus ? -tbl.n[r][j][0] : tbl.n[r][j][0];
      spl[4 * threadIdx.x + 1] = minus ? -tbl.n[r][j][1] : tbl.n[r][j][1];
      spl[4 * threadIdx.x + 2] = minus ? -tbl.n[r][j][2] : tbl.n[r][j][2];
      spl[4 * threadIdx.x + 3] = minus ? tbl.dm[r][j] : tbl.dp[r][j];
    }
#pragma unroll
    for (int i = 0; i < HS_NACC; ++i) acc[i * NCONS + threadIdx.x] = 0.0;
    consumers_sync<NCONS>();
    Chains ch;
    ch.clear();
    const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;
    if (gl <= gh) {
      const int64_t head_end = gl * 4, tail_begin = gh * 4;
      const int64_t nh = head_end - lo, ntail = hi - tail_begin;
      if (threadIdx.x < nh) { const int64_t i = lo + threadIdx.x; add_point_exact<NCONS>(acc, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      else if (threadIdx.x >= 32 && threadIdx.x - 32 < ntail) { const int64_t i = tail_begin + threadIdx.x - 32; add_point_exact<NCONS>(acc, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      // full tiles: the hot loop.  Ring position is carried as (byte offset of the stage, parity) so the loop has no
      // 64-bit index arithmetic; `tt` only survives to keep producer and consumers on the same global tile count.
      const int64_t ngroups = gh - gl;
      const int nfull = static_cast<int>(ngroups / NCONS);
      const int rem_groups = static_cast<int>(ngroups - static_cast<int64_t>(nfull) * NCONS);
      constexpr uint32_t TILE_BYTES = NCONS * 48;
      uint32_t stage = static_cast<uint32_t>(tt % STG);
      uint32_t parity = static_cast<uint32_t>((tt / STG) & 1);
      const uint32_t tiles_s = smem_u32(tiles) + threadIdx.x * 12, full_s = smem_u32(full), empty_s = smem_u32(empty);
      int since_flush = 0;
      int t = 0;
      // One trip = TPI tiles = 4*TPI points per thread.  (A register-level software pipeline of the next trip's LDS was
      // measured and bought nothing: the kernel is bound by instruction dispatch, not by waits at tile boundaries.)
      for (; t + TPI <= nfull; t += TPI) {
        float px[4 * TPI], py[4 * TPI], pz[4 * TPI];
#pragma unroll
        for (int u = 0; u < TPI; ++u) {
          mbar_wait_s(full_s + 8 * stage, parity);
          // this thread's 4 points of the tile: i, i+N, i+2N, i+3N (N = NCONS).  Scalar LDS.32 at a 3-word lane stride are
          // bank-conflict free and land the same coordinate of two points in adjacent registers = a packed operand for free.
          const uint32_t base = tiles_s + stage * TILE_BYTES;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            px[4 * u + e] = lds_f32(base + e * NCONS * 12);
            py[4 * u + e] = lds_f32(base + e * NCONS * 12 + 4);
            pz[4 * u + e] = lds_f32(base + e * NCONS * 12 + 8);
          }
          mbar_arrive_s(empty_s + 8 * stage);  // the values are in registers: hand the slot back before the math
          if (++stage == STG) { stage = 0; parity ^= 1u; }
        }
#pragma unroll
        for (int e = 0; e < 4 * TPI; e += 2)
          add_pair<VAR>(ch, R, pack2(px[e], px[e + 1]), pack2(py[e], py[e + 1]), pack2(pz[e], pz[e + 1]));
        since_flush += TPI;
        if (since_flush >= EV_FLUSH_TILES - TPI + 1) { flush_chains<NCONS, VAR>(ch, acc, 4 * since_flush); since_flush = 0; }
      }
      for (; t < nfull; ++t) {  // leftover single tiles when TPI > 1
        mbar_wait_s(full_s + 8 * stage, parity);
        const uint32_t base = tiles_s + stage * TILE_BYTES;
        float px[4], py[4], pz[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          px[e] = lds_f32(base + e * NCONS * 12);
          py[e] = lds_f32(base + e * NCONS * 12 + 4);
          pz[e] = lds_f32(base + e * NCONS * 12 + 8);
        }
        mbar_arrive_s(empty_s + 8 * stage);
        if (++stage == STG) { stage = 0; parity ^= 1u; }
        add_pair<VAR>(ch, R, pack2(px[0], px[1]), pack2(py[0], py[1]), pack2(pz[0], pz[1]));
        add_pair<VAR>(ch, R, pack2(px[2], px[3]), pack2(py[2], py[3]), pack2(pz[2], pz[3]));
        ++since_flush;
      }
      flush_chains<NCONS, VAR>(ch, acc, 4 * since_flush);
      tt += nfull;
      if (rem_groups) {  // partial last tile of the room segment: exact scalar path for the in-range points
        mbar_wait_s(full_s + 8 * stage, parity);
        const float* tile = reinterpret_cast<const float*>(tiles) + static_cast<size_t>(stage) * NCONS * 12;
        const int npts = rem_groups * 4;
        float px[4], py[4], pz[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {  // out-of-range slots read stale but valid shared memory and are masked below
          const int i = threadIdx.x + e * NCONS;
          px[e] = tile[3 * i]; py[e] = tile[3 * i + 1]; pz[e] = tile[3 * i + 2];
        }
        mbar_arrive_s(empty_s + 8 * stage);
        for (int e = 0; e < 4; ++e)
          if (static_cast<int>(threadIdx.x) + e * NCONS < npts) add_point_exact<NCONS>(acc, spl, px[e], py[e], pz[e]);
        ++tt;
      }
    } else {
      const int64_t i = lo + threadIdx.x;
      if (i < hi) add_point_exact<NCONS>(acc, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
    // consumer-only deterministic block reduction
    {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
      for (int c = 0; c < HS_NACC; ++c) {
        const double sum = warp_sum(acc[c * NCONS + threadIdx.x]);
        if (lane == 0) red[warp * HS_NACC + c] = sum;
      }
      consumers_sync<NCONS>();
      if (threadIdx.x < HS_NACC) {
        double sum = 0;
#pragma unroll
        for (int w = 0; w < NCONS / 32; ++w) sum += red[w * HS_NACC + threadIdx.x];
        partials[(static_cast<int64_t>(blockIdx.x) * nrooms + (r - rfirst)) * HS_NACC + threadIdx.x] = sum;
      }
      consumers_sync<NCONS>();
    }
  }

  // ---------------- last block sums the partials per room in block order
  __threadfence();
  consumers_sync<NCONS>();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  consumers_sync<NCONS>();
  if (!is_last) return;
  __threadfence();
  const int64_t ppb = gpb * 4;
  for (int o = threadIdx.x; o < nrooms * HS_REC; o += NCONS) {
    const int r = o / HS_REC, c = o % HS_REC;
    double s = 0.0;
    if (c < HS_NACC && tbl.off[r + 1] > tbl.off[r]) {
      const int64_t b_lo = tbl.off[r] / ppb, b_hi = (tbl.off[r + 1] - 1) / ppb;
      for (int64_t b = b_lo; b <= b_hi; ++b) {
        const int slot = r - __ldcg(meta + b);
        s += __ldcg(partials + (b * nrooms + slot) * HS_NACC + c);
      }
    }
    out[o] = s;
  }
}

}  // namespace hsk

using namespace hsk;

template <int NCONS, int VAR, int TPI, int STG = EV_STAGES>
static int32_t launch_fast_t(hs_ctx* ctx, const float* xyz, int64_t n, const PairedTable& tbl, double* d_rec_out) {
  const int64_t G = (n + 3) >> 2;
  const int per_sm = ctx->modes[HS_MODE_BLOCKS_PER_SM] > 0 ? ctx->modes[HS_MODE_BLOCKS_PER_SM] : 1;
  int64_t nb = static_cast<int64_t>(ctx->sm_count) * per_sm;
  const int64_t cap = (G + NCONS - 1) / NCONS;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  const int64_t gpb = (G + nb - 1) / nb > 0 ? (G + nb - 1) / nb : 1;
  const size_t need = static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double) + static_cast<size_t>(nb) * sizeof(int) + 64;
  if (int32_t rc = hs_ensure_scratch(ctx, need)) return rc;
  double* partials = reinterpret_cast<double*>(ctx->d_scratch);
  int* meta = reinterpret_cast<int*>(ctx->d_scratch + static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double));
  const size_t smem = static_cast<size_t>(STG) * NCONS * 48 + 2 * STG * 8 + static_cast<size_t>(HS_NACC) * NCONS * 8 +
                      static_cast<size_t>(NCONS / 32) * HS_NACC * 8 + 6 * 4 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_rooms_cuboid_sums_fast<NCONS, VAR, TPI, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  k_rooms_cuboid_sums_fast<NCONS, VAR, TPI, STG><<<static_cast<int>(nb), NCONS + 32, smem, ctx->stream>>>(xyz, n, tbl, gpb, partials, meta, ctx->d_ticket, d_rec_out, static_cast<unsigned int>(ctx->modes[HS_MODE_PRODUCER_SLEEP]));
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

// returns HS_EINVAL-free: caller guarantees the planes are paired
int32_t launch_rooms_cuboid_sums_fast(hs_ctx* ctx, const float* xyz, int64_t n, const RoomTable& rt, double* d_rec_out) {
  PairedTable t;
  memset(&t, 0, sizeof t);
  t.nrooms = rt.nrooms;
  t.one = 1.0f;
  for (int r = 0; r <= rt.nrooms; ++r) t.off[r] = rt.off[r];
  for (int r = 0; r < rt.nrooms; ++r)
    for (int j = 0; j < 3; ++j) {
      for (int c = 0; c < 3; ++c) t.n[r][j][c] = rt.pl[r][2 * j][c];
      t.dp[r][j] = rt.pl[r][2 * j][3];
      t.dm[r][j] = rt.pl[r][2 * j + 1][3];
    }
  // default: ALU-select variant, 16 consumer warps + the producer warp per SM (measured fastest, profiles/r01_*.txt);
  // mode key 3 = 1 selects the all-arithmetic variant (15 consumer warps so that each scheduler holds 4 warps at 128 regs)
  if (ctx->modes[HS_MODE_EVAL_VARIANT] == 1) return launch_fast_t<480, VAR_ARITH, 1>(ctx, xyz, n, t, d_rec_out);
  if (ctx->modes[HS_MODE_EVAL_VARIANT] == 4) {
    if (ctx->modes[HS_MODE_EVAL_CONSUMERS] == 256) return launch_fast_t<256, VAR_MNMX, 1>(ctx, xyz, n, t, d_rec_out);
    if (ctx->modes[HS_MODE_EVAL_CONSUMERS] == 480) return launch_fast_t<480, VAR_MNMX, 1>(ctx, xyz, n, t, d_rec_out);
    return launch_fast_t<512, VAR_MNMX, 1>(ctx, xyz, n, t, d_rec_out);
  }
  switch (ctx->modes[HS_MODE_EVAL_CONSUMERS]) {  // ring-depth experiments (tools/prof_eval.py --cons): consumers * 10 + stages
    case 256: return launch_fast_t<256, VAR_PRED, 1>(ctx, xyz, n, t, d_rec_out);  // with mode key 1 = 2: two independent rings per SM
    case 5123: return launch_fast_t<512, VAR_PRED, 1, 3>(ctx, xyz, n, t, d_rec_out);
    case 4484: return launch_fast_t<448, VAR_PRED, 1, 4>(ctx, xyz, n, t, d_rec_out);
    case 4486: return launch_fast_t<448, VAR_PRED, 1, 6>(ctx, xyz, n, t, d_rec_out);
    case 3844: return launch_fast_t<384, VAR_PRED, 1, 4>(ctx, xyz, n, t, d_rec_out);
    case 3848: return launch_fast_t<384, VAR_PRED, 1, 8>(ctx, xyz, n, t, d_rec_out);
    default: break;
  }
  if (ctx->modes[HS_MODE_EVAL_TPI] == 2) return launch_fast_t<480, VAR_PRED, 2>(ctx, xyz, n, t, d_rec_out);
  return launch_fast_t<512, VAR_PRED, 1>(ctx, xyz, n, t, d_rec_out);
}
