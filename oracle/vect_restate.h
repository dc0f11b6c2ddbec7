/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Restatement of the arithmetic of the `vect` Haskell package (dependency
 * `vect >= 0.4.7`, housescan.cabal:29) that the reference's hot path calls.
 * The package source is NOT mounted under /root/reference, so these are the
 * published definitions as recalled: PARITY UNPINNED at this
 * boundary except through the reference's own self-consistency tests
 * (FitCuboidBFGS.hs:134-140, Main.hs:1881, Main.hs:2637) which tests/ ports.
 *
 * Everything is instantiated twice: T=float  (Data.Vect.Float,  Main.hs:39-40)
 *                                   T=double (Data.Vect.Double, FitCuboidBFGS.hs:20-21)
 * Compile with -ffp-contract=off: GHC's x86-64 Float/Double code has no FMA.
 * Operator association follows Haskell: `a*b + c*d + e*f` == ((a*b)+(c*d))+(e*f).
 */
#ifndef ORACLE_VECT_RESTATE_H
#define ORACLE_VECT_RESTATE_H
#include <math.h>

#define VECT_DEFINE(T, S, SQRT, SIN, COS)                                                  \
  typedef struct { T x, y, z; } v3##S;                                                     \
  typedef struct { T x, y, z, w; } v4##S;                                                  \
  typedef struct { v3##S r0, r1, r2; } m3##S; /* rows */                                   \
  static inline v3##S v3##S##_mk(T x, T y, T z) { v3##S v = {x, y, z}; return v; }         \
  /* (&+) (&-) */                                                                          \
  static inline v3##S v3##S##_add(v3##S a, v3##S b) { return v3##S##_mk(a.x + b.x, a.y + b.y, a.z + b.z); } \
  static inline v3##S v3##S##_sub(v3##S a, v3##S b) { return v3##S##_mk(a.x - b.x, a.y - b.y, a.z - b.z); } \
  /* (*&) / (&*) scalarMul */                                                              \
  static inline v3##S v3##S##_scale(T s, v3##S a) { return v3##S##_mk(s * a.x, s * a.y, s * a.z); } \
  /* dotprod: x1*x2 + y1*y2 + z1*z2 */                                                     \
  static inline T v3##S##_dot(v3##S a, v3##S b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; } \
  static inline T v4##S##_dot(v4##S a, v4##S b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; } \
  static inline T v3##S##_normsqr(v3##S a) { return v3##S##_dot(a, a); }                   \
  static inline T v3##S##_norm(v3##S a) { return SQRT(v3##S##_normsqr(a)); }               \
  static inline T v3##S##_distance(v3##S a, v3##S b) { return v3##S##_norm(v3##S##_sub(a, b)); } \
  /* normalize v = v &* (1 / norm v); mkNormal = normalize */                              \
  static inline v3##S v3##S##_normalize(v3##S a) { T r = (T)1 / v3##S##_norm(a); return v3##S##_mk(a.x * r, a.y * r, a.z * r); } \
  static inline v4##S v4##S##_normalize(v4##S a) { T r = (T)1 / SQRT(v4##S##_dot(a, a));   \
    v4##S o = {a.x * r, a.y * r, a.z * r, a.w * r}; return o; }                            \
  /* crossprod */                                                                          \
  static inline v3##S v3##S##_cross(v3##S a, v3##S b) {                                    \
    return v3##S##_mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } \
  /* (.*) row-vector times matrix: component c = v . column c */                           \
  static inline v3##S v3##S##_lmul(v3##S v, m3##S m) {                                     \
    return v3##S##_mk((v.x * m.r0.x + v.y * m.r1.x) + v.z * m.r2.x,                        \
                      (v.x * m.r0.y + v.y * m.r1.y) + v.z * m.r2.y,                        \
                      (v.x * m.r0.z + v.y * m.r1.z) + v.z * m.r2.z); }                     \
  static inline m3##S m3##S##_transpose(m3##S m) { m3##S t = {                             \
      {m.r0.x, m.r1.x, m.r2.x}, {m.r0.y, m.r1.y, m.r2.y}, {m.r0.z, m.r1.z, m.r2.z}}; return t; } \
  /* leftOrthoU (U (Vec4 a b c d)); rightOrthoU = transpose . leftOrthoU; mkU = normalize */ \
  static inline m3##S m3##S##_left_ortho_u(v4##S q) {                                      \
    T a = q.x, b = q.y, c = q.z, d = q.w; m3##S m = {                                      \
      {((a * a + b * b) - c * c) - d * d, (2 * b) * c - (2 * a) * d, (2 * b) * d + (2 * a) * c}, \
      {(2 * b) * c + (2 * a) * d, ((a * a - b * b) + c * c) - d * d, (2 * c) * d - (2 * a) * b}, \
      {(2 * b) * d - (2 * a) * c, (2 * c) * d + (2 * a) * b, ((a * a - b * b) - c * c) + d * d}}; \
    return m; }                                                                            \
  static inline m3##S m3##S##_right_ortho_u(v4##S q) { return m3##S##_transpose(m3##S##_left_ortho_u(q)); } \
  /* rotMatrix3' (unit axis v) a = (1-c) * outer v v  +  [[c, s z, -s y],[-s z, c, s x],[s y, -s x, c]] */ \
  static inline m3##S m3##S##_rot3_unit(v3##S v, T ang) {                                  \
    T c = COS(ang), s = SIN(ang), k = (T)1 - c; m3##S m = {                                \
      {k * (v.x * v.x) + c,        k * (v.x * v.y) + s * v.z,    k * (v.x * v.z) + (-(s * v.y))}, \
      {k * (v.y * v.x) + (-(s * v.z)), k * (v.y * v.y) + c,      k * (v.y * v.z) + s * v.x}, \
      {k * (v.z * v.x) + s * v.y,  k * (v.z * v.y) + (-(s * v.x)), k * (v.z * v.z) + c}};  \
    return m; }                                                                            \
  /* rotMatrix3 v a = rotMatrix3' (mkNormal v) a */                                        \
  static inline m3##S m3##S##_rot3(v3##S v, T ang) { return m3##S##_rot3_unit(v3##S##_normalize(v), ang); } \
  /* rotateAround c R p = ((p &- c) .* R) &+ c   (Main.hs:1582-1583, FitCuboidBFGS.hs:91-92) */ \
  static inline v3##S v3##S##_rotate_around(v3##S c, m3##S R, v3##S p) {                   \
    return v3##S##_add(v3##S##_lmul(v3##S##_sub(p, c), R), c); }

VECT_DEFINE(float, f, sqrtf, sinf, cosf)
VECT_DEFINE(double, d, sqrt, sin, cos)

#endif
