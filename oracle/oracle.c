/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the data-parallel point-cloud path of nh2/housescan
 * (SURVEY.md §8a rows A1..A13).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (housescan_b200/) never does.
 *
 * PARITY STATUS: the reference is a Haskell executable; no GHC exists in this image and
 * the reference does not build as mounted (HmatrixUtils missing), so this file cannot be
 * checked against outputs of the reference itself.  It is pinned against every check the
 * reference's own sources hold for this path (tests/test_oracle_golden.py):
 *   FitCuboidBFGS.hs:134-140 (cuboidFromParams identity), :29-41 (example box),
 *   Main.hs:1881 (4 corners per cuboid plane), Main.hs:2637 + projTest* (roomProj replay),
 *   Bijection.hs:10-15, TranslationOptimizer.hs:22-35, GroupConnectedComponents.hs:54.
 * For the point-cloud-scale generalisations the north-star adds (A4, A6 over clouds) the
 * reference has no code: "parity unpinned by reference", the oracle is a composition of
 * reference primitives.  `vect`/`hmatrix` arithmetic is restated from the published
 * definitions (vect_restate.h).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * All per-point Float math is sequential-order, non-FMA, exactly as GHC would emit.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "vect_restate.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline arm asks for all host threads explicitly */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}


/* ------------------------------------------------------------------------------------
 * A1-A3  depth frame -> points.   Main.hs:1296-1313
 *   V.imap (i,depth) -> Vec3 (fromIntegral x) (fromIntegral y) (fromIntegral depth), (y,x) = i `quotRem` width
 *   V.filter d /= 0            (order preserving)
 *   V.map scalePoints          x/10.0, y/10.0, d/20.0 - 30.0   (Float, true division)
 * Returns number of valid points; mask_out[i] = 1 iff depth[i] != 0.
 * ---------------------------------------------------------------------------------- */
ORC_API int64_t orc_backproject_ref(const uint16_t* depth, int w, int h, float* xyz_out, uint8_t* mask_out) {
  int64_t n = (int64_t)w * h, m = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t y = i / w, x = i % w;
    float fx = (float)x, fy = (float)y, fd = (float)depth[i];
    int valid = fd != 0.0f;
    if (mask_out) mask_out[i] = (uint8_t)valid;
    if (valid) {
      if (xyz_out) {
        xyz_out[3 * m + 0] = fx / 10.0f;
        xyz_out[3 * m + 1] = fy / 10.0f;
        xyz_out[3 * m + 2] = fd / 20.0f - 30.0f;
      }
      ++m;
    }
  }
  return m;
}

/* ------------------------------------------------------------------------------------
 * A5  signedDistanceToPlaneEq (PlaneEq n d) p = fromNormal n `dotprod` p - d   Main.hs:1371-1372
 * plane layout: 4 floats nx,ny,nz,d.
 * ---------------------------------------------------------------------------------- */
static inline float signed_distance(const float* pl, float px, float py, float pz) {
  return ((pl[0] * px + pl[1] * py) + pl[2] * pz) - pl[3];
}

/* Nearest plane = first minimum of |distance| (minimumBy semantics, FitCuboidBFGS.hs:74). */
static inline int nearest_plane(const float* planes, int K, float px, float py, float pz, float* r_out) {
  int best = 0;
  float rb = signed_distance(planes, px, py, pz), ab = fabsf(rb);
  for (int k = 1; k < K; ++k) {
    float r = signed_distance(planes + 4 * k, px, py, pz), a = fabsf(r);
    if (a < ab) { ab = a; rb = r; best = k; }
  }
  *r_out = rb;
  return best;
}

ORC_API void orc_plane_assign(const float* xyz, int64_t n, const float* planes, int K,
                              uint8_t* assign_out, float* resid_out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    float r;
    int k = nearest_plane(planes, K, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &r);
    if (assign_out) assign_out[i] = (uint8_t)k;
    if (resid_out) resid_out[i] = r;
  }
}

/* projectToPlane eq p = p &- (signedDistanceToPlaneEq eq p *& fromNormal n)   Main.hs:1375-1376 */
ORC_API void orc_project_to_plane(const float* xyz, int64_t n, const float* plane, float* out) {
  for (int64_t i = 0; i < n; ++i) {
    float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    float s = signed_distance(plane, px, py, pz);
    out[3 * i] = px - s * plane[0];
    out[3 * i + 1] = py - s * plane[1];
    out[3 * i + 2] = pz - s * plane[2];
  }
}

/* mkPlaneEq abc d = PlaneEq (mkNormal abc) (d / norm abc)    Main.hs:1360-1361 */
static void mk_plane_eq(v3f abc, float d, float* out) {
  v3f n = v3f_normalize(abc);
  out[0] = n.x; out[1] = n.y; out[2] = n.z; out[3] = d / v3f_norm(abc);
}
ORC_API void orc_mk_plane_eq(const float abc[3], float d, float out[4]) { mk_plane_eq(v3f_mk(abc[0], abc[1], abc[2]), d, out); }

/* rotatePlaneEqAround c R (PlaneEq n d)    Main.hs:1571-1578 */
static void rotate_plane_eq_around(v3f c, m3f R, const float* in, float* out) {
  v3f n = v3f_mk(in[0], in[1], in[2]);
  v3f n2 = v3f_lmul(n, R);
  v3f o = v3f_scale(in[3], n);
  v3f o2 = v3f_rotate_around(c, R, o);
  float d2 = v3f_dot(o2, n2);
  mk_plane_eq(n2, d2, out);
}
/* translatePlaneEq off (PlaneEq n d)    Main.hs:1681-1688 */
static void translate_plane_eq(v3f off, const float* in, float* out) {
  v3f n = v3f_mk(in[0], in[1], in[2]);
  v3f o = v3f_scale(in[3], n);
  v3f o2 = v3f_add(o, off);
  float d2 = v3f_dot(o2, n);
  mk_plane_eq(n, d2, out);
}
ORC_API void orc_rotate_plane_eq_around(const float c[3], const float R[9], const float in[4], float out[4]) {
  m3f m = {{R[0], R[1], R[2]}, {R[3], R[4], R[5]}, {R[6], R[7], R[8]}};
  rotate_plane_eq_around(v3f_mk(c[0], c[1], c[2]), m, in, out);
}
ORC_API void orc_translate_plane_eq(const float off[3], const float in[4], float out[4]) {
  translate_plane_eq(v3f_mk(off[0], off[1], off[2]), in, out);
}

/* ------------------------------------------------------------------------------------
 * Cuboid params -> 6 PlaneEq, Float.   Main.hs:1831-1836 (params -> Float, mkU) and
 * makePlanesFromCuboid Main.hs:1852-1874.   Order: +x -x +y -y +z -z.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_planes_from_cuboid(const double params[10], float planes[24]) {
  float p[10];
  for (int i = 0; i < 10; ++i) p[i] = (float)params[i];
  v3f center = v3f_mk(p[0], p[1], p[2]);
  v4f q = {p[6], p[7], p[8], p[9]};
  m3f R = m3f_right_ortho_u(v4f_normalize(q));
  const v3f zero = v3f_mk(0.0f, 0.0f, 0.0f);
  const float axes[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  for (int k = 0; k < 6; ++k) {
    float origin_eq[4], rot_eq[4];
    float half = p[3 + k / 2] / 2;
    mk_plane_eq(v3f_mk(axes[k][0], axes[k][1], axes[k][2]), half, origin_eq);
    rotate_plane_eq_around(zero, R, origin_eq, rot_eq);
    translate_plane_eq(center, rot_eq, planes + 4 * k);
  }
}

/* cuboidFromParams (Double).   FitCuboidBFGS.hs:98-112.   out = 8 corners x 3 */
ORC_API void orc_cuboid_from_params(const double p[10], double out[24]) {
  v4d q = {p[6], p[7], p[8], p[9]};
  m3d R = m3d_right_ortho_u(v4d_normalize(q));
  v3d c = v3d_mk(p[0], p[1], p[2]);
  int i = 0;
  for (int sx = -1; sx <= 1; sx += 2)
    for (int sy = -1; sy <= 1; sy += 2)
      for (int sz = -1; sz <= 1; sz += 2) {
        /* Vec3 (-a/2) .. : note (- a/2) == negate (a/2) */
        v3d v = v3d_mk(sx < 0 ? -(p[3] / 2) : p[3] / 2, sy < 0 ? -(p[4] / 2) : p[4] / 2, sz < 0 ? -(p[5] / 2) : p[5] / 2);
        v3d r = v3d_add(v3d_lmul(v, R), c);
        out[i++] = r.x; out[i++] = r.y; out[i++] = r.z;
      }
}
/* cuboidFromParamsRotateAround.   FitCuboidBFGS.hs:117-131 */
ORC_API void orc_cuboid_from_params_rotate_around(const double p[10], double out[24]) {
  v4d q = {p[6], p[7], p[8], p[9]};
  m3d R = m3d_right_ortho_u(v4d_normalize(q));
  v3d c = v3d_mk(p[0], p[1], p[2]);
  int i = 0;
  for (int sx = -1; sx <= 1; sx += 2)
    for (int sy = -1; sy <= 1; sy += 2)
      for (int sz = -1; sz <= 1; sz += 2) {
        v3d v = v3d_mk(sx < 0 ? p[0] - p[3] / 2 : p[0] + p[3] / 2, sy < 0 ? p[1] - p[4] / 2 : p[1] + p[4] / 2,
                       sz < 0 ? p[2] - p[5] / 2 : p[2] + p[5] / 2);
        v3d r = v3d_rotate_around(c, R, v);
        out[i++] = r.x; out[i++] = r.y; out[i++] = r.z;
      }
}
/* errfun: sum normsqr (p - e) over zipped corners.   FitCuboidBFGS.hs:51-65 */
ORC_API double orc_errfun(const double pts[24], const double params[10]) {
  double est[24], s = 0;
  orc_cuboid_from_params(params, est);
  for (int i = 0; i < 8; ++i) {
    v3d d = v3d_sub(v3d_mk(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), v3d_mk(est[3 * i], est[3 * i + 1], est[3 * i + 2]));
    s += v3d_normsqr(d);
  }
  return s;
}
/* errfunClosest: each point against its closest estimated corner (first minimum by distance).
 * FitCuboidBFGS.hs:68-76 */
ORC_API double orc_errfun_closest(const double* pts, int npts, const double params[10]) {
  double est[24], s = 0;
  orc_cuboid_from_params(params, est);
  for (int i = 0; i < npts; ++i) {
    v3d p = v3d_mk(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    int best = 0;
    double db = v3d_distance(p, v3d_mk(est[0], est[1], est[2]));
    for (int e = 1; e < 8; ++e) {
      double d = v3d_distance(p, v3d_mk(est[3 * e], est[3 * e + 1], est[3 * e + 2]));
      if (d < db) { db = d; best = e; }
    }
    s += v3d_normsqr(v3d_sub(p, v3d_mk(est[3 * best], est[3 * best + 1], est[3 * best + 2])));
  }
  return s;
}
/* guessDims.   FitCuboidBFGS.hs:247-252 */
static int cmp_double(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
ORC_API void orc_guess_dims(const double pts[24], double out[3]) {
  double d[7];
  v3d f = v3d_mk(pts[0], pts[1], pts[2]);
  for (int i = 1; i < 8; ++i) d[i - 1] = v3d_distance(f, v3d_mk(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
  qsort(d, 7, sizeof(double), cmp_double);
  out[0] = d[0]; out[1] = d[1];
  out[2] = sqrt(d[6] * d[6] - d[0] * d[0] - d[1] * d[1]);
}
/* rotMatrix3 (Double) for the example/generator rotations.  FitCuboidBFGS.hs:41,165 */
ORC_API void orc_rot_matrix3_d(const double axis[3], double ang, double out[9]) {
  m3d m = m3d_rot3(v3d_mk(axis[0], axis[1], axis[2]), ang);
  out[0] = m.r0.x; out[1] = m.r0.y; out[2] = m.r0.z; out[3] = m.r1.x; out[4] = m.r1.y; out[5] = m.r1.z;
  out[6] = m.r2.x; out[7] = m.r2.y; out[8] = m.r2.z;
}
ORC_API void orc_rot_matrix3_f(const float axis[3], float ang, float out[9]) {
  m3f m = m3f_rot3(v3f_mk(axis[0], axis[1], axis[2]), ang);
  out[0] = m.r0.x; out[1] = m.r0.y; out[2] = m.r0.z; out[3] = m.r1.x; out[4] = m.r1.y; out[5] = m.r1.z;
  out[6] = m.r2.x; out[7] = m.r2.y; out[8] = m.r2.z;
}
/* rotationBetweenPlaneEqs n1 n2 = rotMatrix3' (crossprod n1 n2 :: Normal3) (acos (n1.n2 / (|n1||n2|)))
 * Main.hs:1553-1560; crossprod on Normal3 re-normalises. */
ORC_API void orc_rotation_between_normals(const float n1[3], const float n2[3], float out[9]) {
  v3f a = v3f_mk(n1[0], n1[1], n1[2]), b = v3f_mk(n2[0], n2[1], n2[2]);
  v3f axis = v3f_normalize(v3f_cross(a, b));
  float costheta = v3f_dot(a, b) / (v3f_norm(a) * v3f_norm(b));
  m3f m = m3f_rot3_unit(axis, acosf(costheta));
  out[0] = m.r0.x; out[1] = m.r0.y; out[2] = m.r0.z; out[3] = m.r1.x; out[4] = m.r1.y; out[5] = m.r1.z;
  out[6] = m.r2.x; out[7] = m.r2.y; out[8] = m.r2.z;
}

/* ------------------------------------------------------------------------------------
 * A6 (generalisation; no reference code - composed of Main.hs:1852-1874 planes +
 * Main.hs:1371 distances + first-minimum assignment): cuboid objective over a cloud.
 *   r_i = signed Float distance of p_i to its nearest cuboid plane
 *   f   = sum (double) r_i^2
 * Raw sums (what the GPU kernel reduces), all accumulated in double from Float terms:
 *   rec[0]      f
 *   rec[1..6]   Sr[k]   = sum_{i in wall k} r_i
 *   rec[7..15]  B[j][c] = sum_{i on axis j} s_i * p_i[c],  s_i = r_i for the + wall, -r_i for the - wall
 *   rec[16..21] count[k]
 * ---------------------------------------------------------------------------------- */
#define ORC_REC 24
static void cuboid_sums_range(const float* xyz, int64_t i0, int64_t i1, const float* planes, double* rec) {
  for (int i = 0; i < ORC_REC; ++i) rec[i] = 0;
  for (int64_t i = i0; i < i1; ++i) {
    float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2], r;
    int k = nearest_plane(planes, 6, px, py, pz, &r);
    double rd = r, s = (k & 1) ? -rd : rd;
    int j = k >> 1;
    rec[0] += rd * rd;
    rec[1 + k] += rd;
    rec[7 + 3 * j + 0] += s * (double)px;
    rec[7 + 3 * j + 1] += s * (double)py;
    rec[7 + 3 * j + 2] += s * (double)pz;
    rec[16 + k] += 1.0;
  }
}
ORC_API void orc_cuboid_sums(const float* xyz, int64_t n, const double params[10], double rec[ORC_REC]) {
  float planes[24];
  orc_planes_from_cuboid(params, planes);
  int nt = orc_num_threads();
  double* part = (double*)calloc((size_t)nt * ORC_REC, sizeof(double));
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num(), T = omp_get_num_threads();
#else
    int t = 0, T = 1;
#endif
    int64_t i0 = n * t / T, i1 = n * (t + 1) / T;
    cuboid_sums_range(xyz, i0, i1, planes, part + (size_t)t * ORC_REC);
  }
  for (int i = 0; i < ORC_REC; ++i) { rec[i] = 0; for (int t = 0; t < nt; ++t) rec[i] += part[(size_t)t * ORC_REC + i]; }
  free(part);
}

/* d(row j of R)/dq_m for R = rightOrthoU (mkU q), in Double.  R = transpose(L(qhat)). */
static void drot_dq(const double q[4], double R[3][3], double dR[4][3][3]) {
  double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double a = q[0] / nq, b = q[1] / nq, c = q[2] / nq, d = q[3] / nq;
  /* L rows (leftOrthoU); R = L^T */
  double L[3][3] = {{a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c},
                    {2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b},
                    {2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d}};
  /* dL/d(unit component) */
  double dL[4][3][3] = {
      {{2 * a, -2 * d, 2 * c}, {2 * d, 2 * a, -2 * b}, {-2 * c, 2 * b, 2 * a}},
      {{2 * b, 2 * c, 2 * d}, {2 * c, -2 * b, -2 * a}, {2 * d, 2 * a, -2 * b}},
      {{-2 * c, 2 * b, 2 * a}, {2 * b, 2 * c, 2 * d}, {-2 * a, 2 * d, -2 * c}},
      {{-2 * d, -2 * a, 2 * b}, {2 * a, -2 * d, 2 * c}, {2 * b, 2 * c, 2 * d}}};
  double u[4] = {a, b, c, d};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = L[j][i];
  /* chain through normalisation: d qhat_n / d q_m = (delta_nm - qhat_n qhat_m) / |q| */
  for (int m = 0; m < 4; ++m)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double acc = 0;
        for (int n = 0; n < 4; ++n) acc += dL[n][j][i] * ((n == m ? 1.0 : 0.0) - u[n] * u[m]) / nq;
        dR[m][i][j] = acc;
      }
}

/* f, gradient (10), counts (6) by direct per-point accumulation in Double of
 *   d f / d theta = sum 2 r_i d r_i / d theta,   r_i = sigma R_j . (p_i - c) - dim_j / 2
 * with the Float assignment and the Float residual r_i of the reference primitives.
 * gscale[m] = sum |2 r_i d r_i / d theta_m|  (the magnitude tolerances are relative to). */
ORC_API void orc_cuboid_residual_grad(const float* xyz, int64_t n, const double params[10], double* f_out,
                                      double grad[10], int64_t counts[6], double gscale[10]) {
  float planes[24];
  orc_planes_from_cuboid(params, planes);
  double R[3][3], dR[4][3][3];
  drot_dq(params + 6, R, dR);
  double f = 0, g[10] = {0}, gs[10] = {0};
  int64_t cnt[6] = {0};
#pragma omp parallel
  {
    double fl = 0, gl[10] = {0}, gsl[10] = {0};
    int64_t cl[6] = {0};
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < n; ++i) {
      float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2], rf;
      int k = nearest_plane(planes, 6, px, py, pz, &rf);
      int j = k >> 1;
      double sg = (k & 1) ? -1.0 : 1.0, r = rf;
      double u[3] = {px - params[0], py - params[1], pz - params[2]};
      double t[10];
      for (int c = 0; c < 3; ++c) t[c] = -sg * R[j][c];
      for (int c = 0; c < 3; ++c) t[3 + c] = (c == j) ? -0.5 : 0.0;
      for (int m = 0; m < 4; ++m) t[6 + m] = sg * (dR[m][j][0] * u[0] + dR[m][j][1] * u[1] + dR[m][j][2] * u[2]);
      fl += r * r;
      for (int m = 0; m < 10; ++m) { double v = 2 * r * t[m]; gl[m] += v; gsl[m] += fabs(v); }
      cl[k]++;
    }
#pragma omp critical
    {
      f += fl;
      for (int m = 0; m < 10; ++m) { g[m] += gl[m]; gs[m] += gsl[m]; }
      for (int k = 0; k < 6; ++k) cnt[k] += cl[k];
    }
  }
  *f_out = f;
  for (int m = 0; m < 10; ++m) { grad[m] = g[m]; if (gscale) gscale[m] = gs[m]; }
  if (counts) for (int k = 0; k < 6; ++k) counts[k] = cnt[k];
}

/* Per-room, per-plane sums for wall alignment (A13 inputs generalised).
 * out[room][k][0..9] = count, sum r, sum r^2, sum p (3), sum r p (3), max |r|.  Generic K planes per room. */
#define ORC_PS 10
ORC_API void orc_plane_sums(const float* xyz, const int64_t* room_offsets, int nrooms, const float* planes, int K, double* out) {
  for (int r = 0; r < nrooms; ++r) {
    double* o = out + (size_t)r * K * ORC_PS;
    for (int i = 0; i < K * ORC_PS; ++i) o[i] = 0;
    const float* pl = planes + (size_t)r * K * 4;
    for (int64_t i = room_offsets[r]; i < room_offsets[r + 1]; ++i) {
      float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2], rf;
      int k = nearest_plane(pl, K, px, py, pz, &rf);
      double* a = o + k * ORC_PS, rd = rf;
      a[0] += 1; a[1] += rd; a[2] += rd * rd;
      a[3] += px; a[4] += py; a[5] += pz;
      a[6] += rd * px; a[7] += rd * py; a[8] += rd * pz;
      if (fabs(rd) > a[9]) a[9] = fabs(rd);
    }
  }
}

/* ------------------------------------------------------------------------------------
 * A4 (north-star addition; no reference code): per-frame point-to-plane normal equations.
 * Composition of A1-A3 back-projection (or pinhole when intr != NULL), optional pose
 * (row-vector x 4x4, translation in row 3: Main.hs:10, :1725-1730), A5 assignment.
 *   J = [crossprod p n, n]  (Float),  r Float;  accumulate in Double:
 *   out[0..20] upper triangle of J^T J row-major, out[21..26] J^T r, out[27] sum r^2, out[28] count.
 * ---------------------------------------------------------------------------------- */
static inline void pixel_to_point(int x, int y, uint16_t d, const float* intr, const float* pose, float p[3]) {
  float fx = (float)x, fy = (float)y, fd = (float)d;
  float X, Y, Z;
  if (intr) { /* pinhole, depth in millimetres: z = d * 0.001; X = (x - cx) * z / fx */
    Z = fd * 0.001f;
    X = ((fx - intr[2]) * Z) / intr[0];
    Y = ((fy - intr[3]) * Z) / intr[1];
  } else {   /* scalePoints, Main.hs:1311-1313 */
    X = fx / 10.0f; Y = fy / 10.0f; Z = fd / 20.0f - 30.0f;
  }
  if (pose) {
    p[0] = ((X * pose[0] + Y * pose[4]) + Z * pose[8]) + pose[12];
    p[1] = ((X * pose[1] + Y * pose[5]) + Z * pose[9]) + pose[13];
    p[2] = ((X * pose[2] + Y * pose[6]) + Z * pose[10]) + pose[14];
  } else { p[0] = X; p[1] = Y; p[2] = Z; }
}
ORC_API void orc_backproject_reduce6x6(const uint16_t* frames, int64_t nframes, int w, int h, const float* intr,
                                       const float* poses /* nframes x 16 or NULL */, const float* planes, int K, double* out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t fr = 0; fr < nframes; ++fr) {
    const uint16_t* dep = frames + (size_t)fr * w * h;
    const float* pose = poses ? poses + 16 * fr : NULL;
    double* o = out + 29 * fr;
    for (int i = 0; i < 29; ++i) o[i] = 0;
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        uint16_t d = dep[(size_t)y * w + x];
        if (d == 0) continue;
        float p[3], r;
        pixel_to_point(x, y, d, intr, pose, p);
        int k = nearest_plane(planes, K, p[0], p[1], p[2], &r);
        const float* n = planes + 4 * k;
        float J[6] = {p[1] * n[2] - p[2] * n[1], p[2] * n[0] - p[0] * n[2], p[0] * n[1] - p[1] * n[0], n[0], n[1], n[2]};
        int t = 0;
        for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) o[t++] += (double)J[a] * (double)J[b];
        for (int a = 0; a < 6; ++a) o[21 + a] += (double)J[a] * (double)r;
        o[27] += (double)r * (double)r;
        o[28] += 1.0;
      }
  }
}

/* ------------------------------------------------------------------------------------
 * A9  pointMean (Float, sequential foldl').   Main.hs:1596-1601
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_point_mean_f32seq(const float* xyz, int64_t n, float out[3]) {
  float sx = 0, sy = 0, sz = 0;
  for (int64_t i = 0; i < n; ++i) { sx = sx + xyz[3 * i]; sy = sy + xyz[3 * i + 1]; sz = sz + xyz[3 * i + 2]; }
  float inv = 1.0f / (float)n;
  out[0] = sx * inv; out[1] = sy * inv; out[2] = sz * inv;
}
/* parity target for the GPU: same mean with Double accumulation (SURVEY §7 hard part 2). */
ORC_API void orc_point_mean_f64(const float* xyz, int64_t n, double out[3]) {
  double sx = 0, sy = 0, sz = 0;
#pragma omp parallel for reduction(+ : sx, sy, sz) schedule(static)
  for (int64_t i = 0; i < n; ++i) { sx += xyz[3 * i]; sy += xyz[3 * i + 1]; sz += xyz[3 * i + 2]; }
  out[0] = sx / (double)n; out[1] = sy / (double)n; out[2] = sz / (double)n;
}
/* V.maximum . V.map (distance m)   Main.hs:1527  (Float) */
ORC_API float orc_max_distance(const float* xyz, int64_t n, const float m[3]) {
  float best = 0.0f;
  v3f mm = v3f_mk(m[0], m[1], m[2]);
  for (int64_t i = 0; i < n; ++i) {
    float d = v3f_distance(mm, v3f_mk(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    if (i == 0 || d > best) best = d;
  }
  return best;
}

/* ------------------------------------------------------------------------------------
 * A7  fitPlane scatter.   Main.hs:1436-1450
 *   m = pointMean (Float) [mode 0: sequential Float as the reference; mode 1: Double-accumulated,
 *   rounded to Float - the GPU parity target];  q_i = toDouble (p_i &- m)  (Float subtract);
 *   scatter = sum q_i q_i^T (Double), returned as xx,xy,xz,yy,yz,zz.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_scatter3x3(const float* xyz, int64_t n, int mean_mode, float mean_out[3], double sc[6]) {
  float m[3];
  if (mean_mode == 0) orc_point_mean_f32seq(xyz, n, m);
  else { double md[3]; orc_point_mean_f64(xyz, n, md); m[0] = (float)md[0]; m[1] = (float)md[1]; m[2] = (float)md[2]; }
  double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
#pragma omp parallel for reduction(+ : xx, xy, xz, yy, yz, zz) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double x = (double)(xyz[3 * i] - m[0]), y = (double)(xyz[3 * i + 1] - m[1]), z = (double)(xyz[3 * i + 2] - m[2]);
    xx += x * x; xy += x * y; xz += x * z; yy += y * y; yz += y * z; zz += z * z;
  }
  mean_out[0] = m[0]; mean_out[1] = m[1]; mean_out[2] = m[2];
  sc[0] = xx; sc[1] = xy; sc[2] = xz; sc[3] = yy; sc[4] = yz; sc[5] = zz;
}

/* ------------------------------------------------------------------------------------
 * A8  rigid transforms on clouds (Float).
 *   rotateCloudAround c R  = V.map (rotateAround c R)         Main.hs:1657-1659, 1582-1583
 *   translateCloud off     = V.map (off &+)                   Main.hs:1697-1699
 *   projectRoom cloud part = translateCloud off . rotateCloudAround zero rotMat, R/off read from the
 *                            rows of the 4x4 (last column must be 0,0,0,1)      Main.hs:1716,1725-1730
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_rotate_cloud_around(const float* xyz, int64_t n, const float c[3], const float R[9], float* out) {
  m3f m = {{R[0], R[1], R[2]}, {R[3], R[4], R[5]}, {R[6], R[7], R[8]}};
  v3f cc = v3f_mk(c[0], c[1], c[2]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    v3f r = v3f_rotate_around(cc, m, v3f_mk(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
  }
}
ORC_API void orc_translate_cloud(const float* xyz, int64_t n, const float off[3], float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    out[3 * i] = off[0] + xyz[3 * i]; out[3 * i + 1] = off[1] + xyz[3 * i + 1]; out[3 * i + 2] = off[2] + xyz[3 * i + 2];
  }
}
/* returns 0 on success, 1 if the last column is not exactly (0,0,0,1) (the reference pattern-fails). */
ORC_API int orc_project_cloud(const float* xyz, int64_t n, const float M[16], float* out) {
  if (M[3] != 0.0f || M[7] != 0.0f || M[11] != 0.0f || M[15] != 1.0f) return 1;
  m3f R = {{M[0], M[1], M[2]}, {M[4], M[5], M[6]}, {M[8], M[9], M[10]}};
  v3f off = v3f_mk(M[12], M[13], M[14]), zero = v3f_mk(0, 0, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    v3f r = v3f_rotate_around(zero, R, v3f_mk(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    r = v3f_add(off, r);
    out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * A12  kthLargestBy / removeCeiling.   VectorUtil.hs:11-19, Main.hs:2643-2664
 * keys are read at xyz[stride_floats*i + comp].  Returns 0 ok, 1 for k<1, 2 for k>n (the reference `error`s).
 * ---------------------------------------------------------------------------------- */
static int cmp_float_desc(const void* a, const void* b) { float x = *(const float*)a, y = *(const float*)b; return (x < y) - (x > y); }
ORC_API int orc_kth_largest_f32(const float* base, int64_t n, int64_t stride_floats, int64_t k, float* out) {
  if (k < 1) return 1;
  if (k > n) return 2;
  float* tmp = (float*)malloc((size_t)n * sizeof(float));
  for (int64_t i = 0; i < n; ++i) tmp[i] = base[i * stride_floats];
  qsort(tmp, (size_t)n, sizeof(float), cmp_float_desc);
  *out = tmp[k - 1];
  free(tmp);
  return 0;
}
/* V.filter ((<= limit) . comp)  — order preserving.  returns number kept. */
ORC_API int64_t orc_filter_le(const float* xyz, int64_t n, int axis, float limit, float* out, const uint8_t* rgb, uint8_t* rgb_out) {
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i)
    if (xyz[3 * i + axis] <= limit) {
      if (out) { out[3 * m] = xyz[3 * i]; out[3 * m + 1] = xyz[3 * i + 1]; out[3 * m + 2] = xyz[3 * i + 2]; }
      if (rgb && rgb_out) { rgb_out[3 * m] = rgb[3 * i]; rgb_out[3 * m + 1] = rgb[3 * i + 1]; rgb_out[3 * m + 2] = rgb[3 * i + 2]; }
      ++m;
    }
  return m;
}

/* ------------------------------------------------------------------------------------
 * A10/A11  connected components on dense (bijected) vertex ids: label = minimum vertex index of the
 * component.  Data.Graph.components visits roots in ascending vertex order (GroupConnectedComponents.hs:46-47),
 * so component order == ascending minimum index == ascending label.
 * ---------------------------------------------------------------------------------- */
static uint32_t uf_find(uint32_t* p, uint32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
ORC_API void orc_cc_label(const uint32_t* src, const uint32_t* dst, int64_t E, uint32_t N, uint32_t* label) {
  for (uint32_t i = 0; i < N; ++i) label[i] = i;
  for (int64_t e = 0; e < E; ++e) {
    uint32_t a = uf_find(label, src[e]), b = uf_find(label, dst[e]);
    if (a < b) label[b] = a; else if (b < a) label[a] = b;
  }
  for (uint32_t i = 0; i < N; ++i) label[i] = uf_find(label, i);
}
