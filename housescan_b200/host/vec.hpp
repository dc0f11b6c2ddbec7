// Host-side vector algebra with the semantics of the reference's `vect` dependency
// (Data.Vect.Float in Main.hs:39-40, Data.Vect.Double in FitCuboidBFGS.hs:20-21).
// Row vectors, right multiplication (Main.hs:10).  Every expression keeps Haskell's
// left-to-right association and is compiled with -ffp-contract=off so that the Float
// plane equations fed to the GPU are the ones the reference would compute.
#pragma once
#include <cmath>

namespace hs {

template <class T> struct V3 {
  T x, y, z;
  T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> struct V4 { T x, y, z, w; };
template <class T> struct M3 {  // rows
  V3<T> r[3];
};

template <class T> inline V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline V3<T> operator*(T s, V3<T> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> inline T dot(V3<T> a, V3<T> b) {
  T s = a.x * b.x;
  s = s + a.y * b.y;
  s = s + a.z * b.z;
  return s;
}
template <class T> inline T dot(V4<T> a, V4<T> b) {
  T s = a.x * b.x;
  s = s + a.y * b.y;
  s = s + a.z * b.z;
  s = s + a.w * b.w;
  return s;
}
template <class T> inline T norm(V3<T> a) { return std::sqrt(dot(a, a)); }
template <class T> inline T distance(V3<T> a, V3<T> b) { return norm(a - b); }
template <class T> inline V3<T> unit(V3<T> a) {  // normalize = (&* (1/norm))
  T inv = T(1) / norm(a);
  return {a.x * inv, a.y * inv, a.z * inv};
}
template <class T> inline V4<T> unit(V4<T> a) {
  T inv = T(1) / std::sqrt(dot(a, a));
  return {a.x * inv, a.y * inv, a.z * inv, a.w * inv};
}
template <class T> inline V3<T> cross(V3<T> a, V3<T> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// v .* M
template <class T> inline V3<T> rowmul(V3<T> v, const M3<T>& m) {
  V3<T> c0{m.r[0].x, m.r[1].x, m.r[2].x}, c1{m.r[0].y, m.r[1].y, m.r[2].y}, c2{m.r[0].z, m.r[1].z, m.r[2].z};
  return {dot(v, c0), dot(v, c1), dot(v, c2)};
}
// fromOrtho (rightOrthoU (mkU q)): transpose of the standard unit-quaternion matrix
template <class T> inline M3<T> rot_from_quat(V4<T> q_raw) {
  V4<T> q = unit(q_raw);
  T a = q.x, b = q.y, c = q.z, d = q.w;
  T two = T(2);
  // left action matrix L, then R = L^T
  T l00 = a * a + b * b - c * c - d * d, l01 = two * b * c - two * a * d, l02 = two * b * d + two * a * c;
  T l10 = two * b * c + two * a * d, l11 = a * a - b * b + c * c - d * d, l12 = two * c * d - two * a * b;
  T l20 = two * b * d - two * a * c, l21 = two * c * d + two * a * b, l22 = a * a - b * b - c * c + d * d;
  return M3<T>{{{l00, l10, l20}, {l01, l11, l21}, {l02, l12, l22}}};
}
// rotateAround c R p
template <class T> inline V3<T> rotate_around(V3<T> c, const M3<T>& R, V3<T> p) { return rowmul(p - c, R) + c; }

// rotMatrix3' (unit axis v) a = (1 - cos a) v v^T + [[c, s z, -s y], [-s z, c, s x], [s y, -s x, c]]: Rodrigues' rotation by +a about
// v for ROW vectors (p' = p .* R), the convention of Main.hs:10 and of the doc comment at Main.hs:1548-1552
template <class T> inline M3<T> rot_matrix3_unit(V3<T> v, T ang) {
  const T c = std::cos(ang), s = std::sin(ang), k = T(1) - c;
  return M3<T>{{{k * (v.x * v.x) + c, k * (v.x * v.y) + s * v.z, k * (v.x * v.z) + (-(s * v.y))},
                {k * (v.y * v.x) + (-(s * v.z)), k * (v.y * v.y) + c, k * (v.y * v.z) + s * v.x},
                {k * (v.z * v.x) + s * v.y, k * (v.z * v.y) + (-(s * v.x)), k * (v.z * v.z) + c}}};
}
// rotationBetweenPlaneEqs (Main.hs:1553-1560): the rotation that turns normal n1 into the direction of n2; `crossprod` on Normal3
// re-normalises, so the axis is unit(n1 x n2); (anti)parallel normals give NaNs exactly as the reference does (no guard there)
inline M3<float> rotation_between_normals(V3<float> n1, V3<float> n2) {
  const V3<float> axis = unit(cross(n1, n2));
  const float costheta = dot(n1, n2) / (norm(n1) * norm(n2));
  return rot_matrix3_unit(axis, std::acos(costheta));
}

// PlaneEq n d (Main.hs:1357) and its constructors / movers
struct PlaneEq {
  V3<float> n;
  float d;
};
inline PlaneEq mk_plane_eq(V3<float> abc, float d) { return {unit(abc), d / norm(abc)}; }             // Main.hs:1360
inline PlaneEq rotate_plane_eq_around(V3<float> c, const M3<float>& R, PlaneEq e) {                      // Main.hs:1571
  V3<float> n2 = rowmul(e.n, R);
  V3<float> o2 = rotate_around(c, R, e.d * e.n);
  return mk_plane_eq(n2, dot(o2, n2));
}
inline PlaneEq translate_plane_eq(V3<float> off, PlaneEq e) {                                             // Main.hs:1681
  V3<float> o2 = e.d * e.n + off;
  return mk_plane_eq(e.n, dot(o2, e.n));
}

// Proj4 algebra (roomProj, Main.hs:314): 4x4 Float matrices for ROW vectors, translation in row 3; `.*.` is the plain matrix
// product with every entry summed left to right ((a0 b0 + a1 b1) + a2 b2) + a3 b3
struct M4 { float m[16]; };
inline M4 proj_identity() { M4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
inline M4 proj_compose(const M4& A, const M4& B) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float acc = A.m[4 * i] * B.m[j];
      for (int k = 1; k < 4; ++k) acc = acc + A.m[4 * i + k] * B.m[4 * k + j];
      r.m[4 * i + j] = acc;
    }
  return r;
}
inline M4 proj_translate4(V3<float> v, const M4& M) {  // translate4 v: post-translation (Main.hs:1708)
  M4 T = proj_identity();
  T.m[12] = v.x; T.m[13] = v.y; T.m[14] = v.z;
  return proj_compose(M, T);
}
inline M4 proj_linear(const M3<float>& R) {
  M4 L = proj_identity();
  for (int r = 0; r < 3; ++r) { L.m[4 * r] = R.r[r].x; L.m[4 * r + 1] = R.r[r].y; L.m[4 * r + 2] = R.r[r].z; }
  return L;
}
inline M4 proj_rotate_around(V3<float> c, const M3<float>& R, const M4& M) {  // translate4 c . (.*. linear R) . translate4 (neg c), Main.hs:1674
  const V3<float> nc{-c.x, -c.y, -c.z};
  return proj_translate4(c, proj_compose(proj_translate4(nc, M), proj_linear(R)));
}

}  // namespace hs
