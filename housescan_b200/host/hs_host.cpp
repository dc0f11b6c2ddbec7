// Host-side mirror of the reference's Haskell helper modules (C++ because GHC is not available in the
// build image; the reference is compiled code).  Same names, argument meaning and error behaviour as
//   FitCuboidBFGS.hs, TranslationOptimizer.hs, GroupConnectedComponents.hs (regrouping), Main.hs export helpers.
// Nothing here touches points: per-point work is on the GPU (csrc/*.cu).  These are the O(10)-parameter
// optimisers, <=25-node solves and formatters that the reference also runs on the host.
#include "hs_host.hpp"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace hs {

// ------------------------------------------------------------------------------------------------
// params -> 6 PlaneEq in Float.  Main.hs:1831-1836 (toFloat params, mkU) + makePlanesFromCuboid Main.hs:1852-1874
// ------------------------------------------------------------------------------------------------
void planes_from_cuboid(const double params[10], float out[24]) {
  float p[10];
  for (int i = 0; i < 10; ++i) p[i] = static_cast<float>(params[i]);
  const V3<float> center{p[0], p[1], p[2]};
  const M3<float> R = rot_from_quat(V4<float>{p[6], p[7], p[8], p[9]});
  const V3<float> zero{0.f, 0.f, 0.f};
  for (int axis = 0; axis < 3; ++axis)
    for (int side = 0; side < 2; ++side) {
      const float sgn = side == 0 ? 1.f : -1.f;
      V3<float> e{axis == 0 ? sgn : 0.f, axis == 1 ? sgn : 0.f, axis == 2 ? sgn : 0.f};
      PlaneEq eq = mk_plane_eq(e, p[3 + axis] / 2);
      eq = translate_plane_eq(center, rotate_plane_eq_around(zero, R, eq));
      float* o = out + 4 * (2 * axis + side);
      o[0] = eq.n.x; o[1] = eq.n.y; o[2] = eq.n.z; o[3] = eq.d;
    }
}

// ------------------------------------------------------------------------------------------------
// Chain rule from the GPU's per-room sums to d f / d params (Double).
//   plane (j, +-):  N = +-R_j(q),  D = dim_j/2 +- c.R_j,   r = N.p - D
//   rec: [0] f, [1..6] Sr[k] = sum_{wall k} r, [7..15] B[j] = sum_{axis j} s p (s = +-r), [16..21] counts
// ------------------------------------------------------------------------------------------------
static void quat_rows_and_derivs(const double q[4], double R[3][3], double dR[4][3][3]) {
  const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double u[4] = {q[0] / nq, q[1] / nq, q[2] / nq, q[3] / nq};
  // R entries are quadratic forms u^T Q_ij u; write R_ij = sum_{mn} C[i][j][m][n] u_m u_n via explicit table of L = R^T
  auto L = [](const double* w, double out[3][3]) {
    const double a = w[0], b = w[1], c = w[2], d = w[3];
    out[0][0] = a * a + b * b - c * c - d * d; out[0][1] = 2 * (b * c - a * d);         out[0][2] = 2 * (b * d + a * c);
    out[1][0] = 2 * (b * c + a * d);         out[1][1] = a * a - b * b + c * c - d * d; out[1][2] = 2 * (c * d - a * b);
    out[2][0] = 2 * (b * d - a * c);         out[2][1] = 2 * (c * d + a * b);         out[2][2] = a * a - b * b - c * c + d * d;
  };
  double Lm[3][3];
  L(u, Lm);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = Lm[j][i];
  // L is a homogeneous quadratic: dL/du_m = L(u + e_m) - L(u - e_m) over 2 exactly (polarisation), no truncation error.
  double dLu[4][3][3];
  for (int m = 0; m < 4; ++m) {
    double up[4] = {u[0], u[1], u[2], u[3]}, um[4] = {u[0], u[1], u[2], u[3]}, Lp[3][3], Lq[3][3];
    up[m] += 1.0; um[m] -= 1.0;
    L(up, Lp); L(um, Lq);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dLu[m][i][j] = 0.5 * (Lp[i][j] - Lq[i][j]);
  }
  for (int m = 0; m < 4; ++m)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double acc = 0;
        for (int n = 0; n < 4; ++n) acc += dLu[n][j][i] * (((n == m) ? 1.0 : 0.0) - u[n] * u[m]) / nq;
        dR[m][i][j] = acc;
      }
}

void cuboid_grad_from_sums(const double params[10], const double* rec, double* f, double grad[10], int64_t counts[6]) {
  double R[3][3], dR[4][3][3];
  quat_rows_and_derivs(params + 6, R, dR);
  const double* Sr = rec + 1;
  const double* B = rec + 7;
  double g[10] = {0};
  for (int j = 0; j < 3; ++j) {
    const double Splus = Sr[2 * j], Sminus = Sr[2 * j + 1];
    const double T = Splus - Sminus;  // sum of s = sigma r over the axis
    for (int c = 0; c < 3; ++c) g[c] += -2.0 * T * R[j][c];
    g[3 + j] = -(Splus + Sminus);  // 2 * r * (-1/2)
    for (int m = 0; m < 4; ++m) {
      double acc = 0;
      for (int c = 0; c < 3; ++c) acc += dR[m][j][c] * (B[3 * j + c] - params[c] * T);
      g[6 + m] += 2.0 * acc;
    }
  }
  if (f) *f = rec[0];
  if (grad) for (int i = 0; i < 10; ++i) grad[i] = g[i];
  if (counts) for (int k = 0; k < 6; ++k) counts[k] = static_cast<int64_t>(std::llround(rec[16 + k]));
}

// ------------------------------------------------------------------------------------------------
// FitCuboidBFGS.hs on 8 corners (Double)
// ------------------------------------------------------------------------------------------------
void cuboid_from_params(const double p[10], double out[24]) {  // FitCuboidBFGS.hs:98-112
  const M3<double> R = rot_from_quat(V4<double>{p[6], p[7], p[8], p[9]});
  const V3<double> c{p[0], p[1], p[2]};
  const double ha = p[3] / 2, hb = p[4] / 2, hc = p[5] / 2;
  int o = 0;
  for (int i = 0; i < 8; ++i) {
    V3<double> v{(i & 4) ? ha : -ha, (i & 2) ? hb : -hb, (i & 1) ? hc : -hc};
    V3<double> r = rowmul(v, R) + c;
    out[o++] = r.x; out[o++] = r.y; out[o++] = r.z;
  }
}
double errfun(const double pts[24], const double params[10]) {  // FitCuboidBFGS.hs:51-65
  double est[24], s = 0;
  cuboid_from_params(params, est);
  for (int i = 0; i < 8; ++i) {
    V3<double> d = V3<double>{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]} - V3<double>{est[3 * i], est[3 * i + 1], est[3 * i + 2]};
    s += dot(d, d);
  }
  return s;
}
double errfun_closest(const double* pts, int npts, const double params[10]) {  // FitCuboidBFGS.hs:68-76
  double est[24], s = 0;
  cuboid_from_params(params, est);
  for (int i = 0; i < npts; ++i) {
    const V3<double> p{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    int best = 0;
    double db = std::numeric_limits<double>::infinity();
    for (int e = 0; e < 8; ++e) {  // minimumBy keeps the first minimum
      double d = distance(p, V3<double>{est[3 * e], est[3 * e + 1], est[3 * e + 2]});
      if (d < db) { db = d; best = e; }
    }
    V3<double> d = p - V3<double>{est[3 * best], est[3 * best + 1], est[3 * best + 2]};
    s += dot(d, d);
  }
  return s;
}
void guess_dims(const double pts[24], double out[3]) {  // FitCuboidBFGS.hs:247-252
  double d[7];
  const V3<double> f{pts[0], pts[1], pts[2]};
  for (int i = 1; i < 8; ++i) d[i - 1] = distance(f, V3<double>{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
  std::sort(d, d + 7);
  out[0] = d[0]; out[1] = d[1];
  out[2] = std::sqrt(d[6] * d[6] - d[0] * d[0] - d[1] * d[1]);
}
static void point_mean8(const double pts[24], double c[3]) {  // FitCuboidBFGS.hs:80-84
  double s[3] = {0, 0, 0};
  for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) s[k] = s[k] + pts[3 * i + k];
  const double inv = 1.0 / 8.0;
  for (int k = 0; k < 3; ++k) c[k] = s[k] * inv;
}

// ------------------------------------------------------------------------------------------------
// Nelder-Mead with the update rules of GSL's nmsimplex2 as hmatrix's `minimize NMSimplex2 eps maxit` drives it
// (FitCuboidBFGS.hs:184,201,233): reflect (-1), expand (-2), contract (0.5), else shrink about the best corner;
// size = rms distance of the corners to their centre, maintained incrementally; stop when size < eps or maxit.
// ------------------------------------------------------------------------------------------------
NMResult nm_simplex2(const std::function<double(const std::vector<double>&)>& f_in, const std::vector<double>& x0,
                     const std::vector<double>& step, double eps, int maxit, bool keep_path, const NMBatch* batch, int* evals_out) {
  const int n = static_cast<int>(x0.size()), P = n + 1;
  int evals = 0;
  auto f = [&](const std::vector<double>& x) { ++evals; return f_in(x); };
  std::vector<std::vector<double>> X(P, x0);
  std::vector<double> y(P), center(n, 0.0);
  for (int i = 0; i < n; ++i) X[i + 1][i] += step[i];
  if (batch) { (*batch)(X, y); evals += P; }  // the n + 1 corners are independent
  else for (int i = 0; i < P; ++i) y[i] = f(X[i]);
  auto recompute = [&]() {
    std::fill(center.begin(), center.end(), 0.0);
    for (auto& x : X) for (int k = 0; k < n; ++k) center[k] += x[k];
    for (int k = 0; k < n; ++k) center[k] /= P;
    double s2 = 0;
    for (auto& x : X) for (int k = 0; k < n; ++k) s2 += (x[k] - center[k]) * (x[k] - center[k]);
    return s2 / P;
  };
  double S2 = recompute();
  auto corner_move = [&](double coeff, int corner, std::vector<double>& xc) {
    const double alpha = (1 - coeff) * P / (P - 1.0), beta = (P * coeff - 1.0) / (P - 1.0);
    for (int k = 0; k < n; ++k) xc[k] = alpha * center[k] + beta * X[corner][k];
    return f(xc);
  };
  auto update_point = [&](int i, const std::vector<double>& x, double val) {
    double d2 = 0, xmcd = 0;
    for (int k = 0; k < n; ++k) {
      const double delta = x[k] - X[i][k];
      d2 += delta * delta;
      xmcd += (X[i][k] - center[k]) * delta;
    }
    S2 += (2.0 / P) * xmcd + ((P - 1.0) / P) * (d2 / P);
    for (int k = 0; k < n; ++k) center[k] += (x[k] - X[i][k]) / P;
    X[i] = x;
    y[i] = val;
  };
  NMResult res;
  std::vector<double> xc(n), xc2(n);
  int it = 0, lo = 0;
  for (;;) {
    ++it;
    int hi = 0, s_hi = 1;
    lo = 0;
    double dhi = y[0], dlo = y[0], ds_hi = y[1];
    for (int i = 1; i < P; ++i) {
      const double v = y[i];
      if (v < dlo) { dlo = v; lo = i; }
      else if (v > dhi) { ds_hi = dhi; s_hi = hi; dhi = v; hi = i; }
      else if (v > ds_hi) { ds_hi = v; s_hi = i; }
    }
    double val = corner_move(-1.0, hi, xc);
    if (std::isfinite(val) && val < y[lo]) {
      double val2 = corner_move(-2.0, hi, xc2);
      if (std::isfinite(val2) && val2 < y[lo]) update_point(hi, xc2, val2);
      else update_point(hi, xc, val);
    } else if (!std::isfinite(val) || val > y[s_hi]) {
      if (std::isfinite(val) && val <= y[hi]) update_point(hi, xc, val);
      double val2 = corner_move(0.5, hi, xc2);
      if (std::isfinite(val2) && val2 <= y[hi]) update_point(hi, xc2, val2);
      else {
        std::vector<std::vector<double>> moved;
        std::vector<int> which;
        for (int i = 0; i < P; ++i)
          if (i != lo) {
            for (int k = 0; k < n; ++k) X[i][k] = 0.5 * (X[i][k] + X[lo][k]);
            if (batch) { moved.push_back(X[i]); which.push_back(i); }
            else y[i] = f(X[i]);
          }
        if (batch) {  // the n shrunken corners are independent
          std::vector<double> vals(moved.size());
          (*batch)(moved, vals);
          evals += static_cast<int>(moved.size());
          for (size_t q = 0; q < which.size(); ++q) y[which[q]] = vals[q];
        }
        S2 = recompute();
      }
    } else {
      update_point(hi, xc, val);
    }
    lo = static_cast<int>(std::min_element(y.begin(), y.end()) - y.begin());
    const double size = std::sqrt(S2 > 0 ? S2 : recompute());
    if (keep_path) {
      res.path.push_back(static_cast<double>(it));
      res.path.push_back(y[lo]);
      res.path.push_back(size);
      res.path.insert(res.path.end(), X[lo].begin(), X[lo].end());
    }
    if (size < eps || it >= maxit) break;
  }
  res.x = X[lo];
  res.fval = y[lo];
  res.iters = it;
  if (evals_out) *evals_out = evals;
  return res;
}

// fitCuboid (0), fitCuboidFromCenter (1), fitCuboidFromCenterFirst (2).  FitCuboidBFGS.hs:172-233
FitResult fit_cuboid(const double pts[24], int variant, bool keep_path) {
  const int maxIt = 2000;
  double dims[3], c[3];
  guess_dims(pts, dims);
  point_mean8(pts, c);
  const double a = dims[0];
  FitResult out;
  auto from_center = [&](FitResult& r) {
    auto errf = [&](const std::vector<double>& s) {
      double p[10] = {c[0], c[1], c[2], s[0], s[1], s[2], s[3], s[4], s[5], s[6]};
      return errfun_closest(pts, 8, p);
    };
    NMResult nm = nm_simplex2(errf, {a, a, a, 0.1, 0.1, 0.1, 0.1}, {a / 10, a / 10, a / 10, 0.1, 0.1, 0.1, 0.1}, 1e-8, maxIt, keep_path);
    r.params = {c[0], c[1], c[2]};
    r.params.insert(r.params.end(), nm.x.begin(), nm.x.end());
    r.steps = nm.iters;
    r.err = errf(nm.x);
    r.path = nm.path;
    r.path_cols = 3 + 7;
  };
  if (variant == 1) { from_center(out); return out; }
  if (variant == 2) {
    FitResult first;
    from_center(first);
    auto errf = [&](const std::vector<double>& s) { return errfun_closest(pts, 8, s.data()); };
    NMResult nm = nm_simplex2(errf, first.params, {0.01, 0.01, 0.01, a / 10, a / 10, a / 10, 0.1, 0.1, 0.1, 0.1}, 1e-8, maxIt, keep_path);
    out.params = nm.x; out.steps = first.steps + nm.iters; out.err = errf(nm.x); out.path = nm.path; out.path_cols = 13;
    return out;
  }
  auto errf = [&](const std::vector<double>& s) { return errfun(pts, s.data()); };
  NMResult nm = nm_simplex2(errf, {c[0], c[1], c[2], dims[0], dims[1], dims[2], 0.1, 0.1, 0.1, 0.1},
                            {0.01, 0.01, 0.01, a / 10, a / 10, a / 10, 0.1, 0.1, 0.1, 0.1}, 1e-8, maxIt, keep_path);
  out.params = nm.x; out.steps = nm.iters; out.err = errf(nm.x); out.path = nm.path; out.path_cols = 13;
  return out;
}

// ------------------------------------------------------------------------------------------------
// BFGS (north-star addition) on an objective with analytic gradient; backtracking Armijo line search.
// ------------------------------------------------------------------------------------------------
BFGSResult bfgs(const std::function<bool(const double*, double*, double*)>& eval, const double* x0, int n, int max_iter, double gtol) {
  BFGSResult r;
  r.x.assign(x0, x0 + n);
  std::vector<double> g(n), gn(n), xn(n), d(n), H(n * n, 0.0), s(n), yv(n), Hy(n);
  for (int i = 0; i < n; ++i) H[i * n + i] = 1.0;
  double f = 0;
  r.ok = eval(r.x.data(), &f, g.data());
  r.evals = 1;
  if (!r.ok) return r;
  bool fresh = true;
  for (r.iters = 0; r.iters < max_iter; ++r.iters) {
    double gmax = 0;
    for (double v : g) gmax = std::max(gmax, std::fabs(v));
    if (gmax <= gtol) break;
    double slope = 0;
    for (int i = 0; i < n; ++i) { d[i] = 0; for (int j = 0; j < n; ++j) d[i] -= H[i * n + j] * g[j]; }
    for (int i = 0; i < n; ++i) slope += d[i] * g[i];
    if (slope >= 0) {  // not a descent direction: reset
      std::fill(H.begin(), H.end(), 0.0);
      for (int i = 0; i < n; ++i) { H[i * n + i] = 1.0; d[i] = -g[i]; }
      slope = 0;
      for (int i = 0; i < n; ++i) slope += d[i] * g[i];
      fresh = true;
    }
    double dn = 0;
    for (double v : d) dn += v * v;
    dn = std::sqrt(dn);
    double t = fresh ? std::min(1.0, 0.1 / std::max(dn, 1e-300)) : 1.0;  // first step: at most 10 cm / 0.1 in q
    double fn = f;
    bool accepted = false;
    for (int ls = 0; ls < 40; ++ls) {
      for (int i = 0; i < n; ++i) xn[i] = r.x[i] + t * d[i];
      if (!eval(xn.data(), &fn, gn.data())) { r.ok = false; return r; }
      ++r.evals;
      if (std::isfinite(fn) && fn <= f + 1e-4 * t * slope) { accepted = true; break; }
      t *= 0.5;
    }
    if (!accepted) break;
    double sy = 0;
    for (int i = 0; i < n; ++i) { s[i] = xn[i] - r.x[i]; yv[i] = gn[i] - g[i]; sy += s[i] * yv[i]; }
    if (sy > 1e-300) {
      if (fresh) {  // scale the initial inverse Hessian
        double yy = 0;
        for (double v : yv) yy += v * v;
        std::fill(H.begin(), H.end(), 0.0);
        for (int i = 0; i < n; ++i) H[i * n + i] = sy / yy;
        fresh = false;
      }
      double yHy = 0;
      for (int i = 0; i < n; ++i) { Hy[i] = 0; for (int j = 0; j < n; ++j) Hy[i] += H[i * n + j] * yv[j]; }
      for (int i = 0; i < n; ++i) yHy += yv[i] * Hy[i];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          H[i * n + j] += (1.0 + yHy / sy) * s[i] * s[j] / sy - (Hy[i] * s[j] + s[i] * Hy[j]) / sy;
    }
    const double fprev = f;
    r.x = xn; g = gn; f = fn;
    if (std::fabs(fprev - f) <= 1e-14 * std::max(1.0, std::fabs(f))) { ++r.iters; break; }
  }
  r.f = f;
  return r;
}

// ------------------------------------------------------------------------------------------------
// TranslationOptimizer.lstSqDistancesI (TranslationOptimizer.hs:48-72) on bijected indices.
// Rows are given in Map.toList order by the caller.  A has a -1 at column i and +1 at column j (i tested first),
// column 0 dropped (x_0 = 0).  Least squares by Householder QR; rank deficiency => singular (Nothing).
// rmse = sqrt( ||A x - b||_2 / m )  -- 2-norm NOT squared, as the reference (TranslationOptimizer.hs:70).
// ------------------------------------------------------------------------------------------------
bool lstsq_distances(const int32_t* ii, const int32_t* jj, const double* d, int m, int n_nodes, double* pos, double* rmse) {
  const int nc = n_nodes - 1;
  std::vector<double> A(static_cast<size_t>(m) * std::max(nc, 1), 0.0), A0, b(d, d + m), b0(d, d + m);
  auto at = [&](std::vector<double>& M, int r, int c) -> double& { return M[static_cast<size_t>(r) * nc + c]; };
  for (int r = 0; r < m; ++r)
    for (int p = 1; p < n_nodes; ++p) at(A, r, p - 1) = (p == ii[r]) ? -1.0 : ((p == jj[r]) ? 1.0 : 0.0);
  A0 = A;
  std::vector<double> x(std::max(nc, 0), 0.0);
  if (nc > 0) {
    if (m < nc) return false;
    std::vector<double> diag(nc);
    for (int k = 0; k < nc; ++k) {
      double nrm = 0;
      for (int r = k; r < m; ++r) nrm += at(A, r, k) * at(A, r, k);
      nrm = std::sqrt(nrm);
      if (nrm <= 1e-12) return false;
      const double alpha = at(A, k, k) > 0 ? -nrm : nrm;
      std::vector<double> v(m - k);
      for (int r = k; r < m; ++r) v[r - k] = at(A, r, k);
      v[0] -= alpha;
      double vv = 0;
      for (double t : v) vv += t * t;
      if (vv > 0) {
        for (int c = k; c < nc; ++c) {
          double s = 0;
          for (int r = k; r < m; ++r) s += v[r - k] * at(A, r, c);
          s = 2 * s / vv;
          for (int r = k; r < m; ++r) at(A, r, c) -= s * v[r - k];
        }
        double s = 0;
        for (int r = k; r < m; ++r) s += v[r - k] * b[r];
        s = 2 * s / vv;
        for (int r = k; r < m; ++r) b[r] -= s * v[r - k];
      }
      diag[k] = at(A, k, k);
    }
    for (int k = nc - 1; k >= 0; --k) {
      double s = b[k];
      for (int c = k + 1; c < nc; ++c) s -= at(A, k, c) * x[c];
      x[k] = s / diag[k];
    }
  }
  pos[0] = 0.0;
  for (int k = 0; k < nc; ++k) pos[k + 1] = x[k];
  double rr = 0;
  for (int r = 0; r < m; ++r) {
    double s = -b0[r];
    for (int c = 0; c < nc; ++c) s += at(A0, r, c) * x[c];
    rr += s * s;
  }
  *rmse = std::sqrt(std::sqrt(rr) / m);
  return true;
}

// ------------------------------------------------------------------------------------------------
// symmetric 3x3 eigen-decomposition (cyclic Jacobi); eigenvalues ascending, vectors as columns
// ------------------------------------------------------------------------------------------------
void eig_sym3(const double sc[6], double evals[3], double evecs[3][3]) {
  double a[3][3] = {{sc[0], sc[1], sc[2]}, {sc[1], sc[3], sc[4]}, {sc[2], sc[4], sc[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double dia = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-32 * dia || off == 0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
      }
  }
  int idx[3] = {0, 1, 2};
  std::sort(idx, idx + 3, [&](int x, int y) { return a[x][x] < a[y][y]; });
  for (int k = 0; k < 3; ++k) { evals[k] = a[idx[k]][idx[k]]; for (int r = 0; r < 3; ++r) evecs[r][k] = v[r][idx[k]]; }
}

// ------------------------------------------------------------------------------------------------
// `show :: Float -> String` (shortest round-trip digits; fixed for 0.1 <= |x| < 10^7 else d.ddde<n>) and the
// roomProj exporters of Main.hs:2271-2302 (transpose to the left-multiplicative form).
// ------------------------------------------------------------------------------------------------
std::string show_float(float x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
  if (x == 0.0f) return std::signbit(x) ? "-0.0" : "0.0";
  char buf[64];
  int prec = 1;
  for (; prec <= 9; ++prec) {  // shortest digit string that reads back to the same Float
    std::snprintf(buf, sizeof buf, "%.*e", prec - 1, static_cast<double>(std::fabs(x)));
    if (std::strtof(buf, nullptr) == std::fabs(x)) break;
  }
  std::string digits;
  const char* e = std::strchr(buf, 'e');
  for (const char* p = buf; p < e; ++p) if (*p != '.') digits.push_back(*p);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int ex = std::atoi(e + 1) + 1;  // value = 0.d1d2.. * 10^ex
  std::string s = x < 0 ? "-" : "";
  if (ex >= 0 && ex <= 7) {
    if (ex == 0) return s + "0." + digits;
    std::string ip = digits.substr(0, std::min<size_t>(ex, digits.size()));
    while (static_cast<int>(ip.size()) < ex) ip.push_back('0');
    std::string fp = digits.size() > static_cast<size_t>(ex) ? digits.substr(ex) : "0";
    return s + ip + "." + fp;
  }
  std::string rest = digits.size() > 1 ? digits.substr(1) : "0";
  return s + digits.substr(0, 1) + "." + rest + "e" + std::to_string(ex - 1);
}
std::string proj_to_string(const float m[16]) {
  std::string s;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      if (r || c) s += ",";
      s += show_float(m[4 * c + r]);
    }
  return s;
}
std::string proj_to_xf(const float m[16]) {
  std::string s;
  for (int r = 0; r < 4; ++r) {
    for (int c = 0; c < 4; ++c) { if (c) s += " "; s += show_float(m[4 * c + r]); }
    s += "\n";
  }
  return s;
}

// binary little-endian PLY
std::string ply_header(int64_t n, bool rgb) {
  std::string h = "ply\nformat binary_little_endian 1.0\ncomment housescan_b200 full-resolution export\nelement vertex " + std::to_string(static_cast<long long>(n)) +
                  "\nproperty float x\nproperty float y\nproperty float z\n";
  if (rgb) h += "property uchar red\nproperty uchar green\nproperty uchar blue\n";
  h += "end_header\n";
  return h;
}

// header + a file of the final size: the body is filled in by write_ply_part, in any order, by any number of writers
bool write_ply_begin(const char* path, int64_t n, bool rgb, std::string* err) {
  FILE* fp = std::fopen(path, "wb");
  if (!fp) { *err = std::string("cannot open ") + path; return false; }
  const std::string h = ply_header(n, rgb);
  bool ok = std::fwrite(h.data(), 1, h.size(), fp) == h.size();
  if (std::fflush(fp) != 0) ok = false;
  if (ok && ftruncate(fileno(fp), static_cast<off_t>(h.size() + static_cast<size_t>(n) * (rgb ? 15 : 12))) != 0) ok = false;
  if (std::fclose(fp) != 0) ok = false;
  if (!ok) *err = std::string("short write to ") + path;
  return ok;
}

// points [first, first + m) of an n-point file (xyz / rgb point to the FIRST of the m points)
bool write_ply_part(int fd, const float* xyz, const uint8_t* rgb, int64_t first, int64_t m, int64_t n, std::string* err) {
  const size_t hdr = ply_header(n, rgb != nullptr).size();
  auto put = [&](const void* p, size_t bytes, size_t at) {
    const char* c = static_cast<const char*>(p);
    while (bytes) {
      const ssize_t w = pwrite(fd, c, bytes, static_cast<off_t>(at));
      if (w <= 0) return false;
      c += w; at += static_cast<size_t>(w); bytes -= static_cast<size_t>(w);
    }
    return true;
  };
  bool ok = true;
  if (!rgb) ok = put(xyz, static_cast<size_t>(m) * 12, hdr + static_cast<size_t>(first) * 12);
  else {
    const size_t chunk = 1 << 16;
    std::vector<uint8_t> buf(chunk * 15);
    for (int64_t i0 = 0; i0 < m && ok; i0 += chunk) {
      const size_t k = static_cast<size_t>(std::min<int64_t>(chunk, m - i0));
      for (size_t i = 0; i < k; ++i) {
        std::memcpy(&buf[i * 15], xyz + 3 * (i0 + i), 12);
        std::memcpy(&buf[i * 15 + 12], rgb + 3 * (i0 + i), 3);
      }
      ok = put(buf.data(), k * 15, hdr + static_cast<size_t>(first + i0) * 15);
    }
  }
  if (!ok) *err = "short write to the .ply body";
  return ok;
}

bool write_ply(const char* path, const float* xyz, const uint8_t* rgb, int64_t n, std::string* err) {
  if (!write_ply_begin(path, n, rgb != nullptr, err)) return false;
  const int fd = open(path, O_WRONLY);
  if (fd < 0) { *err = std::string("cannot open ") + path; return false; }
  const bool ok = write_ply_part(fd, xyz, rgb, 0, n, n, err);
  return (close(fd) == 0) && ok;
}

// GroupConnectedComponents regrouping from min-index vertex labels (GroupConnectedComponents.hs:46-54):
// components in ascending label order; inside a component edges in reverse input order (fromListWith (++)).
void group_edges_by_label(const uint32_t* src, const uint32_t* label, int64_t E, int32_t* comp_out, int64_t* order_out, int32_t* ncomp) {
  std::vector<uint32_t> labs(E);
  for (int64_t e = 0; e < E; ++e) labs[e] = label[src[e]];
  std::vector<uint32_t> uniq(labs);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  std::vector<int64_t> start(uniq.size() + 1, 0);
  for (int64_t e = 0; e < E; ++e) {
    const int32_t c = static_cast<int32_t>(std::lower_bound(uniq.begin(), uniq.end(), labs[e]) - uniq.begin());
    comp_out[e] = c;
    start[c + 1]++;
  }
  for (size_t c = 0; c < uniq.size(); ++c) start[c + 1] += start[c];
  std::vector<int64_t> fill(start.begin(), start.end() - 1);
  for (int64_t e = E - 1; e >= 0; --e) order_out[fill[comp_out[e]]++] = e;
  *ncomp = static_cast<int32_t>(uniq.size());
}

}  // namespace hs
