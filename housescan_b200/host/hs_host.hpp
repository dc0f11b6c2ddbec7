// Declarations of the host-side mirror of the reference's helper modules (see hs_host.cpp).
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "vec.hpp"

namespace hs {

void planes_from_cuboid(const double params[10], float out[24]);
void cuboid_grad_from_sums(const double params[10], const double* rec, double* f, double grad[10], int64_t counts[6]);

void cuboid_from_params(const double p[10], double out[24]);
double errfun(const double pts[24], const double params[10]);
double errfun_closest(const double* pts, int npts, const double params[10]);
void guess_dims(const double pts[24], double out[3]);

struct NMResult {
  std::vector<double> x;
  double fval = 0;
  int iters = 0;
  std::vector<double> path;  // rows of [iter, f, size, x...]
};
NMResult nm_simplex2(const std::function<double(const std::vector<double>&)>& f, const std::vector<double>& x0,
                     const std::vector<double>& step, double eps, int maxit, bool keep_path);

struct FitResult {
  std::vector<double> params;
  int steps = 0;
  double err = 0;
  std::vector<double> path;
  int path_cols = 0;
};
FitResult fit_cuboid(const double pts[24], int variant, bool keep_path);

struct BFGSResult {
  std::vector<double> x;
  double f = 0;
  int iters = 0, evals = 0;
  bool ok = true;
};
BFGSResult bfgs(const std::function<bool(const double*, double*, double*)>& eval, const double* x0, int n, int max_iter, double gtol);

bool lstsq_distances(const int32_t* ii, const int32_t* jj, const double* d, int m, int n_nodes, double* pos, double* rmse);
void eig_sym3(const double sc[6], double evals[3], double evecs[3][3]);

std::string show_float(float x);
std::string proj_to_string(const float m[16]);
std::string proj_to_xf(const float m[16]);
bool write_ply(const char* path, const float* xyz, const uint8_t* rgb, int64_t n, std::string* err);
void group_edges_by_label(const uint32_t* src, const uint32_t* label, int64_t E, int32_t* comp_out, int64_t* order_out, int32_t* ncomp);

}  // namespace hs
