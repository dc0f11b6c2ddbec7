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
// `batch` (optional): evaluates several independent points at once (the initial simplex, a shrink step) - same values, same order
// as calling f on each; an evaluation session runs such a batch back to back on the device.  evals_out counts objective values.
using NMBatch = std::function<void(const std::vector<std::vector<double>>&, std::vector<double>&)>;
NMResult nm_simplex2(const std::function<double(const std::vector<double>&)>& f, const std::vector<double>& x0,
                     const std::vector<double>& step, double eps, int maxit, bool keep_path, const NMBatch* batch = nullptr,
                     int* evals_out = nullptr);

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

// planeCorner (Main.hs:1413-1430): the point where three planes meet, n_k . x = d_k solved in Double (LU with partial pivoting, the
// algorithm of LAPACK dgesv behind hmatrix's linearSolve) and rounded to Float; false if a pivot is exactly zero (safeLinearSolve -> Nothing)
bool plane_corner(const float p1[4], const float p2[4], const float p3[4], float out[3]);

std::string show_float(float x);
std::string proj_to_string(const float m[16]);
std::string proj_to_xf(const float m[16]);
bool write_ply(const char* path, const float* xyz, const uint8_t* rgb, int64_t n, std::string* err);
std::string ply_header(int64_t n, bool rgb);
bool write_ply_begin(const char* path, int64_t n, bool rgb, std::string* err);
bool write_ply_part(int fd, const float* xyz, const uint8_t* rgb, int64_t first, int64_t m, int64_t n, std::string* err);
void group_edges_by_label(const uint32_t* src, const uint32_t* label, int64_t E, int32_t* comp_out, int64_t* order_out, int32_t* ncomp);


// ---- room input formats (hs_roomio.cpp) ------------------------------------------------------------------------------
int parse_planes_txt(const char* text, size_t len, std::vector<float>* planes);  // planes as nx ny nz d; -1 = no plane parsed
void make_inward_facing(const float center[3], const float* plane_means, float* planes, int K);
bool point_mean_f32seq(const float* xyz, int64_t n, float out[3]);

struct PcdHeader {
  std::vector<std::string> fields;
  std::vector<int> size, count, field_offset;
  std::vector<char> type;  // 'F', 'U', 'I'
  int64_t width = 0, height = 1, points = -1;
  int data_kind = -1;      // 0 ascii, 1 binary, 2 binary_compressed
  size_t data_offset = 0;  // first byte after the DATA line
  int point_step = 0;      // bytes per point in DATA binary
  int find(const std::string& name) const;
};
bool pcd_parse_header(const char* buf, size_t len, PcdHeader* h, std::string* err);
bool pcd_layout(const PcdHeader& h, int* fx, int* fy, int* fz, int* frgb, std::string* err);
bool pcd_ascii_records(const char* buf, size_t len, const PcdHeader& h, std::vector<uint32_t>* rec, int* rec_words, std::string* err);
bool lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
bool read_file(const char* path, std::vector<char>* out, std::string* err);

bool parse_transform_text(const char* text, size_t len, float m_rowmajor[16]);  // .xf or comma-separated, transposed back to roomProj
struct PlyHeader {
  bool ascii = false;
  int64_t n = -1;
  std::vector<std::string> names, types;
  std::vector<int> sizes, offsets;
  int step = 0;            // bytes per vertex record (binary)
  size_t data_offset = 0;  // first byte after end_header
  int find(const std::string& name) const;
};
bool ply_parse_header(const char* buf, size_t len, PlyHeader* h, std::string* err);
bool ply_ascii_records(const char* buf, size_t len, const PlyHeader& h, bool want_rgb, std::vector<uint32_t>* rec, std::string* err);
bool write_pcd(const char* path, const float* xyz, const uint8_t* rgb, int64_t n, std::string* err);

}  // namespace hs
