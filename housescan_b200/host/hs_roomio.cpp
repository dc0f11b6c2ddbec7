// Room input formats: the step before the hot path.
//   planeEqsFromFile  (Main.hs:1379-1389)  planes.txt of PCL's plane detection: `a b c d` per line with ax + by + cz + d = 0
//   loadPCDFileXyzFloat / loadPCDFileXyzRgbNormalFloat / cloudFromFile (Main.hs:1318-1345)  PCD point clouds
//   makeInwardFacing  (Main.hs:1746-1751)
// The reference reads PCD through the `pcd-loader` package and parses numbers with attoparsec; neither is mounted, so this file
// follows the published PCD v0.7 layout and attoparsec's documented `double` grammar (see parse_double below).
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "hs_host.hpp"

namespace hs {

// attoparsec `double` (Data.Attoparsec.ByteString.Char8 `scientifically`): optional sign, at least one decimal digit, then a '.'
// is consumed whenever it is there, followed by zero or more digits (`anyWord8 *> takeWhile isDigit`: "1." parses as 1), then an
// optional `e[sign]digits`; an 'e' that is not followed by a number is not consumed (the alternative backtracks).
// Returns the number of characters consumed (0 = no parse).
static size_t parse_double(const char* s, size_t len, double* out) {
  size_t i = 0;
  if (i < len && (s[i] == '-' || s[i] == '+')) ++i;
  const size_t d0 = i;
  while (i < len && s[i] >= '0' && s[i] <= '9') ++i;
  if (i == d0) return 0;
  size_t dot = std::string::npos;
  if (i < len && s[i] == '.') {
    dot = i++;
    while (i < len && s[i] >= '0' && s[i] <= '9') ++i;
  }
  if (i < len && (s[i] == 'e' || s[i] == 'E')) {
    size_t j = i + 1;
    if (j < len && (s[j] == '-' || s[j] == '+')) ++j;
    const size_t e0 = j;
    while (j < len && s[j] >= '0' && s[j] <= '9') ++j;
    if (j > e0) i = j;
  }
  std::string tok(s, i);
  if (dot != std::string::npos && (dot + 1 == tok.size() || tok[dot + 1] < '0' || tok[dot + 1] > '9')) tok.insert(dot + 1, "0");  // "1." / "1.e5" for strtod
  *out = std::strtod(tok.c_str(), nullptr);  // correctly rounded; attoparsec's own conversion may differ in the last Double
  return i;                                  // digit, which realToFrac :: Double -> Float hides (parity unpinned there)
}
static size_t skip_space(const char* s, size_t len, size_t i) {  // attoparsec skipSpace: ' ' and \t \n \v \f \r
  while (i < len && (s[i] == ' ' || (s[i] >= 9 && s[i] <= 13))) ++i;
  return i;
}

// (mkPlaneEqABCD <$> floatS <*> floatS <*> floatS <*> (negate <$> float)) `sepBy1'` endOfLine under parseOnly: parsing stops
// silently at the first line that does not match; zero planes is the reference's "Could not load planes" error (returns -1).
int parse_planes_txt(const char* text, size_t len, std::vector<float>* planes) {
  planes->clear();
  size_t i = 0;
  int n = 0;
  for (;;) {
    double v[4];
    size_t j = i;
    bool ok = true;
    for (int c = 0; c < 4 && ok; ++c) {
      const size_t used = parse_double(text + j, len - j, &v[c]);
      if (!used) { ok = false; break; }
      j += used;
      if (c < 3) j = skip_space(text, len, j);  // floatS = float <* skipSpace (newlines count as space); none after d
    }
    if (!ok) break;
    const PlaneEq e = mk_plane_eq(V3<float>{static_cast<float>(v[0]), static_cast<float>(v[1]), static_cast<float>(v[2])},
                                  -static_cast<float>(v[3]));  // negate <$> float: PCL's d sits on the left-hand side
    planes->push_back(e.n.x); planes->push_back(e.n.y); planes->push_back(e.n.z); planes->push_back(e.d);
    ++n;
    i = j;
    // endOfLine = "\n" or "\r\n", directly after d
    if (i < len && text[i] == '\n') i += 1;
    else if (i + 1 < len && text[i] == '\r' && text[i + 1] == '\n') i += 2;
    else break;
  }
  return n > 0 ? n : -1;
}

// makeInwardFacing: flip (n, d) unless (roomCenter - planeMean) . n > 0, all in Float
void make_inward_facing(const float center[3], const float* plane_means, float* planes, int K) {
  for (int k = 0; k < K; ++k) {
    const V3<float> inward = V3<float>{center[0], center[1], center[2]} - V3<float>{plane_means[3 * k], plane_means[3 * k + 1], plane_means[3 * k + 2]};
    const V3<float> n{planes[4 * k], planes[4 * k + 1], planes[4 * k + 2]};
    if (!(dot(inward, n) > 0.0f)) {
      planes[4 * k] = -n.x; planes[4 * k + 1] = -n.y; planes[4 * k + 2] = -n.z; planes[4 * k + 3] = -planes[4 * k + 3];
    }
  }
}

// planeCorner (Main.hs:1413-1430).  Column-oriented LU with partial pivoting on the 3x3 system, then the two triangular solves: the
// operation order of dgetf2 / dgetrs, so the Doubles agree with LAPACK's and the Floats they round to are the reference's.
bool plane_corner(const float p1[4], const float p2[4], const float p3[4], float out[3]) {
  double A[3][3] = {{p1[0], p1[1], p1[2]}, {p2[0], p2[1], p2[2]}, {p3[0], p3[1], p3[2]}};
  double b[3] = {p1[3], p2[3], p3[3]};
  int piv[3];
  for (int j = 0; j < 3; ++j) {
    int p = j;
    for (int i = j + 1; i < 3; ++i)
      if (std::fabs(A[i][j]) > std::fabs(A[p][j])) p = i;  // idamax: first maximum
    piv[j] = p;
    if (A[p][j] == 0.0) return false;
    if (p != j)
      for (int c = 0; c < 3; ++c) std::swap(A[j][c], A[p][c]);
    const double r = 1.0 / A[j][j];  // dgetf2 scales the column by the reciprocal of the pivot
    for (int i = j + 1; i < 3; ++i) A[i][j] *= r;
    for (int c = j + 1; c < 3; ++c)
      for (int i = j + 1; i < 3; ++i) A[i][c] -= A[i][j] * A[j][c];
  }
  for (int j = 0; j < 3; ++j)
    if (piv[j] != j) std::swap(b[j], b[piv[j]]);
  for (int j = 0; j < 3; ++j)  // L y = b (unit lower), column oriented
    for (int i = j + 1; i < 3; ++i) b[i] -= A[i][j] * b[j];
  for (int j = 2; j >= 0; --j) {  // U x = y, column oriented
    b[j] /= A[j][j];
    for (int i = 0; i < j; ++i) b[i] -= A[i][j] * b[j];
  }
  out[0] = static_cast<float>(b[0]); out[1] = static_cast<float>(b[1]); out[2] = static_cast<float>(b[2]);
  return true;
}

// pointMean (Main.hs:1596-1601): Float left fold, then multiply by 1 / n
bool point_mean_f32seq(const float* xyz, int64_t n, float out[3]) {
  if (n <= 0) return false;
  V3<float> s{0.f, 0.f, 0.f};
  for (int64_t i = 0; i < n; ++i) s = s + V3<float>{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  const float inv = 1.0f / static_cast<float>(n);
  out[0] = s.x * inv; out[1] = s.y * inv; out[2] = s.z * inv;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------------
// PCD v0.7
// ---------------------------------------------------------------------------------------------------------------------------
int PcdHeader::find(const std::string& name) const {
  for (size_t f = 0; f < fields.size(); ++f)
    if (fields[f] == name) return static_cast<int>(f);
  return -1;
}

bool pcd_parse_header(const char* buf, size_t len, PcdHeader* h, std::string* err) {
  *h = PcdHeader();
  size_t i = 0;
  bool have_data = false;
  while (i < len && !have_data) {
    size_t e = i;
    while (e < len && buf[e] != '\n') ++e;
    std::string line(buf + i, e - i);
    i = e < len ? e + 1 : e;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    std::string tok;
    if (key == "VERSION") continue;
    else if (key == "FIELDS" || key == "COLUMNS") { while (ss >> tok) h->fields.push_back(tok); }
    else if (key == "SIZE") { while (ss >> tok) h->size.push_back(std::atoi(tok.c_str())); }
    else if (key == "TYPE") { while (ss >> tok) h->type.push_back(tok.empty() ? '?' : tok[0]); }
    else if (key == "COUNT") { while (ss >> tok) h->count.push_back(std::atoi(tok.c_str())); }
    else if (key == "WIDTH") { ss >> h->width; }
    else if (key == "HEIGHT") { ss >> h->height; }
    else if (key == "VIEWPOINT") continue;
    else if (key == "POINTS") { ss >> h->points; }
    else if (key == "DATA") {
      ss >> tok;
      if (tok == "ascii") h->data_kind = 0;
      else if (tok == "binary") h->data_kind = 1;
      else if (tok == "binary_compressed") h->data_kind = 2;
      else { *err = "PCD: unknown DATA kind '" + tok + "'"; return false; }
      have_data = true;
    } else { *err = "PCD: unknown header entry '" + key + "'"; return false; }
  }
  if (!have_data) { *err = "PCD: no DATA line"; return false; }
  const size_t nf = h->fields.size();
  if (nf == 0 || h->size.size() != nf || h->type.size() != nf) { *err = "PCD: FIELDS / SIZE / TYPE do not match"; return false; }
  if (h->count.empty()) h->count.assign(nf, 1);
  if (h->count.size() != nf) { *err = "PCD: COUNT does not match FIELDS"; return false; }
  if (h->points < 0) {
    if (h->width < 0 || h->height < 0 || (h->height > 0 && h->width > static_cast<int64_t>(len) / h->height)) { *err = "PCD: WIDTH x HEIGHT exceeds the file size"; return false; }
    h->points = h->width * h->height;  // cannot overflow: the product is <= the file length
  }
  if (h->points < 0) { *err = "PCD: no POINTS / WIDTH x HEIGHT"; return false; }
  if (static_cast<uint64_t>(h->points) > len) { *err = "PCD: POINTS exceeds the file size"; return false; }  // every point takes >= 1 byte
  h->field_offset.resize(nf);
  int off = 0;
  for (size_t f = 0; f < nf; ++f) {
    if (h->size[f] != 1 && h->size[f] != 2 && h->size[f] != 4 && h->size[f] != 8) { *err = "PCD: bad SIZE"; return false; }
    if (h->count[f] < 0 || h->count[f] > (1 << 16)) { *err = "PCD: bad COUNT"; return false; }
    h->field_offset[f] = off;
    off += h->size[f] * h->count[f];
  }
  if (off <= 0 || off > (1 << 24)) { *err = "PCD: bad record size"; return false; }
  h->point_step = off;
  h->data_offset = i;
  return true;
}

// the layout the unpack kernel needs: x, y, z must be 4-byte floats; rgb / rgba (optional) a 4-byte F or U
bool pcd_layout(const PcdHeader& h, int* fx, int* fy, int* fz, int* frgb, std::string* err) {
  *fx = h.find("x"); *fy = h.find("y"); *fz = h.find("z");
  *frgb = h.find("rgb");
  if (*frgb < 0) *frgb = h.find("rgba");
  if (*fx < 0 || *fy < 0 || *fz < 0) { *err = "PCD: no x y z fields"; return false; }
  for (int f : {*fx, *fy, *fz})
    if (h.type[f] != 'F' || h.size[f] != 4 || h.count[f] != 1) { *err = "PCD: x y z must be 4-byte floats"; return false; }
  if (*frgb >= 0 && (h.size[*frgb] != 4 || h.count[*frgb] != 1 || h.type[*frgb] == 'I')) *frgb = -1;
  return true;
}

// DATA ascii: one point per line, fields separated by blanks; returned as tightly packed 4-byte records [x y z (rgb bits)]
bool pcd_ascii_records(const char* buf, size_t len, const PcdHeader& h, std::vector<uint32_t>* rec, int* rec_words, std::string* err) {
  int fx, fy, fz, frgb;
  if (!pcd_layout(h, &fx, &fy, &fz, &frgb, err)) return false;
  const int W = frgb >= 0 ? 4 : 3;
  *rec_words = W;
  rec->assign(static_cast<size_t>(h.points) * W, 0u);
  // column of every field's first value in a line
  std::vector<int> col(h.fields.size());
  int ncol = 0;
  for (size_t f = 0; f < h.fields.size(); ++f) { col[f] = ncol; ncol += h.count[f]; }
  const char* p = buf + h.data_offset;
  const char* end = buf + len;
  for (int64_t i = 0; i < h.points; ++i) {
    for (int c = 0; c < ncol; ++c) {
      while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
      if (p >= end) { *err = "PCD: ascii data ends early"; return false; }
      const char* t0 = p;
      while (p < end && !(*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
      int which = c == col[fx] ? 0 : (c == col[fy] ? 1 : (c == col[fz] ? 2 : ((frgb >= 0 && c == col[frgb]) ? 3 : -1)));
      if (which < 0) continue;
      const std::string tok(t0, p - t0);
      uint32_t bits;
      if (which == 3 && h.type[frgb] == 'U') {
        bits = static_cast<uint32_t>(std::strtoul(tok.c_str(), nullptr, 10));
      } else {
        const float v = std::strtof(tok.c_str(), nullptr);
        std::memcpy(&bits, &v, 4);
      }
      (*rec)[static_cast<size_t>(i) * W + which] = bits;
    }
  }
  return true;
}

// LZF (the codec of DATA binary_compressed): literal runs and back references, as published with liblzf
bool lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  size_t ip = 0, op = 0;
  while (ip < in_len) {
    unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      const size_t run = ctrl + 1;
      if (op + run > out_len || ip + run > in_len) return false;
      std::memcpy(out + op, in + ip, run);
      op += run; ip += run;
    } else {
      size_t l = ctrl >> 5;
      if (l == 7) { if (ip >= in_len) return false; l += in[ip++]; }
      if (ip >= in_len) return false;
      const size_t back = ((ctrl & 0x1f) << 8) + in[ip++] + 1;
      l += 2;
      if (back > op || op + l > out_len) return false;
      for (size_t k = 0; k < l; ++k, ++op) out[op] = out[op - back];  // may overlap: byte by byte
    }
  }
  return op == out_len;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Transform files written by the reference: roomProjectionToXfFormat (Main.hs:2287-2302, four lines of
// four numbers) and roomProjectionToString (Main.hs:2271-2284, sixteen numbers separated by commas).  Both hold the
// LEFT-multiplicative matrix, i.e. the transpose of roomProj; the result here is roomProj again (row vectors, p' = p .* M).
// ---------------------------------------------------------------------------------------------------------------------------
bool parse_transform_text(const char* text, size_t len, float m_rowmajor[16]) {
  float L[16];
  size_t i = 0;
  for (int k = 0; k < 16; ++k) {
    while (i < len && (text[i] == ' ' || text[i] == ',' || (text[i] >= 9 && text[i] <= 13))) ++i;
    if (i >= len) return false;
    const std::string rest(text + i, std::min<size_t>(len - i, 64));
    char* end = nullptr;
    const float v = std::strtof(rest.c_str(), &end);  // Haskell `show` of a Float: decimal or d.ddde-n, both strtof syntax
    if (end == rest.c_str()) return false;
    L[k] = v;
    i += static_cast<size_t>(end - rest.c_str());
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) m_rowmajor[4 * r + c] = L[4 * c + r];
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------------
// PLY (the format of the full-resolution KinFu meshes / clouds the external `plyxform` tool transforms, Main.hs:2287-2325)
// ---------------------------------------------------------------------------------------------------------------------------
static int ply_type_size(const std::string& t) {
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
  if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
  if (t == "double" || t == "float64") return 8;
  return 0;
}
int PlyHeader::find(const std::string& name) const {
  for (size_t f = 0; f < names.size(); ++f)
    if (names[f] == name) return static_cast<int>(f);
  return -1;
}
bool ply_parse_header(const char* buf, size_t len, PlyHeader* h, std::string* err) {
  *h = PlyHeader();
  size_t i = 0;
  int element = -1;  // 0: inside the vertex element, 1: a later element
  bool first = true, done = false;
  while (i < len && !done) {
    size_t e = i;
    while (e < len && buf[e] != '\n') ++e;
    std::string line(buf + i, e - i);
    i = e < len ? e + 1 : e;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (first) { if (line != "ply") { *err = "PLY: missing magic"; return false; } first = false; continue; }
    std::istringstream ss(line);
    std::string key, a, b;
    ss >> key;
    if (key == "format") {
      ss >> a;
      if (a == "ascii") h->ascii = true;
      else if (a == "binary_little_endian") h->ascii = false;
      else { *err = "PLY: unsupported format '" + a + "'"; return false; }
    } else if (key == "element") {
      ss >> a >> b;
      if (element < 0) {
        if (a != "vertex") { *err = "PLY: the first element must be `vertex`"; return false; }
        h->n = std::atoll(b.c_str());
        element = 0;
      } else element = 1;
    } else if (key == "property" && element == 0) {
      ss >> a;
      if (a == "list") { *err = "PLY: list property in the vertex element"; return false; }
      ss >> b;
      const int sz = ply_type_size(a);
      if (!sz) { *err = "PLY: unknown property type '" + a + "'"; return false; }
      h->names.push_back(b); h->types.push_back(a); h->sizes.push_back(sz); h->offsets.push_back(h->step);
      h->step += sz;
    } else if (key == "end_header") done = true;
  }
  if (!done || element < 0 || h->n < 0) { *err = "PLY: incomplete header"; return false; }
  if (static_cast<uint64_t>(h->n) > len) { *err = "PLY: vertex count exceeds the file size"; return false; }
  h->data_offset = i;
  return true;
}
// vertex records of an ascii PLY as packed 4-byte records [x y z (0x00RRGGBB)]
bool ply_ascii_records(const char* buf, size_t len, const PlyHeader& h, bool want_rgb, std::vector<uint32_t>* rec, std::string* err) {
  const int fx = h.find("x"), fy = h.find("y"), fz = h.find("z"), fr = h.find("red"), fg = h.find("green"), fb = h.find("blue");
  const int W = want_rgb ? 4 : 3;
  rec->assign(static_cast<size_t>(h.n) * W, 0u);
  const char* p = buf + h.data_offset;
  const char* end = buf + len;
  const int nprop = static_cast<int>(h.names.size());
  for (int64_t i = 0; i < h.n; ++i) {
    uint32_t rgb = 0;
    for (int c = 0; c < nprop; ++c) {
      while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
      if (p >= end) { *err = "PLY: ascii data ends early"; return false; }
      const char* t0 = p;
      while (p < end && !(*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
      const std::string tok(t0, p - t0);
      if (c == fx || c == fy || c == fz) {
        const float v = std::strtof(tok.c_str(), nullptr);
        std::memcpy(&(*rec)[static_cast<size_t>(i) * W + (c == fx ? 0 : (c == fy ? 1 : 2))], &v, 4);
      } else if (want_rgb && (c == fr || c == fg || c == fb)) {
        const uint32_t v = static_cast<uint32_t>(std::strtoul(tok.c_str(), nullptr, 10)) & 255u;
        rgb |= v << (c == fr ? 16 : (c == fg ? 8 : 0));
      }
    }
    if (want_rgb) (*rec)[static_cast<size_t>(i) * W + 3] = rgb;
  }
  return true;
}

// binary PCD v0.7 with FIELDS x y z [rgb] (what pcl_transform_point_cloud would have written, Main.hs:2305-2313)
bool write_pcd(const char* path, const float* xyz, const uint8_t* rgb, int64_t n, std::string* err) {
  FILE* fp = std::fopen(path, "wb");
  if (!fp) { *err = std::string("cannot open ") + path; return false; }
  std::fprintf(fp, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z%s\nSIZE 4 4 4%s\nTYPE F F F%s\nCOUNT 1 1 1%s\n"
                   "WIDTH %lld\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %lld\nDATA binary\n",
               rgb ? " rgb" : "", rgb ? " 4" : "", rgb ? " F" : "", rgb ? " 1" : "", static_cast<long long>(n), static_cast<long long>(n));
  bool ok = true;
  if (!rgb) ok = std::fwrite(xyz, 12, static_cast<size_t>(n), fp) == static_cast<size_t>(n);
  else {
    const size_t chunk = 1 << 16;
    std::vector<uint8_t> buf(chunk * 16);
    for (int64_t i0 = 0; i0 < n && ok; i0 += chunk) {
      const size_t m = static_cast<size_t>(std::min<int64_t>(chunk, n - i0));
      for (size_t i = 0; i < m; ++i) {
        std::memcpy(&buf[i * 16], xyz + 3 * (i0 + i), 12);
        const uint8_t* c = rgb + 3 * (i0 + i);
        const uint32_t packed = (static_cast<uint32_t>(c[0]) << 16) | (static_cast<uint32_t>(c[1]) << 8) | c[2];
        std::memcpy(&buf[i * 16 + 12], &packed, 4);
      }
      ok = std::fwrite(buf.data(), 16, m, fp) == m;
    }
  }
  if (std::fclose(fp) != 0) ok = false;
  if (!ok) *err = std::string("short write to ") + path;
  return ok;
}

bool read_file(const char* path, std::vector<char>* out, std::string* err) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { *err = std::string(path) + ": " + std::strerror(errno); return false; }
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out->resize(sz > 0 ? static_cast<size_t>(sz) : 0);
  const size_t got = sz > 0 ? std::fread(out->data(), 1, out->size(), f) : 0;
  std::fclose(f);
  if (got != out->size()) { *err = std::string(path) + ": short read"; return false; }
  return true;
}

}  // namespace hs
