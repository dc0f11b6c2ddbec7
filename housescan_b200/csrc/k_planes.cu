// Point-to-plane kernels: assignment (A5), cuboid objective/gradient sums per room (A6), generic per-plane sums (A13 inputs).
// Reference semantics: signedDistanceToPlaneEq Main.hs:1371-1372; nearest plane = first minimum of |distance|
// (minimumBy, FitCuboidBFGS.hs:74); cuboid planes from makePlanesFromCuboid Main.hs:1852-1874 (built on the host).
// All per-point geometry is Float without FMA contraction (bit-exact assignment); all sums are Double.
#include "k_common.cuh"
#include "k_ring.cuh"

namespace hsk {

// ------------------------------------------------------------------------------------------------------------------
// nearest of 6 planes held in registers; returns plane index, signed residual in r
// ------------------------------------------------------------------------------------------------------------------
struct Planes6 {
  float nx[6], ny[6], nz[6], d[6];
};
__device__ __forceinline__ Planes6 load_planes6(const RoomTable& tbl, int r) {
  Planes6 P;
#pragma unroll
  for (int k = 0; k < 6; ++k) { P.nx[k] = tbl.pl[r][k][0]; P.ny[k] = tbl.pl[r][k][1]; P.nz[k] = tbl.pl[r][k][2]; P.d[k] = tbl.pl[r][k][3]; }
  return P;
}
__device__ __forceinline__ int nearest6(const Planes6& P, float x, float y, float z, float& r) {
  float rb = plane_dist(P.nx[0], P.ny[0], P.nz[0], P.d[0], x, y, z);
  float ab = fabsf(rb);
  int kb = 0;
#pragma unroll
  for (int k = 1; k < 6; ++k) {
    const float rk = plane_dist(P.nx[k], P.ny[k], P.nz[k], P.d[k], x, y, z);
    const float ak = fabsf(rk);
    const bool lt = ak < ab;  // strict: ties keep the lower index
    ab = lt ? ak : ab;
    rb = lt ? rk : rb;
    kb = lt ? k : kb;
  }
  r = rb;
  return kb;
}

// ------------------------------------------------------------------------------------------------------------------
// Kernel EXACT (mode 0): every per-point product is formed in Double (exact for Float operands), sums in Double.
//   acc[0] f = sum r^2; acc[1..6] Sr[k]; acc[7..15] B[j][c] = sum s p_c (s = r on the + wall, -r on the - wall); acc[16..21] counts
// ------------------------------------------------------------------------------------------------------------------
struct AccExact {
  double v[HS_NACC];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < HS_NACC; ++i) v[i] = 0.0;
  }
  __device__ __forceinline__ void add(const Planes6& P, float x, float y, float z) {
    float r;
    const int k = nearest6(P, x, y, z, r);
    const double rd = static_cast<double>(r);
    v[0] = fma(rd, rd, v[0]);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const bool m = (k == q);
      v[1 + q] += m ? rd : 0.0;
      v[16 + q] += m ? 1.0 : 0.0;
    }
    const double s = (k & 1) ? -rd : rd;
    const int j = k >> 1;
    const double xd = x, yd = y, zd = z;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double w = (j == a) ? s : 0.0;
      v[7 + 3 * a + 0] = fma(w, xd, v[7 + 3 * a + 0]);
      v[7 + 3 * a + 1] = fma(w, yd, v[7 + 3 * a + 1]);
      v[7 + 3 * a + 2] = fma(w, zd, v[7 + 3 * a + 2]);
    }
  }
};

// Work split: the cloud is cut into groups of 4 points (48 B); block b owns groups [b*gpb, (b+1)*gpb).  For every room
// overlapping its range the block reduces one partial record; the block that finishes last sums the partials per room in
// block order (deterministic for a fixed grid).
template <class Acc>
__global__ void __launch_bounds__(HS_TPB, 2)
k_rooms_cuboid_sums(const float* __restrict__ xyz, int64_t n, const __grid_constant__ RoomTable tbl, int64_t gpb,
                    double* __restrict__ partials, int* __restrict__ meta, unsigned int* ticket, double* __restrict__ out) {
  __shared__ double smem[(HS_TPB / 32) * HS_NACC];
  const int nrooms = tbl.nrooms;
  const int64_t G = (n + 3) >> 2;
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * gpb;
  const int64_t g1 = min(g0 + gpb, G);
  const int64_t p0 = g0 * 4, p1 = min(g1 * 4, n);
  int rfirst = -1, rlast = -2;
  for (int r = 0; r < nrooms; ++r)
    if (tbl.off[r] < p1 && tbl.off[r + 1] > p0) { if (rfirst < 0) rfirst = r; rlast = r; }
  if (threadIdx.x == 0) meta[blockIdx.x] = rfirst;

  for (int r = rfirst; r <= rlast; ++r) {
    const int64_t lo = max(tbl.off[r], p0), hi = min(tbl.off[r + 1], p1);
    const Planes6 P = load_planes6(tbl, r);
    Acc A;
    A.clear();
    const int64_t gl = (lo + 3) >> 2, gh = hi >> 2;  // whole groups inside [lo, hi)
    if (gl <= gh) {
      // ragged head / tail points (at most 3 each), one per thread
      const int64_t head_end = gl * 4, tail_begin = gh * 4;
      const int64_t nh = head_end - lo, nt = hi - tail_begin;
      if (threadIdx.x < nh) { const int64_t i = lo + threadIdx.x; A.add(P, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      else if (threadIdx.x >= 32 && threadIdx.x - 32 < nt) { const int64_t i = tail_begin + threadIdx.x - 32; A.add(P, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      int64_t g = gl + threadIdx.x;
      for (; g + HS_TPB < gh; g += 2 * HS_TPB) {  // two groups in flight per thread
        const Pts4 a = load_group(xyz, g), b = load_group(xyz, g + HS_TPB);
#pragma unroll
        for (int e = 0; e < 4; ++e) A.add(P, a.x[e], a.y[e], a.z[e]);
#pragma unroll
        for (int e = 0; e < 4; ++e) A.add(P, b.x[e], b.y[e], b.z[e]);
      }
      for (; g < gh; g += HS_TPB) {
        const Pts4 a = load_group(xyz, g);
#pragma unroll
        for (int e = 0; e < 4; ++e) A.add(P, a.x[e], a.y[e], a.z[e]);
      }
    } else {
      // the whole overlap lies inside one group
      const int64_t i = lo + threadIdx.x;
      if (i < hi) A.add(P, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
    block_sum_store<HS_NACC>(A.v, partials + (static_cast<int64_t>(blockIdx.x) * nrooms + (r - rfirst)) * HS_NACC, smem);
  }

  if (!last_block_arrives(ticket, gridDim.x)) return;
  const int64_t ppb = gpb * 4;
  for (int o = threadIdx.x; o < nrooms * HS_REC; o += HS_TPB) {
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

// ------------------------------------------------------------------------------------------------------------------
// plane assignment, generic K <= 16 planes (A5): writes uint8 index and optionally the Float residual
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearestK(const PlaneTable& t, float x, float y, float z, float& r) {
  float rb = plane_dist(t.pl[0][0], t.pl[0][1], t.pl[0][2], t.pl[0][3], x, y, z);
  float ab = fabsf(rb);
  int kb = 0;
  for (int k = 1; k < t.K; ++k) {
    const float rk = plane_dist(t.pl[k][0], t.pl[k][1], t.pl[k][2], t.pl[k][3], x, y, z);
    const float ak = fabsf(rk);
    const bool lt = ak < ab;
    ab = lt ? ak : ab; rb = lt ? rk : rb; kb = lt ? k : kb;
  }
  r = rb;
  return kb;
}

// nearest plane (k_common.cuh nearest_plane) + the winner's residual recomputed from its plane: one LDS.128 and the same six
// operations as in the chain, hence the same bits
template <int KT, bool PAIRED>
__device__ __forceinline__ int nearestK_lds(const PlaneTable& t, const float4* __restrict__ spl, float x, float y, float z, float& r) {
  float ab;
  const int kb = nearest_plane<KT, PAIRED>(t, x, y, z, ab);
  const float4 nn = spl[kb];
  r = plane_dist(nn.x, nn.y, nn.z, nn.w, x, y, z);
  return kb;
}
#define nearestK(tbl, x, y, z, r) nearestK_lds<KT, PAIRED>(tbl, spl, x, y, z, r)

template <int KT, bool PAIRED>
__global__ void __launch_bounds__(HS_TPB)
k_plane_assign(const float* __restrict__ xyz, int64_t n, const __grid_constant__ PlaneTable tbl, uint8_t* __restrict__ assign,
               float* __restrict__ resid) {
  __shared__ float4 spl[16];
  if (threadIdx.x < tbl.K) spl[threadIdx.x] = make_float4(tbl.pl[threadIdx.x][0], tbl.pl[threadIdx.x][1], tbl.pl[threadIdx.x][2], tbl.pl[threadIdx.x][3]);
  __syncthreads();
  const int64_t gfull = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; g < gfull; g += stride) {
    const Pts4 p = load_group(xyz, g);
    float r[4];
    uchar4 k;
    k.x = nearestK(tbl, p.x[0], p.y[0], p.z[0], r[0]);
    k.y = nearestK(tbl, p.x[1], p.y[1], p.z[1], r[1]);
    k.z = nearestK(tbl, p.x[2], p.y[2], p.z[2], r[2]);
    k.w = nearestK(tbl, p.x[3], p.y[3], p.z[3], r[3]);
    if (assign) reinterpret_cast<uchar4*>(assign)[g] = k;
    if (resid) reinterpret_cast<float4*>(resid)[g] = make_float4(r[0], r[1], r[2], r[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = gfull * 4 + threadIdx.x;
    float r;
    const int k = nearestK(tbl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], r);
    if (assign) assign[i] = static_cast<uint8_t>(k);
    if (resid) resid[i] = r;
  }
}
#undef nearestK

// ------------------------------------------------------------------------------------------------------------------
// generic per-plane sums over one point range [i0, i1): out[k] = count, sum r, sum r^2, sum p(3), sum r p(3), max|r|
// (one launch per room; wall-alignment inputs, not the throughput path)
// ------------------------------------------------------------------------------------------------------------------
// one HS_PS record (9 sums + max|r|) per block at a stride: all threads load, slices are combined in slice order
__device__ __forceinline__ void last_block_sum_strided(const double* __restrict__ partials, unsigned int nblocks, int stride, double* __restrict__ dst, double* smem) {
  constexpr int S = HS_TPB / HS_PS;
  const int c = threadIdx.x % HS_PS, sl = threadIdx.x / HS_PS;
  double acc = 0.0;
  if (sl < S)
    for (unsigned int b = sl; b < nblocks; b += S) {
      const double v = __ldcg(partials + static_cast<int64_t>(b) * stride + c);
      acc = (c == 9) ? fmax(acc, v) : acc + v;
    }
  smem[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < HS_PS) {
    double t = 0.0;
    for (int q = 0; q < S; ++q) { const double v = smem[q * HS_PS + threadIdx.x]; t = (threadIdx.x == 9) ? fmax(t, v) : t + v; }
    dst[threadIdx.x] = t;
  }
  __syncthreads();
}

template <int K>
__global__ void __launch_bounds__(HS_TPB)
k_plane_sums(const float* __restrict__ xyz, int64_t i0, int64_t i1, const __grid_constant__ PlaneTable tbl,
             double* __restrict__ partials, unsigned int* ticket, double* __restrict__ out) {
  __shared__ double smem[(HS_TPB / 32) * 9];
  __shared__ float smax[HS_TPB / 32];
  double acc[K][9];
  float mx[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { mx[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[k][c] = 0.0; }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
  for (int64_t i = i0 + static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x; i < i1; i += stride) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    float r;
    const int kk = nearestK(tbl, x, y, z, r);
    const double rd = r, xd = x, yd = y, zd = z;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool m = (kk == k);
      const double w = m ? 1.0 : 0.0, wr = m ? rd : 0.0;
      acc[k][0] += w; acc[k][1] += wr; acc[k][2] = fma(wr, rd, acc[k][2]);
      acc[k][3] = fma(w, xd, acc[k][3]); acc[k][4] = fma(w, yd, acc[k][4]); acc[k][5] = fma(w, zd, acc[k][5]);
      acc[k][6] = fma(wr, xd, acc[k][6]); acc[k][7] = fma(wr, yd, acc[k][7]); acc[k][8] = fma(wr, zd, acc[k][8]);
      mx[k] = m ? fmaxf(mx[k], fabsf(r)) : mx[k];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    block_sum_store<9>(acc[k], partials + (static_cast<int64_t>(blockIdx.x) * K + k) * HS_PS, smem);
    float m = mx[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) smax[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < HS_TPB / 32; ++w) t = fmaxf(t, smax[w]);
      partials[(static_cast<int64_t>(blockIdx.x) * K + k) * HS_PS + 9] = t;
    }
    __syncthreads();
  }
  if (!last_block_arrives(ticket, gridDim.x)) return;
  __shared__ double fin[HS_TPB];
#pragma unroll
  for (int k = 0; k < K; ++k)  // record k of block b sits at partials[(b * K + k) * HS_PS]: stride K * HS_PS between blocks
    last_block_sum_strided(partials + static_cast<int64_t>(k) * HS_PS, gridDim.x, K * HS_PS, out + k * HS_PS, fin);
}


// Throughput form of the same record (default for K <= 6).  The plane choice and the residual are the same Float operations as
// above (bit-identical); the nine sums per plane run as predicated Float chains over 32 points per thread (the product inside an
// FMA is exact, only the short partial sums are rounded to Float), and are then added into per-thread Doubles that live in shared
// memory, so the registers hold the 10 K chains and four points in flight instead of 9 K Doubles.  Points are read as 4-point
// groups (3 x LDG.128); the ragged head and tail of the range (room offsets are arbitrary) are taken one point per thread.
// One axis of a cuboid room's wall pair (K == 6, antiparallel pairs).  The axis wins if its nearer wall is the first minimum of the
// six |distances| (strict <, ties keep the lower wall index: the scan order of nearest_plane); inside the pair the - wall wins only
// if it is strictly nearer.  The two walls' predicates come straight from the comparisons (no index, no six setp.eq), every wall
// accumulates its own signed residual (no select), and an axis is one PTX block of 30 operands.
//   axis 0: not (a1 < a0) and not (a2 < min(a0, a1));   axis 1: (a1 < a0) and not (a2 < min(a0, a1));   axis 2: a2 < min(a0, a1)
#define HS_PS_PAIR_BODY(E_SETUP)                                                                                                   \
  asm("{\n .reg .pred E2, E, Wp, Wm;\n .reg .f32 a01;\n"                                                                          \
      " min.f32 a01, %27, %28;\n setp.lt.f32 E2, %29, a01;\n" E_SETUP                                                              \
      " setp.lt.and.f32 Wm, %26, %25, E;\n setp.geu.and.f32 Wp, %26, %25, E;\n"                                                    \
      "@Wp add.rn.f32 %0, %0, 0f3F800000;\n @Wp add.rn.f32 %1, %1, %23;\n @Wp fma.rn.f32 %2, %23, %23, %2;\n"                     \
      "@Wp add.rn.f32 %3, %3, %20;\n @Wp add.rn.f32 %4, %4, %21;\n @Wp add.rn.f32 %5, %5, %22;\n"                                 \
      "@Wp fma.rn.f32 %6, %23, %20, %6;\n @Wp fma.rn.f32 %7, %23, %21, %7;\n @Wp fma.rn.f32 %8, %23, %22, %8;\n"                  \
      "@Wp max.f32 %9, %9, %25;\n"                                                                                                 \
      "@Wm add.rn.f32 %10, %10, 0f3F800000;\n @Wm add.rn.f32 %11, %11, %24;\n @Wm fma.rn.f32 %12, %24, %24, %12;\n"               \
      "@Wm add.rn.f32 %13, %13, %20;\n @Wm add.rn.f32 %14, %14, %21;\n @Wm add.rn.f32 %15, %15, %22;\n"                           \
      "@Wm fma.rn.f32 %16, %24, %20, %16;\n @Wm fma.rn.f32 %17, %24, %21, %17;\n @Wm fma.rn.f32 %18, %24, %22, %18;\n"            \
      "@Wm max.f32 %19, %19, %26;\n}\n"                                                                                            \
      : "+f"(ap[0]), "+f"(ap[1]), "+f"(ap[2]), "+f"(ap[3]), "+f"(ap[4]), "+f"(ap[5]), "+f"(ap[6]), "+f"(ap[7]), "+f"(ap[8]), "+f"(mxp),  \
        "+f"(am[0]), "+f"(am[1]), "+f"(am[2]), "+f"(am[3]), "+f"(am[4]), "+f"(am[5]), "+f"(am[6]), "+f"(am[7]), "+f"(am[8]), "+f"(mxm)   \
      : "f"(x), "f"(y), "f"(z), "f"(sp), "f"(rm), "f"(asp), "f"(arm), "f"(a0), "f"(a1), "f"(a2))
template <int AXIS>
__device__ __forceinline__ void ps_add_pair(float (&ap)[9], float& mxp, float (&am)[9], float& mxm, float x, float y, float z, float sp, float rm,
                                            float asp, float arm, float a0, float a1, float a2) {
  if (AXIS == 0) HS_PS_PAIR_BODY(" setp.geu.and.f32 E, %28, %27, !E2;\n");
  else if (AXIS == 1) HS_PS_PAIR_BODY(" setp.lt.and.f32 E, %28, %27, !E2;\n");
  else HS_PS_PAIR_BODY(" setp.lt.f32 E, %29, a01;\n");
}
#undef HS_PS_PAIR_BODY

template <int K, bool PAIRED>
__device__ __forceinline__ void ps_add(float (&a)[K][9], float (&mx)[K], const PlaneTable& tbl, const float4* __restrict__ spl, float x, float y, float z) {
  if constexpr (K == 6 && PAIRED) {
    float sp[3], rm[3], asp[3], arm[3], aj[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {  // the very operations of nearest_plane's paired form: same bits, same winner
      const float tj = __fadd_rn(__fadd_rn(__fmul_rn(tbl.pl[2 * j][0], x), __fmul_rn(tbl.pl[2 * j][1], y)), __fmul_rn(tbl.pl[2 * j][2], z));
      sp[j] = __fsub_rn(tj, tbl.pl[2 * j][3]);
      rm[j] = __fsub_rn(-tj, tbl.pl[2 * j + 1][3]);
      asp[j] = fabsf(sp[j]);
      arm[j] = fabsf(rm[j]);
      aj[j] = fminf(asp[j], arm[j]);
    }
    ps_add_pair<0>(a[0], mx[0], a[1], mx[1], x, y, z, sp[0], rm[0], asp[0], arm[0], aj[0], aj[1], aj[2]);
    ps_add_pair<1>(a[2], mx[2], a[3], mx[3], x, y, z, sp[1], rm[1], asp[1], arm[1], aj[0], aj[1], aj[2]);
    ps_add_pair<2>(a[4], mx[4], a[5], mx[5], x, y, z, sp[2], rm[2], asp[2], arm[2], aj[0], aj[1], aj[2]);
    return;
  }
  float ab;
  const int kb = nearest_plane<K, PAIRED>(tbl, x, y, z, ab);
  const float4 nn = spl[kb];
  const float r = plane_dist(nn.x, nn.y, nn.z, nn.w, x, y, z);  // the winner's residual: same operations, same bits
  // predicated FP instructions, no branches and no selects: the four points a thread holds stay independent streams
#pragma unroll
  for (int k = 0; k < K; ++k)
    asm("{\n .reg .pred p;\n setp.eq.s32 p, %14, %15;\n"
        "@p add.rn.f32 %0, %0, 0f3F800000;\n @p add.rn.f32 %1, %1, %10;\n @p fma.rn.f32 %2, %10, %10, %2;\n"
        "@p add.rn.f32 %3, %3, %11;\n @p add.rn.f32 %4, %4, %12;\n @p add.rn.f32 %5, %5, %13;\n"
        "@p fma.rn.f32 %6, %10, %11, %6;\n @p fma.rn.f32 %7, %10, %12, %7;\n @p fma.rn.f32 %8, %10, %13, %8;\n"
        "@p max.f32 %9, %9, %16;\n}\n"
        : "+f"(a[k][0]), "+f"(a[k][1]), "+f"(a[k][2]), "+f"(a[k][3]), "+f"(a[k][4]), "+f"(a[k][5]), "+f"(a[k][6]), "+f"(a[k][7]), "+f"(a[k][8]), "+f"(mx[k])
        : "f"(r), "f"(x), "f"(y), "f"(z), "r"(kb), "r"(k), "f"(ab));
}

template <int K, bool PAIRED>
__global__ void __launch_bounds__(HS_TPB, 2)
k_plane_sums_f32(const float* __restrict__ xyz, int64_t i0, int64_t i1, const __grid_constant__ PlaneTable tbl,
                 double* __restrict__ partials, unsigned int* ticket, double* __restrict__ out) {
  extern __shared__ double sdacc[];  // [K * 9][HS_TPB]
  __shared__ double smem[(HS_TPB / 32) * 9];
  __shared__ float smax[HS_TPB / 32];
  __shared__ float4 spl[HS_MAX_PLANES];
  if (threadIdx.x < K) spl[threadIdx.x] = make_float4(tbl.pl[threadIdx.x][0], tbl.pl[threadIdx.x][1], tbl.pl[threadIdx.x][2], tbl.pl[threadIdx.x][3]);
  float a[K][9], mx[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    mx[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c) { a[k][c] = 0.f; sdacc[(k * 9 + c) * HS_TPB + threadIdx.x] = 0.0; }
  }
  __syncthreads();
  const int64_t gl = (i0 + 3) >> 2, gh = i1 >> 2;  // whole groups inside [i0, i1)
  if (gl <= gh) {
    if (blockIdx.x == 0) {  // ragged head / tail points (at most 3 each)
      const int64_t nh = gl * 4 - i0, nt = i1 - gh * 4;
      if (threadIdx.x < nh) { const int64_t i = i0 + threadIdx.x; ps_add<K, PAIRED>(a, mx, tbl, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      else if (threadIdx.x >= 32 && threadIdx.x - 32 < nt) { const int64_t i = gh * 4 + threadIdx.x - 32; ps_add<K, PAIRED>(a, mx, tbl, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
    }
    const int64_t stride = static_cast<int64_t>(gridDim.x) * HS_TPB;
    int it = 0;
    int64_t g = gl + static_cast<int64_t>(blockIdx.x) * HS_TPB + threadIdx.x;
    Pts4 p, nx;
    if (g < gh) nx = load_group(xyz, g);
    for (; g < gh; g += stride) {
      p = nx;
      if (g + stride < gh) nx = load_group(xyz, g + stride);  // the next group is in flight while this one is evaluated
#pragma unroll
      for (int e = 0; e < 4; ++e) ps_add<K, PAIRED>(a, mx, tbl, spl, p.x[e], p.y[e], p.z[e]);
      if ((++it & 7) == 0) {  // 32 points per chain
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int c = 0; c < 9; ++c) { sdacc[(k * 9 + c) * HS_TPB + threadIdx.x] += static_cast<double>(a[k][c]); a[k][c] = 0.f; }
      }
    }
  } else if (blockIdx.x == 0) {  // the whole range lies inside one group
    const int64_t i = i0 + threadIdx.x;
    if (i < i1) ps_add<K, PAIRED>(a, mx, tbl, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double v[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) v[c] = sdacc[(k * 9 + c) * HS_TPB + threadIdx.x] + static_cast<double>(a[k][c]);
    block_sum_store<9>(v, partials + (static_cast<int64_t>(blockIdx.x) * K + k) * HS_PS, smem);
    float m = mx[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) smax[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < HS_TPB / 32; ++w) t = fmaxf(t, smax[w]);
      partials[(static_cast<int64_t>(blockIdx.x) * K + k) * HS_PS + 9] = t;
    }
    __syncthreads();
  }
  if (!last_block_arrives(ticket, gridDim.x)) return;
  // all K records at once: value idx = k * HS_PS + c of block b sits at partials[b * NV + idx]; slices of blocks are summed in
  // slice order (a fixed tree for a fixed grid); component 9 is max |r|
  __shared__ double fin[HS_TPB];
  constexpr int NV = K * HS_PS, S = HS_TPB / NV;
  const int idx = threadIdx.x % NV, sl = threadIdx.x / NV;
  const bool is_max = (idx % HS_PS) == 9;
  double acc = 0.0;
  if (sl < S)
    for (unsigned int b = sl; b < gridDim.x; b += S) {
      const double v = __ldcg(partials + static_cast<int64_t>(b) * NV + idx);
      acc = is_max ? fmax(acc, v) : acc + v;
    }
  fin[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
    for (int q = 0; q < S; ++q) { const double v = fin[q * NV + threadIdx.x]; t = is_max ? fmax(t, v) : t + v; }
    out[threadIdx.x] = t;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Ring form of the per-plane sums (default for K <= 6).  Same per-point block as k_plane_sums_f32 (ps_add), but
//   * the points arrive through warp-private rings of 1-D bulk async copies (the TMA engine): lane 0 of a warp keeps D - 1 tiles
//     of 32 groups (1536 B) in flight for ITS warp, so ~70 KB per SM are always on their way instead of one 48-byte group per
//     thread, and nothing couples the warps of a block inside the streaming loop;
//   * the Float chains are summed across the warp with shuffles every 64 points per lane and only the 9 K warp totals are
//     added in Double (a [warps][9 K] table): no per-thread Doubles, so the 110 KB they took are the rings' now.
// Block b owns a contiguous range of groups; its warps take the tiles of that range round-robin.
// ------------------------------------------------------------------------------------------------------------------
#define PSR_D 4                 // ring slots per warp
#define PSR_TILE_BYTES 1536u    // 32 groups x 48 B
#define PSR_FLUSH_TILES 64      // 256 points per lane between flushes (the chain length of the evaluation kernel)

template <int K, bool PAIRED>
__global__ void __launch_bounds__(HS_TPB, 2)
k_plane_sums_ring(const float* __restrict__ xyz, int64_t i0, int64_t i1, const __grid_constant__ PlaneTable tbl, int64_t gpb,
                  double* __restrict__ partials, unsigned int* ticket, double* __restrict__ out) {
  constexpr int NW = HS_TPB / 32, NV = K * 9;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* rings = smem_raw;                                                                        // [NW][D][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(NW) * PSR_D * PSR_TILE_BYTES);  // [NW][D]
  double* wacc = reinterpret_cast<double*>(full + NW * PSR_D);                                            // [NW][NV]
  float* wtmp = reinterpret_cast<float*>(wacc + NW * NV);                                                  // [NW][64]
  float* wmax = wtmp + NW * 64;                                                                            // [NW][8]
  __shared__ float4 spl[HS_MAX_PLANES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < K) spl[threadIdx.x] = make_float4(tbl.pl[threadIdx.x][0], tbl.pl[threadIdx.x][1], tbl.pl[threadIdx.x][2], tbl.pl[threadIdx.x][3]);
  uint64_t* my_full = full + warp * PSR_D;
  if (lane == 0) {
    for (int sl = 0; sl < PSR_D; ++sl) mbar_init(my_full + sl, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double* my_wacc = wacc + warp * NV;
  float* my_wtmp = wtmp + warp * 64;
  for (int q = lane; q < NV; q += 32) my_wacc[q] = 0.0;
  __syncthreads();

  float a[K][9], mx[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    mx[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 9; ++c) a[k][c] = 0.f;
  }
  auto flush = [&]() {  // chains of the whole warp -> the warp's Double table
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        float v = a[k][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) my_wtmp[k * 9 + c] = v;
        a[k][c] = 0.f;
      }
    __syncwarp();
    for (int q = lane; q < NV; q += 32) my_wacc[q] += static_cast<double>(my_wtmp[q]);
    __syncwarp();
  };

  const int64_t gl = (i0 + 3) >> 2, gh = i1 >> 2;  // whole groups inside [i0, i1)
  if (gl <= gh) {
    if (blockIdx.x == 0) {  // ragged head / tail points (at most 3 each)
      const int64_t nh = gl * 4 - i0, nt = i1 - gh * 4;
      if (threadIdx.x < nh) { const int64_t i = i0 + threadIdx.x; ps_add<K, PAIRED>(a, mx, tbl, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
      else if (threadIdx.x >= 32 && threadIdx.x - 32 < nt) { const int64_t i = gh * 4 + threadIdx.x - 32; ps_add<K, PAIRED>(a, mx, tbl, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
    }
    const int64_t b0 = gl + static_cast<int64_t>(blockIdx.x) * gpb, b1 = min(b0 + gpb, gh);  // this block's groups
    const int64_t ngroups = b1 > b0 ? b1 - b0 : 0;
    const int64_t ntiles = (ngroups + 31) / 32;
    const int64_t mine = ntiles > warp ? (ntiles - warp + NW - 1) / NW : 0;  // tiles warp, warp + NW, ...
    const float4* src = reinterpret_cast<const float4*>(xyz) + 3 * b0;
    const uint32_t ring_s = smem_u32(rings) + warp * (PSR_D * PSR_TILE_BYTES), full_s = smem_u32(my_full);
    auto issue = [&](int64_t i) {  // lane 0: bulk copy of my i-th tile into slot i % D
      const int64_t tg = (warp + i * NW) * 32;
      const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(32), ngroups - tg) * 48);
      const uint32_t slot = static_cast<uint32_t>(i % PSR_D);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_s + 8 * slot), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ring_s + slot * PSR_TILE_BYTES), "l"(src + 3 * tg), "r"(bytes), "r"(full_s + 8 * slot) : "memory");
    };
    if (lane == 0)
      for (int64_t i = 0; i < PSR_D - 1 && i < mine; ++i) issue(i);
    int since_flush = 0;
    for (int64_t i = 0; i < mine; ++i) {
      __syncwarp();  // every lane has read the slot that tile i + D - 1 overwrites (it held tile i - 1)
      if (lane == 0 && i + PSR_D - 1 < mine) issue(i + PSR_D - 1);
      const uint32_t slot = static_cast<uint32_t>(i % PSR_D);
      mbar_wait_s_spin(full_s + 8 * slot, static_cast<uint32_t>((i / PSR_D) & 1));
      const int64_t left = ngroups - (warp + i * NW) * 32;  // groups in this tile (>= 1)
      const uint32_t base = ring_s + slot * PSR_TILE_BYTES + lane * 48;
      const float4 q0 = lds_v4(base), q1 = lds_v4(base + 16), q2 = lds_v4(base + 32);  // lanes past a partial tile read stale (valid) smem
      if (lane < left) {
        ps_add<K, PAIRED>(a, mx, tbl, spl, q0.x, q0.y, q0.z);
        ps_add<K, PAIRED>(a, mx, tbl, spl, q0.w, q1.x, q1.y);
        ps_add<K, PAIRED>(a, mx, tbl, spl, q1.z, q1.w, q2.x);
        ps_add<K, PAIRED>(a, mx, tbl, spl, q2.y, q2.z, q2.w);
      }
      if (++since_flush == PSR_FLUSH_TILES) { flush(); since_flush = 0; }
    }
  } else if (blockIdx.x == 0) {  // the whole range lies inside one group
    const int64_t i = i0 + threadIdx.x;
    if (i < i1) ps_add<K, PAIRED>(a, mx, tbl, spl, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }
  flush();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float m = mx[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) wmax[warp * 8 + k] = m;
  }
  __syncthreads();
  // the block's K records (HS_PS layout: 9 sums + max |r|), warp tables combined in warp order
  if (threadIdx.x < K * HS_PS) {
    const int k = threadIdx.x / HS_PS, c = threadIdx.x % HS_PS;
    double v = 0.0;
    if (c < 9) { for (int w = 0; w < NW; ++w) v += wacc[w * NV + k * 9 + c]; }
    else { float m = 0.f; for (int w = 0; w < NW; ++w) m = fmaxf(m, wmax[w * 8 + k]); v = m; }
    partials[static_cast<int64_t>(blockIdx.x) * (K * HS_PS) + threadIdx.x] = v;
  }
  if (!last_block_arrives(ticket, gridDim.x)) return;
  __shared__ double fin[HS_TPB];
  constexpr int NVO = K * HS_PS, S = HS_TPB / NVO;
  const int idx = threadIdx.x % NVO, sl = threadIdx.x / NVO;
  const bool is_max = (idx % HS_PS) == 9;
  double acc = 0.0;
  if (sl < S)
    for (unsigned int b = sl; b < gridDim.x; b += S) {
      const double v = __ldcg(partials + static_cast<int64_t>(b) * NVO + idx);
      acc = is_max ? fmax(acc, v) : acc + v;
    }
  fin[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < NVO) {
    double t = 0.0;
    for (int q = 0; q < S; ++q) { const double v = fin[q * NVO + threadIdx.x]; t = is_max ? fmax(t, v) : t + v; }
    out[threadIdx.x] = t;
  }
}

}  // namespace hsk

// ======================================================== launchers ==================================================
using namespace hsk;

static int pick_blocks(const hs_ctx* ctx, int64_t work_items, int64_t min_items_per_block, int per_sm) {
  int64_t nb = static_cast<int64_t>(ctx->sm_count) * per_sm;
  const int64_t cap = (work_items + min_items_per_block - 1) / min_items_per_block;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return static_cast<int>(nb);
}

int32_t launch_rooms_cuboid_sums(hs_ctx* ctx, const float* xyz, int64_t n, const RoomTable& tbl, double* d_rec_out) {
  const int64_t G = (n + 3) >> 2;
  const int per_sm = ctx->modes[HS_MODE_BLOCKS_PER_SM] > 0 ? ctx->modes[HS_MODE_BLOCKS_PER_SM] : 2;
  const int nb = pick_blocks(ctx, G, 2 * HS_TPB, per_sm);
  const int64_t gpb = (G + nb - 1) / nb > 0 ? (G + nb - 1) / nb : 1;
  const size_t need = static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double) + static_cast<size_t>(nb) * sizeof(int) + 64;
  if (int32_t rc = hs_ensure_scratch(ctx, need)) return rc;
  double* partials = reinterpret_cast<double*>(ctx->d_scratch);
  int* meta = reinterpret_cast<int*>(ctx->d_scratch + static_cast<size_t>(nb) * tbl.nrooms * HS_NACC * sizeof(double));
  k_rooms_cuboid_sums<AccExact><<<nb, HS_TPB, 0, ctx->stream>>>(xyz, n, tbl, gpb, partials, meta, ctx->d_ticket, d_rec_out);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

int32_t launch_plane_assign(hs_ctx* ctx, const float* xyz, int64_t n, const PlaneTable& tbl, uint8_t* d_assign, float* d_resid) {
  const int nb = pick_blocks(ctx, (n + 3) >> 2, HS_TPB, 8);
  if (tbl.K == 6 && tbl.paired) k_plane_assign<6, true><<<nb, HS_TPB, 0, ctx->stream>>>(xyz, n, tbl, d_assign, d_resid);  // a cuboid room's walls
  else if (tbl.K == 6) k_plane_assign<6, false><<<nb, HS_TPB, 0, ctx->stream>>>(xyz, n, tbl, d_assign, d_resid);
  else k_plane_assign<0, false><<<nb, HS_TPB, 0, ctx->stream>>>(xyz, n, tbl, d_assign, d_resid);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

template <int K>
static int32_t launch_plane_sums_k(hs_ctx* ctx, const float* xyz, int64_t i0, int64_t i1, const PlaneTable& tbl, double* d_out) {
  const int nb = pick_blocks(ctx, i1 - i0, 4 * HS_TPB, 2);
  if (int32_t rc = hs_ensure_scratch(ctx, ctx->ps_scratch_off + static_cast<size_t>(nb) * K * HS_PS * sizeof(double))) return rc;
  unsigned int* ticket = ctx->d_ticket + ctx->ps_ticket_off;  // lane of hs_plane_sums (0: the shared ticket)
  if constexpr (K <= 6) {
    if (ctx->modes[HS_MODE_PS_KERNEL] == 0 && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0) {  // ring form
      constexpr int NW = HS_TPB / 32;
      const int dsm = NW * PSR_D * static_cast<int>(PSR_TILE_BYTES) + NW * PSR_D * 8 + NW * K * 9 * 8 + NW * 64 * 4 + NW * 8 * 4;
      const int64_t ngroups = (i1 >> 2) - ((i0 + 3) >> 2);
      const int64_t gpb = ngroups > 0 ? (ngroups + nb - 1) / nb : 1;
      double* part = reinterpret_cast<double*>(ctx->d_scratch + ctx->ps_scratch_off);
      if (K == 6 && tbl.paired) {
        HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_plane_sums_ring<K, (K == 6)>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm));
        k_plane_sums_ring<K, (K == 6)><<<nb, HS_TPB, dsm, ctx->stream>>>(xyz, i0, i1, tbl, gpb, part, ticket, d_out);
      } else {
        HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_plane_sums_ring<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm));
        k_plane_sums_ring<K, false><<<nb, HS_TPB, dsm, ctx->stream>>>(xyz, i0, i1, tbl, gpb, part, ticket, d_out);
      }
      ctx->launches++;
      HS_CUDA_TRY(ctx, cudaGetLastError());
      return HS_OK;
    }
    if (ctx->modes[HS_MODE_PS_KERNEL] != 1 && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0) {
      const int dsm = K * 9 * HS_TPB * static_cast<int>(sizeof(double));
      double* part = reinterpret_cast<double*>(ctx->d_scratch + ctx->ps_scratch_off);
      if (K == 6 && tbl.paired) {
        HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_plane_sums_f32<K, (K == 6)>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm));
        k_plane_sums_f32<K, (K == 6)><<<nb, HS_TPB, dsm, ctx->stream>>>(xyz, i0, i1, tbl, part, ticket, d_out);
      } else {
        HS_CUDA_TRY(ctx, cudaFuncSetAttribute(k_plane_sums_f32<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsm));
        k_plane_sums_f32<K, false><<<nb, HS_TPB, dsm, ctx->stream>>>(xyz, i0, i1, tbl, part, ticket, d_out);
      }
      ctx->launches++;
      HS_CUDA_TRY(ctx, cudaGetLastError());
      return HS_OK;
    }
  }
  k_plane_sums<K><<<nb, HS_TPB, 0, ctx->stream>>>(xyz, i0, i1, tbl, reinterpret_cast<double*>(ctx->d_scratch + ctx->ps_scratch_off), ticket, d_out);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
int32_t launch_plane_sums(hs_ctx* ctx, const float* xyz, int64_t i0, int64_t i1, const PlaneTable& tbl, double* d_out) {
  switch (tbl.K) {
    case 1: return launch_plane_sums_k<1>(ctx, xyz, i0, i1, tbl, d_out);
    case 2: return launch_plane_sums_k<2>(ctx, xyz, i0, i1, tbl, d_out);
    case 3: return launch_plane_sums_k<3>(ctx, xyz, i0, i1, tbl, d_out);
    case 4: return launch_plane_sums_k<4>(ctx, xyz, i0, i1, tbl, d_out);
    case 5: return launch_plane_sums_k<5>(ctx, xyz, i0, i1, tbl, d_out);
    case 6: return launch_plane_sums_k<6>(ctx, xyz, i0, i1, tbl, d_out);
    case 7: return launch_plane_sums_k<7>(ctx, xyz, i0, i1, tbl, d_out);
    case 8: return launch_plane_sums_k<8>(ctx, xyz, i0, i1, tbl, d_out);
    default: ctx->err = "hs_plane_sums: K must be in 1..8"; return HS_EINVAL;
  }
}
