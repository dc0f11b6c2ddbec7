// PCD records -> cloud (loadPCDFileXyzFloat / loadPCDFileXyzRgbNormalFloat, Main.hs:1318-1329).
// The file's DATA section is copied to the device as it lies on disk; this kernel picks x, y, z (4-byte floats) and the packed
// rgb out of every record and writes the AoS `Vector Vec3` cloud plus, when the file has colours, the Float colour cloud
// `Vec3 (r / 255) (g / 255) (b / 255)` (rgbToFloats, Main.hs:1327).  Field f of point i sits at base + off[f] + i * stride[f]:
// DATA binary has one stride (the point step) and per-field offsets; DATA binary_compressed (after LZF) and the host-parsed
// ascii records are the same formula with other numbers.  HBM-bound: point_step + 12 (+ 12) bytes per point.
#include "k_common.cuh"

namespace hsk {

struct PcdLayout {
  int64_t off[6];     // x, y, z, then either the packed rgb word (rgb_bytes == 0) or the red, green, blue bytes (PLY)
  int64_t stride[6];
  int aligned;        // every 4-byte address is a multiple of 4
  int rgb_bytes;      // 1: colours are three separate uint8 properties
};

__device__ __forceinline__ uint32_t load_u32(const uint8_t* p, bool aligned) {
  if (aligned) return __ldg(reinterpret_cast<const uint32_t*>(p));
  return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) | (static_cast<uint32_t>(p[3]) << 24);
}

__global__ void __launch_bounds__(HS_TPB)
k_pcd_unpack(const uint8_t* __restrict__ raw, int64_t n, const PcdLayout L, float* __restrict__ xyz, float* __restrict__ rgbf) {
  __shared__ float sx[HS_TPB * 3], sc[HS_TPB * 3];
  const bool al = L.aligned != 0;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * HS_TPB; base < n; base += static_cast<int64_t>(gridDim.x) * HS_TPB) {
    const int64_t i = base + threadIdx.x;
    if (i < n) {
      sx[3 * threadIdx.x + 0] = __uint_as_float(load_u32(raw + L.off[0] + i * L.stride[0], al));
      sx[3 * threadIdx.x + 1] = __uint_as_float(load_u32(raw + L.off[1] + i * L.stride[1], al));
      sx[3 * threadIdx.x + 2] = __uint_as_float(load_u32(raw + L.off[2] + i * L.stride[2], al));
      if (rgbf) {
        uint32_t r, g, b;
        if (L.rgb_bytes) {
          r = raw[L.off[3] + i * L.stride[3]]; g = raw[L.off[4] + i * L.stride[4]]; b = raw[L.off[5] + i * L.stride[5]];
        } else {
          const uint32_t c = load_u32(raw + L.off[3] + i * L.stride[3], al);  // 0x00RRGGBB
          r = (c >> 16) & 255u; g = (c >> 8) & 255u; b = c & 255u;
        }
        sc[3 * threadIdx.x + 0] = __fdiv_rn(static_cast<float>(r), 255.0f);
        sc[3 * threadIdx.x + 1] = __fdiv_rn(static_cast<float>(g), 255.0f);
        sc[3 * threadIdx.x + 2] = __fdiv_rn(static_cast<float>(b), 255.0f);
      }
    }
    __syncthreads();
    const int64_t cnt = min(static_cast<int64_t>(HS_TPB), n - base) * 3;  // the tile's floats leave as one contiguous run
    for (int q = threadIdx.x; q < cnt; q += HS_TPB) {
      xyz[3 * base + q] = sx[q];
      if (rgbf) rgbf[3 * base + q] = sc[q];
    }
    __syncthreads();
  }
}

}  // namespace hsk

using namespace hsk;

int32_t launch_pcd_unpack(hs_ctx* ctx, const uint8_t* d_raw, int64_t n, const int64_t off[6], const int64_t stride[6], int rgb_bytes, float* d_xyz, float* d_rgbf) {
  if (n == 0) return HS_OK;
  PcdLayout L;
  bool al = (reinterpret_cast<uintptr_t>(d_raw) & 3) == 0;
  for (int c = 0; c < 6; ++c) {
    L.off[c] = off[c]; L.stride[c] = stride[c];
    if (c < 3 || (c == 3 && d_rgbf && !rgb_bytes)) al = al && (off[c] % 4 == 0) && (stride[c] % 4 == 0);
  }
  L.aligned = al ? 1 : 0;
  L.rgb_bytes = rgb_bytes;
  int64_t nb = (n + HS_TPB - 1) / HS_TPB;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
  if (nb > cap) nb = cap;
  k_pcd_unpack<<<static_cast<int>(nb), HS_TPB, 0, ctx->stream>>>(d_raw, n, L, d_xyz, d_rgbf);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}
