// C ABI of libhousescan_b200.so (include/housescan_b200.h): context/cloud management and the glue from the
// reference-shaped entry points to the CUDA launchers.  No CPU fallback: every compute entry needs a live ctx.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <vector>

#include <fcntl.h>
#include <limits>
#include <unistd.h>

#include "../host/hs_host.hpp"
#include "hs_internal.cuh"
#include "k_peer.cuh"

static thread_local std::string g_create_err;

// temporary device buffers of one call: freed on every path out of the scope (cudaFree waits for work that still uses them)
struct DevTemps {
  std::vector<void*> ptrs;
  template <class T>
  cudaError_t alloc(T** out, size_t bytes) {
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = static_cast<T*>(p);
    return e;
  }
  ~DevTemps() { for (void* p : ptrs) cudaFree(p); }
  DevTemps() = default;
  DevTemps(const DevTemps&) = delete;
  DevTemps& operator=(const DevTemps&) = delete;
};

// Every entry point that touches the device goes through here: one in-flight call per ctx; an open evaluation session owns the
// stream (anything enqueued behind its resident kernel would wait for hs_eval_session_end); and a peer exchange that timed out
// in an earlier asynchronous call is reported by the next call as HS_ENCCL (the kernels raise the mapped status word).
#define HS_LOCK(ctx)                                     \
  if (!(ctx)) return HS_EINVAL;                          \
  std::lock_guard<std::mutex> lock__((ctx)->mu);         \
  if (cudaSetDevice((ctx)->device) != cudaSuccess) { (ctx)->err = "cudaSetDevice failed"; return HS_ECUDA; } \
  if ((ctx)->session) { (ctx)->err = "an evaluation session is open on this context: hs_eval_session_end first"; return HS_EINVAL; } \
  if ((ctx)->h_status && *(volatile uint32_t*)(ctx)->h_status) { *(volatile uint32_t*)(ctx)->h_status = 0; (ctx)->err = "a peer did not deliver its records in an earlier exchange (timed out); the records of that call are NaN"; return HS_ENCCL; }

#define HS_FAIL(ctx, code, msg) \
  do { (ctx)->err = (msg); return (code); } while (0)

int32_t hs_ensure_scratch(hs_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return HS_OK;
  size_t want = std::max(bytes, static_cast<size_t>(1) << 22);
  if (ctx->d_scratch) { HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); HS_CUDA_TRY(ctx, cudaFree(ctx->d_scratch)); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0; }
  HS_CUDA_TRY(ctx, cudaMalloc(&ctx->d_scratch, want));
  ctx->scratch_bytes = want;
  return HS_OK;
}
int32_t hs_ensure_pinned(hs_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_bytes) return HS_OK;
  size_t want = std::max(bytes, static_cast<size_t>(1) << 20);
  if (ctx->h_pinned) { HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); HS_CUDA_TRY(ctx, cudaFreeHost(ctx->h_pinned)); ctx->h_pinned = nullptr; ctx->pinned_bytes = 0; }
  HS_CUDA_TRY(ctx, cudaMallocHost(&ctx->h_pinned, want));
  ctx->pinned_bytes = want;
  return HS_OK;
}

// Host buffers handed to us by the caller are pageable (Haskell Storable vectors, numpy arrays): stage large copies
// through the pinned buffer in chunks so the DMA engine runs at PCIe speed; small ones go directly.
static int32_t copy_h2d(hs_ctx* ctx, void* dst, const void* src, size_t bytes) {
  HS_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return HS_OK;
}
static int32_t copy_d2h_sync(hs_ctx* ctx, void* dst, const void* src, size_t bytes) {
  HS_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return HS_OK;
}

extern "C" {

const char* hs_version(void) { return "housescan_b200 0.1 (sm_100a)"; }

int32_t hs_ctx_create(int32_t device, hs_ctx** out) {
  if (!out) return HS_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                   " (housescan_b200 has no CPU fallback)";
    return HS_ECUDA;
  }
  if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return HS_EINVAL; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return HS_ECUDA; }
  if (prop.major != 10) {
    g_create_err = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ", this library carries sm_100a code only";
    return HS_ECUDA;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return HS_ECUDA; }
  hs_ctx* ctx = new hs_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  auto fail = [&](cudaError_t err) { g_create_err = cudaGetErrorString(err); delete ctx; return HS_ECUDA; };
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaMalloc(&ctx->d_ticket, 256)) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(ctx->d_ticket, 0, 256)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&ctx->d_small, sizeof(double) * (HS_MAX_ROOMS * HS_REC + 64))) != cudaSuccess) return fail(e);
  if ((e = cudaHostAlloc(&ctx->h_status, 64, cudaHostAllocMapped)) != cudaSuccess) return fail(e);
  std::memset(ctx->h_status, 0, 64);
  if ((e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->d_status), ctx->h_status, 0)) != cudaSuccess) return fail(e);
  if (hs_ensure_scratch(ctx, 1 << 22) != HS_OK || hs_ensure_pinned(ctx, 1 << 20) != HS_OK) { g_create_err = ctx->err; delete ctx; return HS_ECUDA; }
  // tuning knobs can also come from the environment (HS_MODE_<key>=<value>, keys as in hs_ctx_set_mode)
  for (int k = 0; k < 16; ++k) {
    const std::string name = "HS_MODE_" + std::to_string(k);
    if (const char* v = std::getenv(name.c_str())) ctx->modes[k] = std::atoi(v);
  }
  *out = ctx;
  return HS_OK;
}

int32_t hs_ctx_destroy(hs_ctx* ctx) {
  if (!ctx) return HS_OK;
  if (ctx->session) hs_eval_session_end(ctx->session);  // stops the resident kernel
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  hs_eval_state_free(ctx);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_ticket) cudaFree(ctx->d_ticket);
  if (ctx->d_small) cudaFree(ctx->d_small);
  for (int p = 0; p < HS_PEER_MAX; ++p) if (ctx->peer_mapped[p]) cudaIpcCloseMemHandle(ctx->peer_mapped[p]);
  if (ctx->d_mailbox) cudaFree(ctx->d_mailbox);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  for (auto& e : ctx->ps_ev) if (e) cudaEventDestroy(e);
  for (auto& st : ctx->ps_aux) if (st) cudaStreamDestroy(st);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return HS_OK;
}

const char* hs_last_error(const hs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int32_t hs_ctx_set_stream(hs_ctx* ctx, void* s) {
  HS_LOCK(ctx);
  ctx->stream = s ? static_cast<cudaStream_t>(s) : ctx->own_stream;
  return HS_OK;
}
int32_t hs_ctx_sync(hs_ctx* ctx) {
  HS_LOCK(ctx);
  HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return HS_OK;
}
int32_t hs_ctx_device(const hs_ctx* ctx) { return ctx ? ctx->device : -1; }
int32_t hs_ctx_sm_count(const hs_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int64_t hs_ctx_launch_count(const hs_ctx* ctx) { return ctx ? ctx->launches : 0; }
int32_t hs_ctx_set_mode(hs_ctx* ctx, int32_t key, int32_t value) {
  if (!ctx || key < 0 || key >= 16) return HS_EINVAL;
  ctx->modes[key] = value;
  return HS_OK;
}

// ---- clouds ----------------------------------------------------------------------------------------------------
static int32_t cloud_alloc_impl(hs_ctx* ctx, int64_t n, hs_cloud** out) {
  if (!out || n < 0) HS_FAIL(ctx, HS_EINVAL, "hs_cloud_alloc: bad arguments");
  hs_cloud* c = new hs_cloud();
  const size_t bytes = ((static_cast<size_t>(n) * 12 + 47) / 48) * 48 + 64;  // whole 4-point groups + slack
  cudaError_t e = cudaMalloc(&c->d, bytes);
  if (e != cudaSuccess) { delete c; ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? HS_ENOMEM : HS_ECUDA; }
  c->n = n; c->cap = static_cast<int64_t>((bytes - 64) / 12); c->owned = true;
  *out = c;
  return HS_OK;
}
int32_t hs_cloud_alloc(hs_ctx* ctx, int64_t n, hs_cloud** out) {
  HS_LOCK(ctx);
  return cloud_alloc_impl(ctx, n, out);
}
int32_t hs_cloud_upload(hs_ctx* ctx, const float* xyz, int64_t n, hs_cloud** out) {
  HS_LOCK(ctx);
  if (!xyz && n > 0) HS_FAIL(ctx, HS_EINVAL, "hs_cloud_upload: null points");
  if (int32_t rc = cloud_alloc_impl(ctx, n, out)) return rc;
  if (n > 0) {
    if (int32_t rc = copy_h2d(ctx, (*out)->d, xyz, static_cast<size_t>(n) * 12)) return rc;
    HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return HS_OK;
}
int32_t hs_cloud_wrap_device(hs_ctx* ctx, void* dptr, int64_t n, hs_cloud** out) {
  HS_LOCK(ctx);
  if (!out || n < 0 || (!dptr && n > 0)) HS_FAIL(ctx, HS_EINVAL, "hs_cloud_wrap_device: bad arguments");
  if (reinterpret_cast<uintptr_t>(dptr) & 15) HS_FAIL(ctx, HS_EINVAL, "hs_cloud_wrap_device: pointer must be 16-byte aligned");
  hs_cloud* c = new hs_cloud();
  c->d = static_cast<float*>(dptr); c->n = n; c->cap = n; c->owned = false;
  *out = c;
  return HS_OK;
}
int32_t hs_cloud_write(hs_ctx* ctx, hs_cloud* cloud, const float* xyz, int64_t n) {
  HS_LOCK(ctx);
  if (!cloud || n < 0 || n > cloud->cap) HS_FAIL(ctx, HS_EINVAL, "hs_cloud_write: size exceeds capacity");
  cloud->n = n;
  if (n > 0) return copy_h2d(ctx, cloud->d, xyz, static_cast<size_t>(n) * 12);
  return HS_OK;
}
int32_t hs_cloud_download(hs_ctx* ctx, const hs_cloud* cloud, float* xyz_out) {
  HS_LOCK(ctx);
  if (!cloud || (!xyz_out && cloud->n > 0)) HS_FAIL(ctx, HS_EINVAL, "hs_cloud_download: bad arguments");
  if (cloud->n == 0) return HS_OK;
  return copy_d2h_sync(ctx, xyz_out, cloud->d, static_cast<size_t>(cloud->n) * 12);
}
int64_t hs_cloud_size(const hs_cloud* c) { return c ? c->n : -1; }
void* hs_cloud_device_ptr(const hs_cloud* c) { return c ? c->d : nullptr; }
int32_t hs_cloud_free(hs_ctx* ctx, hs_cloud* cloud) {
  HS_LOCK(ctx);
  if (!cloud) return HS_OK;
  if (cloud->owned && cloud->d) { HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); HS_CUDA_TRY(ctx, cudaFree(cloud->d)); }
  delete cloud;
  return HS_OK;
}

// ---- (1) depth frames ----------------------------------------------------------------------------------------------
int32_t hs_backproject_ref_dev(hs_ctx* ctx, const void* d_depth, int32_t w, int32_t h, hs_cloud* cloud_out, void* d_mask, int64_t* n_valid) {
  HS_LOCK(ctx);
  if (w <= 0 || h <= 0 || w > (1 << 24) || h > (1 << 24) || !d_depth) HS_FAIL(ctx, HS_EINVAL, "hs_backproject_ref: bad frame");  // pixel coordinates must be exact in Float
  const int64_t npx = static_cast<int64_t>(w) * h;
  if (cloud_out && cloud_out->cap < npx) HS_FAIL(ctx, HS_EINVAL, "hs_backproject_ref: output cloud smaller than w*h");
  int64_t* d_n = reinterpret_cast<int64_t*>(ctx->d_small);
  if (int32_t rc = launch_backproject(ctx, static_cast<const uint16_t*>(d_depth), w, h, cloud_out ? cloud_out->d : nullptr,
                                      static_cast<uint8_t*>(d_mask), d_n)) return rc;
  int64_t nv = 0;
  if (int32_t rc = copy_d2h_sync(ctx, &nv, d_n, sizeof nv)) return rc;
  if (cloud_out) cloud_out->n = nv;
  if (n_valid) *n_valid = nv;
  return HS_OK;
}

int32_t hs_backproject_ref(hs_ctx* ctx, const uint16_t* depth, int32_t w, int32_t h, float* xyz_out, uint8_t* mask_out, int64_t* n_valid) {
  if (!ctx) return HS_EINVAL;
  if (w <= 0 || h <= 0 || !depth) { ctx->err = "hs_backproject_ref: bad frame"; return HS_EINVAL; }
  const int64_t npx = static_cast<int64_t>(w) * h;
  hs_cloud* cl = nullptr;
  uint16_t* d_depth = nullptr;
  uint8_t* d_mask = nullptr;
  int32_t rc = HS_OK;
  DevTemps tmp;
  {
    HS_LOCK(ctx);
    HS_CUDA_TRY(ctx, tmp.alloc(&d_depth, npx * 2 + 16));
    if (mask_out) HS_CUDA_TRY(ctx, tmp.alloc(&d_mask, npx + 16));
    rc = copy_h2d(ctx, d_depth, depth, npx * 2);
  }
  if (rc == HS_OK && xyz_out) rc = hs_cloud_alloc(ctx, npx, &cl);
  int64_t nv = 0;
  if (rc == HS_OK) rc = hs_backproject_ref_dev(ctx, d_depth, w, h, cl, d_mask, &nv);
  if (rc == HS_OK && xyz_out && nv > 0) { HS_LOCK(ctx); rc = copy_d2h_sync(ctx, xyz_out, cl->d, nv * 12); }
  if (rc == HS_OK && mask_out) { HS_LOCK(ctx); rc = copy_d2h_sync(ctx, mask_out, d_mask, npx); }
  if (n_valid) *n_valid = nv;
  if (cl) hs_cloud_free(ctx, cl);
  { HS_LOCK(ctx); cudaStreamSynchronize(ctx->stream); }
  return rc;
}

static int32_t fill_plane_table(hs_ctx* ctx, const float* planes, int32_t K, int32_t maxK, PlaneTable& t, const char* who) {
  if (!planes || K < 1 || K > maxK) { ctx->err = std::string(who) + ": K out of range"; return HS_EINVAL; }
  t.K = K;
  std::memset(t.pl, 0, sizeof t.pl);
  std::memcpy(t.pl, planes, sizeof(float) * 4 * K);
  plane_table_mark_pairs(t);
  return HS_OK;
}

int32_t hs_backproject_reduce6x6_dev(hs_ctx* ctx, const void* d_frames, int64_t nframes, int32_t w, int32_t h, const float* intr,
                                     const float* poses, const float* planes, int32_t K, void* d_out) {
  HS_LOCK(ctx);
  if (!d_frames || nframes < 0 || w <= 0 || h <= 0 || !d_out) HS_FAIL(ctx, HS_EINVAL, "hs_backproject_reduce6x6: bad arguments");
  PlaneTable t;
  if (int32_t rc = fill_plane_table(ctx, planes, K, 16, t, "hs_backproject_reduce6x6")) return rc;
  if (w > (1 << 24) || h > (1 << 24)) HS_FAIL(ctx, HS_EINVAL, "hs_backproject_reduce6x6: frame sides are limited to 2^24");
  float* d_poses = nullptr;
  const size_t pose_bytes = poses ? static_cast<size_t>(nframes) * 64 : 0;  // multiple of 64: the work area stays aligned
  if (int32_t rc = hs_ensure_scratch(ctx, pose_bytes + reduce6x6_work_bytes(ctx, nframes, w, h))) return rc;
  if (poses) {
    d_poses = reinterpret_cast<float*>(ctx->d_scratch);
    if (int32_t rc = copy_h2d(ctx, d_poses, poses, pose_bytes)) return rc;
    HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // caller's pose buffer is not retained
  }
  if (nframes == 0) return HS_OK;
  return launch_reduce6x6(ctx, static_cast<const uint16_t*>(d_frames), nframes, w, h, intr, d_poses, t, static_cast<double*>(d_out),
                          ctx->d_scratch + pose_bytes);
}

int32_t hs_backproject_reduce6x6(hs_ctx* ctx, const uint16_t* frames, int64_t nframes, int32_t w, int32_t h, const float* intr,
                                 const float* poses, const float* planes, int32_t K, double* out) {
  if (!ctx) return HS_EINVAL;
  if (!frames || nframes < 0 || w <= 0 || h <= 0 || !out) { ctx->err = "hs_backproject_reduce6x6: bad arguments"; return HS_EINVAL; }
  if (nframes == 0) return HS_OK;
  const size_t fbytes = static_cast<size_t>(nframes) * w * h * 2;
  uint16_t* d_frames = nullptr;
  double* d_out = nullptr;
  DevTemps tmp;
  {
    HS_LOCK(ctx);
    HS_CUDA_TRY(ctx, tmp.alloc(&d_frames, fbytes + 16));
    HS_CUDA_TRY(ctx, tmp.alloc(&d_out, static_cast<size_t>(nframes) * HS_NE * sizeof(double)));
    if (int32_t rc = copy_h2d(ctx, d_frames, frames, fbytes)) return rc;
  }
  int32_t rc = hs_backproject_reduce6x6_dev(ctx, d_frames, nframes, w, h, intr, poses, planes, K, d_out);
  { HS_LOCK(ctx);
    if (rc == HS_OK) rc = copy_d2h_sync(ctx, out, d_out, static_cast<size_t>(nframes) * HS_NE * sizeof(double));
    cudaStreamSynchronize(ctx->stream); }
  return rc;
}

// ---- (2) planes --------------------------------------------------------------------------------------------------------
int32_t hs_plane_assign_dev(hs_ctx* ctx, const hs_cloud* cloud, const float* planes, int32_t K, void* d_assign, void* d_resid) {
  HS_LOCK(ctx);
  if (!cloud) HS_FAIL(ctx, HS_EINVAL, "hs_plane_assign: null cloud");
  PlaneTable t;
  if (int32_t rc = fill_plane_table(ctx, planes, K, 16, t, "hs_plane_assign")) return rc;
  if (cloud->n == 0) return HS_OK;
  return launch_plane_assign(ctx, cloud->d, cloud->n, t, static_cast<uint8_t*>(d_assign), static_cast<float*>(d_resid));
}

int32_t hs_plane_assign(hs_ctx* ctx, const hs_cloud* cloud, const float* planes, int32_t K, uint8_t* assign_out, float* resid_out) {
  if (!ctx) return HS_EINVAL;
  if (!cloud) { ctx->err = "hs_plane_assign: null cloud"; return HS_EINVAL; }
  const int64_t n = cloud->n;
  uint8_t* d_a = nullptr;
  float* d_r = nullptr;
  DevTemps tmp;
  {
    HS_LOCK(ctx);
    if (assign_out) HS_CUDA_TRY(ctx, tmp.alloc(&d_a, n + 16));
    if (resid_out) HS_CUDA_TRY(ctx, tmp.alloc(&d_r, n * 4 + 16));
  }
  int32_t rc = hs_plane_assign_dev(ctx, cloud, planes, K, d_a, d_r);
  { HS_LOCK(ctx);
    if (rc == HS_OK && assign_out && n) rc = copy_d2h_sync(ctx, assign_out, d_a, n);
    if (rc == HS_OK && resid_out && n) rc = copy_d2h_sync(ctx, resid_out, d_r, n * 4);
    cudaStreamSynchronize(ctx->stream); }
  return rc;
}

int32_t hs_planes_from_cuboid(const double params[10], float planes_out[24]) {
  if (!params || !planes_out) return HS_EINVAL;
  hs::planes_from_cuboid(params, planes_out);
  return HS_OK;
}

static int32_t rooms_sums_enqueue(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms, const double* params, double* d_out, bool exchange = false) {
  if (!cloud || !room_offsets || !params || nrooms < 1) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums: bad arguments");
  for (int r = 0; r < nrooms; ++r)
    if (room_offsets[r] > room_offsets[r + 1]) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums: room offsets must be non-decreasing");
  if (room_offsets[0] < 0 || room_offsets[nrooms] > cloud->n) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums: room offsets outside the cloud");
  for (int r0 = 0; r0 < nrooms; r0 += HS_MAX_ROOMS) {
    RoomTable t;
    std::memset(&t, 0, sizeof t);
    t.nrooms = std::min(HS_MAX_ROOMS, nrooms - r0);
    t.paired = 1;
    for (int r = 0; r < t.nrooms; ++r) {
      t.off[r] = room_offsets[r0 + r];
      float pl[24];
      hs::planes_from_cuboid(params + 10 * (r0 + r), pl);
      std::memcpy(t.pl[r], pl, sizeof pl);
      for (int j = 0; j < 3; ++j)
        for (int c = 0; c < 3; ++c)
          if (pl[8 * j + c] != -pl[8 * j + 4 + c]) t.paired = 0;
    }
    t.off[t.nrooms] = room_offsets[r0 + t.nrooms];
    const int mode = ctx->modes[HS_MODE_EVAL_KERNEL];
    if (mode == HS_EVAL_FAST && !t.paired) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums: fast kernel needs antiparallel plane pairs");
    const bool want_exchange = exchange && ctx->px.world > 1;
    double* d_chunk = d_out + static_cast<size_t>(r0) * HS_REC;
    if (t.paired && mode != HS_EVAL_EXACT) {  // throughput kernel (k_eval.cuh): the exchange happens inside the same launch
      if (int32_t rc = launch_eval(ctx, cloud->d, cloud->n, t, d_chunk, want_exchange)) return rc;
    } else {  // exact-products Double kernel (k_planes.cu) + one more single-warp launch for the exchange
      if (int32_t rc = launch_rooms_cuboid_sums(ctx, cloud->d, cloud->n, t, d_chunk)) return rc;
      if (want_exchange)
        if (int32_t rc = launch_peer_allreduce(ctx, d_chunk, t.nrooms * HS_REC)) return rc;
    }
  }
  return HS_OK;
}

int32_t hs_rooms_cuboid_sums_async(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms, const double* params, void* d_rec_out) {
  HS_LOCK(ctx);
  if (!d_rec_out) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums_async: null output");
  return rooms_sums_enqueue(ctx, cloud, room_offsets, nrooms, params, static_cast<double*>(d_rec_out));
}

int32_t hs_rooms_cuboid_sums_allreduce_async(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms, const double* params, void* d_rec_out) {
  HS_LOCK(ctx);
  if (!d_rec_out) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums_allreduce_async: null output");
  if (ctx->px.world < 1) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums_allreduce_async: no peer group (hs_peer_mailbox_create / hs_peer_mailbox_connect first)");
  return rooms_sums_enqueue(ctx, cloud, room_offsets, nrooms, params, static_cast<double*>(d_rec_out), true);
}

// ---- peer group: CUDA IPC mailboxes between the per-GPU processes of one node ------------------------------------------
int32_t hs_peer_mailbox_create(hs_ctx* ctx, int32_t rank, int32_t world, uint8_t handle_out[64]) {
  HS_LOCK(ctx);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!handle_out || world < 1 || world > HS_PEER_MAX || rank < 0 || rank >= world) HS_FAIL(ctx, HS_EINVAL, "hs_peer_mailbox_create: need 0 <= rank < world <= 8");
  if (ctx->d_mailbox) HS_FAIL(ctx, HS_EINVAL, "hs_peer_mailbox_create: this context already has a mailbox");
  const size_t bytes = hsk::PEER_MAILBOX_BYTES;
  HS_CUDA_TRY(ctx, cudaMalloc(&ctx->d_mailbox, bytes));
  HS_CUDA_TRY(ctx, cudaMemset(ctx->d_mailbox, 0, bytes));
  HS_CUDA_TRY(ctx, cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  HS_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->d_mailbox));
  std::memcpy(handle_out, &h, 64);
  ctx->px = PeerExchange{};
  ctx->px.rank = rank;
  ctx->px.world = 0;  // not usable until connected
  ctx->px.pad = static_cast<uint32_t>(world);
  return HS_OK;
}

int32_t hs_peer_mailbox_connect(hs_ctx* ctx, const uint8_t* handles) {
  HS_LOCK(ctx);
  if (!handles || !ctx->d_mailbox) HS_FAIL(ctx, HS_EINVAL, "hs_peer_mailbox_connect: create the local mailbox first");
  const int world = static_cast<int>(ctx->px.pad), rank = ctx->px.rank;
  for (int p = 0; p < world; ++p) {
    if (p == rank) { ctx->px.mailbox[p] = reinterpret_cast<unsigned long long>(ctx->d_mailbox); continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * p, 64);
    void* ptr = nullptr;
    HS_CUDA_TRY(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_mapped[p] = ptr;
    ctx->px.mailbox[p] = reinterpret_cast<unsigned long long>(ptr);
  }
  ctx->px.world = world;
  ctx->px.epoch = 0;
  ctx->px.pad = 0;
  return HS_OK;
}

// Same-process peer group (one Haskell executable driving several GPUs, housescan.cabal:15-43): ctxs[i] becomes rank i of an
// n-rank group.  CUDA IPC cannot open a handle in the process that created it, so the mailboxes are addressed directly after
// cudaDeviceEnablePeerAccess.  The asynchronous entry points may then be issued for all ranks from one host thread.
int32_t hs_peer_group_create_local(hs_ctx* const* ctxs, int32_t n) {
  if (!ctxs || n < 1 || n > HS_PEER_MAX) return HS_EINVAL;
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return HS_EINVAL;
    for (int j = 0; j < i; ++j)
      if (ctxs[j] == ctxs[i] || ctxs[j]->device == ctxs[i]->device) { ctxs[i]->err = "hs_peer_group_create_local: one context per device"; return HS_EINVAL; }
    if (ctxs[i]->d_mailbox || ctxs[i]->session) { ctxs[i]->err = "hs_peer_group_create_local: context already has a mailbox or an open session"; return HS_EINVAL; }
  }
  for (int i = 0; i < n; ++i) {
    hs_ctx* ctx = ctxs[i];
    std::lock_guard<std::mutex> lock(ctx->mu);
    HS_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      int can = 0;
      HS_CUDA_TRY(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[j]->device));
      if (!can) { ctx->err = "hs_peer_group_create_local: device " + std::to_string(ctx->device) + " cannot access device " + std::to_string(ctxs[j]->device); return HS_ECUDA; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return HS_ECUDA; }
      (void)cudaGetLastError();
    }
    HS_CUDA_TRY(ctx, cudaMalloc(&ctx->d_mailbox, hsk::PEER_MAILBOX_BYTES));
    HS_CUDA_TRY(ctx, cudaMemset(ctx->d_mailbox, 0, hsk::PEER_MAILBOX_BYTES));
    HS_CUDA_TRY(ctx, cudaDeviceSynchronize());
  }
  for (int i = 0; i < n; ++i) {
    hs_ctx* ctx = ctxs[i];
    ctx->px = PeerExchange{};
    for (int j = 0; j < n; ++j) ctx->px.mailbox[j] = reinterpret_cast<unsigned long long>(ctxs[j]->d_mailbox);
    ctx->px.rank = i;
    ctx->px.world = n;
    ctx->peer_local = true;
  }
  return HS_OK;
}

int32_t hs_rooms_cuboid_sums(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms, const double* params, double* rec_out) {
  HS_LOCK(ctx);
  if (!rec_out || nrooms < 1) HS_FAIL(ctx, HS_EINVAL, "hs_rooms_cuboid_sums: bad arguments");
  double* d_out = ctx->d_small;
  DevTemps tmp;  // more rooms than the ctx's resident record buffer holds (HS_MAX_ROOMS): a buffer for this call
  if (nrooms > HS_MAX_ROOMS) HS_CUDA_TRY(ctx, tmp.alloc(&d_out, sizeof(double) * HS_REC * nrooms));
  int32_t rc = rooms_sums_enqueue(ctx, cloud, room_offsets, nrooms, params, d_out);
  if (rc == HS_OK) rc = copy_d2h_sync(ctx, rec_out, d_out, sizeof(double) * HS_REC * nrooms);
  return rc;
}

int32_t hs_cuboid_grad_from_sums(const double params[10], const double rec[HS_REC], double* f, double grad[10], int64_t counts[6]) {
  if (!params || !rec) return HS_EINVAL;
  hs::cuboid_grad_from_sums(params, rec, f, grad, counts);
  return HS_OK;
}

int32_t hs_cuboid_residual_grad(hs_ctx* ctx, const hs_cloud* cloud, const double params[10], double* f, double grad[10], int64_t counts[6]) {
  if (!ctx) return HS_EINVAL;
  if (!cloud || !params) { ctx->err = "hs_cuboid_residual_grad: bad arguments"; return HS_EINVAL; }
  double rec[HS_REC];
  const int64_t off[2] = {0, cloud->n};
  if (int32_t rc = hs_rooms_cuboid_sums(ctx, cloud, off, 1, params, rec)) return rc;
  hs::cuboid_grad_from_sums(params, rec, f, grad, counts);
  return HS_OK;
}

int32_t hs_plane_sums(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms, const float* planes, int32_t K, double* out) {
  HS_LOCK(ctx);
  if (!cloud || !room_offsets || !planes || !out || nrooms < 1) HS_FAIL(ctx, HS_EINVAL, "hs_plane_sums: bad arguments");
  if (room_offsets[0] < 0 || room_offsets[nrooms] > cloud->n) HS_FAIL(ctx, HS_EINVAL, "hs_plane_sums: room offsets outside the cloud");
  if (K < 1 || K > HS_MAX_PLANES) HS_FAIL(ctx, HS_EINVAL, "hs_plane_sums: need 1 <= K <= 8 planes per room");
  // one launch per room, all enqueued back to back; the K x HS_PS records land side by side and come back in ONE copy
  const size_t total = static_cast<size_t>(nrooms) * K * HS_PS;
  double* d_out = ctx->d_small;
  DevTemps tmp;
  if (total > static_cast<size_t>(HS_MAX_ROOMS) * HS_REC) HS_CUDA_TRY(ctx, tmp.alloc(&d_out, total * sizeof(double)));
  int32_t rc = HS_OK;
  HS_CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, total * sizeof(double), ctx->stream));
  // one launch per room (the planes ride in the kernel's constant bank), but three lanes: the ctx stream and two helper streams,
  // each with its own partial-record region and ticket, rooms round-robin - the tail of a launch (last blocks, last-block sums)
  // overlaps the next room's start instead of idling the GPU
  const int lanes = nrooms >= 3 ? 3 : 1;
  const size_t region = ((static_cast<size_t>(ctx->sm_count) * 2 + 8) * K * HS_PS * sizeof(double) + 255) & ~static_cast<size_t>(255);
  cudaStream_t lane_stream[3] = {ctx->stream, nullptr, nullptr};
  if (lanes > 1) {
    if (!ctx->ps_aux[0]) {
      for (int i = 0; i < 2; ++i) HS_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->ps_aux[i], cudaStreamNonBlocking));
      for (int i = 0; i < 3; ++i) HS_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ps_ev[i], cudaEventDisableTiming));
    }
    if (int32_t rc2 = hs_ensure_scratch(ctx, 3 * region)) return rc2;  // before any lane is in flight: growing would free what a lane uses
    lane_stream[1] = ctx->ps_aux[0];
    lane_stream[2] = ctx->ps_aux[1];
    HS_CUDA_TRY(ctx, cudaEventRecord(ctx->ps_ev[0], ctx->stream));  // the helpers start behind the memset (and whatever precedes it)
    HS_CUDA_TRY(ctx, cudaStreamWaitEvent(lane_stream[1], ctx->ps_ev[0], 0));
    HS_CUDA_TRY(ctx, cudaStreamWaitEvent(lane_stream[2], ctx->ps_ev[0], 0));
  }
  const cudaStream_t main_stream = ctx->stream;
  for (int r = 0; r < nrooms && rc == HS_OK; ++r) {
    if (room_offsets[r] > room_offsets[r + 1]) { ctx->err = "hs_plane_sums: room offsets must be non-decreasing"; rc = HS_EINVAL; break; }
    PlaneTable t;
    if ((rc = fill_plane_table(ctx, planes + static_cast<size_t>(r) * K * 4, K, HS_MAX_PLANES, t, "hs_plane_sums")) != HS_OK) break;
    if (room_offsets[r] == room_offsets[r + 1]) continue;  // stays zero
    const int lane = lanes > 1 ? r % 3 : 0;
    ctx->stream = lane_stream[lane];
    ctx->ps_scratch_off = lanes > 1 ? lane * region : 0;
    ctx->ps_ticket_off = lanes > 1 ? 40 + lane : 0;
    rc = launch_plane_sums(ctx, cloud->d, room_offsets[r], room_offsets[r + 1], t, d_out + static_cast<size_t>(r) * K * HS_PS);
  }
  ctx->stream = main_stream;
  ctx->ps_scratch_off = 0;
  ctx->ps_ticket_off = 0;
  if (lanes > 1) {  // join: the ctx stream continues behind both helpers (also on an error path: nothing may still run when we return)
    for (int i = 1; i < 3; ++i) {
      cudaEventRecord(ctx->ps_ev[i], lane_stream[i]);
      cudaStreamWaitEvent(main_stream, ctx->ps_ev[i], 0);
    }
  }
  if (rc == HS_OK) rc = copy_d2h_sync(ctx, out, d_out, total * sizeof(double));
  return rc;
}

static int32_t mean_impl(hs_ctx* ctx, const hs_cloud* cloud, double mean[3]) {
  if (cloud->n == 0) HS_FAIL(ctx, HS_EINVAL, "pointMean: empty");  // Main.hs:1597
  if (int32_t rc = launch_mean(ctx, cloud->d, cloud->n, ctx->d_small)) return rc;
  double s[3];
  if (int32_t rc = copy_d2h_sync(ctx, s, ctx->d_small, sizeof s)) return rc;
  for (int i = 0; i < 3; ++i) mean[i] = s[i] / static_cast<double>(cloud->n);
  return HS_OK;
}

int32_t hs_scatter3x3(hs_ctx* ctx, const hs_cloud* cloud, double mean[3], double scatter[6]) {
  HS_LOCK(ctx);
  if (!cloud || !mean || !scatter) HS_FAIL(ctx, HS_EINVAL, "hs_scatter3x3: bad arguments");
  if (int32_t rc = mean_impl(ctx, cloud, mean)) return rc;
  const float m[3] = {static_cast<float>(mean[0]), static_cast<float>(mean[1]), static_cast<float>(mean[2])};
  if (int32_t rc = launch_scatter(ctx, cloud->d, cloud->n, m, ctx->d_small)) return rc;
  return copy_d2h_sync(ctx, scatter, ctx->d_small, 6 * sizeof(double));
}

int32_t hs_fit_plane(hs_ctx* ctx, const hs_cloud* cloud, float plane_out[4]) {
  if (!ctx) return HS_EINVAL;
  if (!cloud || !plane_out) { ctx->err = "hs_fit_plane: bad arguments"; return HS_EINVAL; }
  if (cloud->n < 3) { ctx->err = "fitPlane: " + std::to_string(cloud->n) + " points given, need at least 3"; return HS_EINVAL; }  // Main.hs:1438
  double mean[3], sc[6], ev[3], evec[3][3];
  if (int32_t rc = hs_scatter3x3(ctx, cloud, mean, sc)) return rc;
  hs::eig_sym3(sc, ev, evec);
  const hs::V3<float> nrm = hs::unit(hs::V3<float>{static_cast<float>(evec[0][0]), static_cast<float>(evec[1][0]), static_cast<float>(evec[2][0])});
  const hs::V3<float> m{static_cast<float>(mean[0]), static_cast<float>(mean[1]), static_cast<float>(mean[2])};
  plane_out[0] = nrm.x; plane_out[1] = nrm.y; plane_out[2] = nrm.z;
  plane_out[3] = hs::dot(nrm, m) - 0.0f;  // signedDistanceToPlaneEq (PlaneEq n 0) m
  return HS_OK;
}

// ---- (3) transforms --------------------------------------------------------------------------------------------------------
static int32_t check_out(hs_ctx* ctx, const hs_cloud* in, hs_cloud* out, const char* who) {
  if (!in || !out) { ctx->err = std::string(who) + ": null cloud"; return HS_EINVAL; }
  if (out->cap < in->n) { ctx->err = std::string(who) + ": output cloud too small"; return HS_EINVAL; }
  out->n = in->n;
  return HS_OK;
}
int32_t hs_transform(hs_ctx* ctx, const hs_cloud* in, const float m[16], hs_cloud* out) {
  HS_LOCK(ctx);
  if (!m) HS_FAIL(ctx, HS_EINVAL, "hs_transform: null matrix");
  if (int32_t rc = check_out(ctx, in, out, "hs_transform")) return rc;
  if (m[3] != 0.0f || m[7] != 0.0f || m[11] != 0.0f || m[15] != 1.0f)
    HS_FAIL(ctx, HS_EINVAL, "projectRoom: last column of the projection is not (0,0,0,1)");  // Main.hs:1725-1728
  const float R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
  const float off[3] = {m[12], m[13], m[14]};
  if (in->n == 0) return HS_OK;
  return launch_affine(ctx, in->d, out->d, in->n, R, nullptr, off, 1);
}
int32_t hs_rotate_around(hs_ctx* ctx, const hs_cloud* in, const float c[3], const float R[9], hs_cloud* out) {
  HS_LOCK(ctx);
  if (!c || !R) HS_FAIL(ctx, HS_EINVAL, "hs_rotate_around: null arguments");
  if (int32_t rc = check_out(ctx, in, out, "hs_rotate_around")) return rc;
  if (in->n == 0) return HS_OK;
  return launch_affine(ctx, in->d, out->d, in->n, R, c, c, 0);
}
int32_t hs_translate(hs_ctx* ctx, const hs_cloud* in, const float off[3], hs_cloud* out) {
  HS_LOCK(ctx);
  if (!off) HS_FAIL(ctx, HS_EINVAL, "hs_translate: null offset");
  if (int32_t rc = check_out(ctx, in, out, "hs_translate")) return rc;
  if (in->n == 0) return HS_OK;
  return launch_affine(ctx, in->d, out->d, in->n, nullptr, nullptr, off, 2);
}
int32_t hs_mean_extent(hs_ctx* ctx, const hs_cloud* cloud, double mean[3], float* maxdist) {
  HS_LOCK(ctx);
  if (!cloud || !mean) HS_FAIL(ctx, HS_EINVAL, "hs_mean_extent: bad arguments");
  if (int32_t rc = mean_impl(ctx, cloud, mean)) return rc;
  if (maxdist) {
    const float m[3] = {static_cast<float>(mean[0]), static_cast<float>(mean[1]), static_cast<float>(mean[2])};
    unsigned int* d_bits = reinterpret_cast<unsigned int*>(ctx->d_small);
    if (int32_t rc = launch_max_nsq(ctx, cloud->d, cloud->n, m, d_bits)) return rc;
    unsigned int bits = 0;
    if (int32_t rc = copy_d2h_sync(ctx, &bits, d_bits, sizeof bits)) return rc;
    float q;
    std::memcpy(&q, &bits, 4);
    *maxdist = std::sqrt(q);
  }
  return HS_OK;
}
int32_t hs_write_ply(hs_ctx* ctx, const hs_cloud* cloud, const uint8_t* rgb, const char* path) {
  if (!ctx) return HS_EINVAL;
  if (!cloud || !path) { ctx->err = "hs_write_ply: bad arguments"; return HS_EINVAL; }
  std::vector<float> host(static_cast<size_t>(cloud->n) * 3);
  if (int32_t rc = hs_cloud_download(ctx, cloud, host.data())) return rc;
  std::string err;
  if (!hs::write_ply(path, host.data(), rgb, cloud->n, &err)) { ctx->err = err; return HS_EIO; }
  return HS_OK;
}
// Sharded full-resolution export (Main.hs:1716-1730 transforms a room's cloud, README.md:16 step 4 exports
// it): every rank transforms its point range on its own GPU and writes its part of ONE file.  hs_write_ply_begin (one caller)
// creates the file with the header at its final size; after that any number of hs_write_ply_part calls - different ranks,
// processes, any order - fill in the body.  The device -> host copy runs in two pinned halves so that the copy of one chunk
// overlaps the write of the previous one.  The file is byte-identical to hs_write_ply's.
int32_t hs_write_ply_begin(const char* path, int64_t n_total, int32_t has_rgb) {
  if (!path || n_total < 0) return HS_EINVAL;
  std::string err;
  return hs::write_ply_begin(path, n_total, has_rgb != 0, &err) ? HS_OK : HS_EIO;
}
int32_t hs_write_ply_part_host(const char* path, const float* xyz, const uint8_t* rgb, int64_t first, int64_t n, int64_t n_total) {
  if (!path || (!xyz && n > 0) || first < 0 || n < 0 || first + n > n_total) return HS_EINVAL;
  const int fd = open(path, O_WRONLY);
  if (fd < 0) return HS_EIO;
  std::string err;
  const bool ok = hs::write_ply_part(fd, xyz, rgb, first, n, n_total, &err);
  return (close(fd) == 0 && ok) ? HS_OK : HS_EIO;
}
int32_t hs_write_ply_part(hs_ctx* ctx, const hs_cloud* cloud, const uint8_t* rgb, const char* path, int64_t first, int64_t n_total) {
  HS_LOCK(ctx);
  if (!cloud || !path || first < 0 || first + cloud->n > n_total) HS_FAIL(ctx, HS_EINVAL, "hs_write_ply_part: need 0 <= first and first + n <= n_total");
  const int fd = open(path, O_WRONLY);
  if (fd < 0) HS_FAIL(ctx, HS_EIO, std::string("hs_write_ply_part: cannot open ") + path + " (hs_write_ply_begin first)");
  const int64_t CH = 1 << 20;  // points per chunk: 12 MB
  int32_t rc = hs_ensure_pinned(ctx, static_cast<size_t>(2 * CH) * 12);
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (rc == HS_OK && (cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess)) { ctx->err = "cudaEventCreate failed"; rc = HS_ECUDA; }
  std::string err;
  float* half[2] = {reinterpret_cast<float*>(ctx->h_pinned), reinterpret_cast<float*>(ctx->h_pinned) + 3 * CH};
  const int64_t nchunks = (cloud->n + CH - 1) / CH;
  auto issue = [&](int64_t c) {
    const int64_t i0 = c * CH, m = std::min(CH, cloud->n - i0);
    if (cudaMemcpyAsync(half[c & 1], cloud->d + 3 * i0, static_cast<size_t>(m) * 12, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaEventRecord(ev[c & 1], ctx->stream) != cudaSuccess) { ctx->err = "hs_write_ply_part: device -> host copy failed"; return false; }
    return true;
  };
  if (rc == HS_OK && nchunks > 0 && !issue(0)) rc = HS_ECUDA;
  for (int64_t c = 0; c < nchunks && rc == HS_OK; ++c) {
    if (c + 1 < nchunks && !issue(c + 1)) { rc = HS_ECUDA; break; }
    if (cudaEventSynchronize(ev[c & 1]) != cudaSuccess) { ctx->err = "hs_write_ply_part: device -> host copy failed"; rc = HS_ECUDA; break; }
    const int64_t i0 = c * CH, m = std::min(CH, cloud->n - i0);
    if (!hs::write_ply_part(fd, half[c & 1], rgb ? rgb + 3 * i0 : nullptr, first + i0, m, n_total, &err)) { ctx->err = err; rc = HS_EIO; }
  }
  cudaStreamSynchronize(ctx->stream);
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  if (close(fd) != 0 && rc == HS_OK) { ctx->err = "hs_write_ply_part: close failed"; rc = HS_EIO; }
  return rc;
}

static int32_t put_string(const std::string& s, char* buf, int32_t buflen) {
  if (!buf || buflen <= static_cast<int32_t>(s.size())) return HS_EINVAL;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return HS_OK;
}
int32_t hs_proj_to_string(const float m[16], char* buf, int32_t buflen) { return m ? put_string(hs::proj_to_string(m), buf, buflen) : HS_EINVAL; }
int32_t hs_proj_to_xf(const float m[16], char* buf, int32_t buflen) { return m ? put_string(hs::proj_to_xf(m), buf, buflen) : HS_EINVAL; }

// ---- (4) connected components ------------------------------------------------------------------------------------------------
int32_t hs_cc_label_dev(hs_ctx* ctx, const void* d_src, const void* d_dst, int64_t E, uint32_t N, void* d_label) {
  HS_LOCK(ctx);
  if (E < 0 || (E > 0 && (!d_src || !d_dst)) || (N > 0 && !d_label)) HS_FAIL(ctx, HS_EINVAL, "hs_cc_label: bad arguments");
  return launch_cc(ctx, static_cast<const uint32_t*>(d_src), static_cast<const uint32_t*>(d_dst), E, N, static_cast<uint32_t*>(d_label));
}
int32_t hs_cc_label(hs_ctx* ctx, const uint32_t* src, const uint32_t* dst, int64_t E, uint32_t N, uint32_t* label_out) {
  if (!ctx) return HS_EINVAL;
  if (E < 0 || (E > 0 && (!src || !dst)) || (N > 0 && !label_out)) { ctx->err = "hs_cc_label: bad arguments"; return HS_EINVAL; }
  for (int64_t e = 0; e < E; ++e)
    if (src[e] >= N || dst[e] >= N) { ctx->err = "hs_cc_label: vertex id out of range (vertices must be contiguous, GroupConnectedComponents.hs:38)"; return HS_EINVAL; }
  if (N == 0) return HS_OK;
  uint32_t *d_s = nullptr, *d_d = nullptr, *d_l = nullptr;
  DevTemps tmp;
  {
    HS_LOCK(ctx);
    HS_CUDA_TRY(ctx, tmp.alloc(&d_l, static_cast<size_t>(N) * 4));
    if (E > 0) {
      HS_CUDA_TRY(ctx, tmp.alloc(&d_s, static_cast<size_t>(E) * 4));
      HS_CUDA_TRY(ctx, tmp.alloc(&d_d, static_cast<size_t>(E) * 4));
      if (int32_t rc = copy_h2d(ctx, d_s, src, static_cast<size_t>(E) * 4)) return rc;
      if (int32_t rc = copy_h2d(ctx, d_d, dst, static_cast<size_t>(E) * 4)) return rc;
    }
  }
  int32_t rc = hs_cc_label_dev(ctx, d_s, d_d, E, N, d_l);
  { HS_LOCK(ctx);
    if (rc == HS_OK) rc = copy_d2h_sync(ctx, label_out, d_l, static_cast<size_t>(N) * 4);
    cudaStreamSynchronize(ctx->stream); }
  return rc;
}
int32_t hs_group_cc(hs_ctx* ctx, const uint32_t* src, const uint32_t* dst, int64_t E, uint32_t N, int32_t* comp_out, int64_t* order_out, int32_t* ncomp_out) {
  if (!ctx) return HS_EINVAL;
  if (E == 0) { if (ncomp_out) *ncomp_out = 0; return HS_OK; }  // groupCCContiguous [] = []
  if (!comp_out || !order_out || !ncomp_out) { ctx->err = "hs_group_cc: null outputs"; return HS_EINVAL; }
  std::vector<uint32_t> label(N);
  if (int32_t rc = hs_cc_label(ctx, src, dst, E, N, label.data())) return rc;
  hs::group_edges_by_label(src, label.data(), E, comp_out, order_out, ncomp_out);
  return HS_OK;
}

// ---- VectorUtil ------------------------------------------------------------------------------------------------------------------
static int32_t kth_impl(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, int64_t k, bool largest, float* out) {
  if (!cloud || !out || axis < 0 || axis > 2) HS_FAIL(ctx, HS_EINVAL, "hs_kth: bad arguments");
  if (k < 1) HS_FAIL(ctx, HS_EINVAL, "kLargestBy: k must be >= 1 if the vector is not empty");   // VectorUtil.hs:13
  if (k > cloud->n) HS_FAIL(ctx, HS_EINVAL, "kLargestBy: k must bet be > length of the vector");  // VectorUtil.hs:14 (sic)
  if (cloud->n >= (static_cast<int64_t>(1) << 32)) HS_FAIL(ctx, HS_EINVAL, "hs_kth: clouds of 2^32 points or more are not supported (32-bit histogram bins); select per shard with hs_kth_shard_pass");
  float* d_out = reinterpret_cast<float*>(ctx->d_small);
  if (int32_t rc = launch_kth(ctx, cloud->d, cloud->n, axis, k, largest, d_out)) return rc;
  return copy_d2h_sync(ctx, out, d_out, sizeof(float));
}
int32_t hs_kth_largest(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, int64_t k, float* out) { HS_LOCK(ctx); return kth_impl(ctx, cloud, axis, k, true, out); }
int32_t hs_kth_smallest(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, int64_t k, float* out) { HS_LOCK(ctx); return kth_impl(ctx, cloud, axis, k, false, out); }

// sharded k-th: one pass = this rank's histogram of the next digit; the caller sums the histograms of all ranks
// (one all-reduce of 2048 counters per pass), picks the digit and calls the next pass with the extended (prefix, mask)
int32_t hs_kth_shard_pass(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, int32_t pass, uint32_t prefix, uint32_t mask, uint32_t* hist_out) {
  HS_LOCK(ctx);
  if (!cloud || !hist_out || axis < 0 || axis > 2 || pass < 0 || pass > 2) HS_FAIL(ctx, HS_EINVAL, "hs_kth_shard_pass: bad arguments");
  if (cloud->n >= (1ll << 32)) HS_FAIL(ctx, HS_EINVAL, "hs_kth_shard_pass: at most 2^32 - 1 points per shard");
  if (cloud->n == 0) { std::memset(hist_out, 0, sizeof(uint32_t) * 2048); return HS_OK; }  // an empty shard adds nothing
  uint32_t* d_buf = nullptr;
  DevTemps tmp;
  HS_CUDA_TRY(ctx, tmp.alloc(&d_buf, sizeof(uint32_t) * 2048));
  int32_t rc = launch_kth_shard_pass(ctx, cloud->d, cloud->n, axis, pass, prefix, mask, d_buf);
  if (rc == HS_OK) rc = copy_d2h_sync(ctx, hist_out, d_buf, sizeof(uint32_t) * 2048);
  cudaStreamSynchronize(ctx->stream);
  return rc;
}
uint32_t hs_kth_key_of_float(float v) { uint32_t b; std::memcpy(&b, &v, 4); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
float hs_kth_float_of_key(uint32_t u) { const uint32_t b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u; float v; std::memcpy(&v, &b, 4); return v; }

static int32_t filter_impl(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, float limit, const hs_cloud* colors, hs_cloud* out, hs_cloud* colors_out, int64_t* n_out) {
  if (!cloud || !out || axis < 0 || axis > 2) HS_FAIL(ctx, HS_EINVAL, "hs_filter_le: bad arguments");
  if (out->cap < cloud->n || out->d == cloud->d) HS_FAIL(ctx, HS_EINVAL, "hs_filter_le: output cloud too small or aliases the input");
  if ((colors != nullptr) != (colors_out != nullptr)) HS_FAIL(ctx, HS_EINVAL, "hs_filter_le: colors in/out must both be given");
  if (colors && (colors->n != cloud->n || colors_out->cap < cloud->n)) HS_FAIL(ctx, HS_EINVAL, "hs_filter_le: colors must be same size as the cloud");  // Main.hs:114
  if (cloud->n >= (static_cast<int64_t>(1) << 32)) HS_FAIL(ctx, HS_EINVAL, "hs_filter_le: clouds of 2^32 points or more are not supported (32-bit compaction offsets); filter per shard");
  int64_t* d_n = reinterpret_cast<int64_t*>(ctx->d_small);
  if (int32_t rc = launch_filter_le(ctx, cloud->d, cloud->n, axis, limit, colors ? colors->d : nullptr, out->d, colors_out ? colors_out->d : nullptr, d_n)) return rc;
  int64_t m = 0;
  if (int32_t rc = copy_d2h_sync(ctx, &m, d_n, sizeof m)) return rc;
  out->n = m;
  if (colors_out) colors_out->n = m;
  if (n_out) *n_out = m;
  return HS_OK;
}
int32_t hs_filter_le(hs_ctx* ctx, const hs_cloud* cloud, int32_t axis, float limit, const hs_cloud* colors, hs_cloud* out, hs_cloud* colors_out, int64_t* n_out) {
  HS_LOCK(ctx);
  return filter_impl(ctx, cloud, axis, limit, colors, out, colors_out, n_out);
}
int32_t hs_remove_ceiling(hs_ctx* ctx, const hs_cloud* cloud, const hs_cloud* colors, hs_cloud* out, hs_cloud* colors_out, int64_t* n_out, float* y_limit) {
  HS_LOCK(ctx);
  if (!cloud || !out) HS_FAIL(ctx, HS_EINVAL, "hs_remove_ceiling: bad arguments");
  if (cloud->n == 0) { out->n = 0; if (colors_out) colors_out->n = 0; if (n_out) *n_out = 0; return HS_OK; }  // V.null guard, Main.hs:2657
  float ylim = 0.f;
  if (int32_t rc = kth_impl(ctx, cloud, 1, cloud->n / 5, true, &ylim)) return rc;  // n < 5 => k = 0 => the reference errors too
  if (y_limit) *y_limit = ylim;
  return filter_impl(ctx, cloud, 1, ylim, colors, out, colors_out, n_out);
}

// ---- room input formats ------------------------------------------------------------------------------------
int32_t hs_plane_eqs_from_text(const char* text, int64_t len, float* planes_out, int32_t cap, int32_t* n_out) {
  if (!text || len < 0 || !n_out || (cap > 0 && !planes_out)) return HS_EINVAL;
  std::vector<float> pl;
  const int n = hs::parse_planes_txt(text, static_cast<size_t>(len), &pl);
  if (n < 0) { *n_out = 0; return HS_EIO; }  // "Could not load planes" (Main.hs:1388)
  *n_out = n;
  if (n > cap) return HS_EINVAL;  // n_out tells the caller how much room it needs
  std::memcpy(planes_out, pl.data(), sizeof(float) * 4 * n);
  return HS_OK;
}

int32_t hs_plane_eqs_from_file(const char* path, float* planes_out, int32_t cap, int32_t* n_out) {
  if (!path) return HS_EINVAL;
  std::vector<char> buf;
  std::string err;
  if (!hs::read_file(path, &buf, &err)) return HS_EIO;
  return hs_plane_eqs_from_text(buf.data(), static_cast<int64_t>(buf.size()), planes_out, cap, n_out);
}

int32_t hs_make_inward_facing(const float room_center[3], const float* plane_means, float* planes_inout, int32_t K) {
  if (!room_center || !plane_means || !planes_inout || K < 0) return HS_EINVAL;
  hs::make_inward_facing(room_center, plane_means, planes_inout, K);
  return HS_OK;
}

// host side of a PCD load: header, then the DATA section in the form the unpack kernel takes
struct PcdStaged {
  hs::PcdHeader h;
  std::vector<char> file;
  std::vector<uint8_t> soa;     // binary_compressed after LZF
  std::vector<uint32_t> rec;    // ascii records
  const uint8_t* raw = nullptr;
  size_t raw_bytes = 0;
  int64_t off[6] = {0, 0, 0, -1, -1, -1}, stride[6] = {0, 0, 0, 0, 0, 0};
  int rgb_bytes = 0;  // colours as three separate bytes (PLY) instead of one packed word (PCD)
  bool has_rgb = false;
};
static bool pcd_stage(const char* path, PcdStaged* st, std::string* err) {
  if (!hs::read_file(path, &st->file, err)) return false;
  if (!hs::pcd_parse_header(st->file.data(), st->file.size(), &st->h, err)) return false;
  const hs::PcdHeader& h = st->h;
  int f[4];
  if (!hs::pcd_layout(h, &f[0], &f[1], &f[2], &f[3], err)) return false;
  st->has_rgb = f[3] >= 0;
  const int64_t n = h.points;
  if (h.data_kind == 0) {
    int W = 3;
    if (!hs::pcd_ascii_records(st->file.data(), st->file.size(), h, &st->rec, &W, err)) return false;
    st->raw = reinterpret_cast<const uint8_t*>(st->rec.data());
    st->raw_bytes = st->rec.size() * 4;
    for (int c = 0; c < 4; ++c) { st->off[c] = 4 * c; st->stride[c] = 4 * W; }
  } else if (h.data_kind == 1) {
    st->raw = reinterpret_cast<const uint8_t*>(st->file.data()) + h.data_offset;
    st->raw_bytes = static_cast<size_t>(n) * h.point_step;
    if (h.data_offset + st->raw_bytes > st->file.size()) { *err = "PCD: binary data ends early"; return false; }
    for (int c = 0; c < 4; ++c) { st->off[c] = f[c] >= 0 ? h.field_offset[f[c]] : -1; st->stride[c] = h.point_step; }
  } else {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(st->file.data()) + h.data_offset;
    if (h.data_offset + 8 > st->file.size()) { *err = "PCD: compressed data ends early"; return false; }
    uint32_t csz, usz;
    std::memcpy(&csz, p, 4); std::memcpy(&usz, p + 4, 4);
    if (h.data_offset + 8 + csz > st->file.size() || usz != static_cast<uint64_t>(n) * h.point_step) { *err = "PCD: bad compressed sizes"; return false; }
    st->soa.resize(usz);
    if (!hs::lzf_decompress(p + 8, csz, st->soa.data(), usz)) { *err = "PCD: LZF stream is corrupt"; return false; }
    st->raw = st->soa.data();
    st->raw_bytes = usz;
    for (int c = 0; c < 4; ++c) {  // field-major: all values of field 0, then field 1, ...
      st->off[c] = f[c] >= 0 ? static_cast<int64_t>(h.field_offset[f[c]]) * n : -1;
      st->stride[c] = f[c] >= 0 ? h.size[f[c]] * h.count[f[c]] : 0;
    }
  }
  return true;
}

int32_t hs_pcd_info(const char* path, int64_t* n_points, int32_t* has_rgb, int32_t* data_kind) {
  if (!path) return HS_EINVAL;
  std::vector<char> buf;
  std::string err;
  hs::PcdHeader h;
  int f[4];
  try {
    if (!hs::read_file(path, &buf, &err) || !hs::pcd_parse_header(buf.data(), buf.size(), &h, &err) || !hs::pcd_layout(h, &f[0], &f[1], &f[2], &f[3], &err)) return HS_EIO;
  } catch (const std::exception&) { return HS_ENOMEM; }
  if (n_points) *n_points = h.points;
  if (has_rgb) *has_rgb = f[3] >= 0;
  if (data_kind) *data_kind = h.data_kind;
  return HS_OK;
}

// staged file -> device clouds: the DATA section goes up as it is and is unpacked on the GPU
static int32_t staged_to_device(hs_ctx* ctx, const PcdStaged& st, int64_t n, const char* path, hs_cloud** cloud_out, hs_cloud** colors_out) {
  if (n == 0) { ctx->err = std::string("File ") + path + " contains no points!"; return HS_EIO; }  // Main.hs:1344
  hs_cloud *cl = nullptr, *col = nullptr;
  int32_t rc = hs_cloud_alloc(ctx, n, &cl);
  const bool want_rgb = colors_out && st.has_rgb;
  if (rc == HS_OK && want_rgb) rc = hs_cloud_alloc(ctx, n, &col);
  if (rc == HS_OK) {
    HS_LOCK(ctx);
    uint8_t* d_raw = nullptr;
    cudaError_t e = cudaMalloc(&d_raw, st.raw_bytes + 16);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); rc = HS_ENOMEM; }
    if (rc == HS_OK) rc = copy_h2d(ctx, d_raw, st.raw, st.raw_bytes);
    if (rc == HS_OK) rc = launch_pcd_unpack(ctx, d_raw, n, st.off, st.stride, st.rgb_bytes, cl->d, want_rgb ? col->d : nullptr);
    cudaStreamSynchronize(ctx->stream);  // the staged host buffers die with this call
    if (d_raw) cudaFree(d_raw);
  }
  if (rc != HS_OK) { if (cl) hs_cloud_free(ctx, cl); if (col) hs_cloud_free(ctx, col); return rc; }
  *cloud_out = cl;
  if (colors_out) *colors_out = col;
  return HS_OK;
}

int32_t hs_ply_info(const char* path, int64_t* n_vertices, int32_t* has_rgb, int32_t* is_ascii) {
  if (!path) return HS_EINVAL;
  try {
    std::vector<char> buf;
    std::string err;
    hs::PlyHeader h;
    if (!hs::read_file(path, &buf, &err) || !hs::ply_parse_header(buf.data(), buf.size(), &h, &err)) return HS_EIO;
    const int f[6] = {h.find("x"), h.find("y"), h.find("z"), h.find("red"), h.find("green"), h.find("blue")};
    for (int c = 0; c < 3; ++c)
      if (f[c] < 0 || h.sizes[f[c]] != 4 || (h.types[f[c]] != "float" && h.types[f[c]] != "float32")) return HS_EIO;
    if (n_vertices) *n_vertices = h.n;
    if (has_rgb) *has_rgb = f[3] >= 0 && f[4] >= 0 && f[5] >= 0 && h.sizes[f[3]] == 1 && h.sizes[f[4]] == 1 && h.sizes[f[5]] == 1;
    if (is_ascii) *is_ascii = h.ascii ? 1 : 0;
    return HS_OK;
  } catch (const std::exception&) { return HS_ENOMEM; }
}

int32_t hs_cloud_from_pcd(hs_ctx* ctx, const char* path, hs_cloud** cloud_out, hs_cloud** colors_out) {
  if (!ctx) return HS_EINVAL;
  if (!path || !cloud_out) { ctx->err = "hs_cloud_from_pcd: bad arguments"; return HS_EINVAL; }
  *cloud_out = nullptr;
  if (colors_out) *colors_out = nullptr;
  try {  // nothing may unwind through the C ABI (a hostile header can ask for any amount of memory)
    PcdStaged st;
    std::string err;
    if (!pcd_stage(path, &st, &err)) { ctx->err = err; return HS_EIO; }
    return staged_to_device(ctx, st, st.h.points, path, cloud_out, colors_out);
  } catch (const std::exception& e) { ctx->err = std::string("hs_cloud_from_pcd: ") + e.what(); return HS_ENOMEM; }
}

int32_t hs_cloud_from_ply(hs_ctx* ctx, const char* path, hs_cloud** cloud_out, hs_cloud** colors_out) {
  if (!ctx) return HS_EINVAL;
  if (!path || !cloud_out) { ctx->err = "hs_cloud_from_ply: bad arguments"; return HS_EINVAL; }
  *cloud_out = nullptr;
  if (colors_out) *colors_out = nullptr;
  try {
  PcdStaged st;
  hs::PlyHeader h;
  std::string err;
  if (!hs::read_file(path, &st.file, &err) || !hs::ply_parse_header(st.file.data(), st.file.size(), &h, &err)) { ctx->err = err; return HS_EIO; }
  const int f[6] = {h.find("x"), h.find("y"), h.find("z"), h.find("red"), h.find("green"), h.find("blue")};
  for (int c = 0; c < 3; ++c)
    if (f[c] < 0 || h.sizes[f[c]] != 4 || (h.types[f[c]] != "float" && h.types[f[c]] != "float32")) { ctx->err = "PLY: x y z must be float properties"; return HS_EIO; }
  st.has_rgb = f[3] >= 0 && f[4] >= 0 && f[5] >= 0 && h.sizes[f[3]] == 1 && h.sizes[f[4]] == 1 && h.sizes[f[5]] == 1;
  if (h.ascii) {
    if (!hs::ply_ascii_records(st.file.data(), st.file.size(), h, st.has_rgb, &st.rec, &err)) { ctx->err = err; return HS_EIO; }
    const int W = st.has_rgb ? 4 : 3;
    st.raw = reinterpret_cast<const uint8_t*>(st.rec.data());
    st.raw_bytes = st.rec.size() * 4;
    for (int c = 0; c < 4; ++c) { st.off[c] = 4 * c; st.stride[c] = 4 * W; }
  } else {
    st.raw = reinterpret_cast<const uint8_t*>(st.file.data()) + h.data_offset;
    st.raw_bytes = static_cast<size_t>(h.n) * h.step;
    if (h.data_offset + st.raw_bytes > st.file.size()) { ctx->err = "PLY: binary data ends early"; return HS_EIO; }
    for (int c = 0; c < 6; ++c) { st.off[c] = f[c] >= 0 ? h.offsets[f[c]] : -1; st.stride[c] = h.step; }
    st.rgb_bytes = 1;
  }
  return staged_to_device(ctx, st, h.n, path, cloud_out, colors_out);
  } catch (const std::exception& e) { ctx->err = std::string("hs_cloud_from_ply: ") + e.what(); return HS_ENOMEM; }
}

int32_t hs_write_pcd(hs_ctx* ctx, const hs_cloud* cloud, const uint8_t* rgb, const char* path) {
  if (!ctx) return HS_EINVAL;
  if (!cloud || !path) { ctx->err = "hs_write_pcd: bad arguments"; return HS_EINVAL; }
  std::vector<float> host(static_cast<size_t>(cloud->n) * 3);
  if (int32_t rc = hs_cloud_download(ctx, cloud, host.data())) return rc;
  std::string err;
  if (!hs::write_pcd(path, host.data(), rgb, cloud->n, &err)) { ctx->err = err; return HS_EIO; }
  return HS_OK;
}

int32_t hs_transform_from_text(const char* text, int64_t len, float m_rowmajor[16]) {
  if (!text || len < 0 || !m_rowmajor) return HS_EINVAL;
  return hs::parse_transform_text(text, static_cast<size_t>(len), m_rowmajor) ? HS_OK : HS_EIO;
}

// the plane hulls are a few dozen points each: read on the host, mean as the reference's Float fold (planeMean, Main.hs:1608)
static bool pcd_host_mean(const char* path, float mean[3], std::string* err) {
  PcdStaged st;
  if (!pcd_stage(path, &st, err)) return false;
  const int64_t n = st.h.points;
  std::vector<float> xyz(static_cast<size_t>(n) * 3);
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) std::memcpy(&xyz[3 * i + c], st.raw + st.off[c] + i * st.stride[c], 4);
  if (!hs::point_mean_f32seq(xyz.data(), n, mean)) { *err = "pointMean: empty"; return false; }  // Main.hs:1597
  return true;
}

int32_t hs_load_room(hs_ctx* ctx, const char* dir, hs_cloud** cloud_out, hs_cloud** colors_out, float* planes_out, int32_t cap, int32_t* K_out) {
  if (!ctx) return HS_EINVAL;
  if (!dir || !cloud_out || !planes_out || !K_out) { ctx->err = "hs_load_room: bad arguments"; return HS_EINVAL; }
  const std::string d(dir);
  if (int32_t rc = hs_cloud_from_pcd(ctx, (d + "/cloud_downsampled.pcd").c_str(), cloud_out, colors_out)) return rc;  // Main.hs:1742-1744
  auto fail = [&](int32_t rc, const std::string& msg) {
    ctx->err = msg;
    hs_cloud_free(ctx, *cloud_out); *cloud_out = nullptr;
    if (colors_out && *colors_out) { hs_cloud_free(ctx, *colors_out); *colors_out = nullptr; }
    return rc;
  };
  int32_t K = 0;
  int32_t rc = hs_plane_eqs_from_file((d + "/planes.txt").c_str(), planes_out, cap, &K);
  if (rc != HS_OK) return fail(rc, rc == HS_EIO ? "Could not load planes: " + d + "/planes.txt" : "hs_load_room: more planes than the caller has room for");
  std::vector<float> means(static_cast<size_t>(K) * 3);
  for (int k = 0; k < K; ++k) {  // planesFromDir, Main.hs:1391-1404
    std::string err;
    bool ok = false;
    try { ok = pcd_host_mean((d + "/cloud_plane_hull" + std::to_string(k) + ".pcd").c_str(), &means[3 * k], &err); }
    catch (const std::exception& e) { err = e.what(); }
    if (!ok) return fail(HS_EIO, err);
  }
  double m[3];
  float md;
  if ((rc = hs_mean_extent(ctx, *cloud_out, m, &md)) != HS_OK) return fail(rc, ctx->err);  // roomCenter = cloudMean cloud, Double sums on the GPU
  const float center[3] = {static_cast<float>(m[0]), static_cast<float>(m[1]), static_cast<float>(m[2])};
  hs::make_inward_facing(center, means.data(), planes_out, K);
  *K_out = K;
  return HS_OK;
}

// ---- plane algebra of the room-editing actions (Main.hs:1553-1578, :1681-1688): host only, Float, reference evaluation order --------
int32_t hs_rotation_between_plane_eqs(const float plane1[4], const float plane2[4], float R_out[9]) {
  if (!plane1 || !plane2 || !R_out) return HS_EINVAL;
  const hs::M3<float> R = hs::rotation_between_normals(hs::V3<float>{plane1[0], plane1[1], plane1[2]}, hs::V3<float>{plane2[0], plane2[1], plane2[2]});
  for (int r = 0; r < 3; ++r) { R_out[3 * r] = R.r[r].x; R_out[3 * r + 1] = R.r[r].y; R_out[3 * r + 2] = R.r[r].z; }
  return HS_OK;
}
int32_t hs_rotate_plane_eq_around(const float center[3], const float R[9], const float plane_in[4], float plane_out[4]) {
  if (!center || !R || !plane_in || !plane_out) return HS_EINVAL;
  const hs::M3<float> M{{{R[0], R[1], R[2]}, {R[3], R[4], R[5]}, {R[6], R[7], R[8]}}};
  const hs::PlaneEq e = hs::rotate_plane_eq_around(hs::V3<float>{center[0], center[1], center[2]}, M,
                                                   hs::PlaneEq{hs::V3<float>{plane_in[0], plane_in[1], plane_in[2]}, plane_in[3]});
  plane_out[0] = e.n.x; plane_out[1] = e.n.y; plane_out[2] = e.n.z; plane_out[3] = e.d;
  return HS_OK;
}
int32_t hs_translate_plane_eq(const float off[3], const float plane_in[4], float plane_out[4]) {
  if (!off || !plane_in || !plane_out) return HS_EINVAL;
  const hs::PlaneEq e = hs::translate_plane_eq(hs::V3<float>{off[0], off[1], off[2]}, hs::PlaneEq{hs::V3<float>{plane_in[0], plane_in[1], plane_in[2]}, plane_in[3]});
  plane_out[0] = e.n.x; plane_out[1] = e.n.y; plane_out[2] = e.n.z; plane_out[3] = e.d;
  return HS_OK;
}

// roomProj bookkeeping of rotateRoomAround / translateRoom / projectRoom (Main.hs:1665-1720)
int32_t hs_proj_compose(const float a[16], const float b[16], float out[16]) {
  if (!a || !b || !out) return HS_EINVAL;
  hs::M4 A, B;
  std::memcpy(A.m, a, sizeof A.m); std::memcpy(B.m, b, sizeof B.m);
  const hs::M4 r = hs::proj_compose(A, B);
  std::memcpy(out, r.m, sizeof r.m);
  return HS_OK;
}
int32_t hs_proj_translate(const float proj[16], const float off[3], float out[16]) {
  if (!proj || !off || !out) return HS_EINVAL;
  hs::M4 P;
  std::memcpy(P.m, proj, sizeof P.m);
  const hs::M4 r = hs::proj_translate4(hs::V3<float>{off[0], off[1], off[2]}, P);
  std::memcpy(out, r.m, sizeof r.m);
  return HS_OK;
}
int32_t hs_proj_rotate_around(const float proj[16], const float center[3], const float R[9], float out[16]) {
  if (!proj || !center || !R || !out) return HS_EINVAL;
  hs::M4 P;
  std::memcpy(P.m, proj, sizeof P.m);
  const hs::M3<float> M{{{R[0], R[1], R[2]}, {R[3], R[4], R[5]}, {R[6], R[7], R[8]}}};
  const hs::M4 r = hs::proj_rotate_around(hs::V3<float>{center[0], center[1], center[2]}, M, P);
  std::memcpy(out, r.m, sizeof r.m);
  return HS_OK;
}

int32_t hs_plane_corner(const float plane1[4], const float plane2[4], const float plane3[4], float corner_out[3]) {
  if (!plane1 || !plane2 || !plane3 || !corner_out) return HS_EINVAL;
  return hs::plane_corner(plane1, plane2, plane3, corner_out) ? HS_OK : HS_ESINGULAR;  // Nothing (Main.hs:1428-1430)
}

// ---- host-side module mirrors ------------------------------------------------------------------------------------------------------
int32_t hs_cuboid_from_params(const double params[10], double out[24]) { if (!params || !out) return HS_EINVAL; hs::cuboid_from_params(params, out); return HS_OK; }
double hs_errfun(const double corners[24], const double params[10]) { return hs::errfun(corners, params); }
double hs_errfun_closest(const double* pts, int32_t npts, const double params[10]) { return hs::errfun_closest(pts, npts, params); }
int32_t hs_guess_dims(const double corners[24], double out[3]) { if (!corners || !out) return HS_EINVAL; hs::guess_dims(corners, out); return HS_OK; }

int32_t hs_fit_cuboid(const double corners[24], int32_t variant, double params_out[10], int32_t* steps, double* err, double* path_out, int32_t path_cap) {
  if (!corners || !params_out || variant < 0 || variant > 2) return HS_EINVAL;
  hs::FitResult r = hs::fit_cuboid(corners, variant, path_out != nullptr);
  for (int i = 0; i < 10; ++i) params_out[i] = r.params[i];
  if (steps) *steps = r.steps;
  if (err) *err = r.err;
  if (path_out && r.path_cols > 0) {
    const size_t rows = std::min<size_t>(r.path.size() / r.path_cols, static_cast<size_t>(std::max(path_cap, 0)));
    std::memcpy(path_out, r.path.data(), rows * r.path_cols * sizeof(double));
  }
  return HS_OK;
}

int32_t hs_fit_cuboid_cloud_bfgs(hs_ctx* ctx, const hs_cloud* cloud, const double init[10], int32_t max_iter, double gtol,
                                 double params_out[10], double* f_out, int32_t* iters, int32_t* evals) {
  if (!ctx) return HS_EINVAL;
  if (!cloud || !init || !params_out) { ctx->err = "hs_fit_cuboid_cloud_bfgs: bad arguments"; return HS_EINVAL; }
  int32_t last_rc = HS_OK;
  // The optimiser loop talks to ONE resident kernel (an evaluation session): every objective evaluation is a parameter set posted
  // to the device and a 24-Double record coming back, not a kernel launch.  The exact-products mode keeps the launch-per-evaluation form.
  hs_eval_session* sess = nullptr;
  const int64_t off[2] = {0, cloud->n};
  if (ctx->modes[HS_MODE_EVAL_KERNEL] != HS_EVAL_EXACT) {
    if (int32_t rc = hs_eval_session_begin(ctx, cloud, off, 1, 0, &sess)) return rc;
  }
  auto eval = [&](const double* x, double* f, double* g) {
    if (sess) {
      double rec[HS_REC];
      last_rc = hs_eval_session_eval(sess, x, rec);
      if (last_rc == HS_OK) hs::cuboid_grad_from_sums(x, rec, f, g, nullptr);
    } else {
      last_rc = hs_cuboid_residual_grad(ctx, cloud, x, f, g, nullptr);
    }
    return last_rc == HS_OK;
  };
  hs::BFGSResult r = hs::bfgs(eval, init, 10, max_iter > 0 ? max_iter : 200, gtol > 0 ? gtol : 1e-6);
  if (sess) {
    const std::string msg = ctx->err;
    const int32_t rc_end = hs_eval_session_end(sess);
    if (last_rc != HS_OK) ctx->err = msg;  // the evaluation's own message, not the close's
    else if (rc_end != HS_OK) return rc_end;
  }
  if (!r.ok) return last_rc != HS_OK ? last_rc : HS_ECUDA;
  for (int i = 0; i < 10; ++i) params_out[i] = r.x[i];
  if (f_out) *f_out = r.f;
  if (iters) *iters = r.iters;
  if (evals) *evals = r.evals;
  return HS_OK;
}

int32_t hs_bfgs_minimize(hs_objective_fn eval, void* user, const double* x0, int32_t n, int32_t max_iter, double gtol, double* x_out, double* f_out,
                         int32_t* iters, int32_t* evals) {
  if (!eval || !x0 || !x_out || n < 1 || n > 4096) return HS_EINVAL;
  auto fn = [&](const double* x, double* f, double* g) { return eval(user, x, f, g) == 0; };
  hs::BFGSResult r = hs::bfgs(fn, x0, n, max_iter > 0 ? max_iter : 200, gtol > 0 ? gtol : 1e-6);
  if (!r.ok) return HS_EINVAL;
  for (int i = 0; i < n; ++i) x_out[i] = r.x[i];
  if (f_out) *f_out = r.f;
  if (iters) *iters = r.iters;
  if (evals) *evals = r.evals;
  return HS_OK;
}

int32_t hs_nm_minimize(hs_value_fn value, void* user, const double* x0, const double* step, int32_t n, double eps, int32_t maxit, double* x_out,
                       double* f_out, int32_t* iters, int32_t* evals) {
  if (!value || !x0 || !step || !x_out || n < 1 || n > 4096) return HS_EINVAL;
  auto f = [&](const std::vector<double>& x) { return value(user, x.data()); };
  int ev = 0;
  hs::NMResult r = hs::nm_simplex2(f, std::vector<double>(x0, x0 + n), std::vector<double>(step, step + n), eps > 0 ? eps : 1e-8,
                                   maxit > 0 ? maxit : 2000, false, nullptr, &ev);
  for (int i = 0; i < n; ++i) x_out[i] = r.x[i];
  if (f_out) *f_out = r.fval;
  if (iters) *iters = r.iters;
  if (evals) *evals = ev;
  return HS_OK;
}

int32_t hs_fit_cuboid_cloud_nm(hs_ctx* ctx, const hs_cloud* cloud, const double init[10], const double step[10], double eps, int32_t maxit,
                               double params_out[10], double* f_out, int32_t* iters, int32_t* evals) {
  if (!ctx) return HS_EINVAL;
  if (!cloud || !init || !step || !params_out) { ctx->err = "hs_fit_cuboid_cloud_nm: bad arguments"; return HS_EINVAL; }
  hs_eval_session* sess = nullptr;
  const int64_t off[2] = {0, cloud->n};
  if (int32_t rc = hs_eval_session_begin(ctx, cloud, off, 1, 0, &sess)) return rc;
  int32_t last_rc = HS_OK;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  auto f = [&](const std::vector<double>& x) {
    double rec[HS_REC];
    if (last_rc == HS_OK) last_rc = hs_eval_session_eval(sess, x.data(), rec);
    return last_rc == HS_OK ? rec[0] : nan;
  };
  int64_t posted = 0;
  hs::NMBatch batch = [&](const std::vector<std::vector<double>>& xs, std::vector<double>& ys) {
    std::vector<double> flat;
    for (auto& x : xs) flat.insert(flat.end(), x.begin(), x.end());
    if (last_rc == HS_OK) last_rc = hs_eval_session_post(sess, flat.data(), static_cast<int32_t>(xs.size()));
    for (size_t q = 0; q < xs.size(); ++q) {
      double rec[HS_REC];
      if (last_rc == HS_OK) last_rc = hs_eval_session_wait(sess, posted + static_cast<int64_t>(q), rec);
      ys[q] = last_rc == HS_OK ? rec[0] : nan;
    }
    posted += static_cast<int64_t>(xs.size());
  };
  // evaluations posted one by one go through f -> hs_eval_session_eval, which posts and waits: keep `posted` in step
  auto f_counted = [&](const std::vector<double>& x) { const double v = f(x); ++posted; return v; };
  int ev = 0;
  hs::NMResult r = hs::nm_simplex2(f_counted, std::vector<double>(init, init + 10), std::vector<double>(step, step + 10), eps > 0 ? eps : 1e-8,
                                   maxit > 0 ? maxit : 2000, false, &batch, &ev);
  const std::string msg = ctx->err;
  const int32_t rc_end = hs_eval_session_end(sess);
  if (last_rc != HS_OK) { ctx->err = msg; return last_rc; }
  if (rc_end != HS_OK) return rc_end;
  for (int i = 0; i < 10; ++i) params_out[i] = r.x[i];
  if (f_out) *f_out = r.fval;
  if (iters) *iters = r.iters;
  if (evals) *evals = ev;
  return HS_OK;
}

int32_t hs_lstsq_distances(const int32_t* i_idx, const int32_t* j_idx, const double* d, int32_t m, int32_t n_nodes, double* pos_out, double* rmse_out) {
  if (!i_idx || !j_idx || !d || !pos_out || !rmse_out || m < 1 || n_nodes < 1) return HS_EINVAL;
  for (int r = 0; r < m; ++r)
    if (i_idx[r] < 0 || j_idx[r] < 0 || i_idx[r] >= n_nodes || j_idx[r] >= n_nodes) return HS_EINVAL;
  return hs::lstsq_distances(i_idx, j_idx, d, m, n_nodes, pos_out, rmse_out) ? HS_OK : HS_ESINGULAR;
}

}  // extern "C"
