// Host side of the evaluation kernel (k_eval.cuh): the weighted static partition, the one-shot launcher behind
// hs_rooms_cuboid_sums*, the standalone peer exchange, and the evaluation session API (hs_eval_session_*): a resident kernel
// that evaluates one parameter set after the other without relaunching — what FitCuboidBFGS's optimiser loop
// (FitCuboidBFGS.hs:184,201,233: up to 2000 objective evaluations per stage) needs from the device.
#include <algorithm>
#include <cmath>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <new>
#include <thread>

#include "../host/hs_host.hpp"
#include "k_eval.cuh"

namespace hsk {

// product configuration: 12 consumer warps, 3 ring stages of 72 KB, 16 points per thread per tile, Float chains of 256 points
constexpr int EVK_NCONS = 384, EVK_STAGES = 3, EVK_GPT = 4, EVK_FLUSH = 16;
constexpr int EVK_SEG_COST_DEFAULT = 1536;  // an extra room segment in a block costs about this many 4-point groups (flush + block sum)

constexpr size_t eval_smem_bytes() {
  return static_cast<size_t>(EVK_STAGES) * EVK_GPT * EVK_NCONS * 48 + 2 * EVK_STAGES * 8 + static_cast<size_t>(EV_WS) * (EVK_NCONS / 32) * EV_NRAW * 8;
}

// standalone exchange for the kernels that do not carry it in their tail (one warp)
__global__ void __launch_bounds__(32) k_peer_allreduce(double* buf, int count, uint32_t epoch, uint32_t* h_status, const __grid_constant__ PeerExchange px) {
  peer_push_warp(px, epoch, buf, count);
  if (!peer_collect_warp(px, epoch, buf, count) && threadIdx.x == 0 && h_status) st_sys_u32(h_status, static_cast<uint32_t>(HS_ENCCL));
}

}  // namespace hsk

using namespace hsk;

struct hs_eval_state {
  EvalPlan h_plan;
  bool plan_valid = false;
  int plan_seg_cost = 0;
  EvalPlan* d_plan = nullptr;
  EvalCtl* d_ctl = nullptr;
  double* d_partials = nullptr;
  size_t partials_bytes = 0;
  double* d_local_rec = nullptr;
  bool attr_set[2] = {false, false};  // MaxDynamicSharedMemorySize set on this ctx's device: one-launch form, session form
  // session rings (allocated at the first session of the ctx, sized for HS_MAX_ROOMS, kept: begin must cost a launch, not five allocations)
  EvalCmd* h_cmds = nullptr;     // mapped
  EvalCmd* d_cmds = nullptr;
  EvalHostCtl* h_ctl = nullptr;  // mapped
  double* h_results = nullptr;   // mapped
  double* d_results = nullptr;
  unsigned long long* h_times = nullptr;  // mapped
};

struct hs_eval_session {
  hs_ctx* ctx = nullptr;
  int32_t nrooms = 0;
  bool exchange = false;
  EvalCmd* h_cmds = nullptr;     // the rings of the ctx's hs_eval_state (one session per ctx at a time)
  EvalHostCtl* h_ctl = nullptr;
  double* h_results = nullptr;
  double* d_results = nullptr;
  unsigned long long* h_times = nullptr;
  uint32_t posted = 0;
  EvalArgs args;          // the resident kernel's arguments
  unsigned long long* d_trace = nullptr;  // HS_EVAL_TRACE=<file>: per-block stamps of the first evaluations, written to <file> at end
  int nblocks = 0;
  bool launched = false;  // the kernel starts with the first post (its commands are in the ring before the kernel looks)
  bool stopped = false;
  bool empty = false;  // no point in any room: records are zero, no kernel runs
};

// ---------------------------------------------------------------------------------------------------------------------------
// static partition: blocks get contiguous ranges of 4-point groups; a block pays `seg_cost` groups for every room boundary inside
// its range, so that blocks with two segments finish together with the others
static int64_t plan_walk(int64_t G, int nb, int64_t target, int seg_cost, const int64_t* bounds, int nbounds, int64_t* g0_out) {
  int64_t g = 0;
  for (int b = 0; b < nb; ++b) {
    if (g0_out) g0_out[b] = g;
    // a block streams `target` groups minus seg_cost for every room boundary strictly inside its range.  If shortening the range
    // would push a boundary out of it, the block ends AT that boundary instead (the room starts the next block: no second segment).
    int64_t end = std::min(G, g + target);
    for (int it = 0; it <= nbounds; ++it) {
      int k = 0;
      for (int i = 0; i < nbounds; ++i) k += (bounds[i] > 4 * g && bounds[i] < 4 * end);
      const int64_t want = std::max(std::min(G, g + 1), std::min(G, g + target - static_cast<int64_t>(k) * seg_cost));
      if (want >= end) break;
      int64_t first_dropped = -1;
      for (int i = 0; i < nbounds; ++i)
        if (bounds[i] > 4 * want && bounds[i] < 4 * end && (first_dropped < 0 || bounds[i] < first_dropped)) first_dropped = bounds[i];
      if (first_dropped < 0) { end = want; break; }
      end = std::max(std::max(want, first_dropped / 4), g + 1);
    }
    g = end;
  }
  if (g0_out) g0_out[nb] = G;
  return g;
}

static void build_plan(EvalPlan& P, int64_t n, const int64_t* off, int nrooms, int nb_max, int seg_cost) {
  std::memset(&P, 0, sizeof P);
  P.n = n;
  P.nrooms = nrooms;
  for (int r = 0; r <= nrooms; ++r) P.off[r] = off[r];
  const int64_t G = (n + 3) >> 2;
  int64_t nb = std::min<int64_t>(nb_max, (G + EVK_NCONS - 1) / EVK_NCONS);
  nb = std::max<int64_t>(1, std::min<int64_t>(nb, EV_MAXB));
  P.nblocks = static_cast<int32_t>(nb);
  int64_t bounds[HS_MAX_ROOMS];
  int nbounds = 0;
  for (int r = 0; r < nrooms; ++r) {
    if (off[r + 1] <= off[r]) P.empty_mask |= 1u << r;
    if (off[r + 1] > off[r]) {
      ++P.nrooms_nonempty;
      if (off[r] > off[0] && off[r] > 0) bounds[nbounds++] = off[r];  // start of a non-empty room that is not the first point
    }
  }
  // smallest target for which nb blocks cover all groups
  int64_t lo = (G + nb - 1) / nb, hi = lo + static_cast<int64_t>(seg_cost) * (nbounds + 1) + 1;
  lo = std::max<int64_t>(lo, 1);
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (plan_walk(G, static_cast<int>(nb), mid, seg_cost, bounds, nbounds, nullptr) >= G) hi = mid; else lo = mid + 1;
  }
  plan_walk(G, static_cast<int>(nb), lo, seg_cost, bounds, nbounds, P.blk_g0);
  for (int r = 0; r < nrooms; ++r) { P.room_blo[r] = 0; P.room_nb[r] = 0; }
  for (int b = 0; b < nb; ++b) {
    const int64_t p0 = P.blk_g0[b] * 4, p1 = std::min(P.blk_g0[b + 1] * 4, n);
    int rfirst = 0, rlast = -1;
    bool any = false;
    for (int r = 0; r < nrooms; ++r)
      if (off[r] < p1 && off[r + 1] > p0 && off[r + 1] > off[r] && p1 > p0) {
        if (!any) { rfirst = r; any = true; }
        rlast = r;
        if (P.room_nb[r] == 0) P.room_blo[r] = b;
        P.room_nb[r] += 1;
      }
    P.blk_rfirst[b] = rfirst;
    P.blk_rlast[b] = rlast;
  }
}

static int32_t eval_state(hs_ctx* ctx, hs_eval_state** out) {
  if (!ctx->eval) {
    hs_eval_state* st = new (std::nothrow) hs_eval_state();
    if (!st) { ctx->err = "out of host memory"; return HS_ENOMEM; }
    ctx->eval = st;
    HS_CUDA_TRY(ctx, cudaMalloc(&st->d_plan, sizeof(EvalPlan)));
    HS_CUDA_TRY(ctx, cudaMalloc(&st->d_ctl, sizeof(EvalCtl)));
    HS_CUDA_TRY(ctx, cudaMemset(st->d_ctl, 0, sizeof(EvalCtl)));
    HS_CUDA_TRY(ctx, cudaMalloc(&st->d_local_rec, sizeof(double) * EV_D * HS_MAX_ROOMS * HS_REC));
  }
  *out = ctx->eval;
  return HS_OK;
}

void hs_eval_state_free(hs_ctx* ctx) {
  hs_eval_state* st = ctx->eval;
  if (!st) return;
  if (st->d_plan) cudaFree(st->d_plan);
  if (st->d_ctl) cudaFree(st->d_ctl);
  if (st->d_partials) cudaFree(st->d_partials);
  if (st->d_local_rec) cudaFree(st->d_local_rec);
  if (st->h_cmds) cudaFreeHost(st->h_cmds);
  if (st->h_ctl) cudaFreeHost(st->h_ctl);
  if (st->h_results) cudaFreeHost(st->h_results);
  if (st->h_times) cudaFreeHost(st->h_times);
  if (st->d_cmds) cudaFree(st->d_cmds);
  if (st->d_results) cudaFree(st->d_results);
  delete st;
  ctx->eval = nullptr;
}

// plan for (n, offsets) on this ctx: rebuilt and uploaded only when it changes
static int32_t eval_prepare(hs_ctx* ctx, int64_t n, const int64_t* off, int nrooms, hs_eval_state** out) {
  hs_eval_state* st = nullptr;
  if (int32_t rc = eval_state(ctx, &st)) return rc;
  const int seg_cost = ctx->modes[HS_MODE_EVAL_SEG_COST] > 0 ? ctx->modes[HS_MODE_EVAL_SEG_COST] : (ctx->modes[HS_MODE_EVAL_SEG_COST] < 0 ? 0 : EVK_SEG_COST_DEFAULT);
  bool same = st->plan_valid && st->h_plan.n == n && st->h_plan.nrooms == nrooms && st->plan_seg_cost == seg_cost;
  for (int r = 0; same && r <= nrooms; ++r) same = st->h_plan.off[r] == off[r];
  if (!same) {
    build_plan(st->h_plan, n, off, nrooms, ctx->sm_count, seg_cost);
    st->plan_seg_cost = seg_cost;
    HS_CUDA_TRY(ctx, cudaMemcpyAsync(st->d_plan, &st->h_plan, sizeof(EvalPlan), cudaMemcpyHostToDevice, ctx->stream));
    HS_CUDA_TRY(ctx, cudaMemsetAsync(st->d_ctl, 0, sizeof(EvalCtl), ctx->stream));  // tickets start from zero with every new plan
    st->plan_valid = true;
  }
  const size_t need = sizeof(double) * EV_D * static_cast<size_t>(nrooms) * st->h_plan.nblocks * EV_NRAW;
  if (need > st->partials_bytes) {
    if (st->d_partials) { HS_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); HS_CUDA_TRY(ctx, cudaFree(st->d_partials)); st->d_partials = nullptr; st->partials_bytes = 0; }
    HS_CUDA_TRY(ctx, cudaMalloc(&st->d_partials, need));
    st->partials_bytes = need;
  }
  *out = st;
  return HS_OK;
}

static void fill_cmd(EvalCmd& c, int r, const float pl[24], uint32_t seq = 0) {
  for (int j = 0; j < 3; ++j) {
    for (int k = 0; k < 3; ++k) c.c[r][3 * j + k] = pl[8 * j + k];
    c.c[r][9 + j] = pl[8 * j + 3];
    c.c[r][12 + j] = pl[8 * j + 7];
  }
  const uint32_t stamp = seq + 1u;  // the kernel's prefetch buffer is valid for evaluation seq iff it carries this stamp
  std::memcpy(&c.c[r][15], &stamp, 4);
}

int32_t launch_peer_allreduce(hs_ctx* ctx, double* d_buf, int count) {
  const uint32_t epoch = ++ctx->px.epoch;
  k_peer_allreduce<<<1, 32, 0, ctx->stream>>>(d_buf, count, epoch, ctx->d_status, ctx->px);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

// caller guarantees the planes are paired (cuboid rooms)
int32_t launch_eval(hs_ctx* ctx, const float* xyz, int64_t n, const RoomTable& rt, double* d_rec_out, bool exchange) {
  hs_eval_state* st = nullptr;
  if (int32_t rc = eval_prepare(ctx, n, rt.off, rt.nrooms, &st)) return rc;
  exchange = exchange && ctx->px.world > 1;
  if (st->h_plan.nrooms_nonempty == 0) {  // nothing to stream: zero records (and still this rank's part in the exchange)
    HS_CUDA_TRY(ctx, cudaMemsetAsync(d_rec_out, 0, sizeof(double) * rt.nrooms * HS_REC, ctx->stream));
    return exchange ? launch_peer_allreduce(ctx, d_rec_out, rt.nrooms * HS_REC) : HS_OK;
  }
  EvalCmd cmd;
  std::memset(&cmd, 0, sizeof cmd);
  for (int r = 0; r < rt.nrooms; ++r) fill_cmd(cmd, r, &rt.pl[r][0][0]);
  EvalArgs a;
  std::memset(&a, 0, sizeof a);
  a.xyz = xyz; a.plan = st->d_plan; a.ctl = st->d_ctl; a.partials = st->d_partials; a.local_rec = st->d_local_rec; a.out = d_rec_out;
  a.h_status = ctx->d_status;
  if (exchange) { a.px = ctx->px; a.epoch0 = ++ctx->px.epoch; }
  auto kern = k_eval<EVK_NCONS, EVK_STAGES, EVK_GPT, EVK_FLUSH, false>;
  if (!st->attr_set[0]) {
    HS_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(eval_smem_bytes())));
    st->attr_set[0] = true;
  }
  kern<<<st->h_plan.nblocks, EVK_NCONS + 96, eval_smem_bytes(), ctx->stream>>>(a, cmd);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  return HS_OK;
}

// ===========================================================================================================================
// evaluation sessions
// ===========================================================================================================================
#define HS_SLOCK(ctx)                                    \
  if (!(ctx)) return HS_EINVAL;                          \
  std::lock_guard<std::mutex> lock__((ctx)->mu);         \
  if (cudaSetDevice((ctx)->device) != cudaSuccess) { (ctx)->err = "cudaSetDevice failed"; return HS_ECUDA; }

static void session_free(hs_eval_session* s) { delete s; }

static int32_t session_rings(hs_ctx* ctx, hs_eval_state* st) {
  if (st->h_cmds) return HS_OK;
  const size_t rec_bytes = sizeof(double) * HS_MAX_ROOMS * HS_REC;
  HS_CUDA_TRY(ctx, cudaHostAlloc(&st->h_ctl, 64, cudaHostAllocMapped));
  HS_CUDA_TRY(ctx, cudaHostAlloc(&st->h_results, rec_bytes * EV_QCAP, cudaHostAllocMapped));
  HS_CUDA_TRY(ctx, cudaMalloc(&st->d_cmds, sizeof(EvalCmd) * EV_QCAP));
  HS_CUDA_TRY(ctx, cudaMalloc(&st->d_results, rec_bytes * EV_QCAP));
  HS_CUDA_TRY(ctx, cudaHostAlloc(&st->h_times, sizeof(unsigned long long) * 2 * EV_QCAP, cudaHostAllocMapped));
  std::memset(st->h_times, 0, sizeof(unsigned long long) * 2 * EV_QCAP);
  HS_CUDA_TRY(ctx, cudaHostAlloc(&st->h_cmds, sizeof(EvalCmd) * EV_QCAP, cudaHostAllocMapped));
  return HS_OK;
}

static inline uint32_t host_load(const uint32_t* p) { return reinterpret_cast<const std::atomic<uint32_t>*>(p)->load(std::memory_order_acquire); }
static inline void host_store(uint32_t* p, uint32_t v) { reinterpret_cast<std::atomic<uint32_t>*>(p)->store(v, std::memory_order_release); }

static int32_t session_check(hs_eval_session* s, const char* who) {
  hs_ctx* ctx = s->ctx;
  const uint32_t err = host_load(&s->h_ctl->error);
  if (err == 1) { ctx->err = std::string(who) + ": a peer did not deliver its records (exchange timed out)"; return HS_ENCCL; }
  if (err == 2) { ctx->err = std::string(who) + ": the session was idle for too long and shut itself down"; return HS_ECUDA; }
  return HS_OK;
}

// spin until `cond` holds; notices a dead kernel and gives up after 60 s
template <class Cond>
static int32_t session_spin(hs_eval_session* s, const char* who, Cond cond) {
  hs_ctx* ctx = s->ctx;
  const auto t0 = std::chrono::steady_clock::now();
  for (uint64_t it = 0;; ++it) {
    if (cond()) return HS_OK;
    if (int32_t rc = session_check(s, who)) return rc;
    if ((it & 0xfff) == 0xfff) {
      const cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q != cudaErrorNotReady && !cond()) {
        ctx->err = std::string(who) + ": the session kernel is not running" + (q == cudaSuccess ? "" : std::string(" (") + cudaGetErrorString(q) + ")");
        return HS_ECUDA;
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) { ctx->err = std::string(who) + ": timed out"; return HS_ECUDA; }
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

// The resident kernel is launched by the first post (or, in deferred mode, when the host first needs the device).  Its control block
// is zero at that point: zeroed at creation, after every plan change and by every hs_eval_session_end.
static int32_t session_launch(hs_eval_session* s) {
  if (s->launched || s->empty) return HS_OK;
  hs_ctx* ctx = s->ctx;
  hs_eval_state* st = ctx->eval;
  auto kern = k_eval<EVK_NCONS, EVK_STAGES, EVK_GPT, EVK_FLUSH, true>;
  if (!st->attr_set[1]) {
    HS_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(eval_smem_bytes())));
    st->attr_set[1] = true;
  }
  EvalCmd none;
  std::memset(&none, 0, sizeof none);
  kern<<<s->nblocks, EVK_NCONS + 128, eval_smem_bytes(), ctx->stream>>>(s->args, none);
  ctx->launches++;
  HS_CUDA_TRY(ctx, cudaGetLastError());
  s->launched = true;
  return HS_OK;
}

extern "C" {

// The evaluation kernel's static partition, host only (no device, no ctx): what `hs_rooms_cuboid_sums*` and the sessions would use
// for this cloud layout on a GPU with `sm_count` SMs.  For tests and for callers that want to see how a layout is cut.
int32_t hs_eval_plan(int64_t n, const int64_t* room_offsets, int32_t nrooms, int32_t sm_count, int32_t seg_cost, int32_t* nblocks_out,
                     int64_t* block_first_group_out /* nblocks + 1 */, int32_t* block_room_first_out, int32_t* block_room_last_out,
                     int32_t* room_first_block_out, int32_t* room_nblocks_out) {
  if (n < 0 || !room_offsets || nrooms < 1 || nrooms > HS_MAX_ROOMS || sm_count < 1 || !nblocks_out) return HS_EINVAL;
  for (int r = 0; r < nrooms; ++r)
    if (room_offsets[r] > room_offsets[r + 1]) return HS_EINVAL;
  if (room_offsets[0] < 0 || room_offsets[nrooms] > n) return HS_EINVAL;
  EvalPlan P;
  build_plan(P, n, room_offsets, nrooms, sm_count, seg_cost > 0 ? seg_cost : (seg_cost < 0 ? 0 : EVK_SEG_COST_DEFAULT));
  *nblocks_out = P.nblocks;
  for (int b = 0; b <= P.nblocks && block_first_group_out; ++b) block_first_group_out[b] = P.blk_g0[b];
  for (int b = 0; b < P.nblocks; ++b) {
    if (block_room_first_out) block_room_first_out[b] = P.blk_rfirst[b];
    if (block_room_last_out) block_room_last_out[b] = P.blk_rlast[b];
  }
  for (int r = 0; r < nrooms; ++r) {
    if (room_first_block_out) room_first_block_out[r] = P.room_blo[r];
    if (room_nblocks_out) room_nblocks_out[r] = P.room_nb[r];
  }
  return HS_OK;
}

int32_t hs_eval_session_begin(hs_ctx* ctx, const hs_cloud* cloud, const int64_t* room_offsets, int32_t nrooms, int32_t allreduce, hs_eval_session** out) {
  HS_SLOCK(ctx);
  if (!out || !cloud || !room_offsets || nrooms < 1 || nrooms > HS_MAX_ROOMS) { ctx->err = "hs_eval_session_begin: bad arguments (1 <= nrooms <= 32)"; return HS_EINVAL; }
  *out = nullptr;
  if (ctx->session) { ctx->err = "hs_eval_session_begin: this context already has an open session"; return HS_EINVAL; }
  for (int r = 0; r < nrooms; ++r)
    if (room_offsets[r] > room_offsets[r + 1]) { ctx->err = "hs_eval_session_begin: room offsets must be non-decreasing"; return HS_EINVAL; }
  if (room_offsets[0] < 0 || room_offsets[nrooms] > cloud->n) { ctx->err = "hs_eval_session_begin: room offsets outside the cloud"; return HS_EINVAL; }
  if (allreduce && ctx->px.world < 1) { ctx->err = "hs_eval_session_begin: no peer group (hs_peer_mailbox_create / _connect or hs_peer_group_create_local first)"; return HS_EINVAL; }
  hs_eval_state* st = nullptr;
  if (int32_t rc = eval_prepare(ctx, cloud->n, room_offsets, nrooms, &st)) return rc;
  if (int32_t rc = session_rings(ctx, st)) return rc;
  hs_eval_session* s = new (std::nothrow) hs_eval_session();
  if (!s) { ctx->err = "out of host memory"; return HS_ENOMEM; }
  s->ctx = ctx;
  s->h_cmds = st->h_cmds; s->h_ctl = st->h_ctl; s->h_results = st->h_results; s->d_results = st->d_results; s->h_times = st->h_times;
  s->nrooms = nrooms;
  s->exchange = allreduce && ctx->px.world > 1;
  s->empty = st->h_plan.nrooms_nonempty == 0;
  const size_t rec_bytes = sizeof(double) * nrooms * HS_REC;
  auto fail = [&](cudaError_t e) { ctx->err = std::string("hs_eval_session_begin: ") + cudaGetErrorString(e); session_free(s); return HS_ECUDA; };
  cudaError_t e;
  std::memset(s->h_ctl, 0, 64);
  if (s->empty) std::memset(s->h_results, 0, rec_bytes * EV_QCAP);  // no kernel: every record is zero
  if (!s->empty) {
    EvalArgs& a = s->args;
    std::memset(&a, 0, sizeof a);
    a.xyz = cloud->d; a.plan = st->d_plan; a.ctl = st->d_ctl; a.partials = st->d_partials; a.local_rec = st->d_local_rec; a.out = s->d_results;
    a.d_cmds = st->d_cmds;
    void* dp = nullptr;
    if ((e = cudaHostGetDevicePointer(&dp, s->h_cmds, 0)) != cudaSuccess) return fail(e);
    a.h_cmds = static_cast<const EvalCmd*>(dp);
    if ((e = cudaHostGetDevicePointer(&dp, s->h_ctl, 0)) != cudaSuccess) return fail(e);
    a.h_ctl = static_cast<EvalHostCtl*>(dp);
    if ((e = cudaHostGetDevicePointer(&dp, s->h_results, 0)) != cudaSuccess) return fail(e);
    a.h_results = static_cast<double*>(dp);
    if ((e = cudaHostGetDevicePointer(&dp, s->h_times, 0)) != cudaSuccess) return fail(e);
    a.h_times = static_cast<unsigned long long*>(dp);
    a.h_status = ctx->d_status;
    a.idle_timeout_ns = 20ull * 1000000000ull;
    if (s->exchange) { a.px = ctx->px; a.epoch0 = ctx->px.epoch + 1; }
    s->nblocks = st->h_plan.nblocks;
    if (const char* tf = std::getenv("HS_EVAL_TRACE")) {  // tracing: where the time of an evaluation goes, block by block
      const size_t bytes = sizeof(unsigned long long) * EV_TRACE_EVALS * s->nblocks * 4;
      if (tf[0] && cudaMalloc(&s->d_trace, bytes) == cudaSuccess) { cudaMemsetAsync(s->d_trace, 0, bytes, ctx->stream); a.trace = s->d_trace; }
    }
  }
  ctx->session = s;
  *out = s;
  return HS_OK;
}

int32_t hs_eval_session_post(hs_eval_session* s, const double* params, int32_t count) {
  if (!s) return HS_EINVAL;
  hs_ctx* ctx = s->ctx;
  HS_SLOCK(ctx);
  if (!params || count < 0) { ctx->err = "hs_eval_session_post: bad arguments"; return HS_EINVAL; }
  if (s->stopped) { ctx->err = "hs_eval_session_post: the session has been stopped"; return HS_EINVAL; }
  // the resident kernel starts before the first parameter set is converted: its start-up (~10 us) hides the conversion, and its
  // dispatcher finds the command on one of its first polls
  if (count > 0 && !s->launched && ctx->modes[HS_MODE_SESSION_LAUNCH] == 0)
    if (int32_t rc = session_launch(s)) return rc;
  for (int32_t i = 0; i < count; ++i) {
    if (!s->empty && s->posted - host_load(&s->h_ctl->done) >= static_cast<uint32_t>(EV_QCAP)) {
      // a ring entry is free again once its evaluation is done (the kernel must be running for that)
      if (int32_t rc = session_launch(s)) return rc;
      if (int32_t rc = session_spin(s, "hs_eval_session_post", [&] { return s->posted - host_load(&s->h_ctl->done) < static_cast<uint32_t>(EV_QCAP); })) return rc;
    }
    EvalCmd& c = s->h_cmds[s->posted % EV_QCAP];
    for (int r = 0; r < s->nrooms; ++r) {
      float pl[24];
      hs::planes_from_cuboid(params + (static_cast<size_t>(i) * s->nrooms + r) * 10, pl);
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k)
          if (!(pl[8 * j + k] == -pl[8 * j + 4 + k])) { ctx->err = "hs_eval_session_post: parameters do not give antiparallel wall pairs (NaN?)"; return HS_EINVAL; }
      fill_cmd(c, r, pl, s->posted);
    }
    ++s->posted;
    if (s->empty) host_store(&s->h_ctl->done, s->posted);  // zero records, already in h_results
    host_store(&s->h_ctl->posted, s->posted);
  }
  return HS_OK;
}

int64_t hs_eval_session_done(const hs_eval_session* s) { return s ? static_cast<int64_t>(host_load(&s->h_ctl->done)) : -1; }

int32_t hs_eval_session_wait(hs_eval_session* s, int64_t seq, double* rec_out) {
  if (!s) return HS_EINVAL;
  hs_ctx* ctx = s->ctx;
  HS_SLOCK(ctx);
  if (seq < 0 || seq >= static_cast<int64_t>(s->posted)) { ctx->err = "hs_eval_session_wait: evaluation has not been posted"; return HS_EINVAL; }
  if (static_cast<int64_t>(s->posted) - seq > EV_QCAP) { ctx->err = "hs_eval_session_wait: record no longer in the result ring (256 evaluations)"; return HS_EINVAL; }
  if (int32_t rc = session_launch(s)) return rc;  // deferred-launch mode: the kernel starts when the host first needs a result
  if (int32_t rc = session_spin(s, "hs_eval_session_wait", [&] { return static_cast<int64_t>(host_load(&s->h_ctl->done)) > seq; })) return rc;
  if (int32_t rc = session_check(s, "hs_eval_session_wait")) return rc;
  if (rec_out) std::memcpy(rec_out, s->h_results + static_cast<size_t>(seq % EV_QCAP) * s->nrooms * HS_REC, sizeof(double) * s->nrooms * HS_REC);
  return HS_OK;
}

int32_t hs_eval_session_eval(hs_eval_session* s, const double* params, double* rec_out) {
  if (!s) return HS_EINVAL;
  if (int32_t rc = hs_eval_session_post(s, params, 1)) return rc;
  return hs_eval_session_wait(s, static_cast<int64_t>(s->posted) - 1, rec_out);
}

int32_t hs_eval_session_stop(hs_eval_session* s) {
  if (!s) return HS_EINVAL;
  hs_ctx* ctx = s->ctx;
  HS_SLOCK(ctx);
  s->stopped = true;
  host_store(&s->h_ctl->stop, 1u);
  if (s->posted > 0)  // deferred-launch mode: the kernel starts now, with every command and the stop flag in place
    if (int32_t rc = session_launch(s)) return rc;
  return HS_OK;
}

int32_t hs_eval_session_times(const hs_eval_session* s, int64_t seq, uint64_t* seen_ns, uint64_t* done_ns) {
  if (!s || seq < 0 || seq >= static_cast<int64_t>(s->posted) || static_cast<int64_t>(s->posted) - seq > EV_QCAP) return HS_EINVAL;
  if (static_cast<int64_t>(host_load(&s->h_ctl->done)) <= seq && !s->empty) return HS_EINVAL;
  if (seen_ns) *seen_ns = s->h_times[2 * (seq % EV_QCAP)];
  if (done_ns) *done_ns = s->h_times[2 * (seq % EV_QCAP) + 1];
  return HS_OK;
}

void* hs_eval_session_device_results(const hs_eval_session* s) { return s ? s->d_results : nullptr; }

int32_t hs_eval_session_end(hs_eval_session* s) {
  if (!s) return HS_OK;
  hs_ctx* ctx = s->ctx;
  HS_SLOCK(ctx);
  s->stopped = true;
  host_store(&s->h_ctl->stop, 1u);
  int32_t rc = s->posted > 0 ? session_launch(s) : HS_OK;
  const cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ctx->err = std::string("hs_eval_session_end: ") + cudaGetErrorString(e); rc = HS_ECUDA; }
  if (rc == HS_OK) rc = session_check(s, "hs_eval_session_end");
  if (rc == HS_OK && !s->empty && host_load(&s->h_ctl->done) != s->posted) { ctx->err = "hs_eval_session_end: the kernel ended before all posted evaluations were done"; rc = HS_ECUDA; }
  if (s->d_trace) {
    const size_t nwords = static_cast<size_t>(EV_TRACE_EVALS) * s->nblocks * 4;
    std::vector<unsigned long long> host(nwords);
    if (cudaMemcpy(host.data(), s->d_trace, nwords * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess)
      if (FILE* fp = std::fopen(std::getenv("HS_EVAL_TRACE"), "wb")) {
        const int32_t hdr[4] = {EV_TRACE_EVALS, s->nblocks, 4, static_cast<int32_t>(s->posted)};
        std::fwrite(hdr, sizeof hdr, 1, fp);
        std::fwrite(host.data(), sizeof(unsigned long long), nwords, fp);
        std::fclose(fp);
      }
    cudaFree(s->d_trace);
  }
  if (s->launched) cudaMemsetAsync(ctx->eval->d_ctl, 0, sizeof(EvalCtl), ctx->stream);  // counters back to zero for the next session / launch
  if (s->exchange) ctx->px.epoch += s->posted;  // every rank posted the same evaluations
  if (ctx->h_status) *ctx->h_status = 0;  // reported through this call
  ctx->session = nullptr;
  session_free(s);
  return rc;
}

}  // extern "C"
