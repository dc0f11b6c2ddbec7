// Internal declarations shared by the CUDA translation units of libhousescan_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>

#include "../../include/housescan_b200.h"

#define HS_MAX_ROOMS 32   // rooms per launch (kernel-parameter table); larger calls are chunked
#define HS_MAX_PLANES 8   // generic planes per room
#define HS_TPB 256        // threads per block of the streaming kernels
#define HS_PEER_MAX 8     // ranks in a peer-memory all-reduce group (one NVSwitch domain)
#define HS_NACC 22        // accumulators of a cuboid-sums record actually reduced (f, Sr[6], B[9], cnt[6])

// peer-memory all-reduce of the record (k_peer.cuh); world <= 1: no exchange
struct PeerExchange {
  unsigned long long mailbox[HS_PEER_MAX];  // device address of every rank's mailbox; [rank] is the local one
  int32_t rank, world;
  uint32_t epoch;  // host side: evaluations exchanged so far (the next one uses epoch + 1); identical on all ranks
  uint32_t pad;
};

struct hs_eval_state;    // k_eval.cu: plan cache, control block, partial-record ring of the evaluation kernel
struct hs_eval_session;  // k_eval.cu: a resident multi-evaluation kernel

struct hs_cloud {
  float* d = nullptr;
  int64_t n = 0;
  int64_t cap = 0;  // points the allocation can hold
  bool owned = true;
};

struct hs_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;
  int64_t launches = 0;
  // scratch (device) for partial reductions, tickets, small results
  char* d_scratch = nullptr;
  size_t scratch_bytes = 0;
  unsigned int* d_ticket = nullptr;  // zero between launches
  double* d_small = nullptr;         // HS_MAX_ROOMS*HS_REC doubles + misc
  // pinned host staging
  char* h_pinned = nullptr;
  size_t pinned_bytes = 0;
  int modes[16] = {0};
  // peer mailbox (multi-GPU): local allocation, the peers' mappings, the running epoch, and whether the next reduction exchanges
  char* d_mailbox = nullptr;
  void* peer_mapped[HS_PEER_MAX] = {nullptr};
  PeerExchange px = {};
  bool peer_local = false;           // mailboxes of a same-process group: addressed directly, nothing to close
  uint32_t* h_status = nullptr;      // mapped pinned word the kernels raise (HS_ENCCL on a peer timeout); d_status is its device alias
  uint32_t* d_status = nullptr;
  // per-plane sums of several rooms: the rooms' launches go round-robin over three lanes (the ctx stream + two helpers) with their own
  // partial-record regions and tickets, so that the tail of one room's launch overlaps the start of the next room's
  cudaStream_t ps_aux[2] = {nullptr, nullptr};
  cudaEvent_t ps_ev[3] = {nullptr, nullptr, nullptr};
  size_t ps_scratch_off = 0;         // byte offset of the current lane's partial-record region inside d_scratch
  int ps_ticket_off = 0;             // word offset of the current lane's ticket inside d_ticket (0 = the shared one)
  hs_eval_state* eval = nullptr;     // evaluation kernel state (lazily created)
  hs_eval_session* session = nullptr;  // open evaluation session: it owns the stream until it ends
  std::mutex mu;
};

enum { HS_MODE_EVAL_KERNEL = 0, HS_MODE_BLOCKS_PER_SM = 1,
       HS_MODE_EVAL_SEG_COST = 2,  // evaluation kernel: cost of an extra room segment in 4-point groups for the weighted partition (0 = default)
       HS_MODE_SESSION_LAUNCH = 4, // evaluation sessions: 0 = the resident kernel starts with the first posted command (default); 1 = it starts when the host
                                   // first needs the device (wait / stop / full ring): for tools that serialise kernel launches (ncu blocks in the launch call)
       HS_MODE_PS_KERNEL = 7,   // per-plane sums: 0 = ring form (bulk-async tiles, warp-level flushes), 1 = all-Double form, 2 = direct-load Float chains
       HS_MODE_SEL_KERNEL = 8,  // k-th: 0 = 11/11/10-bit passes over compacted keys, 1 = four 8-bit passes over the cloud
       HS_MODE_FILTER_KERNEL = 12,  // order-preserving filter: 0 = count pass + scatter pass, 1 = single pass (decoupled look-back)
       HS_MODE_CC_CHUNKS = 11,  // connected components: slices of the edge list with a flatten in between (0 = default)
       HS_MODE_BP_KERNEL = 10,  // back-projection: 0 = single pass (decoupled look-back), 1 = count pass + scatter pass
       HS_MODE_NE_KERNEL = 9    // 6x6 record: 0 = throughput form (Float chains), 1 = all-Double form
};
enum { HS_EVAL_AUTO = 0, HS_EVAL_EXACT = 1, HS_EVAL_FAST = 2 };  // values of modes[HS_MODE_EVAL_KERNEL]

// table of rooms passed by value to the evaluation kernels
struct RoomTable {
  int32_t nrooms;
  int32_t paired;                       // 1: planes 2j+1 have exactly the negated normal of 2j (cuboid rooms)
  int64_t off[HS_MAX_ROOMS + 1];        // point offsets (local to the cloud)
  float pl[HS_MAX_ROOMS][6][4];         // 6 PlaneEq per room
};

struct PlaneTable {
  int32_t K;
  int32_t paired;  // K == 6 and planes 2j+1 have exactly the negated normal of 2j (set by plane_table_mark_pairs)
  float pl[16][4];
};
inline void plane_table_mark_pairs(PlaneTable& t) {
  t.paired = 0;
  if (t.K != 6) return;
  for (int j = 0; j < 3; ++j)
    for (int c = 0; c < 3; ++c)
      if (!(t.pl[2 * j][c] == -t.pl[2 * j + 1][c])) return;  // NaN never pairs
  t.paired = 1;
}

#define HS_CUDA_TRY(ctx, call)                                                                                  \
  do {                                                                                                          \
    cudaError_t e__ = (call);                                                                                   \
    if (e__ != cudaSuccess) {                                                                                   \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                         \
      return HS_ECUDA;                                                                                          \
    }                                                                                                           \
  } while (0)

// ---- launchers (each enqueues on ctx->stream and bumps ctx->launches) -------------------------------------
int32_t hs_ensure_scratch(hs_ctx* ctx, size_t bytes);
int32_t hs_ensure_pinned(hs_ctx* ctx, size_t bytes);

int32_t launch_rooms_cuboid_sums(hs_ctx* ctx, const float* xyz, int64_t n, const RoomTable& tbl, double* d_rec_out);       // exact Double products
// throughput kernel (k_eval.cu): paired planes only; exchange = sum the records over the peer group inside the same launch
int32_t launch_eval(hs_ctx* ctx, const float* xyz, int64_t n, const RoomTable& tbl, double* d_rec_out, bool exchange);
int32_t launch_peer_allreduce(hs_ctx* ctx, double* d_buf, int count);  // standalone exchange kernel (next epoch)
void hs_eval_state_free(hs_ctx* ctx);
int32_t launch_plane_assign(hs_ctx* ctx, const float* xyz, int64_t n, const PlaneTable& tbl, uint8_t* d_assign, float* d_resid);
int32_t launch_plane_sums(hs_ctx* ctx, const float* xyz, int64_t i0, int64_t i1, const PlaneTable& tbl, double* d_out /*K*HS_PS*/);

int32_t launch_affine(hs_ctx* ctx, const float* in, float* out, int64_t n, const float R[9], const float pre[3], const float post[3], int kind);
int32_t launch_mean(hs_ctx* ctx, const float* xyz, int64_t n, double* d_sum3);
int32_t launch_max_nsq(hs_ctx* ctx, const float* xyz, int64_t n, const float m[3], unsigned int* d_maxbits);
int32_t launch_scatter(hs_ctx* ctx, const float* xyz, int64_t n, const float m[3], double* d_sc6);

int32_t launch_backproject(hs_ctx* ctx, const uint16_t* d_depth, int32_t w, int32_t h, float* d_xyz, uint8_t* d_mask, int64_t* d_nvalid);
size_t reduce6x6_work_bytes(const hs_ctx* ctx, int64_t nframes, int32_t w, int32_t h);  // d_work of launch_reduce6x6
int32_t launch_reduce6x6(hs_ctx* ctx, const uint16_t* d_frames, int64_t nframes, int32_t w, int32_t h, const float* intr,
                         const float* d_poses, const PlaneTable& tbl, double* d_out, char* d_work);

int32_t launch_pcd_unpack(hs_ctx* ctx, const uint8_t* d_raw, int64_t n, const int64_t off[6], const int64_t stride[6], int rgb_bytes, float* d_xyz, float* d_rgbf);

int32_t launch_cc(hs_ctx* ctx, const uint32_t* d_src, const uint32_t* d_dst, int64_t E, uint32_t N, uint32_t* d_label);

int32_t launch_kth_shard_pass(hs_ctx* ctx, const float* xyz, int64_t n, int axis, int pass, uint32_t prefix, uint32_t mask, uint32_t* d_hist_out /* 2048 */);
int32_t launch_kth(hs_ctx* ctx, const float* xyz, int64_t n, int axis, int64_t k, bool largest, float* d_out);
int32_t launch_filter_le(hs_ctx* ctx, const float* xyz, int64_t n, int axis, float limit, const float* extra_in, float* out,
                         float* extra_out, int64_t* d_nout);
