// All-reduce of the per-room records over NVLink peer memory, done by ONE WARP of the GPU that finishes a reduction.
//
// The path's only exchange is nrooms x 24 doubles (2.3 KB for the 12-room apartment).  Through NCCL that costs a second
// launch and ~12 us per evaluation (measured at 2 GPUs), which is as long as the 8-GPU shard's whole kernel; here a warp writes
// the rank's record straight into a mailbox slot on every peer (posted NVLink stores), raises a flag there, waits for the
// other ranks' flags in its own mailbox and adds the slots in rank order — deterministic, and the same result on every rank.
// Mailboxes are plain cudaMalloc memory: shared between per-GPU processes with CUDA IPC handles (hs_peer_mailbox_create /
// hs_peer_mailbox_connect), or addressed directly when all ranks live in one process (hs_peer_group_create_local).
//
// Evaluations are numbered by an epoch that every rank counts identically; epoch e uses slot e % PEER_SLOTS.  Slots are
// re-used only after every rank has consumed them: an evaluation session keeps at most EV_D (= PEER_SLOTS / 2) evaluations in
// flight per rank, and a rank finalises evaluations in order (k_eval.cuh), so nobody can be PEER_SLOTS epochs ahead of a reader.
#pragma once
#include "hs_internal.cuh"

namespace hsk {

constexpr int PEER_SLOTS = 16;
constexpr size_t PEER_SLOT_DOUBLES = static_cast<size_t>(HS_MAX_ROOMS) * HS_REC;
constexpr size_t PEER_DATA_DOUBLES = static_cast<size_t>(PEER_SLOTS) * HS_PEER_MAX * PEER_SLOT_DOUBLES;  // [slot][source rank][record]
constexpr size_t PEER_FLAG_STRIDE = 32;                                                                  // uint32 units: one 128-byte line per flag
constexpr size_t PEER_MAILBOX_BYTES = PEER_DATA_DOUBLES * sizeof(double) + static_cast<size_t>(PEER_SLOTS) * HS_PEER_MAX * PEER_FLAG_STRIDE * sizeof(uint32_t);
constexpr unsigned long long PEER_TIMEOUT_NS = 2000000000ull;  // a dead peer must not hang the GPU

__device__ __forceinline__ unsigned long long peer_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ double* peer_data(unsigned long long mailbox, uint32_t slot, int src) {
  return reinterpret_cast<double*>(mailbox) + (static_cast<size_t>(slot) * HS_PEER_MAX + src) * PEER_SLOT_DOUBLES;
}
__device__ __forceinline__ volatile uint32_t* peer_flag(unsigned long long mailbox, uint32_t slot, int src) {
  return reinterpret_cast<volatile uint32_t*>(reinterpret_cast<double*>(mailbox) + PEER_DATA_DOUBLES) + (static_cast<size_t>(slot) * HS_PEER_MAX + src) * PEER_FLAG_STRIDE;
}

constexpr int PEER_VPL = (static_cast<int>(PEER_SLOT_DOUBLES) + 31) / 32;  // record values per lane of the exchanging warp (24 for 32 rooms)

// Step 1 (never blocks): this rank's record src[0..count) -> slot [epoch][rank] of every mailbox, then the flags.  One full warp.
// The record is read ONCE into registers (all loads in flight together), then stored to every peer: posted NVLink stores, nothing
// waits until the single system-scope fence in front of the flags.
__device__ __forceinline__ void peer_push_warp(const PeerExchange& px, uint32_t epoch, const double* src, int count) {
  const int lane = static_cast<int>(threadIdx.x & 31);
  const uint32_t slot = epoch % PEER_SLOTS;
  double v[PEER_VPL];
#pragma unroll
  for (int k = 0; k < PEER_VPL; ++k) v[k] = (lane + 32 * k < count) ? __ldcg(src + lane + 32 * k) : 0.0;
  for (int p = 0; p < px.world; ++p) {
    double* dst = peer_data(px.mailbox[p], slot, px.rank);
#pragma unroll
    for (int k = 0; k < PEER_VPL; ++k)
      if (lane + 32 * k < count) dst[lane + 32 * k] = v[k];
  }
  __threadfence_system();  // every lane's record stores are performed system-wide ...
  __syncwarp();            // ... before any lane raises a flag
  if (lane < px.world) *peer_flag(px.mailbox[lane], slot, px.rank) = epoch;  // lane p tells rank p that this rank's record has landed
}

// Step 2: wait for every rank's record of this epoch in the local mailbox and add them in rank order into dst[0..count).
// Returns false (and fills dst with NaN) when a peer did not show up within PEER_TIMEOUT_NS.  One full warp.  The HS_PEER_MAX
// loads of a value are issued together (the slots of absent ranks are not read), then added in rank order: identical on every rank.
__device__ __forceinline__ bool peer_collect_warp(const PeerExchange& px, uint32_t epoch, double* dst, int count) {
  const int lane = static_cast<int>(threadIdx.x & 31);
  const uint32_t slot = epoch % PEER_SLOTS;
  bool ok = true;
  if (lane < px.world) {
    volatile uint32_t* f = peer_flag(px.mailbox[px.rank], slot, lane);
    const unsigned long long t0 = peer_now_ns();
    while (static_cast<int32_t>(*f - epoch) < 0) {  // long naps: this warp shares its scheduler with warps that stream the cloud
      if (peer_now_ns() - t0 > PEER_TIMEOUT_NS) { ok = false; break; }
      __nanosleep(250);
    }
    __threadfence_system();
  }
  ok = __all_sync(0xffffffffu, ok);
  const double* base = peer_data(px.mailbox[px.rank], slot, 0);
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  for (int i = lane; i < count; i += 64) {  // two values per round: 2 x HS_PEER_MAX loads in flight
    double v[2][HS_PEER_MAX];
    const bool second = i + 32 < count;
#pragma unroll
    for (int r = 0; r < HS_PEER_MAX; ++r) {  // the peers wrote these lines over NVLink: read them from L2, never from a stale L1 line
      v[0][r] = (r < px.world) ? __ldcv(base + static_cast<size_t>(r) * PEER_SLOT_DOUBLES + i) : 0.0;
      v[1][r] = (r < px.world && second) ? __ldcv(base + static_cast<size_t>(r) * PEER_SLOT_DOUBLES + i + 32) : 0.0;
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int r = 0; r < HS_PEER_MAX; ++r)
      if (r < px.world) { s0 += v[0][r]; s1 += v[1][r]; }  // rank order
    dst[i] = ok ? s0 : nan;
    if (second) dst[i + 32] = ok ? s1 : nan;
  }
  return ok;
}

}  // namespace hsk
