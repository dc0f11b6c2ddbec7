// All-reduce of the per-room records over NVLink peer memory, done by the block that finishes a GPU's reduction.
//
// The path's only exchange is nrooms x 24 doubles (2.3 KB for the 12-room apartment).  Through NCCL that costs a second
// launch and ~12 us per evaluation (measured at 2 GPUs), which is as long as the 8-GPU shard's whole kernel; here the last
// block of the reduction kernel writes its record straight into a mailbox slot on every peer (posted NVLink stores), raises a
// flag there, waits for the other ranks' flags in its own mailbox and adds the slots in rank order — deterministic, and the
// same result on every rank.  Mailboxes are plain cudaMalloc memory shared between the per-GPU processes with CUDA IPC
// handles (hs_peer_mailbox_create / hs_peer_mailbox_connect).
#pragma once
#include "hs_internal.cuh"

namespace hsk {

constexpr size_t PEER_SLOT_DOUBLES = static_cast<size_t>(HS_MAX_ROOMS) * HS_REC;
constexpr size_t PEER_DATA_DOUBLES = 2 * HS_PEER_MAX * PEER_SLOT_DOUBLES;  // [epoch parity][source rank][record]
constexpr size_t PEER_FLAG_STRIDE = 32;                                     // uint32 units: one 128-byte line per flag
constexpr size_t PEER_MAILBOX_BYTES = PEER_DATA_DOUBLES * sizeof(double) + HS_PEER_MAX * PEER_FLAG_STRIDE * sizeof(uint32_t);

__device__ __forceinline__ unsigned long long peer_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// buf[0..count) of this rank -> sum over ranks, in place.  Called by ALL `nthreads` threads of one block; `sync` is a barrier
// over exactly those threads.  Two data parities suffice: nobody can send epoch e+2 before everybody has sent e+1, and a rank
// sends e+1 only after it has finished reading e.
template <class Sync>
__device__ __forceinline__ void peer_allreduce(const PeerExchange& px, double* buf, int count, int tid, int nthreads, Sync sync) {
  const int par = static_cast<int>(px.epoch & 1u);
  double* mine = reinterpret_cast<double*>(px.mailbox[px.rank]);
  volatile uint32_t* my_flags = reinterpret_cast<volatile uint32_t*>(mine + PEER_DATA_DOUBLES);
  for (int p = 0; p < px.world; ++p) {  // my record into slot [par][rank] of every mailbox (my own included)
    double* dst = reinterpret_cast<double*>(px.mailbox[p]) + (static_cast<size_t>(par) * HS_PEER_MAX + px.rank) * PEER_SLOT_DOUBLES;
    for (int i = tid; i < count; i += nthreads) dst[i] = buf[i];
  }
  sync();  // the block's stores happen-before the flag threads' system-scope fence below (cumulative release)
  if (tid < px.world) {  // thread p tells rank p that this rank's record has landed
    __threadfence_system();
    volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(reinterpret_cast<double*>(px.mailbox[tid]) + PEER_DATA_DOUBLES) + px.rank * PEER_FLAG_STRIDE;
    *f = px.epoch;
    // and waits for rank p's record in this rank's mailbox (bounded: a dead peer must not hang the GPU)
    const unsigned long long t0 = peer_now_ns();
    while (static_cast<int32_t>(my_flags[tid * PEER_FLAG_STRIDE] - px.epoch) < 0) {
      if (peer_now_ns() - t0 > 2000000000ull) { my_flags[HS_PEER_MAX * PEER_FLAG_STRIDE - 1] = px.epoch; break; }  // timeout marker
    }
    __threadfence_system();
  }
  sync();
  const bool timed_out = my_flags[HS_PEER_MAX * PEER_FLAG_STRIDE - 1] == px.epoch;
  const volatile double* slots = mine + static_cast<size_t>(par) * HS_PEER_MAX * PEER_SLOT_DOUBLES;
  for (int i = tid; i < count; i += nthreads) {
    double s = 0.0;
    for (int r = 0; r < px.world; ++r) s += slots[r * PEER_SLOT_DOUBLES + i];  // rank order: identical on every rank
    buf[i] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : s;
  }
}

}  // namespace hsk
