// comm.cuh -- shared definitions of the NVSwitch peer-memory communicator
// (control block layout, flags, hbComm) used by comm.cu and sharded.cu.
#pragma once
#include "common.cuh"

namespace hb {

constexpr int kMaxWorld = 16;
constexpr int kMaxA2aTensors = 256;
constexpr int kSnapSlots = 4;
constexpr uint32_t kTokenMagic = 0x48423230u;  // "HB20"

// ---- control block (device memory, at the start of the symmetric allocation) --
struct SegEntry {      // one (tensor k, peer) segment
  uint64_t src_off;    // byte offset in the local input k   (push) / local window (copy-out)
  uint64_t dst_off;    // byte offset in the peer window     (push) / output k     (copy-out)
  uint64_t bytes;
  uint64_t chunk_begin;  // first chunk index of this segment
};
struct Snapshot {
  int32_t matrix[kMaxWorld * kMaxA2aTensors * kMaxWorld];  // S[q][k][r]
  SegEntry push[kMaxA2aTensors * kMaxWorld];               // [k][r]
  SegEntry pull[kMaxA2aTensors * kMaxWorld];               // [k][q] copy-out
  uint64_t push_chunks, pull_chunks;
  uint64_t overflow;
  int32_t recv_sizes[kMaxA2aTensors * kMaxWorld];          // [k][q]
};
struct Control {
  uint32_t barrier_flags[kMaxWorld];                // epoch of last barrier seen from q
  uint32_t sizes_flags[2][kMaxWorld];               // per mailbox parity
  uint32_t data_flags[2][kMaxWorld];                // per window half
  uint32_t plan_flags[8][kMaxWorld];                // sharded-plan phases (sharded.cu)
  uint32_t done_counter[8];                         // last-CTA-done counters
  int32_t mailbox[2][kMaxWorld * kMaxA2aTensors * kMaxWorld];  // [parity][q][k][r]
  int32_t plan_mailbox[2][kMaxWorld * kMaxA2aTensors * kMaxWorld];  // sharded plan sizes
  Snapshot snap[kSnapSlots];
};

constexpr uint64_t kChunkBytes = 16384;

}  // namespace hb

struct hbComm {
  int rank, world, local;
  size_t window_bytes;   // data window size
  size_t reserved_bytes; // leading part of the window owned by a sharded plan
  size_t alloc_bytes;
  unsigned char* base;   // local allocation
  unsigned char* peer[hb::kMaxWorld];  // mapped peer allocations (own = base)
  bool opened[hb::kMaxWorld];
  bool connected;
  uint32_t barrier_epoch;
  uint32_t sizes_calls;  // number of hbAlltoallvNSizes issued
  uint32_t data_calls;   // number of hbAlltoallvN issued
  int n_of_call[hb::kSnapSlots];
  cudaIpcMemHandle_t handle;
};

namespace hb {

static inline size_t control_bytes() { return align_up(sizeof(Control), 4096); }

struct PeerPtrs {
  unsigned char* p[kMaxWorld];
};

static inline PeerPtrs peer_ptrs(const hbComm* c) {
  PeerPtrs pp;
  for (int i = 0; i < kMaxWorld; ++i) pp.p[i] = i < c->world ? c->peer[i] : nullptr;
  return pp;
}

// spin until flag (epoch-valued, monotonically increasing) reaches `epoch`
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch) {
  return spin_until(flag, epoch);
}

}  // namespace hb
