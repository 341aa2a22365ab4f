// comm.cuh -- shared definitions of the NVSwitch peer-memory communicator
// (control block layout, flags, hbComm, phased submission) used by comm.cu and
// sharded.cu.
#pragma once
#include <functional>

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
  uint32_t data_flags[2][kMaxWorld];                // per window half (alltoallv payload, allreduce)
  uint32_t plan_flags[8][kMaxWorld];                // sharded-plan phases (sharded.cu)
  uint32_t done_counter[8];                         // last-CTA-done counters
  int32_t mailbox[2][kMaxWorld * kMaxA2aTensors * kMaxWorld];  // [parity][q][k][r]
  int32_t plan_mailbox[2][kMaxWorld * kMaxA2aTensors * kMaxWorld];  // sharded plan sizes
  Snapshot snap[kSnapSlots];
};

constexpr uint64_t kChunkBytes = 16384;

struct LocalGroup;  // in-process rendezvous of W communicators on one device (comm.cu)

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
  uint32_t win_seq;      // window-half users so far (alltoallv payloads + allreduces)
  uint32_t plan_epoch;   // sharded-plan steps issued on this communicator (monotonic ACROSS plans)
  int n_of_call[hb::kSnapSlots];
  int32_t* d_status;     // optional sticky status word for kernels of ops without one
  hb::LocalGroup* group; // != nullptr: member of an in-process group
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

// A collective is a fixed sequence of PHASES per rank.  Kernels of phase p only
// spin on flags that kernels of phases < p (of any rank, same or earlier op)
// publish, and a phase that publishes never waits.  A multi-process
// communicator runs the phases back to back on `stream` (the spins resolve as the
// peers' kernels run on their GPUs).  A member of an in-process group blocks
// until every rank of the group has submitted the same op; the last arrival
// then issues all ranks' phases phase-major (with event fences when the ranks
// use different streams), so that on ONE device no kernel ever spins on a flag
// whose producer has not been launched.  op_code identifies the op for the
// rendezvous sanity check.
using PhaseFn = std::function<int(int phase)>;
int comm_submit(hbComm* c, int op_code, int nphases, cudaStream_t stream, const PhaseFn& run);

enum { kOpBarrier = 1, kOpA2aSizes = 2, kOpA2aData = 3, kOpAllreduce = 4,
       kOpShardedFwd = 5, kOpShardedBwd = 6 };

const char* get_last_error();

}  // namespace hb
