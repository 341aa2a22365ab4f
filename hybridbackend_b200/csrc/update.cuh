// update.cuh -- internal interface of the sort -> runs -> apply machinery
// (sparse_update.cu) shared with the sharded path (sharded.cu).
#pragma once
#include "comm.cuh"

namespace hb {

// What differs from the plain single-rank call per feature.
struct UpdExtra {
  // sort input
  const uint32_t* keys32 = nullptr;  // key_kind 2: uint32 keys (local rows) instead of int64 ids
  const int32_t* n_dev = nullptr;    // device-side number of entries (<= the static nnz)
  int32_t key_kind = 0;              // 0: id / id_div   1: (id % W) << lbits | id / W   2: keys32
  int32_t lbits = 0;                 // key_kind 1
  // run detection (requester side of the sharded path)
  int32_t* inv = nullptr;            // [nnz] unique index of every input position (-1: invalid id)
  int32_t* owner_start1 = nullptr;   // [W]   1 + index of the first unique of each owner (0: none)
  // emit sink: the per-unique gradient sum goes to the owner's window instead of
  // into a table (requester side of the sharded backward)
  const int32_t* emit_send_off = nullptr;     // [W+1] bucket starts in my unique list
  const int32_t* emit_remote_base = nullptr;  // [W]   my segment start in owner r's window
  uint64_t emit_off = 0;                      // byte offset of the feature's grads_in region
  int32_t emit_cap = 0;                       // rows of that region
};

struct EmitCtx {       // peers of the emit sink
  PeerPtrs peers;
  uint64_t window_off;
  int32_t world;
};

enum { kPhaseSort = 1, kPhaseApply = 2 };

// Per-feature device outputs of the sort phase that other kernels read.
struct UpdViews {
  const uint32_t* ukey;    // [U] unique keys, ascending
  const int32_t* ustart;   // [U+1] first sorted entry of each unique
  const int32_t* counts;   // [0] = U, [1] = number of valid entries
};

// phases: kPhaseSort = bag map + radix sort + run detection; kPhaseApply = the
// fused duplicate-sum + sink (optimizer apply or emit).  `extras` may be nullptr
// (plain call) or an array of n.  `views` (optional, n entries) receives the
// device pointers of the run arrays inside the workspace.
int sparse_update_run(int n, const hbUpdateFeature* feats, const hbOptimizer* opt, void* ws,
                      size_t ws_bytes, int32_t* d_status, cudaStream_t stream, const WaitSpec* wait,
                      const UpdExtra* extras, const EmitCtx* emit, int phases, UpdViews* views);
size_t sparse_update_workspace_bytes(int n, const hbUpdateFeature* feats);

}  // namespace hb
