// partition.cu -- K1: HbPartitionByModulo[N] / HbPartitionByDualModuloStage{One,Two}[N]
// on the stable bucket machinery of bucket.cuh.  C-ABI in include/hb_b200.h.
#include "bucket.cuh"

namespace hb {

template <typename Tr>
static int run_partition_typed(int n, const void* const* d_inputs, const int32_t* lens,
                               int32_t P, int32_t M, void* const* d_outputs,
                               int32_t* const* d_sizes, int32_t* const* d_indices,
                               uint32_t* scratch, size_t scratch_words, cudaStream_t stream) {
  // one memset zeroes histograms, status words and tickets of every chunk
  HB_CUDA_OK(cudaMemsetAsync(scratch, 0, scratch_words * sizeof(uint32_t), stream));
  size_t off = 0;
  for (int c0 = 0; c0 < n; c0 += kMaxSegs) {
    BucketParams bp;
    const int nc = (n - c0 < kMaxSegs) ? n - c0 : kMaxSegs;
    int tiles = 0;
    for (int k = 0; k < nc; ++k) {
      BucketSeg& s = bp.seg[k];
      s.in_keys = d_inputs[c0 + k];
      s.in_vals = nullptr;
      s.out_keys = d_outputs[c0 + k];
      s.out_vals = nullptr;
      s.out_inv = d_indices[c0 + k];
      s.out_sizes = d_sizes[c0 + k];
      s.n = lens[c0 + k];
      s.n_dev = nullptr;
      s.tile_begin = tiles;
      s.shift = 0;
      s.key_limit = 0;
      s.hist_slot = k;
      s.passes = 1;
      tiles += bucket_tiles(s.n);
    }
    BucketScratch sc = bucket_scratch_carve(scratch + off, nc, (size_t)tiles, P, 1);
    bp.hist = sc.hist;
    bp.status = sc.status[0];
    bp.ticket = sc.ticket[0];
    bp.nsegs = nc;
    bp.nbins = P;
    bp.total_tiles = tiles;
    bp.npass = 1;
    bp.pass = 0;
    bp.digit_bits = 0;
    bp.p = P;
    bp.m = M;
    bp.pow2_mask = ((P & (P - 1)) == 0) ? P - 1 : -1;
    bp.div = 1;
    bp.div_shift = 0;
    int rc = bucket_hist_launch<Tr>(bp, stream, HB_K_PART_HIST);
    if (rc != HB_OK) return rc;
    rc = bucket_pass_launch<Tr>(bp, stream, HB_K_PART_PASS);
    if (rc != HB_OK) return rc;
    off += bucket_scratch_words(nc, (size_t)tiles, P, 1);
  }
  return HB_OK;
}

template <typename T>
static int run_partition(int stage, int n, const void* const* d_inputs, const int32_t* lens,
                         int32_t P, int32_t M, void* const* d_outputs, int32_t* const* d_sizes,
                         int32_t* const* d_indices, uint32_t* scratch, size_t scratch_words,
                         cudaStream_t stream) {
  switch (stage) {
    case 0: return run_partition_typed<ModuloTraits<T>>(n, d_inputs, lens, P, M, d_outputs, d_sizes, d_indices, scratch, scratch_words, stream);
    case 1: return run_partition_typed<DualModuloTraits<T, 1>>(n, d_inputs, lens, P, M, d_outputs, d_sizes, d_indices, scratch, scratch_words, stream);
    default: return run_partition_typed<DualModuloTraits<T, 2>>(n, d_inputs, lens, P, M, d_outputs, d_sizes, d_indices, scratch, scratch_words, stream);
  }
}

static int partition_entry(int dtype, int stage, int n, const void* const* d_inputs,
                           const int32_t* lens, int32_t P, int32_t M, void* const* d_outputs,
                           int32_t* const* d_sizes, int32_t* const* d_indices, void* ws,
                           size_t ws_bytes, cudaStream_t stream) {
  HB_REQUIRE(n >= 1, "partition: N must be >= 1 (got %d)", n);
  HB_REQUIRE(P >= 1, "partition: num_partitions must be >= 1 (got %d)", P);
  HB_REQUIRE(P <= kMaxBins, "partition: num_partitions %d exceeds %d", P, kMaxBins);
  HB_REQUIRE(stage >= 0 && stage <= 2, "partition: bad stage %d", stage);
  HB_REQUIRE(stage == 0 || M >= 1, "partition: modulus must be >= 1 (got %d)", M);
  HB_REQUIRE(stage == 0 || (int64_t)P * M <= INT32_MAX, "partition: num_partitions*modulus overflows int32");
  HB_REQUIRE(d_inputs && lens && d_outputs && d_sizes && d_indices, "partition: null argument");
  HB_REQUIRE(dtype == HB_I32 || dtype == HB_I64 || dtype == HB_U32 || dtype == HB_U64,
             "partition: unsupported dtype %d (int32/int64/uint32/uint64 only)", dtype);
  for (int k = 0; k < n; ++k) HB_REQUIRE(lens[k] >= 0, "partition: negative length for input %d", k);
  size_t need = 0;
  int rc = hbPartitionWorkspaceBytes(n, lens, P, &need);
  if (rc != HB_OK) return rc;
  if (ws_bytes < need || (need > 0 && ws == nullptr)) {
    set_last_error("partition: workspace %zu < required %zu bytes", ws_bytes, need);
    return HB_ERR_WORKSPACE;
  }
  uint32_t* scratch = reinterpret_cast<uint32_t*>(ws);
  const size_t scratch_words = need / sizeof(uint32_t);
  switch (dtype) {
    case HB_I32: return run_partition<int32_t>(stage, n, d_inputs, lens, P, M, d_outputs, d_sizes, d_indices, scratch, scratch_words, stream);
    case HB_I64: return run_partition<int64_t>(stage, n, d_inputs, lens, P, M, d_outputs, d_sizes, d_indices, scratch, scratch_words, stream);
    case HB_U32: return run_partition<uint32_t>(stage, n, d_inputs, lens, P, M, d_outputs, d_sizes, d_indices, scratch, scratch_words, stream);
    case HB_U64: return run_partition<uint64_t>(stage, n, d_inputs, lens, P, M, d_outputs, d_sizes, d_indices, scratch, scratch_words, stream);
  }
  set_last_error("partition: unsupported dtype %d (int32/int64/uint32/uint64 only)", dtype);
  return HB_ERR_INVALID;
}

}  // namespace hb

extern "C" {

int hbPartitionWorkspaceBytes(int n, const int32_t* lens, int32_t num_partitions, size_t* bytes) {
  HB_REQUIRE(n >= 1 && lens && bytes && num_partitions >= 1, "hbPartitionWorkspaceBytes: bad argument");
  size_t words = 0;
  for (int c0 = 0; c0 < n; c0 += hb::kMaxSegs) {
    const int nc = (n - c0 < hb::kMaxSegs) ? n - c0 : hb::kMaxSegs;
    size_t tiles = 0;
    for (int k = c0; k < c0 + nc; ++k) {
      HB_REQUIRE(lens[k] >= 0, "hbPartitionWorkspaceBytes: negative length");
      tiles += hb::bucket_tiles(lens[k]);
    }
    words += hb::bucket_scratch_words(nc, tiles, num_partitions, 1);
  }
  *bytes = hb::align_up(words * sizeof(uint32_t), 256);
  return HB_OK;
}

int hbPartitionByModuloN(int dtype, int n, const void* const* d_inputs, const int32_t* lens,
                         int32_t num_partitions, void* const* d_outputs, int32_t* const* d_sizes,
                         int32_t* const* d_indices, void* d_workspace, size_t workspace_bytes,
                         hbStream stream) {
  return hb::partition_entry(dtype, 0, n, d_inputs, lens, num_partitions, 1, d_outputs, d_sizes,
                             d_indices, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int hbPartitionByDualModuloN(int dtype, int stage, int n, const void* const* d_inputs,
                             const int32_t* lens, int32_t num_partitions, int32_t modulus,
                             void* const* d_outputs, int32_t* const* d_sizes,
                             int32_t* const* d_indices, void* d_workspace, size_t workspace_bytes,
                             hbStream stream) {
  HB_REQUIRE(stage == 1 || stage == 2, "hbPartitionByDualModuloN: stage must be 1 or 2");
  return hb::partition_entry(dtype, stage, n, d_inputs, lens, num_partitions, modulus, d_outputs,
                             d_sizes, d_indices, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
