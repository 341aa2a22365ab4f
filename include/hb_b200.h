/* hb_b200.h -- C-ABI of libhb_b200.so: the B200-native sharded-embedding hot path
 * behind HybridBackend's operator surface.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / a CUDA stream
 * (passed as void* so the header needs no CUDA include), enqueues work on that
 * stream and returns an int status (HB_OK = 0).  No per-step entry point allocates an
 * output, synchronises the device, or throws (the Create / Destroy / Connect calls of
 * communicators and plans do allocate and synchronise: setup time); scratch comes from caller-provided
 * workspaces sized by the matching *WorkspaceBytes query (the TF shim uses
 * allocate_temp, the Python harness torch.empty).  Host arrays of device pointers
 * are read before the call returns.  Callable from any thread.
 *
 * "Replaces" comments cite the reference interface (paths relative to
 * /root/reference/hybridbackend/) that a TF-1.15 OpKernel shim maps onto each
 * call -- see INTEGRATION.md for the binding code.
 */
#ifndef HB_B200_H_
#define HB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_ERR_INVALID 1   /* bad argument (message in hbGetLastErrorString) */
#define HB_ERR_CUDA 2      /* CUDA runtime error */
#define HB_ERR_WORKSPACE 3 /* workspace too small */
#define HB_ERR_COMM 4      /* communicator / peer-mapping error */

typedef void* hbStream; /* cudaStream_t */

/* element types (partition ids: the first four, as
 * tensorflow/distribute/partition/partition_by_modulo_ops.cc:52) */
enum { HB_I32 = 0, HB_I64 = 1, HB_U32 = 2, HB_U64 = 3, HB_F32 = 4, HB_F16 = 5,
       HB_I8 = 6, HB_U8 = 7, HB_F64 = 8 };
/* embedding_lookup_sparse combiners */
enum { HB_SUM = 0, HB_MEAN = 1, HB_SQRTN = 2 };
/* sparse optimizers */
enum { HB_OPT_SGD = 0, HB_OPT_ADAGRAD = 1, HB_OPT_LAZY_ADAM = 2 };
/* bits of the sticky device status word */
enum { HB_STATUS_ID_OUT_OF_RANGE = 1, HB_STATUS_WINDOW_OVERFLOW = 2,
       HB_STATUS_BAD_OFFSETS = 4, HB_STATUS_PEER_TIMEOUT = 8 };

/* kernel ids for the launch counter / event profiler */
enum { HB_K_PART_HIST = 1, HB_K_PART_PASS = 2, HB_K_RESERVED3 = 3,
       HB_K_SORT_HIST = 4, HB_K_SORT_PASS = 5, HB_K_RESERVED6 = 6,
       HB_K_BAG_MAP = 7, HB_K_LOOKUP_FWD = 8, HB_K_SPARSE_UPDATE = 9,
       HB_K_SPARSE_FIXUP = 10, HB_K_CAST = 11, HB_K_CACHE_LOOKUP = 12,
       HB_K_BARRIER = 13, HB_K_A2A_SIZES = 14, HB_K_A2A_TABLES = 15,
       HB_K_A2A_PUSH = 16, HB_K_A2A_COPYOUT = 17, HB_K_SH_EXCHANGE = 18,
       HB_K_SH_PUSH_IDS = 19, HB_K_SH_OWNER_GATHER = 20, HB_K_SH_STITCH = 21,
       HB_K_SH_PUSH_GRADS = 22, HB_K_SH_PAD = 23, HB_K_AR_PUSH = 24,
       HB_K_AR_REDUCE = 25, HB_K_RUNS = 26, HB_K_UPDATE_LONG = 27,
       HB_K_SH_PUBLISH = 28, HB_K_SH_UNIQUE = 29, HB_K_H2D_STAGE = 30,
       HB_K_COUNT = 32 };

const char* hbGetLastErrorString(void);
/* number of kernels this library has launched in this process */
int64_t hbGetLaunchCount(void);
/* per-kernel CUDA-event timing on the launching stream (off by default) */
int hbProfileEnable(int on);
int hbProfileReset(void);
/* synchronises the recorded events; total device ms and launches of kernel_id */
int hbProfileGet(int kernel_id, double* total_ms, int64_t* launches);
const char* hbKernelName(int kernel_id);
int hbGetVersion(void);
/* "sm_100a" etc: the architectures compiled into this library. */
const char* hbGetBuildInfo(void);

/* ---------------------------------------------------------------------------
 * K1  Stable multi-input partition.
 * Replaces HbPartitionByModulo / HbPartitionByModuloN
 *   (tensorflow/distribute/partition/partition_by_modulo_ops.cc:46-207, GPU
 *   functors partition_by_modulo_functors.cu.cc:44-350) and, with stage != 0,
 *   HbPartitionByDualModuloStage{One,Two}[N]
 *   (partition_by_dual_modulo_ops.cc:46-324, ...functors.cu.cc:44-421).
 * For each input k (device vector d_inputs[k] of lens[k] ids of `dtype`):
 *   d_outputs[k][.] ids grouped by shard, STABLE (bit-identical to the
 *                   reference CPU functor, partition_by_modulo_functors.cc:48-69),
 *   d_sizes[k][num_partitions] int32 bucket sizes,
 *   d_indices[k][i] int32 position of input i in d_outputs[k].
 * stage: 0 = modulo, 1 / 2 = dual modulo stage one / two (uses `modulus`).
 * ------------------------------------------------------------------------- */
int hbPartitionWorkspaceBytes(int n, const int32_t* lens, int32_t num_partitions,
                              size_t* bytes);
int hbPartitionByModuloN(int dtype, int n, const void* const* d_inputs,
                         const int32_t* lens, int32_t num_partitions,
                         void* const* d_outputs, int32_t* const* d_sizes,
                         int32_t* const* d_indices, void* d_workspace,
                         size_t workspace_bytes, hbStream stream);
int hbPartitionByDualModuloN(int dtype, int stage, int n,
                             const void* const* d_inputs, const int32_t* lens,
                             int32_t num_partitions, int32_t modulus,
                             void* const* d_outputs, int32_t* const* d_sizes,
                             int32_t* const* d_indices, void* d_workspace,
                             size_t workspace_bytes, hbStream stream);

/* ---------------------------------------------------------------------------
 * K3  Fused multi-table gather + segment pooling (single rank / local shard).
 * Replaces, per feature, the TF-1.15 chain behind tf.nn.embedding_lookup_sparse
 *   Unique -> GatherV2 -> SparseSegment{Sum,Mean,SqrtN}
 *   (call sites embedding/sharding.py:171-203, docs/tutorial/ranking/data.py:189).
 * One launch covers all n features.  Bag b of feature k pools
 *   ids[offsets[b] .. offsets[b+1])  (offsets == NULL: one id per bag),
 * local row = id / id_div (id_div = world size on a row-interleaved shard,
 * embedding/sharding.py:185-186; 1 otherwise).  out row b is written at
 * out + b*out_stride (so features can share one concatenated [B, sum D] buffer).
 * dim must be a multiple of 4 and <= 1024; table/out 16-byte aligned.
 * Out-of-range ids contribute zeros and raise HB_STATUS_ID_OUT_OF_RANGE in
 * *d_status (may be NULL).
 * ------------------------------------------------------------------------- */
typedef struct hbLookupFeature {
  const float* table;     /* [rows, dim] fp32 row-major, device */
  int64_t rows;
  const int64_t* ids;     /* [nnz] device */
  const int64_t* offsets; /* [nbags+1] device, or NULL */
  int64_t nbags;
  float* out;             /* [nbags, out_stride] device */
  int64_t out_stride;     /* in floats, >= dim */
  int32_t dim;
  int32_t combiner;       /* HB_SUM / HB_MEAN / HB_SQRTN */
  int64_t id_div;         /* >= 1 */
  int64_t nnz;            /* number of ids; == nbags when offsets == NULL.  A bag
                           * reaching beyond it raises HB_STATUS_BAD_OFFSETS. */
} hbLookupFeature;

int hbGroupLookupForward(int n, const hbLookupFeature* feats, int32_t* d_status,
                         hbStream stream);

/* ---------------------------------------------------------------------------
 * K5  Backward of the pooled lookup fused with the sparse optimizer apply.
 * Replaces SparseSegment*Grad -> UnsortedSegmentSum (IndexedSlices dedup,
 *   TF optimizer._apply_sparse_duplicate_indices) -> SparseApplyAdagrad /
 *   LazyAdam row update, i.e. SURVEY.md 3.4 / 8(a) rows a12-a14; routing rule
 *   training/gradient.py:180-218 (sharded grads applied locally, no 1/W).
 * Deterministic: contributions to one row are added in ascending position
 * order (fixed tile tree for rows spanning tiles); no float atomics.
 *   grad     [nbags, grad_stride] upstream gradient of the pooled output
 *   slot0    Adagrad accumulator / Adam m ; slot1 Adam v (NULL if unused)
 * ------------------------------------------------------------------------- */
typedef struct hbUpdateFeature {
  float* table;           /* [rows, dim] updated in place */
  float* slot0;
  float* slot1;
  int64_t rows;
  const int64_t* ids;     /* [nnz] */
  const int64_t* offsets; /* [nbags+1] or NULL (one id per bag) */
  int64_t nbags;
  int64_t nnz;            /* == nbags when offsets == NULL */
  const float* grad;
  int64_t grad_stride;    /* floats */
  int32_t dim;
  int32_t combiner;
  int64_t id_div;
} hbUpdateFeature;

/* hbOptimizer.flags: HB_OPT_FLAG_FAST_MATH computes the step with the GPU's
 * approximate sqrt / divide (sqrt.approx, div.approx: <= 2 ulp each -- the class of
 * arithmetic TF's GPU kernels use, which call rsqrt) instead of the IEEE
 * sqrt-then-divide sequence of TF's CPU kernels; default 0 = bit-exact with the
 * CPU semantics. */
enum { HB_OPT_FLAG_FAST_MATH = 1 };
typedef struct hbOptimizer {
  int32_t kind;           /* HB_OPT_* */
  float lr;
  float beta1, beta2, eps; /* LazyAdam */
  int32_t flags;          /* HB_OPT_FLAG_* */
  int64_t step;           /* 1-based, LazyAdam bias correction */
} hbOptimizer;

int hbGroupSparseUpdateWorkspaceBytes(int n, const hbUpdateFeature* feats,
                                      size_t* bytes);
int hbGroupLookupBackwardUpdate(int n, const hbUpdateFeature* feats,
                                const hbOptimizer* opt, void* d_workspace,
                                size_t workspace_bytes, int32_t* d_status,
                                hbStream stream);
/* The same in two halves, so the id sort (which needs only ids / offsets) can be
 * enqueued on a side stream while the forward gather runs, and only the fused
 * duplicate-sum + optimizer apply stays on the backward's critical path.  Both
 * calls take the same feats / workspace; table, slots and grad may be NULL for
 * hbGroupSparseSort.  hbGroupLookupBackwardUpdate == Sort followed by Apply. */
int hbGroupSparseSort(int n, const hbUpdateFeature* feats, void* d_workspace,
                      size_t workspace_bytes, int32_t* d_status, hbStream stream);
int hbGroupSparseApply(int n, const hbUpdateFeature* feats, const hbOptimizer* opt,
                       void* d_workspace, size_t workspace_bytes, int32_t* d_status,
                       hbStream stream);

/* ---------------------------------------------------------------------------
 * fp32 <-> fp16 wire casts.  Replaces functor::Cast / CastN
 *   (tensorflow/common/cast.cu.cc:37-495), used when comm_wire_dtype=float16.
 * ------------------------------------------------------------------------- */
int hbCastN(int n, const void* const* d_inputs, void* const* d_outputs,
            const int64_t* counts, int from_dtype, int to_dtype, hbStream stream);

/* ---------------------------------------------------------------------------
 * Slab-hash cache probe.  Replaces HbLookup (embedding/lookup_ops.cc:38-145,
 *   lookup_functors.cu.cc:53-161).  d_hit_and_miss_keys_indices[n] and
 *   d_hit_cache_indices_and_miss_keys[n] are filled from the front with hits
 *   and from the back with misses; d_miss_count[0] receives the miss count and
 *   d_miss_count[1] the hit count (so the buffer holds TWO int32).
 * ------------------------------------------------------------------------- */
int hbCacheLookup(const int64_t* d_keys_cache, int64_t cache_slab_count,
                  const int64_t* d_keys, int32_t key_count,
                  int32_t* d_hit_and_miss_keys_indices,
                  int64_t* d_hit_cache_indices_and_miss_keys,
                  int32_t* d_miss_count, hbStream stream);

/* ---------------------------------------------------------------------------
 * Communicator: one per process/GPU.  Replaces NcclCollective + bootstrap ops
 *   (tensorflow/distribute/nccl/nccl_collective.cc:40-65,:434-465,
 *   nccl_create.cc:32-137, nccl_get_id.cc:35-70).  Instead of an NCCL id the
 *   128-byte token carries a CUDA IPC handle of this rank's symmetric window;
 *   the caller all-gathers the tokens out of band (the reference broadcasts its
 *   id over TF gRPC, distribute/collective.py:108-115) and passes all of them
 *   to hbCommConnect, which maps every peer window (NVLink P2P).
 * window_bytes: size of the symmetric window every rank allocates.
 * ------------------------------------------------------------------------- */
typedef struct hbComm hbComm;
#define HB_COMM_TOKEN_BYTES 128

int hbCommCreate(int rank, int world_size, int local_size, size_t window_bytes,
                 hbComm** comm, unsigned char token_out[HB_COMM_TOKEN_BYTES]);
int hbCommConnect(hbComm* comm, const unsigned char* all_tokens /* world*128 */);
/* The reference's own bootstrap protocol: ONE 128-byte id made on rank 0 (replaces
 * HbGetNcclId, nccl_get_id.cc:35-70: `id: int64[16]`), broadcast by the caller
 * (distribute/collective.py:108-115), then hbCommCreateFromId on every rank (replaces
 * HbCreateNcclCollective(handle, id; world_size, local_size, rank), nccl_create.cc:45-62).
 * The id names a rendezvous directory on the node ($HB_B200_RENDEZVOUS_DIR, default
 * /dev/shm) through which the ranks exchange their IPC tokens; == hbCommCreate + the
 * out-of-band all-gather + hbCommConnect.  Blocks until all ranks arrived (120 s). */
int hbGetUniqueId(unsigned char id_out[HB_COMM_TOKEN_BYTES]);
int hbCommCreateFromId(const unsigned char id[HB_COMM_TOKEN_BYTES], int rank, int world_size,
                       int local_size, size_t window_bytes, hbComm** comm);
/* In-process group: world_size communicators on the CURRENT device whose peer
 * windows are each other's allocations (no IPC, no second GPU).  One host thread
 * per rank then drives comms[r] through the very same entry points and kernels
 * as one process per GPU would; every collective entry point of a group member
 * blocks until all ranks of the group have issued the same call (see
 * csrc/comm.cuh comm_submit), so ranks MUST run on separate threads.  Built so
 * that a single-GPU CI box exercises the multi-rank kernels; also usable to run
 * W logical shards on one GPU. */
int hbCommCreateLocalGroup(int world_size, size_t window_bytes, hbComm** comms /* [world] */);
int hbCommDestroy(hbComm* comm);
/* Sticky device status word (HB_STATUS_* bits) for the kernels of entry points
 * that take none (barrier, size exchange); NULL detaches it. */
int hbCommSetStatusWord(hbComm* comm, int32_t* d_status);
int hbCommRank(const hbComm* comm);
int hbCommWorldSize(const hbComm* comm);
/* device pointer of the local window (tests / zero-copy producers) */
void* hbCommWindow(hbComm* comm);
size_t hbCommWindowBytes(const hbComm* comm);
/* all ranks rendezvous on the stream (flag exchange over NVLink) */
int hbCommBarrier(hbComm* comm, hbStream stream);

/* ---------------------------------------------------------------------------
 * K2  AlltoallvN over NVSwitch peer stores (no NCCL on this path).
 * Replaces HbNcclAlltoallv / HbNcclAlltoallvN and the size pre-exchange
 *   HbNcclAlltoall[N] (tensorflow/distribute/nccl/nccl_alltoallv.cc:200-580,
 *   nccl_collective.cc:112-384).  Two-phase so the caller can allocate outputs:
 *   1. hbAlltoallvNSizes: d_send_sizes[k][W] int32 (device) -> every peer;
 *      d_recv_sizes[k][W] (device) receives recv_sizes[q] = what rank q sends
 *      me; the same values are written to h_recv_sizes (pinned host, n*W int32)
 *      once the stream reaches that point (the reference also blocks the host
 *      here: nccl_alltoallv.cc:533).
 *   2. hbAlltoallvN: d_inputs[k] = [sum send, common[k]] elements of
 *      elem_bytes[k]; d_outputs[k] = concat over source ranks (ascending) of
 *      the segment addressed to me (nccl_collective.cc:250-288).  The j-th
 *      hbAlltoallvN pairs with the j-th hbAlltoallvNSizes (sizes stay on the
 *      device in a FIFO of 4 snapshots); at most 3 size exchanges may be
 *      outstanding.  Needs window_bytes/2 >= this rank's receive volume.
 * ------------------------------------------------------------------------- */
int hbAlltoallvNSizes(hbComm* comm, int n, const int32_t* const* d_send_sizes,
                      int32_t* const* d_recv_sizes, int32_t* h_recv_sizes,
                      hbStream stream);
int hbAlltoallvN(hbComm* comm, int n, const void* const* d_inputs,
                 const int64_t* common_sizes, const int32_t* elem_bytes,
                 void* const* d_outputs, int32_t* d_status, hbStream stream);

/* ---------------------------------------------------------------------------
 * Small dense all-reduce (fp32 sum, then * scale) over the peer windows.
 * Replaces Collective.allreduce for the DENSE gradients of replicated "small"
 *   embedding tables (training/gradient.py:132-141 densify, :157-160 allreduce,
 *   :77-97 + :216 the 1/W mean -> pass scale = 1.0f / world).  Every rank adds the
 *   W contributions in rank order, so all replicas get bit-identical results.
 *   Moves W x count floats per rank: for small tables, not for model gradients.
 *   Needs (window_bytes - plan bytes)/2 >= world * count * 4.  d_in == d_out ok.
 * ------------------------------------------------------------------------- */
int hbAllreduceSumF32(hbComm* comm, const float* d_in, float* d_out, int64_t count,
                      float scale, int32_t* d_status, hbStream stream);

/* ---------------------------------------------------------------------------
 * K1+K2+K3+K4 fused: sharded GroupLookup (additive op; its oracle is the
 *   composition embedding/sharding.py:171-203 of the ops above).
 * Forward, per rank: sort the ids by (owner = id % W, local row = id / W) and
 * deduplicate them -> push the UNIQUE local rows (32 bit) + counts to their owners
 * -> owners gather the rows and store them straight into the requester's window at
 * the slot of the unique -> requester expands through the inverse map while it
 * pools into out.  Zero host synchronisation, static shapes; only unique ids, rows
 * and gradient sums cross NVLink.  The owner's sort of the received rows (needed by
 * the backward) is enqueued at forward time on a plan-owned side stream.
 * Backward: requester sums the row gradients of each unique id (position order)
 * straight into the owner's window; owners sum the per-rank contributions of a row
 * in rank order and apply the optimizer (sharded gradients are not averaged:
 * training/gradient.py:216-217).
 * ------------------------------------------------------------------------- */
typedef struct hbShardedFeature {
  float* shard;           /* local shard [shard_rows, dim] */
  float* slot0;
  float* slot1;
  int64_t shard_rows;
  const int64_t* ids;     /* this rank's global ids [nnz] */
  const int64_t* offsets; /* [nbags+1] or NULL */
  int64_t nbags;
  int64_t nnz;
  float* out;             /* forward: [nbags, out_stride] */
  int64_t out_stride;
  const float* grad;      /* backward: [nbags, grad_stride] */
  int64_t grad_stride;
  int32_t dim;
  int32_t combiner;
} hbShardedFeature;

typedef struct hbShardedPlan hbShardedPlan;
/* max_nnz[k]: static upper bound of nnz for feature k on ANY rank;
 * capacity_factor >= 1: owner-side receive capacity per feature =
 * ceil(capacity_factor * max_nnz[k]) ids (W*max_nnz is always safe; receive counts
 * stay on the device, so capacity costs memory, not work).  Overflow raises
 * HB_STATUS_WINDOW_OVERFLOW on EVERY rank.  Limits: n <= 256 features per plan, dims
 * multiples of 4 up to 1024, one plan per communicator at a time (PlanDestroy frees the
 * slot; epochs are monotonic across plans).  PlanCreate / PlanDestroy allocate and
 * synchronise the device (setup time, not per step).  Every rank
 * must issue the same Forward / BackwardUpdate sequence; a peer that never
 * arrives raises HB_STATUS_PEER_TIMEOUT after 20 s instead of hanging. */
int hbShardedPlanCreate(hbComm* comm, int n, const int64_t* max_nnz,
                        const int32_t* dims, double capacity_factor,
                        hbShardedPlan** plan);
int hbShardedPlanDestroy(hbShardedPlan* plan);
size_t hbShardedPlanWindowBytes(int world, int n, const int64_t* max_nnz,
                                const int32_t* dims, double capacity_factor);
int hbShardedLookupForward(hbShardedPlan* plan, const hbShardedFeature* feats,
                           int32_t* d_status, hbStream stream);
int hbShardedLookupBackwardUpdate(hbShardedPlan* plan,
                                  const hbShardedFeature* feats,
                                  const hbOptimizer* opt, int32_t* d_status,
                                  hbStream stream);

/* ---------------------------------------------------------------------------
 * Host-buffer entry (what bench.py's e2e leg and a data-loader facing caller
 * use; the reference feeds ids from its host input pipeline).  One pinned host
 * block holding every feature's ids (+offsets) is copied H2D into d_in_block
 * (feats[k].ids / .offsets point INTO d_in_block), the forward runs, and the
 * contiguous device output block (feats[k].out point into it) is copied D2H
 * into h_out_block -- all enqueued on `stream`, nothing synchronised.
 * ------------------------------------------------------------------------- */
int hbGroupLookupForwardHost(int n, const hbLookupFeature* feats,
                             const void* h_in_block, void* d_in_block,
                             size_t in_block_bytes, const void* d_out_block,
                             void* h_out_block, size_t out_block_bytes,
                             int32_t* d_status, hbStream stream);

/* ---------------------------------------------------------------------------
 * Fused multi-tensor host -> device staging of the step's sparse features.
 * Replaces HbH2DTransferN (ops/transfer/transfer.cc:68-80, transfer_functors.cu.cc:38-238):
 * n host tensors of bytes[k] bytes reach n device tensors.  Inputs in pinned (page-locked)
 * host memory are read by ONE kernel launch straight over PCIe (zero copy, 128-bit
 * loads); pageable inputs fall back to one cudaMemcpyAsync each, as in the reference.
 * Everything is enqueued on `stream`; nothing is synchronised.
 * ------------------------------------------------------------------------- */
int hbH2DTransferN(int n, const void* const* h_inputs, void* const* d_outputs,
                   const int64_t* bytes, hbStream stream);

#ifdef __cplusplus
}
#endif
#endif /* HB_B200_H_ */
