/* hb_oracle.h -- CPU restatement of the HybridBackend sharded-embedding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hybridbackend_b200/ may include, link
 * or dlopen this.  Allowed users: tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - partition (modulo / dual modulo): PINNED.  Checked bit-for-bit against the
 *     reference's own CPU functor compiled from /root/reference (oracle/_ref,
 *     built by oracle/build_ref.sh) and against fixtures under tests/golden/.
 *   - alltoallv / alltoallv_n data movement: PINNED against the reference's golden
 *     vectors (hybridbackend/tensorflow/distribute/tests/alltoall_test.py:219-304).
 *   - murmur3_hash32 / slab-hash probe: PINNED (murmur3 vs the reference header
 *     compiled in oracle/_ref; probe vs fixtures generated with it).
 *   - embedding_lookup_sparse pooling, sharded lookup composition, Adagrad /
 *     LazyAdam sparse apply: PARITY UNPINNED -- the arithmetic lives in
 *     tensorflow==1.15.5 (absent); the reference holds no value-asserting test for
 *     it.  Restated from TF-1.15's published semantics and
 *     cross-checked against torch embedding_bag and closed forms.
 */
#ifndef HB_ORACLE_H_
#define HB_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype codes shared with include/hb_b200.h */
enum { HBO_I32 = 0, HBO_I64 = 1, HBO_U32 = 2, HBO_U64 = 3 };
enum { HBO_SUM = 0, HBO_MEAN = 1, HBO_SQRTN = 2 };

/* partition_by_modulo_functors.cc:37-71.  Returns 0 on success. */
int hbo_partition_by_modulo(int dtype, const void* input, int32_t n,
                            int32_t num_partitions, void* output,
                            int32_t* sizes, int32_t* indices);

/* partition_by_dual_modulo_functors.cc:37-91.  stage is 1 or 2. */
int hbo_partition_by_dual_modulo(int dtype, int stage, const void* input,
                                 int32_t n, int32_t num_partitions,
                                 int32_t modulus, void* output, int32_t* sizes,
                                 int32_t* indices);

/* nccl_collective.cc:250-288 restated for W in-process ranks.
 * send[r] is rank r's input of sum(send_sizes[r*W..]) * common_size elements of
 * elem_bytes each; recv[r] must hold sum_q send_sizes[q*W+r]*common_size elems.
 * recv_sizes is W*W, recv_sizes[r*W+q] = send_sizes[q*W+r]. */
int hbo_alltoallv(int world, const void* const* send, const int32_t* send_sizes,
                  int64_t common_size, int elem_bytes, void* const* recv,
                  int32_t* recv_sizes);

/* embedding/variables.py:107-111: rows of shard s of an N-row table over W. */
int64_t hbo_shard_rows(int64_t bucket_size, int num_shards, int shard);
/* embedding/variables.py:112-117: SaveSliceInfo row offset of shard s. */
int64_t hbo_shard_offset(int64_t bucket_size, int num_shards, int shard);
/* embedding/variables.py:95-96: 1 if the table stays replicated. */
int hbo_is_small_table(int64_t bucket_size, int num_shards, int64_t batch_size);

/* tf.unique semantics (first-occurrence order).  Returns number of uniques. */
int64_t hbo_unique_i64(const int64_t* ids, int64_t n, int64_t* uniq,
                       int32_t* inverse);

/* TF-1.15 embedding_lookup_sparse(params, sp_ids, None, combiner) at W=1:
 * unique -> gather -> sparse_segment_{sum,mean,sqrtn}.  CSR form: bag b owns
 * ids[offsets[b]..offsets[b+1]).  out is [nbags, out_stride] (first dim floats
 * per row written; empty bags -> zeros). */
int hbo_embedding_lookup_sparse(const float* table, int64_t rows, int dim,
                                const int64_t* ids, const int64_t* offsets,
                                int64_t nbags, int combiner, float* out,
                                int64_t out_stride);

/* Same, skipping the hash-unique stage (numerically identical; used as the
 * multi-thread CPU baseline inner loop). */
int hbo_embedding_bag(const float* table, int64_t rows, int dim,
                      const int64_t* ids, const int64_t* offsets, int64_t nbags,
                      int combiner, float* out, int64_t out_stride);

/* Backward of the pooled lookup: per-id row gradient (IndexedSlices values),
 * row_grad[p] = grad[bag(p)] * scale(bag) ; [nnz, dim]. */
int hbo_lookup_row_grads(const float* grad, int64_t grad_stride, int dim,
                         const int64_t* offsets, int64_t nbags, int combiner,
                         float* row_grad);

/* TF optimizer._apply_sparse_duplicate_indices: sum duplicates (in position
 * order) then apply once per unique row.  rows_idx are LOCAL row indices. */
int hbo_sparse_apply_adagrad(float* table, float* accum, int64_t rows, int dim,
                             const int64_t* rows_idx, const float* row_grad,
                             int64_t nnz, float lr);
/* tf.contrib.opt.LazyAdamOptimizer._apply_sparse: touched rows only.
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller side of TF; here step
 * t (1-based) is passed and lr_t derived in double then cast to float. */
int hbo_sparse_apply_lazy_adam(float* table, float* m, float* v, int64_t rows,
                               int dim, const int64_t* rows_idx,
                               const float* row_grad, int64_t nnz, float lr,
                               float beta1, float beta2, float eps,
                               int64_t step);

/* embedding/sharding.py:171-203 for W in-process ranks, one feature:
 * shards[s] is shard s's [hbo_shard_rows(N,W,s), dim] table; ids[r]/n[r] the
 * flat ids of rank r; out[r] is [n[r], dim] (row per id, stitched order). */
int hbo_sharded_embedding_lookup(int world, const float* const* shards,
                                 int64_t bucket_size, int dim,
                                 const int64_t* const* ids, const int64_t* n,
                                 float* const* out);

/* hybridbackend/common/murmur3.cu.h:28-77 for an 8-byte key, seed 0. */
uint32_t hbo_murmur3_hash32_i64(int64_t key);

/* embedding/lookup_functors.cu.cc:53-149 restated sequentially.  The GPU
 * kernel's output ORDER inside the hit and miss groups depends on the launch
 * shape; the oracle emits hits in ascending key index and misses in ascending
 * key index, and tests compare as sets of (key index, payload) pairs.
 * keys_cache is slabs*32 entries, empty = INT64_MIN.  Returns miss count;
 * hit_idx/hit_cache sized n, miss_idx/miss_keys sized n. */
int64_t hbo_cache_lookup(const int64_t* keys_cache, int64_t slabs,
                         const int64_t* keys, int64_t n, int32_t* hit_idx,
                         int64_t* hit_cache, int32_t* miss_idx,
                         int64_t* miss_keys, int64_t* n_hit);

/* ---------------------------------------------------------------------------
 * Multi-threaded step of the same CPU semantics (pthreads), used only by
 * bench.py's CPU legs: forward = embedding_lookup_sparse per (feature,
 * bag-chunk) task, backward = dedup + SparseApplyAdagrad per (feature,
 * row-residue part) task (parts own disjoint rows, so the union equals the
 * unsplit update).  One id per bag.  Returns 0 on success.
 * ------------------------------------------------------------------------- */
typedef struct hbo_mt_feature {
  float* table;
  float* accum;       /* NULL: forward only */
  int64_t rows;
  int32_t dim;
  int32_t parts;      /* backward tasks for this feature (>= 1) */
  const int64_t* ids; /* [nbags] */
  int64_t nbags;
  const float* grad;  /* [nbags, grad_stride] */
  int64_t grad_stride;
  float* out;         /* [nbags, out_stride] */
  int64_t out_stride;
} hbo_mt_feature;

int hbo_mt_step(int nfeat, const hbo_mt_feature* feats, int64_t fwd_chunk, float lr, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* HB_ORACLE_H_ */
