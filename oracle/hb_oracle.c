/* hb_oracle.c -- CPU restatement of the HybridBackend sharded-embedding hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see hb_oracle.h for who may use it and for the
 * parity-pinning status of every function).  Plain C99, no dependencies.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction so the
 * fp32 sequences below are exactly the ones written).
 *
 * Each function cites the reference file:line (relative to
 * /root/reference/hybridbackend/) it restates.
 */
#include "hb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Partition: tensorflow/distribute/partition/partition_by_modulo_functors.cc */
/* :37-71 -- stable counting sort by floor-modulo shard.                      */
/* The C++ reference evaluates (v % P + P) % P in the usual arithmetic        */
/* conversions of T with int32 P: signed T -> signed math (sign-safe modulo), */
/* unsigned T -> P converted to unsigned T.  Restated per type below.         */
/* ------------------------------------------------------------------------- */

#define DEFINE_SHARD_FN(NAME, T)                                              \
  static inline int32_t NAME(T v, int32_t p) {                                \
    return (int32_t)((v % (T)p + (T)p) % (T)p);                               \
  }
DEFINE_SHARD_FN(shard_i32, int32_t)
DEFINE_SHARD_FN(shard_i64, int64_t)
DEFINE_SHARD_FN(shard_u32, uint32_t)
DEFINE_SHARD_FN(shard_u64, uint64_t)

/* Counting sort common to modulo and dual modulo
 * (partition_by_modulo_functors.cc:48-69, dual: :60-89). */
static int counting_sort(const int32_t* shard, int dtype, const void* input,
                         int32_t n, int32_t p, void* output, int32_t* sizes,
                         int32_t* indices) {
  int32_t* local = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  int32_t* offs = (int32_t*)calloc((size_t)p, sizeof(int32_t));
  if (!local || !offs) return 2;
  for (int32_t i = 0; i < n; ++i) {
    local[i] = offs[shard[i]];
    offs[shard[i]]++;
  }
  memcpy(sizes, offs, sizeof(int32_t) * (size_t)p);
  for (int32_t i = 1; i < p; ++i) offs[i] += offs[i - 1];
  for (int32_t i = 0; i < n; ++i) {
    int32_t off = local[i];
    if (shard[i] > 0) off += offs[shard[i] - 1];
    switch (dtype) {
      case HBO_I32: case HBO_U32:
        ((uint32_t*)output)[off] = ((const uint32_t*)input)[i]; break;
      default:
        ((uint64_t*)output)[off] = ((const uint64_t*)input)[i]; break;
    }
    indices[i] = off;
  }
  free(local);
  free(offs);
  return 0;
}

int hbo_partition_by_modulo(int dtype, const void* input, int32_t n,
                            int32_t num_partitions, void* output,
                            int32_t* sizes, int32_t* indices) {
  if (num_partitions < 1 || n < 0) return 1;
  int32_t* shard = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  if (!shard) return 2;
  for (int32_t i = 0; i < n; ++i) {
    switch (dtype) {
      case HBO_I32: shard[i] = shard_i32(((const int32_t*)input)[i], num_partitions); break;
      case HBO_I64: shard[i] = shard_i64(((const int64_t*)input)[i], num_partitions); break;
      case HBO_U32: shard[i] = shard_u32(((const uint32_t*)input)[i], num_partitions); break;
      case HBO_U64: shard[i] = shard_u64(((const uint64_t*)input)[i], num_partitions); break;
      default: free(shard); return 1;
    }
  }
  int rc = counting_sort(shard, dtype, input, n, num_partitions, output, sizes, indices);
  free(shard);
  return rc;
}

/* partition_by_dual_modulo_functors.cc:37-91.
 * pre = (v % (P*M) + P*M) % (P*M); stage 1 shard = (pre % P + P) % P,
 * stage 2 shard = pre / M. */
#define DEFINE_DUAL_FN(NAME, T)                                               \
  static inline int32_t NAME(T v, int32_t p, int32_t m, int stage) {          \
    const int32_t pm = p * m;                                                 \
    const T pre = (v % (T)pm + (T)pm) % (T)pm;                                \
    if (stage == 1) return (int32_t)((pre % (T)p + (T)p) % (T)p);             \
    return (int32_t)(pre / (T)m);                                             \
  }
DEFINE_DUAL_FN(dual_i32, int32_t)
DEFINE_DUAL_FN(dual_i64, int64_t)
DEFINE_DUAL_FN(dual_u32, uint32_t)
DEFINE_DUAL_FN(dual_u64, uint64_t)

int hbo_partition_by_dual_modulo(int dtype, int stage, const void* input,
                                 int32_t n, int32_t num_partitions,
                                 int32_t modulus, void* output, int32_t* sizes,
                                 int32_t* indices) {
  if (num_partitions < 1 || modulus < 1 || n < 0 || (stage != 1 && stage != 2))
    return 1;
  int32_t* shard = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  if (!shard) return 2;
  for (int32_t i = 0; i < n; ++i) {
    switch (dtype) {
      case HBO_I32: shard[i] = dual_i32(((const int32_t*)input)[i], num_partitions, modulus, stage); break;
      case HBO_I64: shard[i] = dual_i64(((const int64_t*)input)[i], num_partitions, modulus, stage); break;
      case HBO_U32: shard[i] = dual_u32(((const uint32_t*)input)[i], num_partitions, modulus, stage); break;
      case HBO_U64: shard[i] = dual_u64(((const uint64_t*)input)[i], num_partitions, modulus, stage); break;
      default: free(shard); return 1;
    }
    if (shard[i] < 0 || shard[i] >= num_partitions) { free(shard); return 3; }
  }
  int rc = counting_sort(shard, dtype, input, n, num_partitions, output, sizes, indices);
  free(shard);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* Alltoallv: tensorflow/distribute/nccl/nccl_collective.cc:250-288.          */
/* Rank r sends segment i of its input (send_sizes[r][i]*common elems) to     */
/* rank i and receives from rank i at recvoffset accumulated in rank order:   */
/* output = concat over source ranks (ascending) of the segment addressed to  */
/* me.  recv_sizes come from the size pre-exchange (nccl_alltoallv.cc:307).   */
/* ------------------------------------------------------------------------- */
int hbo_alltoallv(int world, const void* const* send, const int32_t* send_sizes,
                  int64_t common_size, int elem_bytes, void* const* recv,
                  int32_t* recv_sizes) {
  if (world < 1 || common_size < 0 || elem_bytes < 1) return 1;
  for (int r = 0; r < world; ++r)
    for (int q = 0; q < world; ++q)
      recv_sizes[r * world + q] = send_sizes[q * world + r];
  for (int q = 0; q < world; ++q) { /* sender */
    int64_t sendoff = 0;
    for (int r = 0; r < world; ++r) { /* destination */
      const int64_t bytes =
          (int64_t)send_sizes[q * world + r] * common_size * elem_bytes;
      int64_t recvoff = 0;
      for (int qq = 0; qq < q; ++qq)
        recvoff += (int64_t)send_sizes[qq * world + r] * common_size * elem_bytes;
      if (bytes > 0)
        memcpy((char*)recv[r] + recvoff, (const char*)send[q] + sendoff,
               (size_t)bytes);
      sendoff += bytes;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Sharding rule: tensorflow/embedding/variables.py:95-117.                   */
/* ------------------------------------------------------------------------- */
int64_t hbo_shard_rows(int64_t bucket_size, int num_shards, int shard) {
  int64_t rows = bucket_size / num_shards;              /* :107 */
  if (shard < bucket_size % num_shards) rows += 1;      /* :108-109 */
  return rows;
}

int64_t hbo_shard_offset(int64_t bucket_size, int num_shards, int shard) {
  int64_t off = (bucket_size / num_shards) * shard;     /* :118 */
  const int64_t rem = bucket_size % num_shards;         /* :119 */
  off += (shard < rem) ? shard : rem;                   /* :120-123 */
  return off;
}

int hbo_is_small_table(int64_t bucket_size, int num_shards, int64_t batch_size) {
  return (bucket_size <= num_shards || bucket_size <= batch_size) ? 1 : 0; /* :96 */
}

/* ------------------------------------------------------------------------- */
/* tf.unique (TF-1.15 UniqueOp, CPU: tensorflow/core/kernels/unique_op.cc,    */
/* UniqueOp::Compute -- hash map in input order): first-occurrence order,     */
/* int32 inverse.                                                            */
/* ------------------------------------------------------------------------- */
typedef struct {
  int64_t* keys;
  int64_t* vals;
  uint64_t mask;
} hmap_t;

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33; return x;
}

static int hmap_init(hmap_t* h, int64_t n) {
  uint64_t cap = 16;
  while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
  h->keys = (int64_t*)malloc(sizeof(int64_t) * cap);
  h->vals = (int64_t*)malloc(sizeof(int64_t) * cap);
  if (!h->keys || !h->vals) return 2;
  for (uint64_t i = 0; i < cap; ++i) h->vals[i] = -1;
  h->mask = cap - 1;
  return 0;
}

static void hmap_free(hmap_t* h) { free(h->keys); free(h->vals); }

/* returns slot value if present, else inserts val and returns -1 */
static inline int64_t hmap_get_or_put(hmap_t* h, int64_t key, int64_t val) {
  uint64_t s = mix64((uint64_t)key) & h->mask;
  for (;;) {
    if (h->vals[s] < 0) { h->keys[s] = key; h->vals[s] = val; return -1; }
    if (h->keys[s] == key) return h->vals[s];
    s = (s + 1) & h->mask;
  }
}

int64_t hbo_unique_i64(const int64_t* ids, int64_t n, int64_t* uniq,
                       int32_t* inverse) {
  hmap_t h;
  if (hmap_init(&h, n)) return -1;
  int64_t u = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t got = hmap_get_or_put(&h, ids[i], u);
    if (got < 0) { uniq[u] = ids[i]; got = u++; }
    inverse[i] = (int32_t)got;
  }
  hmap_free(&h);
  return u;
}

/* ------------------------------------------------------------------------- */
/* TF-1.15 embedding_lookup_sparse (python/ops/embedding_ops.py) at W=1:      */
/*   ids, idx = unique(sp_ids.values)                                         */
/*   emb = embedding_lookup(params, ids)          (GatherV2)                  */
/*   out = sparse_segment_{sum,mean,sqrtn}(emb, idx, segment_ids)             */
/* The segment reduction accumulates rows in index order in fp32; mean        */
/* divides the sum by the count, sqrtn by sqrt(count).  PARITY UNPINNED.      */
/* TF-1.15 sources restated (tensorflow==1.15.5, not under /root/reference):  */
/*   python/ops/embedding_ops.py  embedding_lookup_sparse (:unique, :gather,  */
/*     sparse_segment_* with sp_weights None)                                 */
/*   core/kernels/segment_reduction_ops.cc  SparseSegmentReductionOpBase (CPU) */
/*     -- one output row at a time, input rows taken in index order.  How the   */
/*     kernel groups the additions and where it applies the mean / sqrtn scale */
/*     for long bags cannot be checked without the TF sources or binary; the   */
/*     restatement is the sequential sum followed by one division (exact for   */
/*     the bag lengths 1..2, the published semantics beyond)                   */
/* ------------------------------------------------------------------------- */
static inline void finish_bag(float* o, int dim, int64_t cnt, int combiner) {
  if (cnt == 0) return;
  if (combiner == HBO_MEAN) {
    const float c = (float)cnt;
    for (int d = 0; d < dim; ++d) o[d] = o[d] / c;
  } else if (combiner == HBO_SQRTN) {
    const float c = sqrtf((float)cnt);
    for (int d = 0; d < dim; ++d) o[d] = o[d] / c;
  }
}

int hbo_embedding_lookup_sparse(const float* table, int64_t rows, int dim,
                                const int64_t* ids, const int64_t* offsets,
                                int64_t nbags, int combiner, float* out,
                                int64_t out_stride) {
  const int64_t nnz = offsets[nbags];
  int64_t* uniq = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
  int32_t* inv = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  if (!uniq || !inv) return 2;
  const int64_t u = hbo_unique_i64(ids, nnz, uniq, inv);
  if (u < 0) return 2;
  for (int64_t i = 0; i < u; ++i)
    if (uniq[i] < 0 || uniq[i] >= rows) { free(uniq); free(inv); return 3; }
  float* emb = (float*)malloc(sizeof(float) * (size_t)(u > 0 ? u : 1) * (size_t)dim);
  if (!emb) return 2;
  for (int64_t i = 0; i < u; ++i)
    memcpy(emb + i * dim, table + uniq[i] * dim, sizeof(float) * (size_t)dim);
  for (int64_t b = 0; b < nbags; ++b) {
    float* o = out + b * out_stride;
    for (int d = 0; d < dim; ++d) o[d] = 0.0f;
    for (int64_t p = offsets[b]; p < offsets[b + 1]; ++p) {
      const float* row = emb + (int64_t)inv[p] * dim;
      for (int d = 0; d < dim; ++d) o[d] = o[d] + row[d];
    }
    finish_bag(o, dim, offsets[b + 1] - offsets[b], combiner);
  }
  free(emb); free(uniq); free(inv);
  return 0;
}

int hbo_embedding_bag(const float* table, int64_t rows, int dim,
                      const int64_t* ids, const int64_t* offsets, int64_t nbags,
                      int combiner, float* out, int64_t out_stride) {
  for (int64_t b = 0; b < nbags; ++b) {
    float* o = out + b * out_stride;
    for (int d = 0; d < dim; ++d) o[d] = 0.0f;
    for (int64_t p = offsets[b]; p < offsets[b + 1]; ++p) {
      if (ids[p] < 0 || ids[p] >= rows) return 3;
      const float* row = table + ids[p] * dim;
      for (int d = 0; d < dim; ++d) o[d] = o[d] + row[d];
    }
    finish_bag(o, dim, offsets[b + 1] - offsets[b], combiner);
  }
  return 0;
}

/* Gradient of sparse_segment_{sum,mean,sqrtn} w.r.t. the gathered rows
 * (TF python/ops/math_grad.py _SparseSegment{Sum,Mean,SqrtN}Grad ->
 * core/kernels/segment_reduction_ops.cc SparseSegmentGradOpBase): row p receives
 * grad[bag(p)], divided by count (mean) or sqrt(count) (sqrtn).  PARITY UNPINNED. */
int hbo_lookup_row_grads(const float* grad, int64_t grad_stride, int dim,
                         const int64_t* offsets, int64_t nbags, int combiner,
                         float* row_grad) {
  for (int64_t b = 0; b < nbags; ++b) {
    const int64_t cnt = offsets[b + 1] - offsets[b];
    float c = 1.0f;
    if (combiner == HBO_MEAN) c = (float)cnt;
    else if (combiner == HBO_SQRTN) c = sqrtf((float)cnt);
    for (int64_t p = offsets[b]; p < offsets[b + 1]; ++p)
      for (int d = 0; d < dim; ++d) {
        const float g = grad[b * grad_stride + d];
        row_grad[p * dim + d] = (combiner == HBO_SUM) ? g : g / c;
      }
  }
  return 0;
}

/* python/training/optimizer.py _apply_sparse_duplicate_indices ->
 * _deduplicate_indexed_slices: unique + unsorted_segment_sum, then one apply per
 * unique row.  CPU UnsortedSegmentSum (core/kernels/segment_reduction_ops.cc,
 * UnsortedSegmentFunctor<CPUDevice>) walks the input rows 0..N-1 once and adds row i
 * into output[segment_ids[i]]: per unique row the additions happen in POSITION order,
 * which is the order restated here (and the order the GPU kernels bound against).  Returns summed grads in sum_g [u, dim] and unique rows; caller frees. */
static int64_t dedup_sum(const int64_t* rows_idx, const float* row_grad,
                         int64_t nnz, int dim, int64_t** uniq_out,
                         float** sum_out) {
  int64_t* uniq = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
  int32_t* inv = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  if (!uniq || !inv) return -1;
  const int64_t u = hbo_unique_i64(rows_idx, nnz, uniq, inv);
  float* sum = (float*)calloc((size_t)(u > 0 ? u : 1) * (size_t)dim, sizeof(float));
  if (!sum) return -1;
  for (int64_t p = 0; p < nnz; ++p) {
    float* s = sum + (int64_t)inv[p] * dim;
    const float* g = row_grad + p * dim;
    for (int d = 0; d < dim; ++d) s[d] = s[d] + g[d];
  }
  free(inv);
  *uniq_out = uniq;
  *sum_out = sum;
  return u;
}

/* TF-1.15 core/kernels/training_ops.cc SparseApplyAdagradOp<CPUDevice>::Compute
 * (update_slots=true, no epsilon).  The published semantics are restated with IEEE sqrt
 * and divide; TF's vectorised inner loop may evaluate the same expression as a multiply
 * by an (Eigen) rsqrt, whose last-bit behaviour is not reproducible without the binary --
 * one of the reasons this function is labelled PARITY UNPINNED:
 *   accum[r] += g*g ; var[r] -= lr * g / sqrt(accum[r]).  PARITY UNPINNED. */
int hbo_sparse_apply_adagrad(float* table, float* accum, int64_t rows, int dim,
                             const int64_t* rows_idx, const float* row_grad,
                             int64_t nnz, float lr) {
  int64_t* uniq; float* sum;
  const int64_t u = dedup_sum(rows_idx, row_grad, nnz, dim, &uniq, &sum);
  if (u < 0) return 2;
  for (int64_t i = 0; i < u; ++i) {
    if (uniq[i] < 0 || uniq[i] >= rows) { free(uniq); free(sum); return 3; }
    float* w = table + uniq[i] * dim;
    float* a = accum + uniq[i] * dim;
    const float* g = sum + i * dim;
    for (int d = 0; d < dim; ++d) {
      const float acc = a[d] + g[d] * g[d];
      a[d] = acc;
      w[d] = w[d] - (lr * g[d]) / sqrtf(acc);
    }
  }
  free(uniq); free(sum);
  return 0;
}

/* tf.contrib.opt.LazyAdamOptimizer._apply_sparse (TF-1.15:
 * tensorflow/contrib/opt/python/training/lazy_adam_optimizer.py -- gather m, v of the
 * touched rows, update, scatter_update back):
 *   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
 *   m[r] = b1*m[r] + (1-b1)*g ; v[r] = b2*v[r] + (1-b2)*g*g
 *   var[r] -= lr_t * m[r] / (sqrt(v[r]) + eps)          PARITY UNPINNED. */
int hbo_sparse_apply_lazy_adam(float* table, float* m, float* v, int64_t rows,
                               int dim, const int64_t* rows_idx,
                               const float* row_grad, int64_t nnz, float lr,
                               float beta1, float beta2, float eps,
                               int64_t step) {
  int64_t* uniq; float* sum;
  const int64_t u = dedup_sum(rows_idx, row_grad, nnz, dim, &uniq, &sum);
  if (u < 0) return 2;
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) /
                             (1.0 - pow((double)beta1, (double)step)));
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  for (int64_t i = 0; i < u; ++i) {
    if (uniq[i] < 0 || uniq[i] >= rows) { free(uniq); free(sum); return 3; }
    float* w = table + uniq[i] * dim;
    float* mm = m + uniq[i] * dim;
    float* vv = v + uniq[i] * dim;
    const float* g = sum + i * dim;
    for (int d = 0; d < dim; ++d) {
      const float mn = beta1 * mm[d] + omb1 * g[d];
      const float vn = beta2 * vv[d] + omb2 * (g[d] * g[d]);
      mm[d] = mn;
      vv[d] = vn;
      w[d] = w[d] - (lr_t * mn) / (sqrtf(vn) + eps);
    }
  }
  free(uniq); free(sum);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Sharded lookup recipe: tensorflow/embedding/sharding.py:171-203.           */
/* ------------------------------------------------------------------------- */
int hbo_sharded_embedding_lookup(int world, const float* const* shards,
                                 int64_t bucket_size, int dim,
                                 const int64_t* const* ids, const int64_t* n,
                                 float* const* out) {
  const int W = world;
  int rc = 0;
  /* :179-180 partition_by_modulo(ids, num_shards) on every rank */
  int64_t** part = (int64_t**)calloc((size_t)W, sizeof(void*));
  int32_t** pidx = (int32_t**)calloc((size_t)W, sizeof(void*));
  int32_t* sizes = (int32_t*)calloc((size_t)W * W, sizeof(int32_t));
  int32_t* rsizes = (int32_t*)calloc((size_t)W * W, sizeof(int32_t));
  for (int r = 0; r < W; ++r) {
    part[r] = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n[r] > 0 ? n[r] : 1));
    pidx[r] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n[r] > 0 ? n[r] : 1));
    rc |= hbo_partition_by_modulo(HBO_I64, ids[r], (int32_t)n[r], W, part[r],
                                  sizes + r * W, pidx[r]);
  }
  /* :181-182 alltoall(ids_shards, sizes=ids_sizes) */
  int64_t** rids = (int64_t**)calloc((size_t)W, sizeof(void*));
  int64_t* rtotal = (int64_t*)calloc((size_t)W, sizeof(int64_t));
  for (int r = 0; r < W; ++r) {
    for (int q = 0; q < W; ++q) rtotal[r] += sizes[q * W + r];
    rids[r] = (int64_t*)malloc(sizeof(int64_t) * (size_t)(rtotal[r] > 0 ? rtotal[r] : 1));
  }
  rc |= hbo_alltoallv(W, (const void* const*)part, sizes, 1, 8, (void* const*)rids, rsizes);
  /* owner side */
  float** remb = (float**)calloc((size_t)W, sizeof(void*));
  for (int r = 0; r < W && rc == 0; ++r) {
    const int64_t m = rtotal[r];
    int64_t* uniq = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1));
    int32_t* inv = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    const int64_t u = hbo_unique_i64(rids[r], m, uniq, inv);      /* :183-184 */
    const int64_t srows = hbo_shard_rows(bucket_size, W, r);
    remb[r] = (float*)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1) * (size_t)dim);
    for (int64_t i = 0; i < u; ++i) {
      uniq[i] = uniq[i] / W;                                      /* :185-186 */
      if (uniq[i] < 0 || uniq[i] >= srows) rc = 3;
    }
    /* :187-189 fn(params, shard_ids) then :190-192 shard_unique_restore */
    for (int64_t p = 0; p < m && rc == 0; ++p)
      memcpy(remb[r] + p * dim, shards[r] + uniq[inv[p]] * dim,
             sizeof(float) * (size_t)dim);
    free(uniq); free(inv);
  }
  /* :193-196 alltoall(embeddings, sizes=shard_sizes, common_shape=[dim]) */
  float** back = (float**)calloc((size_t)W, sizeof(void*));
  int32_t* bsizes = (int32_t*)calloc((size_t)W * W, sizeof(int32_t));
  for (int r = 0; r < W; ++r)
    back[r] = (float*)malloc(sizeof(float) * (size_t)(n[r] > 0 ? n[r] : 1) * (size_t)dim);
  if (rc == 0)
    rc |= hbo_alltoallv(W, (const void* const*)remb, rsizes, dim, 4,
                        (void* const*)back, bsizes);
  /* :197-199 gather(embeddings, shard_index) -- shard_stitch */
  for (int r = 0; r < W && rc == 0; ++r)
    for (int64_t i = 0; i < n[r]; ++i)
      memcpy(out[r] + i * dim, back[r] + (int64_t)pidx[r][i] * dim,
             sizeof(float) * (size_t)dim);
  for (int r = 0; r < W; ++r) {
    free(part[r]); free(pidx[r]); free(rids[r]); free(remb[r]); free(back[r]);
  }
  free(part); free(pidx); free(sizes); free(rsizes); free(rids); free(rtotal);
  free(remb); free(back); free(bsizes);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* common/murmur3.cu.h:28-77 (MurmurHash3_x86_32) specialised to len = 8.     */
/* ------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t hbo_murmur3_hash32_i64(int64_t key) {
  uint32_t blocks[2];
  memcpy(blocks, &key, 8);
  uint32_t h1 = 0;
  const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  for (int i = 0; i < 2; ++i) {
    uint32_t k1 = blocks[i];
    k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
    h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
  }
  h1 ^= 8u;
  h1 ^= h1 >> 16; h1 *= 0x85ebca6bu;
  h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u;
  h1 ^= h1 >> 16;
  return h1;
}

/* ------------------------------------------------------------------------- */
/* Slab-hash cache probe: embedding/lookup_functors.cu.cc:53-149.             */
/* slab = murmur3(key) % slabs (computed in T=int64 after widening the u32    */
/* hash, :72); probe slabs linearly; a slab containing the key -> hit at      */
/* offset slab*32 + first matching slot; a slab containing an empty slot      */
/* (INT64_MIN) -> miss; all slabs probed -> miss.                             */
/* ------------------------------------------------------------------------- */
int64_t hbo_cache_lookup(const int64_t* keys_cache, int64_t slabs,
                         const int64_t* keys, int64_t n, int32_t* hit_idx,
                         int64_t* hit_cache, int32_t* miss_idx,
                         int64_t* miss_keys, int64_t* n_hit) {
  const int64_t kEmpty = INT64_MIN;
  int64_t nh = 0, nm = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t key = keys[i];
    int64_t slab = (int64_t)hbo_murmur3_hash32_i64(key) % slabs;
    int64_t probed = 0;
    int done = 0;
    while (!done) {
      if (probed >= slabs) { miss_idx[nm] = (int32_t)i; miss_keys[nm++] = key; break; }
      const int64_t off = slab * 32;
      int good = -1, empty = 0;
      for (int s = 0; s < 32; ++s) {
        if (good < 0 && keys_cache[off + s] == key) good = s;
        if (keys_cache[off + s] == kEmpty) empty = 1;
      }
      if (good >= 0) { hit_idx[nh] = (int32_t)i; hit_cache[nh++] = off + good; done = 1; }
      else if (empty) { miss_idx[nm] = (int32_t)i; miss_keys[nm++] = key; done = 1; }
      else { probed++; slab = (slab + 1) % slabs; }
    }
  }
  *n_hit = nh;
  return nm;
}
