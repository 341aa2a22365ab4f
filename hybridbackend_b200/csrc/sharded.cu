// sharded.cu -- the fused sharded GroupLookup over NVSwitch peer memory:
// K1 partition + K2 id push + K3 owner gather fused with the row push (the
// return all-to-all) + K4 stitch/pool, and the backward gradient push fused
// with K5 (owner-side dedup + sparse optimizer).  It is the composition of
// embedding/sharding.py:171-203 with every intermediate kept on the device:
// no host synchronisation, static shapes, no NCCL.
//
// Per rank and step (forward):
//   partition        ids --(id % W, stable)--> part_ids, sizes[W], idx      (K1)
//   exchange         sizes all-gather through peer mailboxes; every rank derives
//                    all offsets it needs from the W x F x W matrix          (1 CTA)
//   push_ids         my bucket r -> owner r's ids_in window (128-bit stores) (K2)
//   owner_gather     for every id I own: read the row of my shard (local row
//                    id / W) and store it into the REQUESTER's rows_in window at
//                    its partitioned position -- gather and return all-to-all
//                    are one kernel, tile by tile over NVLink              (K3+K2')
//   stitch_pool      out[b] = pool_p rows_in[idx[p]]                          (K4)
// Backward:
//   push_grads       row gradient of position p -> owner's grads_in window at
//                    the position its id was received                       (K2'')
//   owner update     radix sort of the received ids + fused dedup/optimizer  (K5)
// Cross-GPU ordering: epoch-valued flags (st.release.sys by the last CTA of the
// producer kernel, ld.acquire.sys spin at the start of the consumer kernel).
#include <math.h>
#include <string.h>

#include <vector>

#include "bucket.cuh"
#include "comm.cuh"

namespace hb {

constexpr int kShMaxFeats = 64;   // features per plan launch group
constexpr int kIdChunk = 2048;    // ids per copy chunk (16 KB)

struct ShFeatMeta {
  int32_t send_off[kMaxWorld + 1];    // my bucket starts in part_ids (prefix of my sizes)
  int32_t remote_base[kMaxWorld];     // my segment start inside owner r's ids_in / grads_in
  int32_t recv_base[kMaxWorld + 1];   // as owner: start of source q's segment (clamped to cap)
  int32_t src_bucket_off[kMaxWorld];  // as owner: start of bucket `me` in q's partitioned order
  int32_t recv_total;                 // unclamped number of ids addressed to me
  int32_t recv_clamped;               // min(recv_total, cap): entries actually received
};

struct ShFeat {
  // static (plan)
  int64_t* part_ids;
  int32_t* part_idx;
  int32_t* sizes;        // [W]
  int32_t* bag_of_pos;   // CSR features
  uint64_t ids_in_off[2];  // byte offsets inside the data window
  uint64_t rows_in_off;
  uint64_t grads_in_off;
  int32_t cap;
  int32_t dim;
  int32_t log2g;
  int32_t max_nnz;
  // per call
  const float* shard;
  const float* grad;
  const int64_t* offsets;
  int64_t shard_rows;
  int64_t grad_stride;
  int32_t nnz;
  int32_t nbags;
  int32_t combiner;
  int32_t cta_begin;
};

struct ShParams {
  ShFeat f[kShMaxFeats];
  ShFeatMeta* meta;      // [n] device
  PeerPtrs peers;
  uint64_t window_off;   // control_bytes()
  int32_t* status;
  int32_t n, me, world, parity;
  uint32_t epoch;
  int32_t total_ctas;
  int32_t div_shift;     // log2(W) or -1
};

__device__ __forceinline__ Control* ctl(const ShParams& P, int r) {
  return reinterpret_cast<Control*>(P.peers.p[r]);
}
__device__ __forceinline__ unsigned char* win(const ShParams& P, int r) {
  return P.peers.p[r] + P.window_off;
}

// last-CTA-done: publish `epoch` into flag slot `phase` of every peer
__device__ __forceinline__ void signal_all_peers(const ShParams& P, int phase, int counter) {
  // CTA barrier, then ONE system fence by thread 0: the barrier orders every
  // thread's peer stores before the fence, the fence makes them visible system
  // wide before the counter/flag updates (cooperative-groups grid.sync pattern)
  __syncthreads();
  __shared__ bool s_last;
  Control* mine = ctl(P, P.me);
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned done = atomicAdd(&mine->done_counter[counter], 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) mine->done_counter[counter] = 0;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < P.world) {
    __threadfence_system();
    st_release_sys_u32(&ctl(P, threadIdx.x)->plan_flags[phase][P.me], P.epoch);
  }
}

__device__ __forceinline__ void wait_all_peers(const ShParams& P, int phase) {
  if ((int)threadIdx.x < P.world)
    if (!wait_flag(&ctl(P, P.me)->plan_flags[phase][threadIdx.x], P.epoch))
      raise_status(P.status, HB_STATUS_PEER_TIMEOUT);
  __syncthreads();
}

// ---- exchange: sizes all-gather + offset derivation (1 CTA) ---------------------
__global__ void __launch_bounds__(256) sh_exchange_kernel(const __grid_constant__ ShParams P) {
  const int W = P.world, me = P.me, n = P.n;
  const int par = P.epoch & 1;
  Control* mine = ctl(P, me);
  for (int i = threadIdx.x; i < n * W; i += blockDim.x) {
    const int f = i / W, r = i % W;
    const int32_t v = P.f[f].sizes[r];
    for (int q = 0; q < W; ++q)
      ctl(P, q)->plan_mailbox[par][(me * kMaxA2aTensors + f) * kMaxWorld + r] = v;
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < W) {
    st_release_sys_u32(&ctl(P, threadIdx.x)->plan_flags[0][me], P.epoch);
    if (!wait_flag(&mine->plan_flags[0][threadIdx.x], P.epoch))
      raise_status(P.status, HB_STATUS_PEER_TIMEOUT);
  }
  __syncthreads();
  auto S = [&](int q, int f, int r) -> int32_t {
    return *reinterpret_cast<volatile int32_t*>(
        &mine->plan_mailbox[par][(q * kMaxA2aTensors + f) * kMaxWorld + r]);
  };
  for (int f = threadIdx.x; f < n; f += blockDim.x) {
    ShFeatMeta m;
    const int cap = P.f[f].cap;
    int acc = 0;
    for (int r = 0; r < W; ++r) { m.send_off[r] = acc; acc += S(me, f, r); }
    m.send_off[W] = acc;
    for (int r = W + 1; r <= kMaxWorld; ++r) m.send_off[r] = acc;
    for (int r = 0; r < W; ++r) {
      int b = 0;
      for (int q = 0; q < me; ++q) b += S(q, f, r);
      m.remote_base[r] = b;
    }
    acc = 0;
    for (int q = 0; q < W; ++q) {
      m.recv_base[q] = acc < cap ? acc : cap;
      acc += S(q, f, me);
      int o = 0;
      for (int r = 0; r < me; ++r) o += S(q, f, r);
      m.src_bucket_off[q] = o;
    }
    m.recv_total = acc;
    for (int q = W; q <= kMaxWorld; ++q) m.recv_base[q] = acc < cap ? acc : cap;
    m.recv_clamped = acc < cap ? acc : cap;
    if (acc > cap) raise_status(P.status, HB_STATUS_WINDOW_OVERFLOW);
    P.meta[f] = m;
  }
}

// ---- push ids ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) sh_push_ids_kernel(const __grid_constant__ ShParams P) {
  const int W = P.world;
  // work items: (f, r, chunk); chunk < ceil(max_nnz / kIdChunk)
  int item = blockIdx.x;
  for (;; item += gridDim.x) {
    // decode item -> (f, r, c) walking features (few); total_ctas carries the item count
    if (item >= P.total_ctas) break;
    int f = 0, rem = item;
    while (true) {
      const int per = W * ((P.f[f].max_nnz + kIdChunk - 1) / kIdChunk);
      if (rem < per) break;
      rem -= per;
      ++f;
    }
    const int chunks = (P.f[f].max_nnz + kIdChunk - 1) / kIdChunk;
    const int r = rem / chunks, c = rem % chunks;
    const ShFeatMeta& m = P.meta[f];
    const int beg = m.send_off[r] + c * kIdChunk;
    int end = m.send_off[r + 1];
    if (beg >= end) continue;
    if (end > beg + kIdChunk) end = beg + kIdChunk;
    // clamp to the owner's capacity
    const int cap_r = P.f[f].cap;
    const int dst0 = m.remote_base[r] + c * kIdChunk;
    int cnt = end - beg;
    if (dst0 + cnt > cap_r) cnt = cap_r - dst0;
    if (cnt <= 0) continue;
    const int64_t* src = P.f[f].part_ids + beg;
    int64_t* dst = reinterpret_cast<int64_t*>(win(P, r) + P.f[f].ids_in_off[P.parity]) + dst0;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
      const int n2 = cnt >> 1;
      for (int i = threadIdx.x; i < n2; i += blockDim.x)
        reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(src)[i];
      if ((cnt & 1) && threadIdx.x == 0) dst[cnt - 1] = src[cnt - 1];
    } else {
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = src[i];
    }
  }
  signal_all_peers(P, 1, 2);
}

// ---- owner gather fused with the row push ---------------------------------------------
constexpr int kShRowsPerGroup = 4;

template <int V>
__global__ void __launch_bounds__(256) sh_owner_gather_kernel(const __grid_constant__ ShParams P) {
  wait_all_peers(P, 1);
  bool oob = false;
  // persistent CTAs: chunk -> (feature, first received position)
  for (int chunk_id = blockIdx.x; chunk_id < P.total_ctas; chunk_id += gridDim.x) {
    int lo = 0, hi = P.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (P.f[mid].cta_begin <= chunk_id) lo = mid; else hi = mid - 1;
    }
    const ShFeat& F = P.f[lo];
    const ShFeatMeta& m = P.meta[lo];
    const int log2g = F.log2g;
    const int groups = 256 >> log2g;
    const int g = threadIdx.x >> log2g;
    const int l = threadIdx.x & ((1 << log2g) - 1);
    const int dim = F.dim;
    const int p0 = (chunk_id - F.cta_begin) * groups * kShRowsPerGroup;
    const int total = m.recv_clamped;
    if (p0 >= total) continue;
    const int64_t* ids_in = reinterpret_cast<const int64_t*>(win(P, P.me) + F.ids_in_off[P.parity]);
    int col[V];
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      col[v] = ((v << log2g) + l) * 4;
      act[v] = col[v] < dim;
    }
    int64_t row[kShRowsPerGroup];
    int q[kShRowsPerGroup];
    int dstrow[kShRowsPerGroup];
    bool ok[kShRowsPerGroup];
    int64_t idv[kShRowsPerGroup];
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {  // all id loads in flight together
      const int p = p0 + u * groups + g;
      idv[u] = ids_in[p < total ? p : p0];
    }
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {
      const int p = p0 + u * groups + g;
      ok[u] = p < total;
      row[u] = -1; q[u] = 0; dstrow[u] = 0;
      if (ok[u]) {
        const int64_t id = idv[u];
        int64_t r = -1;
        if (id >= 0) r = P.div_shift >= 0 ? (int64_t)((uint64_t)id >> P.div_shift) : id / P.world;
        if ((uint64_t)r >= (uint64_t)F.shard_rows) { oob = true; r = -1; }
        row[u] = r;
        int qq = 0;
        while (qq + 1 < P.world && m.recv_base[qq + 1] <= p) ++qq;
        q[u] = qq;
        dstrow[u] = m.src_bucket_off[qq] + (p - m.recv_base[qq]);
      }
    }
    float4 val[kShRowsPerGroup][V];
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u)
#pragma unroll
      for (int v = 0; v < V; ++v) {
        val[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[u] && row[u] >= 0 && act[v])
          val[u][v] = ld_nc_f4(reinterpret_cast<const float4*>(F.shard + row[u] * dim + col[v]));
      }
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {
      if (!ok[u]) continue;
      float* dst = reinterpret_cast<float*>(win(P, q[u]) + F.rows_in_off) + (int64_t)dstrow[u] * dim;
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (act[v]) *reinterpret_cast<float4*>(dst + col[v]) = val[u][v];
    }
  }
  if (oob) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
  signal_all_peers(P, 2, 3);
}

// ---- backward: push row gradients to the owners ----------------------------------------
template <int V>
__global__ void __launch_bounds__(256) sh_push_grads_kernel(const __grid_constant__ ShParams P) {
  for (int chunk_id = blockIdx.x; chunk_id < P.total_ctas; chunk_id += gridDim.x) {
    int lo = 0, hi = P.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (P.f[mid].cta_begin <= chunk_id) lo = mid; else hi = mid - 1;
    }
    const ShFeat& F = P.f[lo];
    const ShFeatMeta& m = P.meta[lo];
    const int log2g = F.log2g;
    const int groups = 256 >> log2g;
    const int g = threadIdx.x >> log2g;
    const int l = threadIdx.x & ((1 << log2g) - 1);
    const int dim = F.dim;
    const int p0 = (chunk_id - F.cta_begin) * groups * kShRowsPerGroup;
    if (p0 >= F.nnz) continue;
    const bool scaled = F.offsets != nullptr && F.combiner != HB_SUM;
    int col[V];
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      col[v] = ((v << log2g) + l) * 4;
      act[v] = col[v] < dim;
    }
    int bag[kShRowsPerGroup], r[kShRowsPerGroup], drow[kShRowsPerGroup];
    float sc[kShRowsPerGroup];
    bool ok[kShRowsPerGroup];
    int jv[kShRowsPerGroup];
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {  // independent loads first
      const int p = p0 + u * groups + g;
      const int pp = p < F.nnz ? p : p0;
      jv[u] = F.part_idx[pp];
      bag[u] = F.offsets != nullptr ? F.bag_of_pos[pp] : pp;
    }
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {
      const int p = p0 + u * groups + g;
      ok[u] = p < F.nnz;
      r[u] = 0; drow[u] = 0; sc[u] = 1.0f;
      if (ok[u]) {
        if (scaled) {
          const int64_t c = F.offsets[bag[u] + 1] - F.offsets[bag[u]];
          sc[u] = (F.combiner == HB_MEAN) ? (float)c : __fsqrt_rn((float)c);
        }
        const int j = jv[u];
        int rr = 0;
        while (rr + 1 < P.world && m.send_off[rr + 1] <= j) ++rr;
        r[u] = rr;
        drow[u] = m.remote_base[rr] + (j - m.send_off[rr]);
        if (drow[u] >= F.cap) ok[u] = false;  // overflow already flagged by the exchange
      }
    }
    float4 val[kShRowsPerGroup][V];
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u)
#pragma unroll
      for (int v = 0; v < V; ++v) {
        val[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[u] && act[v])
          val[u][v] = ld_nc_f4(reinterpret_cast<const float4*>(F.grad + (int64_t)bag[u] * F.grad_stride + col[v]));
      }
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {
      if (!ok[u]) continue;
      float* dst = reinterpret_cast<float*>(win(P, r[u]) + F.grads_in_off) + (int64_t)drow[u] * dim;
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (act[v]) {
          float4 x = val[u][v];
          if (scaled)
            x = make_float4(__fdiv_rn(x.x, sc[u]), __fdiv_rn(x.y, sc[u]), __fdiv_rn(x.z, sc[u]),
                            __fdiv_rn(x.w, sc[u]));
          *reinterpret_cast<float4*>(dst + col[v]) = x;
        }
    }
  }
  signal_all_peers(P, 3, 4);
}

// bag index of every id position (CSR features), thread per bag
__global__ void __launch_bounds__(256) sh_bag_map_kernel(const __grid_constant__ ShParams P) {
  int lo = 0, hi = P.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.f[mid].cta_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const ShFeat& F = P.f[lo];
  if (F.offsets == nullptr || F.bag_of_pos == nullptr) return;
  const int b = (blockIdx.x - F.cta_begin) * 256 + threadIdx.x;
  if (b >= F.nbags) return;
  const int64_t s = F.offsets[b], e = F.offsets[b + 1];
  if (s < 0 || e < s || e > F.nnz) { raise_status(P.status, HB_STATUS_BAD_OFFSETS); return; }
  for (int64_t p = s; p < e; ++p) F.bag_of_pos[p] = b;
}

static int ilog2c32(int64_t x) {
  int l = 0;
  while (((int64_t)1 << l) < x) ++l;
  return l;
}

static void sh_shape(int dim, int* log2g, int* v) {
  const int vecs = dim / 4;
  if (vecs <= 32) { *log2g = ilog2c32(vecs); *v = 1; return; }
  *log2g = 5;
  int vv = (vecs + 31) / 32, p = 1;
  while (p < vv) p <<= 1;
  *v = p;
}

struct WindowLayout {
  std::vector<uint64_t> ids_in[2], rows_in, grads_in;
  std::vector<int64_t> cap;
  uint64_t total;
};

static WindowLayout sh_window_layout(int world, int n, const int64_t* max_nnz, const int32_t* dims,
                                     double cf) {
  WindowLayout L;
  uint64_t o = 0;
  L.cap.resize(n);
  for (int k = 0; k < n; ++k) {
    double c = ceil(cf * (double)max_nnz[k]);
    int64_t cap = (int64_t)c;
    if (cap > (int64_t)world * max_nnz[k]) cap = (int64_t)world * max_nnz[k];
    if (cap < max_nnz[k] && world == 1) cap = max_nnz[k];
    if (cap < 1) cap = 1;
    L.cap[k] = cap;
  }
  for (int par = 0; par < 2; ++par) {
    L.ids_in[par].resize(n);
    for (int k = 0; k < n; ++k) { L.ids_in[par][k] = o; o = align_up(o + (uint64_t)L.cap[k] * 8, 256); }
  }
  L.rows_in.resize(n);
  for (int k = 0; k < n; ++k) { L.rows_in[k] = o; o = align_up(o + (uint64_t)max_nnz[k] * dims[k] * 4, 256); }
  L.grads_in.resize(n);
  for (int k = 0; k < n; ++k) { L.grads_in[k] = o; o = align_up(o + (uint64_t)L.cap[k] * dims[k] * 4, 256); }
  L.total = o;
  return L;
}

}  // namespace hb

struct hbShardedPlan {
  hbComm* comm;
  int n;
  double cf;
  std::vector<int64_t> max_nnz;
  std::vector<int32_t> dims;
  hb::WindowLayout layout;
  // local buffers (one allocation)
  unsigned char* local;
  size_t local_bytes;
  std::vector<int64_t*> part_ids;
  std::vector<int32_t*> part_idx, sizes, bag_of_pos;
  hb::ShFeatMeta* meta;
  void* part_ws;
  size_t part_ws_bytes;
  void* upd_ws;
  size_t upd_ws_bytes;
  uint32_t step;
  bool have_forward;
};

namespace hb {

static int fill_params(const hbShardedPlan* pl, const hbShardedFeature* feats, int c0, int nc,
                       ShParams* P, int32_t* d_status) {
  const hbComm* c = pl->comm;
  P->meta = pl->meta + c0;
  P->peers = peer_ptrs(c);
  P->window_off = control_bytes();
  P->status = d_status;
  P->n = nc;
  P->me = c->rank;
  P->world = c->world;
  P->parity = pl->step & 1;
  P->epoch = pl->step;
  P->total_ctas = 0;
  P->div_shift = ((c->world & (c->world - 1)) == 0) ? ilog2c32(c->world) : -1;
  for (int j = 0; j < nc; ++j) {
    const int k = c0 + j;
    ShFeat& F = P->f[j];
    const hbShardedFeature& f = feats[k];
    int v;
    F.part_ids = pl->part_ids[k];
    F.part_idx = pl->part_idx[k];
    F.sizes = pl->sizes[k];
    F.bag_of_pos = pl->bag_of_pos[k];
    F.ids_in_off[0] = pl->layout.ids_in[0][k];
    F.ids_in_off[1] = pl->layout.ids_in[1][k];
    F.rows_in_off = pl->layout.rows_in[k];
    F.grads_in_off = pl->layout.grads_in[k];
    F.cap = (int32_t)pl->layout.cap[k];
    F.dim = pl->dims[k];
    sh_shape(F.dim, &F.log2g, &v);
    F.max_nnz = (int32_t)pl->max_nnz[k];
    F.shard = f.shard;
    F.grad = f.grad;
    F.offsets = f.offsets;
    F.shard_rows = f.shard_rows;
    F.grad_stride = f.grad_stride;
    F.nnz = (int32_t)f.nnz;
    F.nbags = (int32_t)f.nbags;
    F.combiner = f.combiner;
    F.cta_begin = 0;
  }
  return HB_OK;
}

static int validate_sharded(const hbShardedPlan* pl, const hbShardedFeature* feats, bool backward) {
  for (int k = 0; k < pl->n; ++k) {
    const hbShardedFeature& f = feats[k];
    HB_REQUIRE(f.dim == pl->dims[k], "sharded: feature %d dim %d differs from the plan's %d", k, f.dim, pl->dims[k]);
    HB_REQUIRE(f.nnz >= 0 && f.nnz <= pl->max_nnz[k], "sharded: feature %d nnz %lld exceeds max_nnz %lld", k,
               (long long)f.nnz, (long long)pl->max_nnz[k]);
    HB_REQUIRE(f.nbags >= 0 && f.nbags <= INT32_MAX, "sharded: feature %d bad nbags", k);
    HB_REQUIRE(f.offsets != nullptr || f.nnz == f.nbags, "sharded: feature %d has no offsets, so nnz must equal nbags", k);
    HB_REQUIRE(f.shard_rows >= 0 && f.shard_rows < ((int64_t)1 << 32) - 2, "sharded: feature %d bad shard_rows", k);
    HB_REQUIRE(f.combiner >= HB_SUM && f.combiner <= HB_SQRTN, "sharded: feature %d bad combiner", k);
    HB_REQUIRE(f.shard != nullptr || f.shard_rows == 0, "sharded: feature %d null shard", k);
    HB_REQUIRE(f.ids != nullptr || f.nnz == 0, "sharded: feature %d null ids", k);
    if (!backward) {
      HB_REQUIRE((f.out != nullptr || f.nbags == 0) && f.out_stride >= f.dim && f.out_stride % 4 == 0,
                 "sharded: feature %d bad out / out_stride", k);
    } else {
      HB_REQUIRE((f.grad != nullptr || f.nbags == 0) && f.grad_stride >= f.dim && f.grad_stride % 4 == 0,
                 "sharded: feature %d bad grad / grad_stride", k);
    }
  }
  return HB_OK;
}

}  // namespace hb

extern "C" {

size_t hbShardedPlanWindowBytes(int world, int n, const int64_t* max_nnz, const int32_t* dims,
                                double capacity_factor) {
  if (world < 1 || n < 1 || !max_nnz || !dims) return 0;
  if (capacity_factor < 1.0) capacity_factor = 1.0;
  return (size_t)hb::sh_window_layout(world, n, max_nnz, dims, capacity_factor).total + 4096;
}

int hbShardedPlanCreate(hbComm* comm, int n, const int64_t* max_nnz, const int32_t* dims,
                        double capacity_factor, hbShardedPlan** plan) {
  using namespace hb;
  HB_REQUIRE(comm && comm->connected, "hbShardedPlanCreate: communicator not connected");
  HB_REQUIRE(n >= 1 && n <= kShMaxFeats, "hbShardedPlanCreate: n=%d not in [1,%d] sharded features per plan", n, kShMaxFeats);
  HB_REQUIRE(max_nnz && dims && plan, "hbShardedPlanCreate: null argument");
  HB_REQUIRE(comm->reserved_bytes == 0, "hbShardedPlanCreate: this communicator already hosts a plan");
  if (capacity_factor < 1.0) capacity_factor = 1.0;
  for (int k = 0; k < n; ++k) {
    HB_REQUIRE(max_nnz[k] >= 1 && max_nnz[k] <= INT32_MAX / 2, "hbShardedPlanCreate: bad max_nnz[%d]", k);
    HB_REQUIRE(dims[k] >= 4 && dims[k] % 4 == 0 && dims[k] <= 1024, "hbShardedPlanCreate: dim[%d]=%d must be a multiple of 4 in [4,1024]", k, dims[k]);
  }
  hbShardedPlan* pl = new hbShardedPlan();
  pl->comm = comm;
  pl->n = n;
  pl->cf = capacity_factor;
  pl->max_nnz.assign(max_nnz, max_nnz + n);
  pl->dims.assign(dims, dims + n);
  pl->layout = sh_window_layout(comm->world, n, max_nnz, dims, capacity_factor);
  for (int k = 0; k < n; ++k)
    HB_REQUIRE(pl->layout.cap[k] <= INT32_MAX / 2, "hbShardedPlanCreate: capacity of feature %d too large", k);
  if (pl->layout.total > comm->window_bytes) {
    set_last_error("hbShardedPlanCreate: window %zu B < %llu B needed (see hbShardedPlanWindowBytes)",
                   comm->window_bytes, (unsigned long long)pl->layout.total);
    delete pl;
    return HB_ERR_WORKSPACE;
  }
  // local buffers
  std::vector<int32_t> lens(n);
  for (int k = 0; k < n; ++k) lens[k] = (int32_t)max_nnz[k];
  size_t part_ws = 0;
  int rc = hbPartitionWorkspaceBytes(n, lens.data(), comm->world, &part_ws);
  if (rc != HB_OK) { delete pl; return rc; }
  std::vector<hbUpdateFeature> uf(n);
  for (int k = 0; k < n; ++k) {
    memset(&uf[k], 0, sizeof(hbUpdateFeature));
    uf[k].rows = ((int64_t)1 << 32) - 3;  // worst case number of radix passes
    uf[k].nnz = uf[k].nbags = pl->layout.cap[k];
    uf[k].dim = dims[k];
    uf[k].id_div = comm->world;
  }
  size_t upd_ws = 0;
  rc = hbGroupSparseUpdateWorkspaceBytes(n, uf.data(), &upd_ws);
  if (rc != HB_OK) { delete pl; return rc; }
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  std::vector<size_t> o_ids(n), o_idx(n), o_sizes(n), o_bag(n);
  for (int k = 0; k < n; ++k) {
    o_ids[k] = take((size_t)max_nnz[k] * 8);
    o_idx[k] = take((size_t)max_nnz[k] * 4);
    o_sizes[k] = take((size_t)kMaxWorld * 4);
    o_bag[k] = take((size_t)max_nnz[k] * 4);
  }
  const size_t o_meta = take(sizeof(ShFeatMeta) * n);
  const size_t o_pws = take(part_ws);
  const size_t o_uws = take(upd_ws);
  pl->local_bytes = o;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&pl->local), pl->local_bytes);
  if (e != cudaSuccess) {
    set_last_error("hbShardedPlanCreate: cudaMalloc(%zu) failed: %s", pl->local_bytes, cudaGetErrorString(e));
    delete pl;
    return HB_ERR_CUDA;
  }
  cudaMemset(pl->local, 0, pl->local_bytes);
  pl->part_ids.resize(n); pl->part_idx.resize(n); pl->sizes.resize(n); pl->bag_of_pos.resize(n);
  for (int k = 0; k < n; ++k) {
    pl->part_ids[k] = reinterpret_cast<int64_t*>(pl->local + o_ids[k]);
    pl->part_idx[k] = reinterpret_cast<int32_t*>(pl->local + o_idx[k]);
    pl->sizes[k] = reinterpret_cast<int32_t*>(pl->local + o_sizes[k]);
    pl->bag_of_pos[k] = reinterpret_cast<int32_t*>(pl->local + o_bag[k]);
  }
  pl->meta = reinterpret_cast<ShFeatMeta*>(pl->local + o_meta);
  pl->part_ws = pl->local + o_pws;
  pl->part_ws_bytes = part_ws;
  pl->upd_ws = pl->local + o_uws;
  pl->upd_ws_bytes = upd_ws;
  pl->step = 0;
  pl->have_forward = false;
  comm->reserved_bytes = align_up(pl->layout.total, 4096);
  cudaDeviceSynchronize();
  *plan = pl;
  return HB_OK;
}

int hbShardedPlanDestroy(hbShardedPlan* pl) {
  if (!pl) return HB_OK;
  cudaDeviceSynchronize();
  if (pl->local) cudaFree(pl->local);
  if (pl->comm) pl->comm->reserved_bytes = 0;
  delete pl;
  return HB_OK;
}

int hbShardedLookupForward(hbShardedPlan* pl, const hbShardedFeature* feats, int32_t* d_status,
                           hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(pl && feats, "hbShardedLookupForward: null argument");
  int rc = validate_sharded(pl, feats, false);
  if (rc != HB_OK) return rc;
  hbComm* c = pl->comm;
  const int n = pl->n, W = c->world;
  pl->step++;
  pl->have_forward = true;

  // K1: stable partition of every feature's ids by id % W
  {
    std::vector<const void*> in(n);
    std::vector<void*> out(n);
    std::vector<int32_t*> sz(n), ix(n);
    std::vector<int32_t> lens(n);
    for (int k = 0; k < n; ++k) {
      in[k] = feats[k].ids; out[k] = pl->part_ids[k]; sz[k] = pl->sizes[k]; ix[k] = pl->part_idx[k];
      lens[k] = (int32_t)feats[k].nnz;
    }
    rc = hbPartitionByModuloN(HB_I64, n, in.data(), lens.data(), W, out.data(), sz.data(), ix.data(),
                              pl->part_ws, pl->part_ws_bytes, stream_);
    if (rc != HB_OK) return rc;
  }
  HB_REQUIRE(n <= kShMaxFeats, "hbShardedLookupForward: more than %d sharded features per plan", kShMaxFeats);
  ShParams P;
  fill_params(pl, feats, 0, n, &P, d_status);
  // exchange
  {
    KernelScope ks(HB_K_SH_EXCHANGE, stream);
    sh_exchange_kernel<<<1, 256, 0, stream>>>(P);
  }
  HB_CUDA_OK(cudaGetLastError());
  // push ids
  {
    int items = 0;
    for (int k = 0; k < n; ++k) items += W * (int)((pl->max_nnz[k] + kIdChunk - 1) / kIdChunk);
    ShParams Q = P;
    Q.total_ctas = items;
    const int maxg = device_sm_count() * 4;
    KernelScope ks(HB_K_SH_PUSH_IDS, stream);
    sh_push_ids_kernel<<<items < maxg ? items : maxg, 256, 0, stream>>>(Q);
  }
  HB_CUDA_OK(cudaGetLastError());
  // owner gather + row push, one launch per V
  for (int V = 1; V <= 8; V <<= 1) {
    ShParams Q = P;
    int m = 0, ctas = 0;
    std::vector<int> idx;
    for (int k = 0; k < n; ++k) {
      int l2, v;
      sh_shape(pl->dims[k], &l2, &v);
      if (v != V) continue;
      Q.f[m] = P.f[k];
      Q.f[m].cta_begin = ctas;
      const int per = (256 >> l2) * kShRowsPerGroup;
      ctas += (int)((pl->layout.cap[k] + per - 1) / per);
      idx.push_back(k);
      ++m;
    }
    if (m == 0) continue;
    // meta must follow the compaction: use a per-V meta view only when contiguous
    HB_REQUIRE(m == n, "hbShardedLookupForward: all sharded features of a plan must share dim <= 128 or the same V");
    Q.n = m;
    Q.total_ctas = ctas;
    const int maxg = device_sm_count() * 8;
    const int grid = ctas < maxg ? ctas : maxg;
    KernelScope ks(HB_K_SH_OWNER_GATHER, stream);
    switch (V) {
      case 1: sh_owner_gather_kernel<1><<<grid, 256, 0, stream>>>(Q); break;
      case 2: sh_owner_gather_kernel<2><<<grid, 256, 0, stream>>>(Q); break;
      case 4: sh_owner_gather_kernel<4><<<grid, 256, 0, stream>>>(Q); break;
      default: sh_owner_gather_kernel<8><<<grid, 256, 0, stream>>>(Q); break;
    }
  }
  HB_CUDA_OK(cudaGetLastError());
  // stitch + pool from the local rows_in window
  {
    std::vector<hbLookupFeature> lf(n);
    std::vector<const int32_t*> idx32(n);
    unsigned char* mywin = c->base + control_bytes();
    for (int k = 0; k < n; ++k) {
      lf[k].table = reinterpret_cast<const float*>(mywin + pl->layout.rows_in[k]);
      lf[k].rows = feats[k].nnz;
      lf[k].ids = nullptr;
      lf[k].offsets = feats[k].offsets;
      lf[k].nbags = feats[k].nbags;
      lf[k].out = feats[k].out;
      lf[k].out_stride = feats[k].out_stride;
      lf[k].dim = feats[k].dim;
      lf[k].combiner = feats[k].combiner;
      lf[k].id_div = 1;
      idx32[k] = pl->part_idx[k];
    }
    Control* mine = reinterpret_cast<Control*>(c->base);
    WaitSpec w{&mine->plan_flags[2][0], pl->step, W};
    rc = lookup_forward_run(n, lf.data(), idx32.data(), &w, true, d_status, stream, HB_K_SH_STITCH);
    if (rc != HB_OK) return rc;
  }
  return HB_OK;
}

int hbShardedLookupBackwardUpdate(hbShardedPlan* pl, const hbShardedFeature* feats,
                                  const hbOptimizer* opt, int32_t* d_status, hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(pl && feats && opt, "hbShardedLookupBackwardUpdate: null argument");
  HB_REQUIRE(pl->have_forward, "hbShardedLookupBackwardUpdate: no forward to pair with");
  int rc = validate_sharded(pl, feats, true);
  if (rc != HB_OK) return rc;
  hbComm* c = pl->comm;
  const int n = pl->n, W = c->world;
  pl->have_forward = false;
  ShParams P;
  fill_params(pl, feats, 0, n, &P, d_status);
  // bag map for CSR features
  {
    ShParams Q = P;
    int ctas = 0;
    bool any = false;
    for (int k = 0; k < n; ++k) {
      Q.f[k].cta_begin = ctas;
      ctas += (int)((feats[k].nbags + 255) / 256) + 1;
      any = any || feats[k].offsets != nullptr;
    }
    if (any) {
      KernelScope ks(HB_K_BAG_MAP, stream);
      sh_bag_map_kernel<<<ctas, 256, 0, stream>>>(Q);
      HB_CUDA_OK(cudaGetLastError());
    }
  }
  // push row gradients to the owners
  for (int V = 1; V <= 8; V <<= 1) {
    ShParams Q = P;
    int m = 0, ctas = 0;
    for (int k = 0; k < n; ++k) {
      int l2, v;
      sh_shape(pl->dims[k], &l2, &v);
      if (v != V) continue;
      Q.f[m].cta_begin = ctas;
      const int per = (256 >> l2) * kShRowsPerGroup;
      ctas += (int)((pl->max_nnz[k] + per - 1) / per);
      ++m;
    }
    if (m == 0) continue;
    HB_REQUIRE(m == n, "hbShardedLookupBackwardUpdate: all sharded features of a plan must share V");
    Q.total_ctas = ctas;
    const int maxg = device_sm_count() * 8;
    const int grid = ctas < maxg ? ctas : maxg;
    KernelScope ks(HB_K_SH_PUSH_GRADS, stream);
    switch (V) {
      case 1: sh_push_grads_kernel<1><<<grid, 256, 0, stream>>>(Q); break;
      case 2: sh_push_grads_kernel<2><<<grid, 256, 0, stream>>>(Q); break;
      case 4: sh_push_grads_kernel<4><<<grid, 256, 0, stream>>>(Q); break;
      default: sh_push_grads_kernel<8><<<grid, 256, 0, stream>>>(Q); break;
    }
  }
  HB_CUDA_OK(cudaGetLastError());
  // owner side: sort the received ids, sum duplicates, apply the optimizer
  {
    std::vector<hbUpdateFeature> uf(n);
    unsigned char* mywin = c->base + control_bytes();
    const int par = pl->step & 1;
    for (int k = 0; k < n; ++k) {
      uf[k].table = feats[k].shard;
      uf[k].slot0 = feats[k].slot0;
      uf[k].slot1 = feats[k].slot1;
      uf[k].rows = feats[k].shard_rows;
      uf[k].ids = reinterpret_cast<const int64_t*>(mywin + pl->layout.ids_in[par][k]);
      uf[k].offsets = nullptr;
      uf[k].nbags = uf[k].nnz = pl->layout.cap[k];
      uf[k].grad = reinterpret_cast<const float*>(mywin + pl->layout.grads_in[k]);
      uf[k].grad_stride = pl->dims[k];
      uf[k].dim = pl->dims[k];
      uf[k].combiner = HB_SUM;
      uf[k].id_div = W;
    }
    Control* mine = reinterpret_cast<Control*>(c->base);
    WaitSpec w{&mine->plan_flags[3][0], pl->step, W};
    std::vector<const int32_t*> ndev(n);
    for (int k = 0; k < n; ++k) ndev[k] = &pl->meta[k].recv_clamped;
    rc = sparse_update_run(n, uf.data(), opt, pl->upd_ws, pl->upd_ws_bytes, d_status, stream, &w,
                           ndev.data());
    if (rc != HB_OK) return rc;
  }
  return HB_OK;
}

}  // extern "C"
