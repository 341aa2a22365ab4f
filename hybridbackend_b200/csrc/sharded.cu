// sharded.cu -- the fused sharded GroupLookup over NVSwitch peer memory.  It is the
// composition of embedding/sharding.py:171-203 (partition -> alltoallv(ids) ->
// unique -> gather -> alltoallv(rows) -> stitch) and of its backward with every
// intermediate kept on the device: no host synchronisation, static shapes, no NCCL,
// and only UNIQUE ids / rows / gradient sums on the wire.
//
// Forward, per rank and step:
//   requester sort   ids --radix sort by (id % W, id / W)--> runs            (K1)
//                    one scan -> unique ids grouped by owner (the partition of
//                    HbPartitionByModulo applied to the deduplicated ids), the
//                    inverse map position -> unique, per-owner counts
//   publish / meta   counts all-gather through peer mailboxes; every rank derives
//                    all offsets it needs from the W x F x W matrix
//   push_ids         my unique local rows of owner r -> r's ids_in window      (K2)
//   owner_gather     for every row I own that was asked for: read it from my shard
//                    and store it into the REQUESTER's rows_in window at the slot of
//                    its unique -- gather and return all-to-all are one kernel,
//                    tile by tile over NVLink                                 (K3+K2')
//   stitch_pool      out[b] = pool_p rows_in[inverse[p]]                        (K4)
// Backward:
//   emit             per unique id: sum of the row gradients of its positions (in
//                    position order) stored straight into the owner's grads_in
//                    window (K2'' fused with the requester half of the dedup)
//   owner update     radix sort of the received rows + fused dedup/optimizer    (K5)
// Requester-side dedup changes wire bytes only: the owner still adds the per-rank
// sums of a row in rank order (training/gradient.py:216-217: no 1/W).
// Cross-GPU ordering: epoch-valued flags (st.release.sys by the last CTA of the
// producer kernel, ld.acquire.sys spin at the start of the consumer kernel).  A
// producer only writes a peer's window after a flag wait that is stream-ordered
// behind the peer's last read of it, so every window region is single-buffered.
#include <math.h>
#include <string.h>

#include <vector>

#include "bucket.cuh"
#include "comm.cuh"
#include "update.cuh"

namespace hb {

constexpr int kShMaxFeats = kMaxA2aTensors;  // features per plan (mailbox rows)
constexpr int kShThreads = 256;
constexpr int kIdChunk = 2048;               // unique ids per push work unit

struct ShFeatMeta {
  int32_t send_off[kMaxWorld + 1];    // my bucket starts in my unique list
  int32_t remote_base[kMaxWorld];     // my segment start inside owner r's ids_in / grads_in
  int32_t recv_base[kMaxWorld + 1];   // as owner: start of source q's segment (clamped to cap)
  int32_t src_bucket_off[kMaxWorld];  // as owner: start of bucket `me` in q's unique list
  int32_t recv_total;                 // unclamped number of rows asked of me
  int32_t recv_clamped;               // min(recv_total, cap): entries actually received
};

// One feature of a plan; lives in DEVICE memory (n can be in the hundreds: C4 has
// 200 features), rewritten by one small H2D copy per call.
struct ShFeat {
  const uint32_t* ukey;        // requester: unique composite keys (sort workspace)
  const int32_t* counts;       // requester: [0] = number of uniques
  const int32_t* owner_start1; // requester: [W] first unique of owner r, +1 (0: none)
  const float* shard;          // owner: local shard
  int64_t shard_rows;
  uint64_t ids_in_off;         // byte offsets inside the data window
  uint64_t rows_in_off;
  uint64_t grads_in_off;
  int32_t cap;
  int32_t dim;
  int32_t log2g;
  int32_t lbits;
  int32_t max_nnz;
  int32_t pad;
};

struct ShParams {
  const ShFeat* feats;   // [n] device
  ShFeatMeta* meta;      // [n] device
  PeerPtrs peers;
  uint64_t window_off;   // control_bytes()
  int32_t* status;
  int32_t n, me, world;
  uint32_t epoch;
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

// returns false (and raises the status bit) when a peer never arrived
__device__ __forceinline__ bool wait_all_peers(const ShParams& P, int phase) {
  __shared__ int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  if ((int)threadIdx.x < P.world)
    if (!wait_flag(&ctl(P, P.me)->plan_flags[phase][threadIdx.x], P.epoch)) {
      s_timeout = 1;
      raise_status(P.status, HB_STATUS_PEER_TIMEOUT);
    }
  __syncthreads();
  return s_timeout == 0;
}

// dense work map over up to kShMaxFeats features (see sparse_update.cu seg_scan)
__device__ __forceinline__ int sh_seg_scan(int nsegs, int my_units, int* s_begin /*[kShMaxFeats + 1]*/) {
  if ((int)threadIdx.x < nsegs) s_begin[threadIdx.x] = my_units;
  __syncthreads();
  if (threadIdx.x < 32) {
    int carry = 0;
    for (int b = 0; b < nsegs; b += 32) {
      const int i = b + (int)threadIdx.x;
      const int v = i < nsegs ? s_begin[i] : 0;
      int incl = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, off);
        if ((int)threadIdx.x >= off) incl += y;
      }
      if (i < nsegs) s_begin[i] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) s_begin[nsegs] = carry;
  }
  __syncthreads();
  return s_begin[nsegs];
}
__device__ __forceinline__ int sh_seg_find(const int* s_begin, int nsegs, int unit) {
  int lo = 0, hi = nsegs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (s_begin[mid] <= unit) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---- publish: my per-owner unique counts -> every peer's mailbox (1 CTA) ---------------
__global__ void __launch_bounds__(kShThreads) sh_publish_kernel(const __grid_constant__ ShParams P) {
  const int W = P.world, me = P.me, n = P.n;
  const int par = P.epoch & 1;
  for (int f = threadIdx.x; f < n; f += blockDim.x) {
    const ShFeat& F = P.feats[f];
    // unique list is grouped by owner: bucket r = [start[r], start[r+1]); owners
    // without ids never wrote their start
    int nxt = F.counts[0];
    int32_t sz[kMaxWorld];
    for (int r = W - 1; r >= 0; --r) {
      const int s1 = F.owner_start1[r];
      const int st = s1 > 0 ? s1 - 1 : nxt;
      sz[r] = nxt - st;
      nxt = st;
    }
    for (int r = 0; r < W; ++r)
      for (int q = 0; q < W; ++q)
        ctl(P, q)->plan_mailbox[par][(me * kMaxA2aTensors + f) * kMaxWorld + r] = sz[r];
  }
  __syncthreads();
  if ((int)threadIdx.x < W) {
    __threadfence_system();
    st_release_sys_u32(&ctl(P, threadIdx.x)->plan_flags[0][me], P.epoch);
  }
}

// ---- meta: offset algebra from the all-gathered W x F x W matrix (1 CTA) ---------------
__global__ void __launch_bounds__(kShThreads) sh_meta_kernel(const __grid_constant__ ShParams P) {
  const int W = P.world, me = P.me, n = P.n;
  const int par = P.epoch & 1;
  Control* mine = ctl(P, me);
  wait_all_peers(P, 0);
  auto S = [&](int q, int f, int r) -> int32_t {
    return *reinterpret_cast<volatile int32_t*>(
        &mine->plan_mailbox[par][(q * kMaxA2aTensors + f) * kMaxWorld + r]);
  };
  for (int f = threadIdx.x; f < n; f += blockDim.x) {
    ShFeatMeta m;
    const int cap = P.feats[f].cap;
    int acc = 0;
    for (int r = 0; r < W; ++r) { m.send_off[r] = acc; acc += S(me, f, r); }
    for (int r = W; r <= kMaxWorld; ++r) m.send_off[r] = acc;
    for (int r = 0; r < kMaxWorld; ++r) {
      int b = 0;
      if (r < W)
        for (int q = 0; q < me; ++q) b += S(q, f, r);
      m.remote_base[r] = b;
    }
    acc = 0;
    for (int q = 0; q < kMaxWorld; ++q) {
      m.recv_base[q] = acc < cap ? acc : cap;
      int o = 0;
      if (q < W) {
        acc += S(q, f, me);
        for (int r = 0; r < me; ++r) o += S(q, f, r);
      }
      m.src_bucket_off[q] = o;
    }
    m.recv_base[kMaxWorld] = acc < cap ? acc : cap;
    m.recv_total = acc;
    m.recv_clamped = acc < cap ? acc : cap;
    // every rank sees the whole matrix: an overflow at ANY owner is raised on EVERY
    // rank (its requesters would otherwise read rows that were never gathered)
    bool ovf = false;
    for (int r = 0; r < W; ++r) {
      int t = 0;
      for (int q = 0; q < W; ++q) t += S(q, f, r);
      if (t > cap) ovf = true;
    }
    if (ovf) raise_status(P.status, HB_STATUS_WINDOW_OVERFLOW);
    P.meta[f] = m;
  }
}

// ---- push unique local rows to their owners -------------------------------------------
__global__ void __launch_bounds__(kShThreads) sh_push_ids_kernel(const __grid_constant__ ShParams P) {
  __shared__ int s_begin[kShMaxFeats + 1];
  const int W = P.world;
  int units = 0;
  if ((int)threadIdx.x < P.n) {
    const int U = P.meta[threadIdx.x].send_off[W];
    units = (U + kIdChunk - 1) / kIdChunk;
  }
  const int total = sh_seg_scan(P.n, units, s_begin);
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int f = sh_seg_find(s_begin, P.n, unit);
    const ShFeat& F = P.feats[f];
    const ShFeatMeta& m = P.meta[f];
    const int U = m.send_off[W];
    const uint32_t lmask = (F.lbits >= 32) ? 0xFFFFFFFFu : ((1u << F.lbits) - 1u);
    const int i0 = (unit - s_begin[f]) * kIdChunk;
#pragma unroll
    for (int j = 0; j < kIdChunk / kShThreads; ++j) {
      const int i = i0 + j * kShThreads + (int)threadIdx.x;
      if (i >= U) break;
      const uint32_t key = F.ukey[i];
      int r = 0;
      while (r + 1 < W && m.send_off[r + 1] <= i) ++r;
      const int dst = m.remote_base[r] + (i - m.send_off[r]);
      if (dst < F.cap)
        reinterpret_cast<uint32_t*>(win(P, r) + F.ids_in_off)[dst] = key & lmask;
    }
  }
  signal_all_peers(P, 1, 2);
}

// ---- owner gather fused with the row push ---------------------------------------------
constexpr int kShRowsPerGroup = 4;

__global__ void __launch_bounds__(kShThreads) sh_owner_gather_kernel(const __grid_constant__ ShParams P) {
  __shared__ int s_begin[kShMaxFeats + 1];
  const bool arrived = wait_all_peers(P, 1);
  int units = 0;
  if ((int)threadIdx.x < P.n && arrived) {
    const int per = (kShThreads >> P.feats[threadIdx.x].log2g) * kShRowsPerGroup;
    units = (P.meta[threadIdx.x].recv_clamped + per - 1) / per;
  }
  const int total = sh_seg_scan(P.n, units, s_begin);
  bool oob = false;
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int f = sh_seg_find(s_begin, P.n, unit);
    const ShFeat& F = P.feats[f];
    const ShFeatMeta& m = P.meta[f];
    const int log2g = F.log2g;
    const int G = 1 << log2g;
    const int groups = kShThreads >> log2g;
    const int g = threadIdx.x >> log2g;
    const int l = threadIdx.x & (G - 1);
    const int dim = F.dim;
    const int p0 = (unit - s_begin[f]) * groups * kShRowsPerGroup;
    const int cnt = m.recv_clamped;
    const uint32_t* ids_in = reinterpret_cast<const uint32_t*>(win(P, P.me) + F.ids_in_off);
    uint32_t idv[kShRowsPerGroup];
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {  // all id loads in flight together
      const int p = p0 + u * groups + g;
      idv[u] = ids_in[p < cnt ? p : p0];
    }
    int64_t row[kShRowsPerGroup];
    float* dst[kShRowsPerGroup];
#pragma unroll
    for (int u = 0; u < kShRowsPerGroup; ++u) {
      const int p = p0 + u * groups + g;
      row[u] = -1;
      dst[u] = nullptr;
      if (p < cnt) {
        int q = 0;
        while (q + 1 < P.world && m.recv_base[q + 1] <= p) ++q;
        dst[u] = reinterpret_cast<float*>(win(P, q) + F.rows_in_off) +
                 (int64_t)(m.src_bucket_off[q] + (p - m.recv_base[q])) * dim;
        if ((int64_t)idv[u] < F.shard_rows) row[u] = (int64_t)idv[u];
        else oob = true;  // the requester still gets a (zero) row
      }
    }
    // 16-byte column c of lane l: (l + k * G) * 4 floats, k = 0 .. ; dim <= 128 is one step
    for (int c = l * 4; c < dim; c += G * 4) {
      float4 val[kShRowsPerGroup];
#pragma unroll
      for (int u = 0; u < kShRowsPerGroup; ++u) {
        val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row[u] >= 0) val[u] = ld_nc_f4(reinterpret_cast<const float4*>(F.shard + row[u] * dim + c));
      }
#pragma unroll
      for (int u = 0; u < kShRowsPerGroup; ++u)
        if (dst[u] != nullptr) *reinterpret_cast<float4*>(dst[u] + c) = val[u];
    }
  }
  if (oob) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
  signal_all_peers(P, 2, 3);
}

// ---- backward: the gradient sums are in the owners' windows -> raise the flag ----------
__global__ void __launch_bounds__(32) sh_signal_kernel(const __grid_constant__ ShParams P, int phase) {
  // launched behind the emit kernels on the same stream: their peer stores are
  // complete; one system fence, then the release stores
  if ((int)threadIdx.x < P.world) {
    __threadfence_system();
    st_release_sys_u32(&ctl(P, threadIdx.x)->plan_flags[phase][P.me], P.epoch);
  }
}

static int ilog2c32(int64_t x) {
  int l = 0;
  while (((int64_t)1 << l) < x) ++l;
  return l;
}

static void sh_shape(int dim, int* log2g) {
  const int vecs = dim / 4;
  *log2g = vecs <= 32 ? ilog2c32(vecs) : 5;
}

struct WindowLayout {
  std::vector<uint64_t> ids_in, rows_in, grads_in;
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
  L.ids_in.resize(n);
  for (int k = 0; k < n; ++k) { L.ids_in[k] = o; o = align_up(o + (uint64_t)L.cap[k] * 4, 256); }
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
  std::vector<int32_t*> inv;
  int32_t* owner_start1;   // [n][kMaxWorld], zeroed per step
  hb::ShFeat* d_feats;
  hb::ShFeatMeta* meta;
  void* req_ws;            // requester: sort / runs workspace (kept from forward to backward)
  size_t req_ws_bytes;
  void* own_ws;            // owner: sort / runs / apply workspace
  size_t own_ws_bytes;
  std::vector<hb::UpdViews> req_views;
  uint32_t epoch;          // epoch of the last forward
  bool have_forward;
  // the owner's sort of the received rows needs only ids_in, complete once the owner gather
  // has started: it runs on a side stream next to the stitch instead of in the backward
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  bool owner_sorted;
};

namespace hb {

static int validate_sharded(const hbShardedPlan* pl, const hbShardedFeature* feats, bool backward) {
  for (int k = 0; k < pl->n; ++k) {
    const hbShardedFeature& f = feats[k];
    HB_REQUIRE(f.dim == pl->dims[k], "sharded: feature %d dim %d differs from the plan's %d", k, f.dim, pl->dims[k]);
    HB_REQUIRE(f.nnz >= 0 && f.nnz <= pl->max_nnz[k], "sharded: feature %d nnz %lld exceeds max_nnz %lld", k,
               (long long)f.nnz, (long long)pl->max_nnz[k]);
    HB_REQUIRE(f.nbags >= 0 && f.nbags <= INT32_MAX, "sharded: feature %d bad nbags", k);
    HB_REQUIRE(f.offsets != nullptr || f.nnz == f.nbags, "sharded: feature %d has no offsets, so nnz must equal nbags", k);
    HB_REQUIRE(f.shard_rows >= 0 && f.shard_rows < ((int64_t)1 << 31) / pl->comm->world, "sharded: feature %d bad shard_rows", k);
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

static void requester_jobs(const hbShardedPlan* pl, const hbShardedFeature* feats, bool backward,
                           std::vector<hbUpdateFeature>* uf, std::vector<UpdExtra>* ex) {
  const int n = pl->n, W = pl->comm->world;
  uf->resize(n);
  ex->resize(n);
  for (int k = 0; k < n; ++k) {
    hbUpdateFeature& u = (*uf)[k];
    memset(&u, 0, sizeof(u));
    // any owner's shard has at most my rows + 1 (embedding/variables.py:107-109)
    u.rows = feats[k].shard_rows + 1;
    u.ids = feats[k].ids;
    u.offsets = feats[k].offsets;
    u.nbags = feats[k].nbags;
    u.nnz = feats[k].nnz;
    u.grad = backward ? feats[k].grad : nullptr;
    u.grad_stride = backward ? feats[k].grad_stride : feats[k].dim;
    u.dim = feats[k].dim;
    u.combiner = feats[k].combiner;
    u.id_div = W;
    UpdExtra& e = (*ex)[k];
    e.key_kind = 1;
    e.lbits = ilog2c32(u.rows + 1);
    e.inv = pl->inv[k];
    e.owner_start1 = pl->owner_start1 + (size_t)k * kMaxWorld;
    e.emit_send_off = pl->meta[k].send_off;
    e.emit_remote_base = pl->meta[k].remote_base;
    e.emit_off = pl->layout.grads_in[k];
    e.emit_cap = (int32_t)pl->layout.cap[k];
  }
}

// the owner-side update job: rows received in ids_in / grads_in, counts on the device
static void owner_jobs(const hbShardedPlan* pl, const hbShardedFeature* feats, std::vector<hbUpdateFeature>* of,
                       std::vector<UpdExtra>* oe) {
  const int n = pl->n, W = pl->comm->world;
  of->resize(n);
  oe->resize(n);
  unsigned char* mywin = pl->comm->base + control_bytes();
  for (int k = 0; k < n; ++k) {
    hbUpdateFeature& f = (*of)[k];
    memset(&f, 0, sizeof(hbUpdateFeature));
    f.table = feats[k].shard;
    f.slot0 = feats[k].slot0;
    f.slot1 = feats[k].slot1;
    f.rows = feats[k].shard_rows;
    f.ids = nullptr;
    f.offsets = nullptr;
    f.nbags = f.nnz = pl->layout.cap[k];
    f.grad = reinterpret_cast<const float*>(mywin + pl->layout.grads_in[k]);
    f.grad_stride = pl->dims[k];
    f.dim = pl->dims[k];
    f.combiner = HB_SUM;
    f.id_div = W;
    (*oe)[k].key_kind = 2;
    (*oe)[k].keys32 = reinterpret_cast<const uint32_t*>(mywin + pl->layout.ids_in[k]);
    (*oe)[k].n_dev = &pl->meta[k].recv_clamped;
  }
}

static ShParams sh_params(const hbShardedPlan* pl, int32_t* d_status) {
  const hbComm* c = pl->comm;
  ShParams P;
  P.feats = pl->d_feats;
  P.meta = pl->meta;
  P.peers = peer_ptrs(c);
  P.window_off = control_bytes();
  P.status = d_status;
  P.n = pl->n;
  P.me = c->rank;
  P.world = c->world;
  P.epoch = pl->epoch;
  return P;
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
    HB_REQUIRE(max_nnz[k] >= 1 && max_nnz[k] <= INT32_MAX / 4, "hbShardedPlanCreate: bad max_nnz[%d]", k);
    HB_REQUIRE(dims[k] >= 4 && dims[k] % 4 == 0 && dims[k] <= 1024, "hbShardedPlanCreate: dim[%d]=%d must be a multiple of 4 in [4,1024]", k, dims[k]);
  }
  WindowLayout layout = sh_window_layout(comm->world, n, max_nnz, dims, capacity_factor);
  for (int k = 0; k < n; ++k)
    HB_REQUIRE(layout.cap[k] <= INT32_MAX / 4, "hbShardedPlanCreate: capacity of feature %d too large", k);
  if (layout.total > comm->window_bytes) {
    set_last_error("hbShardedPlanCreate: window %zu B < %llu B needed (see hbShardedPlanWindowBytes)",
                   comm->window_bytes, (unsigned long long)layout.total);
    return HB_ERR_WORKSPACE;
  }
  // workspaces: requester (max_nnz entries, bags possible) and owner (cap entries)
  static const int64_t kDummy = 0;
  std::vector<hbUpdateFeature> rq(n), ow(n);
  for (int k = 0; k < n; ++k) {
    memset(&rq[k], 0, sizeof(hbUpdateFeature));
    rq[k].nnz = rq[k].nbags = max_nnz[k];
    rq[k].offsets = &kDummy;  // size for the CSR form
    rq[k].dim = dims[k];
    ow[k] = rq[k];
    ow[k].offsets = nullptr;
    ow[k].nnz = ow[k].nbags = layout.cap[k];
  }
  const size_t req_ws = sparse_update_workspace_bytes(n, rq.data());
  const size_t own_ws = sparse_update_workspace_bytes(n, ow.data());
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  std::vector<size_t> o_inv(n);
  for (int k = 0; k < n; ++k) o_inv[k] = take((size_t)max_nnz[k] * 4);
  const size_t o_start = take((size_t)n * kMaxWorld * 4);
  const size_t o_feats = take(sizeof(ShFeat) * n);
  const size_t o_meta = take(sizeof(ShFeatMeta) * n);
  const size_t o_rws = take(req_ws);
  const size_t o_ows = take(own_ws);
  hbShardedPlan* pl = new hbShardedPlan();
  pl->comm = comm;
  pl->n = n;
  pl->cf = capacity_factor;
  pl->max_nnz.assign(max_nnz, max_nnz + n);
  pl->dims.assign(dims, dims + n);
  pl->layout = layout;
  pl->local_bytes = o;
  pl->local = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&pl->local), pl->local_bytes);
  if (e == cudaSuccess) e = cudaMemset(pl->local, 0, pl->local_bytes);
  if (e != cudaSuccess) {
    set_last_error("hbShardedPlanCreate: cudaMalloc(%zu) failed: %s", pl->local_bytes, cudaGetErrorString(e));
    if (pl->local) cudaFree(pl->local);
    delete pl;
    return HB_ERR_CUDA;
  }
  pl->inv.resize(n);
  for (int k = 0; k < n; ++k) pl->inv[k] = reinterpret_cast<int32_t*>(pl->local + o_inv[k]);
  pl->owner_start1 = reinterpret_cast<int32_t*>(pl->local + o_start);
  pl->d_feats = reinterpret_cast<ShFeat*>(pl->local + o_feats);
  pl->meta = reinterpret_cast<ShFeatMeta*>(pl->local + o_meta);
  pl->req_ws = pl->local + o_rws;
  pl->req_ws_bytes = req_ws;
  pl->own_ws = pl->local + o_ows;
  pl->own_ws_bytes = own_ws;
  pl->req_views.resize(n);
  pl->epoch = 0;
  pl->have_forward = false;
  pl->owner_sorted = false;
  pl->side = nullptr;
  pl->ev_fork = pl->ev_join = nullptr;
  if (cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    set_last_error("hbShardedPlanCreate: cannot create the side stream");
    cudaFree(pl->local);
    delete pl;
    return HB_ERR_CUDA;
  }
  comm->reserved_bytes = align_up(pl->layout.total, 4096);
  cudaDeviceSynchronize();
  *plan = pl;
  return HB_OK;
}

int hbShardedPlanDestroy(hbShardedPlan* pl) {
  if (!pl) return HB_OK;
  cudaDeviceSynchronize();
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->ev_join) cudaEventDestroy(pl->ev_join);
  if (pl->side) cudaStreamDestroy(pl->side);
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
  // the epoch lives in the communicator: flags left behind by an earlier plan on the
  // same communicator can never satisfy a wait of this one
  pl->epoch = ++c->plan_epoch;
  pl->have_forward = true;
  std::vector<hbUpdateFeature> uf;
  std::vector<UpdExtra> ex;
  requester_jobs(pl, feats, false, &uf, &ex);
  const int sms = device_sm_count();

  return comm_submit(c, kOpShardedFwd, 4, stream, [&](int phase) -> int {
    int rc2 = HB_OK;
    if (phase == 0) {
      // requester: sort by (owner, local row), find the unique ids, publish the counts
      HB_CUDA_OK(cudaMemsetAsync(pl->owner_start1, 0, (size_t)n * kMaxWorld * 4, stream));
      rc2 = sparse_update_run(n, uf.data(), nullptr, pl->req_ws, pl->req_ws_bytes, d_status, stream,
                              nullptr, ex.data(), nullptr, kPhaseSort, pl->req_views.data());
      if (rc2 != HB_OK) return rc2;
      std::vector<ShFeat> hf(n);
      for (int k = 0; k < n; ++k) {
        ShFeat& F = hf[k];
        F.ukey = pl->req_views[k].ukey;
        F.counts = pl->req_views[k].counts;
        F.owner_start1 = pl->owner_start1 + (size_t)k * kMaxWorld;
        F.shard = feats[k].shard;
        F.shard_rows = feats[k].shard_rows;
        F.ids_in_off = pl->layout.ids_in[k];
        F.rows_in_off = pl->layout.rows_in[k];
        F.grads_in_off = pl->layout.grads_in[k];
        F.cap = (int32_t)pl->layout.cap[k];
        F.dim = pl->dims[k];
        sh_shape(F.dim, &F.log2g);
        F.lbits = ex[k].lbits;
        F.max_nnz = (int32_t)pl->max_nnz[k];
        F.pad = 0;
      }
      // pageable source: staged by the driver before the call returns
      HB_CUDA_OK(cudaMemcpyAsync(pl->d_feats, hf.data(), sizeof(ShFeat) * n, cudaMemcpyHostToDevice, stream));
      ShParams P = sh_params(pl, d_status);
      KernelScope ks(HB_K_SH_PUBLISH, stream);
      sh_publish_kernel<<<1, kShThreads, 0, stream>>>(P);
      HB_CUDA_OK(cudaGetLastError());
      return HB_OK;
    }
    ShParams P = sh_params(pl, d_status);
    if (phase == 1) {
      {
        KernelScope ks(HB_K_SH_EXCHANGE, stream);
        sh_meta_kernel<<<1, kShThreads, 0, stream>>>(P);
      }
      HB_CUDA_OK(cudaGetLastError());
      int64_t units = 0;
      for (int k = 0; k < n; ++k) units += (feats[k].nnz + kIdChunk - 1) / kIdChunk;
      const int maxg = sms * 4;
      const int grid = units < 1 ? 1 : (units < maxg ? (int)units : maxg);
      KernelScope ks(HB_K_SH_PUSH_IDS, stream);
      sh_push_ids_kernel<<<grid, kShThreads, 0, stream>>>(P);
      HB_CUDA_OK(cudaGetLastError());
      return HB_OK;
    }
    if (phase == 2) {
      {
        KernelScope ks(HB_K_SH_OWNER_GATHER, stream);
        sh_owner_gather_kernel<<<sms * 8, kShThreads, 0, stream>>>(P);
        HB_CUDA_OK(cudaGetLastError());
      }
      // ids_in is complete behind the gather's flag wait: sort it for the backward now, on
      // the side stream (joined by the owner apply)
      std::vector<hbUpdateFeature> of;
      std::vector<UpdExtra> oe;
      owner_jobs(pl, feats, &of, &oe);
      HB_CUDA_OK(cudaEventRecord(pl->ev_fork, stream));
      HB_CUDA_OK(cudaStreamWaitEvent(pl->side, pl->ev_fork, 0));
      int rc3 = sparse_update_run(n, of.data(), nullptr, pl->own_ws, pl->own_ws_bytes, d_status, pl->side, nullptr,
                                  oe.data(), nullptr, kPhaseSort, nullptr);
      if (rc3 != HB_OK) return rc3;
      HB_CUDA_OK(cudaEventRecord(pl->ev_join, pl->side));
      pl->owner_sorted = true;
      return HB_OK;
    }
    // stitch + pool from the local rows_in window through the inverse map
    std::vector<hbLookupFeature> lf(n);
    std::vector<const int32_t*> idx32(n);
    unsigned char* mywin = c->base + control_bytes();
    for (int k = 0; k < n; ++k) {
      lf[k].table = reinterpret_cast<const float*>(mywin + pl->layout.rows_in[k]);
      lf[k].rows = pl->max_nnz[k];
      lf[k].ids = nullptr;
      lf[k].offsets = feats[k].offsets;
      lf[k].nbags = feats[k].nbags;
      lf[k].out = feats[k].out;
      lf[k].out_stride = feats[k].out_stride;
      lf[k].dim = feats[k].dim;
      lf[k].combiner = feats[k].combiner;
      lf[k].id_div = 1;
      lf[k].nnz = feats[k].nnz;
      idx32[k] = pl->inv[k];
    }
    Control* mine = reinterpret_cast<Control*>(c->base);
    WaitSpec w{&mine->plan_flags[2][0], pl->epoch, W};
    return lookup_forward_run(n, lf.data(), idx32.data(), &w, true, d_status, stream, HB_K_SH_STITCH);
  });
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
  std::vector<hbUpdateFeature> uf;
  std::vector<UpdExtra> ex;
  requester_jobs(pl, feats, true, &uf, &ex);

  return comm_submit(c, kOpShardedBwd, 2, stream, [&](int phase) -> int {
    ShParams P = sh_params(pl, d_status);
    if (phase == 0) {
      // requester: per-unique gradient sums straight into the owners' windows
      EmitCtx em;
      em.peers = P.peers;
      em.window_off = P.window_off;
      em.world = W;
      int rc2 = sparse_update_run(n, uf.data(), nullptr, pl->req_ws, pl->req_ws_bytes, d_status, stream,
                                  nullptr, ex.data(), &em, kPhaseApply, nullptr);
      if (rc2 != HB_OK) return rc2;
      KernelScope ks(HB_K_SH_PUSH_GRADS, stream);
      sh_signal_kernel<<<1, 32, 0, stream>>>(P, 3);
      HB_CUDA_OK(cudaGetLastError());
      return HB_OK;
    }
    // owner: the received rows were sorted at forward time (side stream); sum duplicates
    // across ranks, apply the optimizer
    std::vector<hbUpdateFeature> of;
    std::vector<UpdExtra> oe;
    owner_jobs(pl, feats, &of, &oe);
    int phases = kPhaseApply;
    if (pl->owner_sorted) HB_CUDA_OK(cudaStreamWaitEvent(stream, pl->ev_join, 0));
    else phases |= kPhaseSort;
    pl->owner_sorted = false;
    Control* mine = reinterpret_cast<Control*>(c->base);
    WaitSpec w{&mine->plan_flags[3][0], pl->epoch, W};
    return sparse_update_run(n, of.data(), opt, pl->own_ws, pl->own_ws_bytes, d_status, stream, &w,
                             oe.data(), nullptr, phases, nullptr);
  });
}

}  // extern "C"
