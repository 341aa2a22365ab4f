// sparse_update.cu -- K5: backward of the pooled lookup fused with the sparse
// optimizer apply (Adagrad / LazyAdam / SGD), deterministic, no float atomics.
//
// Pipeline (all features of the group per launch):
//   1. bag_of_position : CSR features only -- bag index of every id position.
//   2. LSD radix sort  : (local row, bag) pairs grouped by row, 9-bit digits,
//                        stable, so entries of a row stay in position order
//                        (bucket.cuh; 1..4 passes depending on the table size).
//   3. update kernel   : a sub-warp group of G lanes walks a tile of C sorted
//                        entries; gradient rows of a run (same row id) are summed
//                        in registers in position order and the optimizer is
//                        applied once per unique row (one read-modify-write of
//                        the table row and its slot rows).  Runs that cross tile
//                        borders are combined in shared memory inside the CTA
//                        ("super-tile"), in tile order.
//   4. fix-up kernel   : rows that span super-tiles (hot keys) are finished by
//                        chaining the per-super-tile partial sums, in order.
// Summation order is a fixed function of the input => bit-reproducible.
//
// Algorithmic HBM bytes per looked-up id without duplicates (Adagrad, fp32):
//   8 (id) + 4*dim (grad) + 2*4*dim (w, acc read) + 2*4*dim (w, acc write).
#include "bucket.cuh"

namespace hb {

constexpr int kUpdThreads = 256;
constexpr int kMaxUpdFeats = 96;

enum { kFirstOpen = 1, kBoth = 2, kLastOpen = 4 };

struct UpdFeat {
  float* table;
  float* slot0;
  float* slot1;
  const float* grad;
  const int64_t* offsets;   // bag sizes for mean/sqrtn (nullptr: one id per bag)
  const uint32_t* keys;     // sorted local rows
  const int32_t* bags;      // bag index of each sorted entry
  float* st_part;           // [nst][2][dim] partial sums of open runs
  uint32_t* st_key;         // [nst][2]
  int32_t* st_flag;         // [nst]
  const int32_t* n_dev;     // may be nullptr: device-side number of entries (<= n)
  int64_t rows;
  int64_t grad_stride;
  int32_t n;
  int32_t dim;
  int32_t combiner;
  int32_t log2g;
  int32_t cta_begin;        // first super-tile (== CTA) of this feature
  int32_t nst;              // number of super-tiles
};

struct UpdParams {
  UpdFeat f[kMaxUpdFeats];
  WaitSpec wait;
  int32_t* status;
  int32_t nfeats;
  int32_t total_ctas;
  int32_t opt;
  float lr;        // Adagrad/SGD: lr ; LazyAdam: bias-corrected lr_t
  float beta1, beta2, eps;
  float omb1, omb2;
};

__device__ __forceinline__ int find_upd_feat(const UpdParams& P, int cta) {
  int lo = 0, hi = P.nfeats - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.f[mid].cta_begin <= cta) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add_rn(const float4& a, const float4& b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z),
                     __fadd_rn(a.w, b.w));
}

// One optimizer step on 4 consecutive elements of a row (values in registers).
template <int OPT>
__device__ __forceinline__ void opt_step4(const UpdParams& P, float4& wv, float4& s0v, float4& s1v,
                                          const float4& g) {
  const float gg[4] = {g.x, g.y, g.z, g.w};
  float ww[4] = {wv.x, wv.y, wv.z, wv.w};
  if constexpr (OPT == HB_OPT_ADAGRAD) {
    // SparseApplyAdagrad: accum += g*g ; var -= lr*g/sqrt(accum)
    float aa[4] = {s0v.x, s0v.y, s0v.z, s0v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      aa[i] = __fadd_rn(aa[i], __fmul_rn(gg[i], gg[i]));
      ww[i] = __fsub_rn(ww[i], __fdiv_rn(__fmul_rn(P.lr, gg[i]), __fsqrt_rn(aa[i])));
    }
    s0v = make_float4(aa[0], aa[1], aa[2], aa[3]);
  } else if constexpr (OPT == HB_OPT_LAZY_ADAM) {
    // LazyAdam: m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; var -= lr_t*m/(sqrt(v)+eps)
    float mm[4] = {s0v.x, s0v.y, s0v.z, s0v.w};
    float v2[4] = {s1v.x, s1v.y, s1v.z, s1v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mm[i] = __fadd_rn(__fmul_rn(P.beta1, mm[i]), __fmul_rn(P.omb1, gg[i]));
      v2[i] = __fadd_rn(__fmul_rn(P.beta2, v2[i]), __fmul_rn(P.omb2, __fmul_rn(gg[i], gg[i])));
      ww[i] = __fsub_rn(ww[i], __fdiv_rn(__fmul_rn(P.lr, mm[i]),
                                         __fadd_rn(__fsqrt_rn(v2[i]), P.eps)));
    }
    s0v = make_float4(mm[0], mm[1], mm[2], mm[3]);
    s1v = make_float4(v2[0], v2[1], v2[2], v2[3]);
  } else {  // SGD
#pragma unroll
    for (int i = 0; i < 4; ++i) ww[i] = __fsub_rn(ww[i], __fmul_rn(P.lr, gg[i]));
  }
  wv = make_float4(ww[0], ww[1], ww[2], ww[3]);
}

// Apply the optimizer to up to N rows at once: all row loads (table + slots) are
// issued before the first use, so N*(1+slots) 128-bit loads are in flight per lane.
template <int V, int N, int OPT>
__device__ __forceinline__ bool apply_rows(const UpdParams& P, const UpdFeat& F, unsigned mask,
                                           const uint32_t (&key)[N], const float4 (&g)[N][V],
                                           const int (&col)[V], const bool (&act)[V]) {
  bool all_ok = true;
  bool ok[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    ok[i] = (mask >> i) & 1u;
    if (ok[i] && key[i] == 0xFFFFFFFFu) ok[i] = false;  // padding entry (sharded path)
    if (ok[i] && (uint64_t)key[i] >= (uint64_t)F.rows) { ok[i] = false; all_ok = false; }
  }
  float4 w[N][V], s0[N][V], s1[N][V];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int v = 0; v < V; ++v) {
      w[i][v] = s0[i][v] = s1[i][v] = f4_zero();
      if (ok[i] && act[v]) {
        const int64_t o = (int64_t)key[i] * F.dim + col[v];
        w[i][v] = *reinterpret_cast<const float4*>(F.table + o);
        if constexpr (OPT != HB_OPT_SGD) s0[i][v] = *reinterpret_cast<const float4*>(F.slot0 + o);
        if constexpr (OPT == HB_OPT_LAZY_ADAM) s1[i][v] = *reinterpret_cast<const float4*>(F.slot1 + o);
      }
    }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (ok[i] && act[v]) {
        const int64_t o = (int64_t)key[i] * F.dim + col[v];
        opt_step4<OPT>(P, w[i][v], s0[i][v], s1[i][v], g[i][v]);
        *reinterpret_cast<float4*>(F.table + o) = w[i][v];
        if constexpr (OPT != HB_OPT_SGD) *reinterpret_cast<float4*>(F.slot0 + o) = s0[i][v];
        if constexpr (OPT == HB_OPT_LAZY_ADAM) *reinterpret_cast<float4*>(F.slot1 + o) = s1[i][v];
      }
  return all_ok;
}

template <int V, int OPT>
__device__ __forceinline__ bool apply_row(const UpdParams& P, const UpdFeat& F, uint32_t key,
                                          const float4 (&acc)[V], const int (&col)[V],
                                          const bool (&act)[V]) {
  const uint32_t k1[1] = {key};
  float4 g1[1][V];
#pragma unroll
  for (int v = 0; v < V; ++v) g1[0][v] = acc[v];
  return apply_rows<V, 1, OPT>(P, F, 1u, k1, g1, col, act);
}

// A warp owns 32 consecutive sorted entries (lane e holds key/bag of entry e: one
// coalesced load each); the run structure of the warp tile is a pair of ballot
// masks.  Each of the 32/G groups walks its G consecutive entries in sub-batches
// of SB: the gradient rows AND the table/slot rows of the runs ending in the
// sub-batch are loaded together (one memory round trip), summed in position
// order, and applied.  smem per CTA: groups * (2*V*G float4 partials + 2 keys + flag).
template <int V, int OPT>
__global__ void __launch_bounds__(kUpdThreads, (V == 1 ? 3 : 1))
sparse_update_kernel(const __grid_constant__ UpdParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  wait_spec(P.wait, P.status);
  const int fi = find_upd_feat(P, blockIdx.x);
  const UpdFeat& F = P.f[fi];
  const int st = blockIdx.x - F.cta_begin;
  const int log2g = F.log2g;
  const int G = 1 << log2g;
  const int groups = kUpdThreads >> log2g;
  const int g = threadIdx.x >> log2g;
  const int l = threadIdx.x & (G - 1);
  const int dim = F.dim;
  int n = F.n;
  if (F.n_dev != nullptr) { const int d = *F.n_dev; n = d < 0 ? 0 : (d < n ? d : n); }
  // smem carve-up
  float4* s_part = reinterpret_cast<float4*>(s_raw);                  // [groups][2][V][G]
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_part + (size_t)groups * 2 * V * G);  // [groups][2]
  int32_t* s_flag = reinterpret_cast<int32_t*>(s_key + groups * 2);   // [groups]

  int col[V];
  bool act[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    col[v] = ((v << log2g) + l) * 4;
    act[v] = col[v] < dim;
  }
  bool oob = false;
  int flag = 0;

  // ---- warp tile: 32 entries, lane e <-> entry e -------------------------------
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int64_t w0 = ((int64_t)st * (kUpdThreads / 32) + warp) * 32;  // first entry of the warp
  const int wcnt = (int)max((int64_t)0, min((int64_t)32, (int64_t)n - w0));
  uint32_t my_key = 0xFFFFFFFFu;
  int32_t my_bag = 0;
  if ((int)lane < wcnt) { my_key = F.keys[w0 + lane]; my_bag = F.bags[w0 + lane]; }
  // neighbours of the warp tile
  uint32_t edge_key = 0;
  bool edge_has = false;
  if (lane == 0 && wcnt > 0 && w0 > 0) { edge_key = F.keys[w0 - 1]; edge_has = true; }
  if (lane == 31 && wcnt == 32 && w0 + 32 < n) { edge_key = F.keys[w0 + 32]; edge_has = true; }
  float my_scale = 1.0f;
  const bool scaled = F.combiner != HB_SUM && F.offsets != nullptr;
  if (scaled && (int)lane < wcnt) {
    const int64_t c = F.offsets[my_bag + 1] - F.offsets[my_bag];
    my_scale = (F.combiner == HB_MEAN) ? (float)c : __fsqrt_rn((float)c);
  }
  // same_prev bit e: entry e has the same key as entry e-1 (globally)
  const uint32_t up_key = __shfl_up_sync(0xffffffffu, my_key, 1);
  bool sp = false;
  if ((int)lane < wcnt) sp = (lane == 0) ? (edge_has && edge_key == my_key) : (up_key == my_key);
  const unsigned same_prev = __ballot_sync(0xffffffffu, sp);
  // does the run of the warp's last entry continue in the next warp tile?
  const unsigned cont_next = __ballot_sync(0xffffffffu, lane == 31 && edge_has && edge_key == my_key);
  const bool warp_open_right = cont_next != 0;

  // ---- group tile: C = G consecutive entries of the warp tile --------------------
  constexpr int SB = (V == 1) ? 4 : (V == 2 ? 2 : 1);
  const int gw = (lane >> log2g);     // group index inside the warp
  const int c0 = gw * G;              // first warp-entry of my group
  const int cnt = max(0, min(G, wcnt - c0));
  float4* my_part = s_part + (size_t)g * 2 * V * G;
  if (wcnt > 0) {  // warp-uniform: every lane takes part in the shuffles below
    const bool first_open_left = (same_prev >> c0) & 1u;
    bool seen_tail = false;
    float4 carry[V];
#pragma unroll
    for (int v = 0; v < V; ++v) carry[v] = f4_zero();
    // trip count G/SB is the same for every group of the warp (groups with fewer
    // valid entries run predicated-off iterations)
#pragma unroll 1
    for (int j0 = 0; j0 < G; j0 += SB) {
      uint32_t k[SB];
      bool valid[SB], head[SB], is_apply[SB];
      int part_slot[SB];  // -1 none, 0 first-open partial, 1 last-open partial
      int part_flag[SB];
      unsigned apply_mask = 0;
      float4 gv[SB][V];
      float sc[SB];
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        const int e = c0 + j0 + i;                 // entry index inside the warp tile
        valid[i] = (j0 + i) < cnt;
        const int src = valid[i] ? e : c0;
        k[i] = __shfl_sync(0xffffffffu, my_key, src);
        const int32_t bag = __shfl_sync(0xffffffffu, my_bag, src);
        sc[i] = __shfl_sync(0xffffffffu, my_scale, src);
        head[i] = !((same_prev >> src) & 1u);
        is_apply[i] = false;
        part_slot[i] = -1;
        part_flag[i] = 0;
        if (valid[i]) {
          const bool last = (j0 + i == cnt - 1);
          // tail: the next entry (inside the warp tile, or beyond it) starts a new run
          bool next_same;
          if (e + 1 < wcnt) next_same = (same_prev >> (e + 1)) & 1u;
          else next_same = warp_open_right;        // e is the warp tile's last entry
          const bool tail = !next_same || last;
          if (tail) {
            const bool ol = first_open_left && !seen_tail;
            const bool orr = last && next_same;
            if (!ol && !orr) { is_apply[i] = true; apply_mask |= 1u << i; }
            else if (ol) { part_slot[i] = 0; part_flag[i] = kFirstOpen | (orr ? kBoth : 0); }
            else { part_slot[i] = 1; part_flag[i] = kLastOpen; }
            seen_tail = true;
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            gv[i][v] = f4_zero();
            if (act[v])
              gv[i][v] = ld_nc_f4(reinterpret_cast<const float4*>(
                  F.grad + (int64_t)bag * F.grad_stride + col[v]));
          }
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) gv[i][v] = f4_zero();
        }
      }
      // table / slot rows of the runs that end (closed) in this sub-batch: issued
      // right behind the gradient loads, consumed after the sums
      bool ok[SB];
      float4 w[SB][V], s0[SB][V], s1[SB][V];
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        ok[i] = is_apply[i];
        if (ok[i] && k[i] == 0xFFFFFFFFu) ok[i] = false;                      // padding entry
        if (ok[i] && (uint64_t)k[i] >= (uint64_t)F.rows) { ok[i] = false; oob = true; }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          w[i][v] = s0[i][v] = s1[i][v] = f4_zero();
          if (ok[i] && act[v]) {
            const int64_t o = (int64_t)k[i] * dim + col[v];
            w[i][v] = *reinterpret_cast<const float4*>(F.table + o);
            if constexpr (OPT != HB_OPT_SGD) s0[i][v] = *reinterpret_cast<const float4*>(F.slot0 + o);
            if constexpr (OPT == HB_OPT_LAZY_ADAM) s1[i][v] = *reinterpret_cast<const float4*>(F.slot1 + o);
          }
        }
      }
      if (scaled) {
#pragma unroll
        for (int i = 0; i < SB; ++i)
#pragma unroll
          for (int v = 0; v < V; ++v)
            gv[i][v] = make_float4(__fdiv_rn(gv[i][v].x, sc[i]), __fdiv_rn(gv[i][v].y, sc[i]),
                                   __fdiv_rn(gv[i][v].z, sc[i]), __fdiv_rn(gv[i][v].w, sc[i]));
      }
      // segmented inclusive sums, in position order
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        if (valid[i]) {
          const bool cont = (j0 + i == 0) ? false : !head[i];  // continues a run of THIS group tile
          if (cont) {
#pragma unroll
            for (int v = 0; v < V; ++v)
              gv[i][v] = f4_add_rn(i == 0 ? carry[v] : gv[i > 0 ? i - 1 : 0][v], gv[i][v]);
          }
        }
      }
      // run ends: apply or park the partial sum
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        if (ok[i]) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (act[v]) {
              const int64_t o = (int64_t)k[i] * dim + col[v];
              opt_step4<OPT>(P, w[i][v], s0[i][v], s1[i][v], gv[i][v]);
              *reinterpret_cast<float4*>(F.table + o) = w[i][v];
              if constexpr (OPT != HB_OPT_SGD) *reinterpret_cast<float4*>(F.slot0 + o) = s0[i][v];
              if constexpr (OPT == HB_OPT_LAZY_ADAM) *reinterpret_cast<float4*>(F.slot1 + o) = s1[i][v];
            }
        } else if (part_slot[i] >= 0) {
#pragma unroll
          for (int v = 0; v < V; ++v) my_part[(part_slot[i] * V + v) * G + l] = gv[i][v];
          if (l == 0) s_key[g * 2 + part_slot[i]] = k[i];
          flag |= part_flag[i];
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) carry[v] = gv[SB - 1][v];
    }
  }
  __shared__ int s_stflag;
  if (threadIdx.x == 0) s_stflag = 0;
  if (l == 0) s_flag[g] = flag;
  __syncthreads();

  // ---- combine runs crossing tile borders inside the super-tile -------------
  float* gpart = F.st_part + (size_t)st * 2 * dim;
  int st_flag = 0;
  if (g == 0 && (flag & kFirstOpen)) {
    // chain entering from the previous super-tile
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = s_part[(size_t)(0 * 2 + 0) * V * G + v * G + l];
    int t = 0;
    bool both = true;
    while (true) {
      if (!(s_flag[t] & kBoth)) { both = false; break; }
      ++t;
      if (t == groups || !(s_flag[t] & kFirstOpen)) break;  // leaves the super-tile
#pragma unroll
      for (int v = 0; v < V; ++v)
        acc[v] = f4_add_rn(acc[v], s_part[(size_t)(t * 2 + 0) * V * G + v * G + l]);
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (act[v]) *reinterpret_cast<float4*>(gpart + 0 * dim + col[v]) = acc[v];
    if (l == 0) F.st_key[(size_t)st * 2 + 0] = s_key[0];
    st_flag |= kFirstOpen | (both ? kBoth : 0);
  }
  if (flag & kLastOpen) {
    // chain starting in my tile
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = s_part[(size_t)(g * 2 + 1) * V * G + v * G + l];
    const uint32_t key = s_key[g * 2 + 1];
    int t = g + 1;
    bool closed = false;
    while (t < groups) {
      if (!(s_flag[t] & kFirstOpen)) break;  // (empty tail tile) -- cannot happen when open
#pragma unroll
      for (int v = 0; v < V; ++v)
        acc[v] = f4_add_rn(acc[v], s_part[(size_t)(t * 2 + 0) * V * G + v * G + l]);
      if (!(s_flag[t] & kBoth)) { closed = true; break; }
      ++t;
    }
    if (closed) {
      if (!apply_row<V, OPT>(P, F, key, acc, col, act)) oob = true;
    } else {
      // still open at the end of the super-tile
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (act[v]) *reinterpret_cast<float4*>(gpart + 1 * dim + col[v]) = acc[v];
      if (l == 0) {
        F.st_key[(size_t)st * 2 + 1] = key;
        atomicOr(&s_stflag, kLastOpen);
      }
    }
  }
  if (g == 0 && l == 0 && st_flag) atomicOr(&s_stflag, st_flag);
  if (oob) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
  __syncthreads();
  if (threadIdx.x == 0) F.st_flag[st] = s_stflag;
}

// Finish rows that span super-tiles (hot keys): one WARP per super-tile that
// starts a chain.  Per round the warp reads the flags of the next 32 super-tiles
// with one coalesced load and finds the chain end with a ballot; its 32/G groups
// then sum the partial rows t+gi, t+gi+ng, ... (independent loads), and the group
// sums are combined in a fixed order with shuffles.  Order of addition is a fixed
// function of the chain length => deterministic.
template <int V, int OPT>
__global__ void __launch_bounds__(kUpdThreads)
sparse_update_fixup_kernel(const __grid_constant__ UpdParams P) {
  const int fi = find_upd_feat(P, blockIdx.x);
  const UpdFeat& F = P.f[fi];
  const int log2g = F.log2g;
  const int G = 1 << log2g;
  const int ng = 32 >> log2g;            // groups per warp
  const unsigned lane = lane_id();
  const int gi = lane >> log2g;          // my group inside the warp
  const int l = lane & (G - 1);
  const int dim = F.dim;
  const int st = (blockIdx.x - F.cta_begin) * (kUpdThreads / 32) + (threadIdx.x >> 5);
  if (st >= F.nst) return;               // warp-uniform
  if (!(F.st_flag[st] & kLastOpen)) return;
  int col[V];
  bool act[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    col[v] = ((v << log2g) + l) * 4;
    act[v] = col[v] < dim;
  }
  // group 0 starts from the chain head's partial, the others from zero
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v)
    acc[v] = (gi == 0 && act[v])
                 ? *reinterpret_cast<const float4*>(F.st_part + ((size_t)st * 2 + 1) * dim + col[v])
                 : f4_zero();
  const uint32_t key = F.st_key[(size_t)st * 2 + 1];
  int t = st + 1;
  bool done = false;
  while (!done && t < F.nst) {
    // flags of super-tiles t .. t+31
    const int fl = (t + (int)lane < F.nst) ? F.st_flag[t + lane] : 0;
    const unsigned is_first = __ballot_sync(0xffffffffu, (fl & kFirstOpen) != 0);
    const unsigned is_both = __ballot_sync(0xffffffffu, (fl & kBoth) != 0);
    // chain covers tiles while FirstOpen; it ends after the first one without Both
    const unsigned stop_a = ~is_first;            // tile does not continue the chain at all
    const unsigned stop_b = is_first & ~is_both;  // last tile of the chain (included)
    int m;                                        // number of tiles of this round to add
    const int pa = stop_a ? __ffs(stop_a) - 1 : 32;
    const int pb = stop_b ? __ffs(stop_b) - 1 : 32;
    if (pb < pa) { m = pb + 1; done = true; }
    else { m = pa; if (pa < 32) done = true; }
    // my group adds tiles t+gi, t+gi+ng, ... (< t+m), 8 loads in flight
    for (int u0 = gi; u0 < m; u0 += ng * 8) {
      float4 x[8][V];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int u = u0 + j * ng;
#pragma unroll
        for (int v = 0; v < V; ++v)
          x[j][v] = (u < m && act[v])
                        ? *reinterpret_cast<const float4*>(F.st_part + ((size_t)(t + u) * 2 + 0) * dim + col[v])
                        : f4_zero();
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (u0 + j * ng < m)
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = f4_add_rn(acc[v], x[j][v]);
    }
    t += 32;
  }
  // combine the group sums in a fixed order: lanes of group gi add group gi+off
#pragma unroll
  for (int v = 0; v < V; ++v) {
    for (int off = ng >> 1; off >= 1; off >>= 1) {
      float4 y;
      y.x = __shfl_down_sync(0xffffffffu, acc[v].x, off << log2g);
      y.y = __shfl_down_sync(0xffffffffu, acc[v].y, off << log2g);
      y.z = __shfl_down_sync(0xffffffffu, acc[v].z, off << log2g);
      y.w = __shfl_down_sync(0xffffffffu, acc[v].w, off << log2g);
      if (gi < off) acc[v] = f4_add_rn(acc[v], y);
    }
  }
  if (gi == 0)
    if (!apply_row<V, OPT>(P, F, key, acc, col, act)) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
}

// bag index of every id position, for CSR features (thread per bag).
struct BagMapFeat {
  const int64_t* offsets;
  int32_t* bag_of_pos;
  int32_t nbags;
  int32_t cta_begin;
  int64_t nnz;
};
struct BagMapParams {
  BagMapFeat f[kMaxUpdFeats];
  int32_t* status;
  int32_t nfeats;
};

__global__ void __launch_bounds__(256) bag_of_position_kernel(const __grid_constant__ BagMapParams P) {
  int lo = 0, hi = P.nfeats - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.f[mid].cta_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const BagMapFeat& F = P.f[lo];
  const int b = (blockIdx.x - F.cta_begin) * 256 + threadIdx.x;
  if (b >= F.nbags) return;
  int64_t s = F.offsets[b], e = F.offsets[b + 1];
  if (s < 0 || e < s || e > F.nnz) { raise_status(P.status, HB_STATUS_BAD_OFFSETS); return; }
  for (int64_t p = s; p < e; ++p) F.bag_of_pos[p] = b;
}

static int ilog2c(int64_t x) {
  int l = 0;
  while (((int64_t)1 << l) < x) ++l;
  return l;
}

static void upd_shape(int dim, int* log2g, int* v) {
  const int vecs = dim / 4;
  if (vecs <= 32) { *log2g = ilog2c(vecs); *v = 1; return; }
  *log2g = 5;
  int vv = (vecs + 31) / 32, p = 1;
  while (p < vv) p <<= 1;
  *v = p;
}

constexpr int kRadixBits = 9;
constexpr int kRadixBins = 1 << kRadixBits;


// per-feature workspace layout
struct UpdLayout {
  size_t keysA, keysB, valsA, valsB, bagmap, st_part, st_key, st_flag, end;
  int passes, log2g, V, C, nst;
};

static UpdLayout upd_layout(const hbUpdateFeature& f, size_t base) {
  UpdLayout L;
  const size_t n = (size_t)f.nnz;
  size_t o = base;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.keysA = take(4 * n); L.keysB = take(4 * n);
  L.valsA = take(4 * n); L.valsB = take(4 * n);
  L.bagmap = take(f.offsets ? 4 * n : 0);
  upd_shape(f.dim, &L.log2g, &L.V);
  L.C = 1 << L.log2g;  // entries per group: a warp owns 32 entries, a CTA 256
  const int64_t per_st = kUpdThreads;
  L.nst = (int)((f.nnz + per_st - 1) / per_st);
  L.st_part = take((size_t)L.nst * 2 * f.dim * 4);
  L.st_key = take((size_t)L.nst * 2 * 4);
  L.st_flag = take((size_t)L.nst * 4);
  const int64_t local_rows = f.rows;
  const int bits = local_rows > 1 ? ilog2c(local_rows) : 1;
  L.passes = (bits + kRadixBits - 1) / kRadixBits;
  if (L.passes < 1) L.passes = 1;
  L.end = o;
  return L;
}

template <int V, int OPT>
static int launch_update_opt(const UpdParams& P, const UpdParams& X, cudaStream_t stream) {
  // smem: groups*(2*V*G*16 + 12) with groups*G == 256
  const size_t smem = (size_t)2 * V * kUpdThreads * 16 + (size_t)kUpdThreads * 12;
  if (smem > 48 * 1024)
    HB_CUDA_OK(cudaFuncSetAttribute(sparse_update_kernel<V, OPT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (P.total_ctas > 0) {
    KernelScope ks(HB_K_SPARSE_UPDATE, stream);
    sparse_update_kernel<V, OPT><<<P.total_ctas, kUpdThreads, smem, stream>>>(P);
  }
  HB_CUDA_OK(cudaGetLastError());
  if (X.total_ctas > 0) {
    KernelScope ks(HB_K_SPARSE_FIXUP, stream);
    sparse_update_fixup_kernel<V, OPT><<<X.total_ctas, kUpdThreads, 0, stream>>>(X);
  }
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

template <int V>
static int launch_update(const UpdParams& P, const UpdParams& X, cudaStream_t stream) {
  switch (P.opt) {
    case HB_OPT_ADAGRAD: return launch_update_opt<V, HB_OPT_ADAGRAD>(P, X, stream);
    case HB_OPT_LAZY_ADAM: return launch_update_opt<V, HB_OPT_LAZY_ADAM>(P, X, stream);
    default: return launch_update_opt<V, HB_OPT_SGD>(P, X, stream);
  }
}

enum { kPhaseSort = 1, kPhaseApply = 2 };

static int validate_upd(int k, const hbUpdateFeature& f, const hbOptimizer* opt, int phases) {
  HB_REQUIRE(f.dim >= 4 && f.dim % 4 == 0 && f.dim <= 1024,
             "update: feature %d dim %d must be a multiple of 4 in [4,1024]", k, f.dim);
  HB_REQUIRE(f.nnz >= 0 && f.nnz <= INT32_MAX && f.nbags >= 0 && f.nbags <= INT32_MAX,
             "update: feature %d bad nnz/nbags", k);
  HB_REQUIRE(f.offsets != nullptr || f.nnz == f.nbags,
             "update: feature %d has no offsets, so nnz must equal nbags", k);
  HB_REQUIRE(f.rows >= 0 && f.rows < ((int64_t)1 << 32) - 1, "update: feature %d rows out of range", k);
  HB_REQUIRE(f.id_div >= 1, "update: feature %d id_div must be >= 1", k);
  HB_REQUIRE(f.grad_stride >= f.dim && f.grad_stride % 4 == 0,
             "update: feature %d grad_stride must be a multiple of 4 and >= dim", k);
  HB_REQUIRE(f.combiner >= HB_SUM && f.combiner <= HB_SQRTN, "update: feature %d bad combiner", k);
  if (f.nnz > 0 && (phases & kPhaseSort)) HB_REQUIRE(f.ids != nullptr, "update: feature %d null ids", k);
  if (f.nnz > 0 && (phases & kPhaseApply)) {
    HB_REQUIRE(f.table && f.grad, "update: feature %d null pointer", k);
    HB_REQUIRE(((uintptr_t)f.table & 15) == 0 && ((uintptr_t)f.grad & 15) == 0,
               "update: feature %d table/grad must be 16-byte aligned", k);
    if (opt->kind == HB_OPT_ADAGRAD)
      HB_REQUIRE(f.slot0 != nullptr, "update: feature %d Adagrad needs slot0 (accumulator)", k);
    if (opt->kind == HB_OPT_LAZY_ADAM)
      HB_REQUIRE(f.slot0 != nullptr && f.slot1 != nullptr, "update: feature %d LazyAdam needs slot0/slot1", k);
  }
  return HB_OK;
}

// Sort (row, bag) pairs of every feature and run the fused update.  `id_div`
// is shared by the features of one call (1 locally, W on a row-interleaved shard).
int sparse_update_run(int n, const hbUpdateFeature* feats, const hbOptimizer* opt, void* ws,
                      size_t ws_bytes, int32_t* d_status, cudaStream_t stream,
                      const WaitSpec* wait, const int32_t* const* n_dev, int phases) {
  static const hbOptimizer kNoOpt = {HB_OPT_SGD, 0.f, 0.f, 0.f, 0.f, 1};
  if (!(phases & kPhaseApply) && opt == nullptr) opt = &kNoOpt;
  HB_REQUIRE(n >= 1 && feats && opt, "update: bad arguments");
  HB_REQUIRE(opt->kind >= HB_OPT_SGD && opt->kind <= HB_OPT_LAZY_ADAM, "update: bad optimizer kind %d", opt->kind);
  for (int k = 0; k < n; ++k) {
    int rc = validate_upd(k, feats[k], opt, phases);
    if (rc != HB_OK) return rc;
    HB_REQUIRE(feats[k].id_div == feats[0].id_div, "update: all features of a call must share id_div");
  }
  size_t need = 0;
  int rc = hbGroupSparseUpdateWorkspaceBytes(n, feats, &need);
  if (rc != HB_OK) return rc;
  if (ws_bytes < need || (need > 0 && ws == nullptr)) {
    set_last_error("update: workspace %zu < required %zu bytes", ws_bytes, need);
    return HB_ERR_WORKSPACE;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);

  for (int c0 = 0; c0 < n; c0 += kMaxUpdFeats) {
    const int nc = (n - c0 < kMaxUpdFeats) ? n - c0 : kMaxUpdFeats;
    // layouts
    UpdLayout L[kMaxUpdFeats];
    size_t off = 0;
    {
      // recompute the running offset of this chunk
      for (int k = 0; k < c0; ++k) off = upd_layout(feats[k], off).end;
    }
    size_t o = off;
    int total_tiles = 0, max_passes = 0;
    for (int k = 0; k < nc; ++k) {
      L[k] = upd_layout(feats[c0 + k], o);
      o = L[k].end;
      total_tiles += bucket_tiles(feats[c0 + k].nnz);
      if (feats[c0 + k].nnz > 0 && L[k].passes > max_passes) max_passes = L[k].passes;
    }
    int32_t* counts = reinterpret_cast<int32_t*>(base + o);  // chunk-shared radix counters

    // 1. bag map for CSR features
    if (phases & kPhaseSort) {
      BagMapParams B;
      B.status = d_status;
      B.nfeats = 0;
      int ctas = 0;
      for (int k = 0; k < nc; ++k) {
        const hbUpdateFeature& f = feats[c0 + k];
        if (f.offsets == nullptr || f.nbags == 0) continue;
        BagMapFeat& bf = B.f[B.nfeats++];
        bf.offsets = f.offsets;
        bf.bag_of_pos = reinterpret_cast<int32_t*>(base + L[k].bagmap);
        bf.nbags = (int32_t)f.nbags;
        bf.cta_begin = ctas;
        bf.nnz = f.nnz;
        ctas += (int)((f.nbags + 255) / 256);
      }
      if (ctas > 0) {
        KernelScope ks(HB_K_BAG_MAP, stream);
        bag_of_position_kernel<<<ctas, 256, 0, stream>>>(B);
        HB_CUDA_OK(cudaGetLastError());
      }
    }

    // 2. LSD radix sort: one memset, one histogram kernel over the ids (all digit
    //    positions), then one kernel per digit position; feature k takes part in
    //    pass p iff p < passes[k]
    if (max_passes > 0 && (phases & kPhaseSort)) {
      const size_t words = bucket_scratch_words(nc, (size_t)total_tiles, kRadixBins, max_passes);
      HB_CUDA_OK(cudaMemsetAsync(counts, 0, words * sizeof(uint32_t), stream));
      BucketScratch sc = bucket_scratch_carve(reinterpret_cast<uint32_t*>(counts), nc,
                                              (size_t)total_tiles, kRadixBins, max_passes);
      for (int p = -1; p < max_passes; ++p) {  // p == -1: histogram launch
        BucketParams bp;
        bp.nsegs = 0;
        int tiles = 0;
        for (int k = 0; k < nc; ++k) {
          const hbUpdateFeature& f = feats[c0 + k];
          if (f.nnz == 0 || (p >= 0 && p >= L[k].passes)) continue;
          BucketSeg& s = bp.seg[bp.nsegs++];
          uint32_t* kA = reinterpret_cast<uint32_t*>(base + L[k].keysA);
          uint32_t* kB = reinterpret_cast<uint32_t*>(base + L[k].keysB);
          int32_t* vA = reinterpret_cast<int32_t*>(base + L[k].valsA);
          int32_t* vB = reinterpret_cast<int32_t*>(base + L[k].valsB);
          if (p <= 0) {
            s.in_keys = f.ids;
            s.in_vals = f.offsets ? reinterpret_cast<int32_t*>(base + L[k].bagmap) : nullptr;
            s.out_keys = kA;
            s.out_vals = vA;
          } else {
            const bool a2b = (p & 1) == 1;  // pass 1: A->B, pass 2: B->A, ...
            s.in_keys = a2b ? kA : kB;
            s.in_vals = a2b ? vA : vB;
            s.out_keys = a2b ? kB : kA;
            s.out_vals = a2b ? vB : vA;
          }
          s.out_inv = nullptr;
          s.out_sizes = nullptr;
          s.n = (int32_t)f.nnz;
          s.n_dev = n_dev ? n_dev[c0 + k] : nullptr;
          s.tile_begin = tiles;
          s.shift = (p < 0 ? 0 : p) * kRadixBits;
          s.key_limit = (uint32_t)f.rows;
          s.hist_slot = k;
          s.passes = L[k].passes;
          tiles += bucket_tiles(f.nnz);
        }
        if (bp.nsegs == 0) break;
        bp.hist = sc.hist;
        bp.status = sc.status[p < 0 ? 0 : p];
        bp.ticket = sc.ticket[p < 0 ? 0 : p];
        bp.nbins = kRadixBins;
        bp.total_tiles = tiles;
        bp.npass = max_passes;
        bp.pass = p < 0 ? 0 : p;
        bp.digit_bits = kRadixBits;
        bp.p = 1; bp.m = 1; bp.pow2_mask = 0;
        bp.div = feats[0].id_div;
        bp.div_shift = ((bp.div & (bp.div - 1)) == 0 && bp.div <= (1 << 30)) ? ilog2c(bp.div) : -1;
        if (p < 0) rc = bucket_hist_launch<RadixFirstTraits>(bp, stream, HB_K_SORT_HIST);
        else if (p == 0) rc = bucket_pass_launch<RadixFirstTraits>(bp, stream, HB_K_SORT_PASS);
        else rc = bucket_pass_launch<RadixNextTraits>(bp, stream, HB_K_SORT_PASS);
        if (rc != HB_OK) return rc;
      }
    }

    // 3./4. fused update + fix-up, one launch per V
    if (!(phases & kPhaseApply)) continue;
    for (int V = 1; V <= 8; V <<= 1) {
      UpdParams U;
      U.wait = wait ? *wait : WaitSpec{nullptr, 0, 0};
      U.status = d_status;
      U.nfeats = 0;
      U.total_ctas = 0;
      U.opt = opt->kind;
      U.lr = opt->lr;
      U.beta1 = opt->beta1; U.beta2 = opt->beta2; U.eps = opt->eps;
      U.omb1 = 1.0f - opt->beta1; U.omb2 = 1.0f - opt->beta2;
      if (opt->kind == HB_OPT_LAZY_ADAM) {
        const double t = (double)(opt->step < 1 ? 1 : opt->step);
        U.lr = (float)((double)opt->lr * sqrt(1.0 - pow((double)opt->beta2, t)) /
                       (1.0 - pow((double)opt->beta1, t)));
      }
      int fix_ctas = 0;
      int fix_begin[kMaxUpdFeats];
      for (int k = 0; k < nc; ++k) {
        const hbUpdateFeature& f = feats[c0 + k];
        if (L[k].V != V || f.nnz == 0) continue;
        UpdFeat& F = U.f[U.nfeats];
        const bool inA = (L[k].passes & 1) == 1;  // 1 pass -> A, 2 -> B, 3 -> A, ...
        F.table = f.table; F.slot0 = f.slot0; F.slot1 = f.slot1;
        F.grad = f.grad; F.offsets = f.offsets;
        F.keys = reinterpret_cast<uint32_t*>(base + (inA ? L[k].keysA : L[k].keysB));
        F.bags = reinterpret_cast<int32_t*>(base + (inA ? L[k].valsA : L[k].valsB));
        F.st_part = reinterpret_cast<float*>(base + L[k].st_part);
        F.st_key = reinterpret_cast<uint32_t*>(base + L[k].st_key);
        F.st_flag = reinterpret_cast<int32_t*>(base + L[k].st_flag);
        F.n_dev = n_dev ? n_dev[c0 + k] : nullptr;
        F.rows = f.rows; F.grad_stride = f.grad_stride;
        F.n = (int32_t)f.nnz; F.dim = f.dim; F.combiner = f.combiner;
        F.log2g = L[k].log2g;
        F.cta_begin = U.total_ctas;
        F.nst = L[k].nst;
        U.total_ctas += L[k].nst;
        fix_begin[U.nfeats] = fix_ctas;
        fix_ctas += (L[k].nst + (kUpdThreads / 32) - 1) / (kUpdThreads / 32);
        U.nfeats++;
      }
      if (U.nfeats == 0) continue;
      // fix-up launch: same params, cta_begin re-based to fix-up CTAs
      UpdParams X = U;
      for (int k = 0; k < X.nfeats; ++k) X.f[k].cta_begin = fix_begin[k];
      X.total_ctas = fix_ctas;
      switch (V) {
        case 1: rc = launch_update<1>(U, X, stream); break;
        case 2: rc = launch_update<2>(U, X, stream); break;
        case 4: rc = launch_update<4>(U, X, stream); break;
        default: rc = launch_update<8>(U, X, stream); break;
      }
      if (rc != HB_OK) return rc;
    }
  }
  return HB_OK;
}

}  // namespace hb

extern "C" {

int hbGroupSparseUpdateWorkspaceBytes(int n, const hbUpdateFeature* feats, size_t* bytes) {
  using namespace hb;
  HB_REQUIRE(n >= 1 && feats && bytes, "hbGroupSparseUpdateWorkspaceBytes: bad argument");
  size_t o = 0;
  size_t max_chunk_counts = 0;
  for (int c0 = 0; c0 < n; c0 += kMaxUpdFeats) {
    const int nc = (n - c0 < kMaxUpdFeats) ? n - c0 : kMaxUpdFeats;
    size_t tiles = 0;
    for (int k = 0; k < nc; ++k) {
      const hbUpdateFeature& f = feats[c0 + k];
      HB_REQUIRE(f.nnz >= 0 && f.nnz <= INT32_MAX && f.dim >= 4 && f.dim % 4 == 0 && f.dim <= 1024,
                 "hbGroupSparseUpdateWorkspaceBytes: feature %d bad nnz/dim", c0 + k);
      o = upd_layout(f, o).end;
      tiles += bucket_tiles(f.nnz);
    }
    const size_t cb = align_up(bucket_scratch_words(nc, tiles, kRadixBins, kMaxPasses) * sizeof(uint32_t), 256);
    if (cb > max_chunk_counts) max_chunk_counts = cb;
  }
  *bytes = o + max_chunk_counts + 256;
  return HB_OK;
}

int hbGroupLookupBackwardUpdate(int n, const hbUpdateFeature* feats, const hbOptimizer* opt,
                                void* d_workspace, size_t workspace_bytes, int32_t* d_status,
                                hbStream stream) {
  return hb::sparse_update_run(n, feats, opt, d_workspace, workspace_bytes, d_status,
                               (cudaStream_t)stream, nullptr, nullptr, hb::kPhaseSort | hb::kPhaseApply);
}

int hbGroupSparseSort(int n, const hbUpdateFeature* feats, void* d_workspace, size_t workspace_bytes,
                      int32_t* d_status, hbStream stream) {
  return hb::sparse_update_run(n, feats, nullptr, d_workspace, workspace_bytes, d_status,
                               (cudaStream_t)stream, nullptr, nullptr, hb::kPhaseSort);
}

int hbGroupSparseApply(int n, const hbUpdateFeature* feats, const hbOptimizer* opt, void* d_workspace,
                       size_t workspace_bytes, int32_t* d_status, hbStream stream) {
  return hb::sparse_update_run(n, feats, opt, d_workspace, workspace_bytes, d_status,
                               (cudaStream_t)stream, nullptr, nullptr, hb::kPhaseApply);
}

}  // extern "C"
