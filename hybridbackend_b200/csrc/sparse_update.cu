// sparse_update.cu -- K5: backward of the pooled lookup fused with the sparse
// optimizer apply (Adagrad / LazyAdam / SGD), deterministic, no float atomics.
//
// Pipeline (all features of the group per launch):
//   1. bag_of_position : CSR features only -- bag index of every id position.
//   2. cluster sort    : ONE launch, a cluster of 8 CTAs per feature (cluster_sort.cuh):
//                        stable LSD radix sort of (row key, bag) pairs, 9-bit digits, the
//                        passes separated by cluster barriers; then, in the same kernel,
//                        one scan over the sorted keys -> unique keys, the first entry and
//                        the first four bags of every run, the unique count (on device).
//   3. short kernel    : work item = one UNIQUE row.  A sub-warp group of G lanes
//                        (G*4 floats = one row) sums the gradient rows of the run in
//                        position order in registers and applies the optimizer once (one
//                        read-modify-write of the table row and its slot rows).  No
//                        cross-thread state, no shuffles, no barriers.  Runs longer than
//                        kShortMax are cut into pieces of piece_rows(dim) entries and queued.
//   4. long kernel     : hot rows, one warp per queued piece: its groups sum interleaved
//                        entries (8 rows per lane in flight), the group sums are combined
//                        in a fixed shuffle tree; multi-piece runs park piece sums in global
//                        memory and the warp that arrives last adds them in piece order.
// The order of every floating-point addition is a fixed function of the run length
// => bit-reproducible (test_determinism), whatever the scheduling.
//
// Algorithmic HBM bytes (Adagrad, fp32): per looked-up id 8 (sorted key + bag) +
// 4*dim (gradient row); per unique row 4*4*dim (table + accumulator, read + write).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "bucket.cuh"
#include "cluster_sort.cuh"
#include "update.cuh"

namespace hb {

constexpr int kUpdThreads = 256;
constexpr int kMaxUpdFeats = kCsMaxFeats;
constexpr int kShortMax = 16;   // runs up to this many entries are summed by one group

enum { kModeApply = 0, kModeEmit = 1 };

struct UpdFeat {
  float* table;
  float* slot0;
  float* slot1;
  const float* grad;
  const int64_t* offsets;   // bag sizes for mean/sqrtn (nullptr: one id per bag)
  const uint32_t* ukey;     // [U] unique keys
  const int32_t* ustart;    // [U+1]
  const int32_t* counts;    // [0] = U
  const int4* ubag4;        // [U] values (bag / position) of the first four entries of each run
  const int32_t* vals;      // bag index of each sorted entry (or its position, see pos2bag)
  const int32_t* pos2bag;   // != nullptr: vals hold input positions, bag = pos2bag[position]
  const int32_t* emit_send_off;
  const int32_t* emit_remote_base;
  uint64_t emit_off;
  int64_t rows;
  int64_t grad_stride;
  int32_t emit_cap;
  int32_t dim;
  int32_t combiner;
  int32_t log2g;
  int32_t max_chunks;       // static bound of the number of warp units of this feature (0: not in this launch)
  int32_t piece;            // entries per piece of a long run (kPieceBytes of rows)
};

struct LongItem {   // one piece of a hot row
  int32_t feat, u;
  int32_t start, count;   // sorted entries [start, start + count)
  uint32_t key;
  int32_t pbase;          // first partial slot of the run (-1: single piece)
  int32_t piece, np;
};

struct UpdParams {
  UpdFeat f[kMaxUpdFeats];
  WaitSpec wait;
  EmitCtx emit;
  int32_t* status;
  int32_t* long_count;      // [0] queued pieces, [1] reserved partial slots
  LongItem* items;
  float* part;              // [part_cap][part_stride] piece sums of multi-piece runs
  int32_t* tickets;         // [part_cap] arrival counters, indexed by the run's pbase
  int32_t item_cap, part_cap, part_stride;
  int32_t ticket_idx;       // long kernel: its work ticket is long_count[ticket_idx] (one per vector class)
  int32_t nfeats;
  int32_t opt;
  int32_t fast;             // approximate sqrt/div (MUFU) instead of the IEEE sequence
  float lr;                 // Adagrad/SGD: lr ; LazyAdam: bias-corrected lr_t
  float beta1, beta2, eps;
  float omb1, omb2;
};

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add_rn(const float4& a, const float4& b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z),
                     __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4_div_rn(const float4& a, float c) {
  return make_float4(__fdiv_rn(a.x, c), __fdiv_rn(a.y, c), __fdiv_rn(a.z, c), __fdiv_rn(a.w, c));
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {  // L2 only (data written by other SMs)
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

// One optimizer step on 4 consecutive elements of a row (values in registers).
// FAST = false is the IEEE sequence of the oracle (bit-exact); FAST = true uses
// the MUFU approximations (sqrt.approx / div.approx, <= 2 ulp each) -- the class of
// arithmetic TF's GPU kernels use (rsqrt) -- selected by HB_OPT_FLAG_FAST_MATH.
template <int OPT, bool FAST>
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
      if constexpr (FAST) ww[i] = __fsub_rn(ww[i], __fdividef(__fmul_rn(P.lr, gg[i]), sqrt_approx(aa[i])));
      else ww[i] = __fsub_rn(ww[i], __fdiv_rn(__fmul_rn(P.lr, gg[i]), __fsqrt_rn(aa[i])));
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
      if constexpr (FAST)
        ww[i] = __fsub_rn(ww[i], __fdividef(__fmul_rn(P.lr, mm[i]), __fadd_rn(sqrt_approx(v2[i]), P.eps)));
      else
        ww[i] = __fsub_rn(ww[i], __fdiv_rn(__fmul_rn(P.lr, mm[i]), __fadd_rn(__fsqrt_rn(v2[i]), P.eps)));
    }
    s0v = make_float4(mm[0], mm[1], mm[2], mm[3]);
    s1v = make_float4(v2[0], v2[1], v2[2], v2[3]);
  } else {  // SGD
#pragma unroll
    for (int i = 0; i < 4; ++i) ww[i] = __fsub_rn(ww[i], __fmul_rn(P.lr, gg[i]));
  }
  wv = make_float4(ww[0], ww[1], ww[2], ww[3]);
}

// bag (= gradient row) of sorted entry i
__device__ __forceinline__ int entry_bag(const UpdFeat& F, int i) {
  const int v = F.vals[i];
  return F.pos2bag != nullptr ? F.pos2bag[v] : v;
}

// gradient scale of the pooled lookup's backward: mean -> 1/count, sqrtn -> 1/sqrt(count)
__device__ __forceinline__ float bag_scale(const UpdFeat& F, int bag) {
  const int64_t c = F.offsets[bag + 1] - F.offsets[bag];
  return (F.combiner == HB_MEAN) ? (float)c : __fsqrt_rn((float)c);
}

// ---- sink of one finished row sum -------------------------------------------------------
// kModeApply: read-modify-write of the table row and its slots (row/slot values may
// have been prefetched into w/s0/s1).  kModeEmit: store the sum into the owner's
// grads_in window at the slot of unique `u` (requester side of the sharded backward).
template <int V, int MODE>
__device__ __forceinline__ float* emit_dst(const UpdParams& P, const UpdFeat& F, int u) {
  int r = 0;
  while (r + 1 < P.emit.world && F.emit_send_off[r + 1] <= u) ++r;
  const int drow = F.emit_remote_base[r] + (u - F.emit_send_off[r]);
  if (drow >= F.emit_cap) return nullptr;  // overflow already flagged by the exchange
  return reinterpret_cast<float*>(P.emit.peers.p[r] + P.emit.window_off + F.emit_off) + (int64_t)drow * F.dim;
}

template <int V, int OPT, int MODE, bool FAST>
__device__ __forceinline__ void sink_row(const UpdParams& P, const UpdFeat& F, uint32_t key, int u,
                                         const float4 (&acc)[V], const int (&col)[V],
                                         const bool (&act)[V], bool& oob) {
  if constexpr (MODE == kModeApply) {
    if ((uint64_t)key >= (uint64_t)F.rows) { oob = true; return; }
    float4 w[V], s0[V], s1[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      w[v] = s0[v] = s1[v] = f4_zero();
      if (act[v]) {
        const int64_t o = (int64_t)key * F.dim + col[v];
        w[v] = *reinterpret_cast<const float4*>(F.table + o);
        if constexpr (OPT != HB_OPT_SGD) s0[v] = *reinterpret_cast<const float4*>(F.slot0 + o);
        if constexpr (OPT == HB_OPT_LAZY_ADAM) s1[v] = *reinterpret_cast<const float4*>(F.slot1 + o);
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (act[v]) {
        const int64_t o = (int64_t)key * F.dim + col[v];
        opt_step4<OPT, FAST>(P, w[v], s0[v], s1[v], acc[v]);
        *reinterpret_cast<float4*>(F.table + o) = w[v];
        if constexpr (OPT != HB_OPT_SGD) *reinterpret_cast<float4*>(F.slot0 + o) = s0[v];
        if constexpr (OPT == HB_OPT_LAZY_ADAM) *reinterpret_cast<float4*>(F.slot1 + o) = s1[v];
      }
  } else {
    float* dst = emit_dst<V, MODE>(P, F, u);
    if (dst == nullptr) return;
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (act[v]) *reinterpret_cast<float4*>(dst + col[v]) = acc[v];
  }
}

// ---- dense work map ---------------------------------------------------------------------
// Work units per feature are data dependent (unique counts, device-side lengths):
// every CTA scans them into shared memory once and walks the dense unit list, so
// no CTA is ever launched for (or iterates over) work that does not exist.
__device__ __forceinline__ int seg_scan(int nsegs, int my_units, int* s_begin /*[kMaxSegs + 1]*/) {
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

__device__ __forceinline__ int seg_find(const int* s_begin, int nsegs, int unit) {
  int lo = 0, hi = nsegs - 1;
  while (lo < hi) {  // last seg with begin <= unit (empty segs are skipped)
    const int mid = (lo + hi + 1) >> 1;
    if (s_begin[mid] <= unit) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---- 4. apply ---------------------------------------------------------------------------
// Two small kernels instead of one large one (the fused kernel of the previous design
// spent a third of its issue slots waiting for instructions: 64 KB of code, two
// divergent paths per CTA, 24 warps per SM):
//   short kernel : one UNIQUE row per sub-warp group, static round-robin over the dense
//                  unit list.  Run bounds, key and the first four bags of the run come
//                  from the run arrays (prefetched one unit ahead), so a run of up to 4
//                  entries costs ONE exposed memory round trip: table row, slot rows and
//                  gradient rows are requested together.  Runs longer than kShortMax are
//                  cut into pieces and queued (their groups do the queueing instead of
//                  idling).
//   long kernel  : one warp per queued piece, kLongRows x 128-bit loads per lane in flight.
constexpr int kPieceBytes = 16384;   // gradient bytes per hot-row piece
constexpr int kMaxPiece = 512;       // entries per piece at most

static inline int piece_rows(int dim) {
  static const int piece_bytes = [] {   // HB_PIECE_BYTES: tuning knob (power of two, 2 KB .. 64 KB)
    const char* e = getenv("HB_PIECE_BYTES");
    const int v = e ? atoi(e) : 0;
    return (v >= 2048 && v <= 65536 && (v & (v - 1)) == 0) ? v : kPieceBytes;
  }();
  int p = piece_bytes / (dim * 4);
  if (p > kMaxPiece) p = kMaxPiece;
  if (p < 8) p = 8;
  return p;
}

// run bounds, key and first bags of the unique row a group owns in warp unit `unit`
struct ShortMeta {
  int fi, u, s, len;
  uint32_t key;
  int4 b;
};

__device__ __forceinline__ void short_meta(const UpdParams& P, const int* s_begin, int unit, unsigned lane,
                                           int& fi_hint, ShortMeta& M) {
  // units of one warp ascend: advance the feature instead of searching for it
  int fi = fi_hint;
  while (unit >= s_begin[fi + 1]) ++fi;
  fi_hint = fi;
  M.fi = fi;
  const UpdFeat& F = P.f[fi];
  const int ng = 32 >> F.log2g;
  const int gi = (int)lane >> F.log2g;
  const int U = F.counts[0];
  M.u = (unit - s_begin[fi]) * ng + gi;
  const bool ok = M.u < U;
  const int uu = ok ? M.u : 0;
  M.s = F.ustart[uu];
  M.len = ok ? F.ustart[uu + 1] - M.s : 0;
  M.key = F.ukey[uu];
  M.b = F.ubag4[uu];
}

template <int V, int OPT, int MODE, bool FAST, int OCC>
__global__ void __launch_bounds__(kUpdThreads, OCC)
update_short_kernel(const __grid_constant__ UpdParams P) {
  __shared__ int s_begin[kMaxUpdFeats + 2];
  constexpr int NG = (V == 1) ? 4 : (V == 2 ? 2 : 1);   // gradient rows in flight per group
  wait_spec(P.wait, P.status);
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  int units = 0;
  if ((int)threadIdx.x < P.nfeats) {
    const int ng = 32 >> P.f[threadIdx.x].log2g;
    units = (P.f[threadIdx.x].counts[0] + ng - 1) / ng;
    if (units > P.f[threadIdx.x].max_chunks) units = P.f[threadIdx.x].max_chunks;
  }
  const int total = seg_scan(P.nfeats, units, s_begin);
  if (threadIdx.x == 0) s_begin[P.nfeats + 1] = 0x7FFFFFFF;  // sentinel for the feature walk
  __syncthreads();
  const int nwarps = gridDim.x * (kUpdThreads / 32);
  int unit = blockIdx.x * (kUpdThreads / 32) + warp;
  bool oob = false;
  int fi_hint = 0;
  ShortMeta M, Mn;
  if (unit < total) short_meta(P, s_begin, unit, lane, fi_hint, M);
  while (unit < total) {
    const int next = unit + nwarps;
    // the NEXT unit's run bounds travel while this unit's rows do
    if (next < total) short_meta(P, s_begin, next, lane, fi_hint, Mn);
    const UpdFeat& F = P.f[M.fi];
    const int log2g = F.log2g;
    const int l = (int)lane & ((1 << log2g) - 1);
    const int dim = F.dim;
    // ---- hot rows: cut into pieces and queue them for the long kernel --------------------
    {
      const bool is_long = M.len > kShortMax && l == 0;
      unsigned todo = __ballot_sync(0xffffffffu, is_long);
      if (todo) {
        const int np = is_long ? (M.len + F.piece - 1) / F.piece : 0;
        int base = -1, pb = -1;
        if (is_long) {
          base = atomicAdd(&P.long_count[0], np);
          if (np > 1) pb = atomicAdd(&P.long_count[1], np);
          if (base + np > P.item_cap || (np > 1 && pb + np > P.part_cap)) {  // cannot happen by the layout bounds
            raise_status(P.status, HB_STATUS_WINDOW_OVERFLOW);
            base = -1;
          }
        }
        todo = __ballot_sync(0xffffffffu, is_long && base >= 0);
        while (todo) {  // the pieces of a hot row are written by the whole warp
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int r_u = __shfl_sync(0xffffffffu, M.u, src);
          const int r_s = __shfl_sync(0xffffffffu, M.s, src);
          const int r_len = __shfl_sync(0xffffffffu, M.len, src);
          const int r_np = __shfl_sync(0xffffffffu, np, src);
          const int r_base = __shfl_sync(0xffffffffu, base, src);
          const int r_pb = __shfl_sync(0xffffffffu, pb, src);
          const uint32_t r_key = __shfl_sync(0xffffffffu, M.key, src);
          for (int j = (int)lane; j < r_np; j += 32) {
            LongItem it;
            it.feat = M.fi; it.u = r_u; it.start = r_s + j * F.piece;
            it.count = min(F.piece, r_len - j * F.piece);
            it.key = r_key; it.pbase = r_pb; it.piece = j; it.np = r_np;
            P.items[r_base + j] = it;
          }
        }
      }
    }
    // ---- short run: sum in position order, one read-modify-write -------------------------
    bool ok = M.len > 0 && M.len <= kShortMax;
    if (MODE == kModeApply && ok && (uint64_t)M.key >= (uint64_t)F.rows) {
      oob = true;
      ok = false;
    }
    if (ok) {
      const bool scaled = F.combiner != HB_SUM && F.offsets != nullptr;
      int col[V];
      bool act[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        col[v] = ((v << log2g) + l) * 4;
        act[v] = col[v] < dim;
      }
      float4 w[V], s0[V], s1[V], acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        w[v] = s0[v] = s1[v] = acc[v] = f4_zero();
        if constexpr (MODE == kModeApply) {
          if (act[v]) {
            const int64_t o = (int64_t)M.key * dim + col[v];
            w[v] = *reinterpret_cast<const float4*>(F.table + o);
            if constexpr (OPT != HB_OPT_SGD) s0[v] = *reinterpret_cast<const float4*>(F.slot0 + o);
            if constexpr (OPT == HB_OPT_LAZY_ADAM) s1[v] = *reinterpret_cast<const float4*>(F.slot1 + o);
          }
        }
      }
      // entries j0 .. j0 + NG - 1 (bag < 0: no such entry), added in position order
      auto batch = [&](int j0, const int (&bag)[NG]) {
        float sc[NG];
        float4 x[NG][V];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          sc[i] = (scaled && bag[i] >= 0) ? bag_scale(F, bag[i]) : 1.0f;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            x[i][v] = f4_zero();
            if (bag[i] >= 0 && act[v])
              x[i][v] = ld_nc_f4(reinterpret_cast<const float4*>(F.grad + (int64_t)bag[i] * F.grad_stride + col[v]));
          }
        }
#pragma unroll
        for (int i = 0; i < NG; ++i)
          if (bag[i] >= 0) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float4 t = scaled ? f4_div_rn(x[i][v], sc[i]) : x[i][v];
              acc[v] = (j0 + i == 0) ? t : f4_add_rn(acc[v], t);
            }
          }
      };
      const int b4[4] = {M.b.x, M.b.y, M.b.z, M.b.w};
      {  // first batch: its bags came with the run bounds
        int bag[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          bag[i] = (i < M.len) ? b4[i] : -1;
          if (F.pos2bag != nullptr && bag[i] >= 0) bag[i] = F.pos2bag[bag[i]];
        }
        batch(0, bag);
      }
      for (int j0 = NG; j0 < M.len; j0 += NG) {
        int bag[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          const int j = j0 + i;
          int vv = -1;
          if (j < M.len) {
            if (NG < 4 && j < 4) vv = (j == 1) ? b4[1] : (j == 2 ? b4[2] : b4[3]);
            else vv = F.vals[M.s + j];
            if (F.pos2bag != nullptr) vv = F.pos2bag[vv];
          }
          bag[i] = vv;
        }
        batch(j0, bag);
      }
      if constexpr (MODE == kModeApply) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (act[v]) {
            const int64_t o = (int64_t)M.key * dim + col[v];
            opt_step4<OPT, FAST>(P, w[v], s0[v], s1[v], acc[v]);
            *reinterpret_cast<float4*>(F.table + o) = w[v];
            if constexpr (OPT != HB_OPT_SGD) *reinterpret_cast<float4*>(F.slot0 + o) = s0[v];
            if constexpr (OPT == HB_OPT_LAZY_ADAM) *reinterpret_cast<float4*>(F.slot1 + o) = s1[v];
          }
      } else {
        sink_row<V, OPT, MODE, FAST>(P, F, M.key, M.u, acc, col, act, oob);
      }
    }
    M = Mn;
    unit = next;
  }
  if (oob) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
}

// Sum of the entries gi, gi + ng, ... of a piece, in that order: R gradient rows per lane
// in flight (128-bit loads), the bags of the next batch are requested before the rows
// of the current one are consumed.
template <int V, int R, bool SCALED>
__device__ __forceinline__ void piece_sum(const UpdFeat& F, const LongItem& cur, int ng, int gi,
                                          const int (&col)[V], const bool (&act)[V], float4 (&acc)[V]) {
  const int m = cur.count;
  const int step = ng * R;
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = f4_zero();
  bool first = true;
  int bag[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = r * ng + gi;
    bag[r] = j < m ? entry_bag(F, cur.start + j) : 0;
  }
  for (int t0 = 0; t0 < m; t0 += step) {
    float4 x[R][V];
    float sc[SCALED ? R : 1];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool valid = t0 + r * ng + gi < m;
      if constexpr (SCALED) sc[r] = valid ? bag_scale(F, bag[r]) : 1.0f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        x[r][v] = f4_zero();
        if (valid && act[v])
          x[r][v] = ld_nc_f4(reinterpret_cast<const float4*>(F.grad + (int64_t)bag[r] * F.grad_stride + col[v]));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {   // bags of the next batch
      const int j = t0 + step + r * ng + gi;
      bag[r] = j < m ? entry_bag(F, cur.start + j) : 0;
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (t0 + r * ng + gi < m) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4 t = x[r][v];
          if constexpr (SCALED) t = f4_div_rn(t, sc[r]);
          acc[v] = first ? t : f4_add_rn(acc[v], t);
        }
        first = false;
      }
  }
}

// Hot rows: one warp per queued piece (dynamic tickets: pieces differ in length).  Group
// gi of the warp adds entries gi, gi + ng, ... of the piece in that order, kLongRows
// gradient rows per lane in flight in registers (128-bit loads; the bags of the next
// batch travel while the rows of the current one do); the group sums are combined in a
// fixed shuffle tree.  Multi-piece runs park piece sums in global memory and the warp
// that arrives last adds them in piece order.
template <int V, int OPT, int MODE, bool FAST>
__global__ void __launch_bounds__(kUpdThreads, (V == 1 ? 3 : 2))
update_long_kernel(const __grid_constant__ UpdParams P) {
  constexpr int R = (V == 1) ? 8 : (V == 2 ? 4 : (V == 4 ? 2 : 1));  // rows in flight per group
  const unsigned lane = lane_id();
  int n_long = P.long_count[0];
  if (n_long > P.item_cap) n_long = P.item_cap;
  bool oob = false;
  // One ticket per piece.  (Handing out 4 pieces per atomic was measured: 81 vs 60 us -- the
  // kernel is bound by its tail, a piece is ~10 us of dependent round trips, so the finest
  // granularity wins although the ticket atomics show up as 27 % of the stall samples.)
  constexpr int kLongChunk = 1;
  auto grab = [&]() -> int {
    int v = 0;
    if (lane == 0) v = atomicAdd(&P.long_count[P.ticket_idx], kLongChunk);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  int chunk = grab();
  int chunk_next = grab();   // one chunk ahead: the atomic's latency is never exposed
  int sub = 0;
  while (chunk < n_long) {
    const int it = chunk + sub;
    auto advance = [&]() {
      if (++sub == kLongChunk || chunk + sub >= n_long) {
        sub = 0;
        chunk = chunk_next;
        chunk_next = grab();
      }
    };
    const LongItem cur = P.items[it];
    const UpdFeat& F = P.f[cur.feat];
    if (F.max_chunks == 0) {   // a feature of another vector class: not this launch's
      advance();
      continue;
    }
    const int log2g = F.log2g;
    const int ng = 32 >> log2g;        // groups per warp
    const int gi = (int)lane >> log2g;
    const int l = (int)lane & ((1 << log2g) - 1);
    const int dim = F.dim;
    const bool scaled = F.combiner != HB_SUM && F.offsets != nullptr;
    int col[V];
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      col[v] = ((v << log2g) + l) * 4;
      act[v] = col[v] < dim;
    }
    float4 acc[V];
    if (scaled) piece_sum<V, (R > 1 ? R / 2 : 1), true>(F, cur, ng, gi, col, act, acc);
    else piece_sum<V, R, false>(F, cur, ng, gi, col, act, acc);
    // group sums -> piece sum, fixed order: g0 += g(ng/2) ... (idle groups hold zeros)
#pragma unroll
    for (int v = 0; v < V; ++v)
      for (int off = ng >> 1; off >= 1; off >>= 1) {
        float4 y;
        y.x = __shfl_down_sync(0xffffffffu, acc[v].x, off << log2g);
        y.y = __shfl_down_sync(0xffffffffu, acc[v].y, off << log2g);
        y.z = __shfl_down_sync(0xffffffffu, acc[v].z, off << log2g);
        y.w = __shfl_down_sync(0xffffffffu, acc[v].w, off << log2g);
        if (gi < off) acc[v] = f4_add_rn(acc[v], y);
      }
    bool last = false;
    if (cur.np == 1) {
      if (gi == 0) sink_row<V, OPT, MODE, FAST>(P, F, cur.key, cur.u, acc, col, act, oob);
    } else {
      // multi-piece run: park the piece sum; the warp arriving last adds all pieces in
      // piece order (which warp that is does not change the order of the additions)
      float* prow = P.part + (size_t)(cur.pbase + cur.piece) * P.part_stride;
      if (gi == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (act[v]) *reinterpret_cast<float4*>(prow + col[v]) = acc[v];
      }
      __threadfence();
      __syncwarp();
      int old = 0;
      if (lane == 0) old = atomicAdd(&P.tickets[cur.pbase], 1);
      old = __shfl_sync(0xffffffffu, old, 0);
      last = (old == cur.np - 1);  // warp-uniform
    }
    if (last) {
      __threadfence();
      // group gi adds pieces gi, gi + ng, ...; the group sums are combined as above
      const float* p0 = P.part + (size_t)cur.pbase * P.part_stride;
      float4 tot[V];
#pragma unroll
      for (int v = 0; v < V; ++v) tot[v] = f4_zero();
      bool first2 = true;
      for (int j0 = gi; j0 - gi < cur.np; j0 += ng * R) {  // warp-uniform trip count
        float4 x[R][V];
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
          for (int v = 0; v < V; ++v)
            x[i][v] = (j0 + i * ng < cur.np && act[v])
                          ? ld_cg_f4(reinterpret_cast<const float4*>(p0 + (size_t)(j0 + i * ng) * P.part_stride + col[v]))
                          : f4_zero();
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (j0 + i * ng < cur.np) {
#pragma unroll
            for (int v = 0; v < V; ++v) tot[v] = first2 ? x[i][v] : f4_add_rn(tot[v], x[i][v]);
            first2 = false;
          }
      }
#pragma unroll
      for (int v = 0; v < V; ++v)
        for (int off = ng >> 1; off >= 1; off >>= 1) {
          float4 y;
          y.x = __shfl_down_sync(0xffffffffu, tot[v].x, off << log2g);
          y.y = __shfl_down_sync(0xffffffffu, tot[v].y, off << log2g);
          y.z = __shfl_down_sync(0xffffffffu, tot[v].z, off << log2g);
          y.w = __shfl_down_sync(0xffffffffu, tot[v].w, off << log2g);
          if (gi < off) tot[v] = f4_add_rn(tot[v], y);
        }
      if (gi == 0) sink_row<V, OPT, MODE, FAST>(P, F, cur.key, cur.u, tot, col, act, oob);
    }
    advance();
  }
  if (oob) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
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

constexpr int kRadixBits = kCsBits;   // digit width of the cluster sort

// per-feature workspace layout (a function of nnz, dim and offsets != NULL only)
struct UpdLayout {
  size_t keysA, keysB, valsA, valsB, bagmap, ukey, ustart, ubag4, counts, cs_hist, cs_uniq, end;
  int log2g, V;
};

static UpdLayout upd_layout(const hbUpdateFeature& f, size_t base) {
  UpdLayout L;
  const size_t n = (size_t)f.nnz;
  size_t o = base;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.keysA = take(4 * n); L.keysB = take(4 * n);
  L.valsA = take(4 * n); L.valsB = take(4 * n);
  L.bagmap = take(f.offsets ? 4 * n : 0);
  L.ukey = take(4 * n);
  L.ustart = take(4 * (n + 1));
  L.ubag4 = take(16 * n);
  L.counts = take(64);
  L.cs_hist = take(sizeof(uint32_t) * kCsBins * (size_t)cs_tiles(f.nnz));
  L.cs_uniq = take(sizeof(int32_t) * ((size_t)cs_tiles(f.nnz) + 1));
  upd_shape(f.dim, &L.log2g, &L.V);
  L.end = o;
  return L;
}

// chunk-shared scratch behind the per-feature regions: the hot-row queue
struct SharedLayout {
  size_t zero_sort, zero_sort_bytes;   // zeroed by the sort kernel: queue counters and tickets
  size_t long_count, tickets;
  size_t items, part;
  size_t end;
  int item_cap, part_cap, part_stride;
};

// Queue bounds of one feature: every queued run is longer than kShortMax, so there are
// fewer than nnz / kShortMax of them, each with at most len / piece + 1 pieces; only
// runs of more than one piece (len > piece) park partial sums.
static inline size_t long_items_of(int64_t nnz, int dim) {
  return (size_t)(nnz / kShortMax + nnz / piece_rows(dim) + 2);
}
static inline size_t long_parts_of(int64_t nnz, int dim) { return (size_t)(2 * (nnz / piece_rows(dim)) + 2); }

static SharedLayout shared_layout(size_t base, size_t item_cap, size_t part_cap, int max_dim) {
  SharedLayout S;
  size_t o = align_up(base, 256);
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  S.zero_sort = o;
  S.item_cap = (int)item_cap;
  S.part_cap = (int)part_cap;
  S.part_stride = (max_dim + 3) / 4 * 4;
  S.long_count = take(64);
  S.tickets = take((size_t)S.part_cap * sizeof(int32_t));
  S.zero_sort_bytes = o - S.zero_sort;
  S.items = take((size_t)S.item_cap * sizeof(LongItem));
  S.part = take((size_t)S.part_cap * S.part_stride * sizeof(float));
  S.end = o;
  return S;
}


template <int V, int OPT, int MODE, bool FAST, int OCC>
static int launch_short(const UpdParams& U, cudaStream_t stream) {
  int per_sm = 0;
  HB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, update_short_kernel<V, OPT, MODE, FAST, OCC>, kUpdThreads, 0));
  const int grid = device_sm_count() * (per_sm > 0 ? per_sm : 1);
  KernelScope ks(HB_K_SPARSE_UPDATE, stream);
  update_short_kernel<V, OPT, MODE, FAST, OCC><<<grid, kUpdThreads, 0, stream>>>(U);
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

// which: 0 = short kernel, 1 = long kernel
template <int V, int OPT, int MODE, bool FAST>
static int launch_apply(const UpdParams& U, int which, cudaStream_t stream) {
  int per_sm = 0;
  if (which == 0) {
    // CTAs per SM: measured at V = 1 (D <= 128) -- 4 (64 regs, no spills) 82 us, 5 (48 regs)
    // 91 us, 6 (40 regs) 98 us on the C2 workload: spills cost more than occupancy gains
    return launch_short<V, OPT, MODE, FAST, (V == 1 ? 4 : (V == 2 ? 3 : 1))>(U, stream);
  } else {
    HB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, update_long_kernel<V, OPT, MODE, FAST>, kUpdThreads, 0));
    const int grid = device_sm_count() * (per_sm > 0 ? per_sm : 1);
    KernelScope ks(HB_K_UPDATE_LONG, stream);
    update_long_kernel<V, OPT, MODE, FAST><<<grid, kUpdThreads, 0, stream>>>(U);
  }
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

template <int V, int MODE>
static int launch_apply_opt(const UpdParams& U, int which, cudaStream_t stream) {
  if (MODE == kModeEmit) return launch_apply<V, HB_OPT_SGD, MODE, false>(U, which, stream);
  switch (U.opt) {
    case HB_OPT_ADAGRAD:
      return U.fast ? launch_apply<V, HB_OPT_ADAGRAD, MODE, true>(U, which, stream)
                    : launch_apply<V, HB_OPT_ADAGRAD, MODE, false>(U, which, stream);
    case HB_OPT_LAZY_ADAM:
      return U.fast ? launch_apply<V, HB_OPT_LAZY_ADAM, MODE, true>(U, which, stream)
                    : launch_apply<V, HB_OPT_LAZY_ADAM, MODE, false>(U, which, stream);
    default:
      return launch_apply<V, HB_OPT_SGD, MODE, false>(U, which, stream);
  }
}

static int launch_apply_v(int V, bool emit, const UpdParams& U, int which, cudaStream_t stream) {
  if (emit) {
    switch (V) {
      case 1: return launch_apply_opt<1, kModeEmit>(U, which, stream);
      case 2: return launch_apply_opt<2, kModeEmit>(U, which, stream);
      case 4: return launch_apply_opt<4, kModeEmit>(U, which, stream);
      default: return launch_apply_opt<8, kModeEmit>(U, which, stream);
    }
  }
  switch (V) {
    case 1: return launch_apply_opt<1, kModeApply>(U, which, stream);
    case 2: return launch_apply_opt<2, kModeApply>(U, which, stream);
    case 4: return launch_apply_opt<4, kModeApply>(U, which, stream);
    default: return launch_apply_opt<8, kModeApply>(U, which, stream);
  }
}

// debug: HB_CS_TIMING=1 allocates a small device buffer the sort kernel stamps
static unsigned long long* g_cs_timing = nullptr;
static unsigned long long* cs_timing_buffer() {
  static const bool on = [] { const char* e = getenv("HB_CS_TIMING"); return e != nullptr && e[0] == '1'; }();
  if (on && g_cs_timing == nullptr) {
    if (cudaMalloc(reinterpret_cast<void**>(&g_cs_timing), 64 * 8) != cudaSuccess) g_cs_timing = nullptr;
    else cudaMemset(g_cs_timing, 0, 64 * 8);
  }
  return g_cs_timing;
}

static int validate_upd(int k, const hbUpdateFeature& f, const hbOptimizer* opt, int phases,
                        const UpdExtra* ex, bool emit) {
  HB_REQUIRE(f.dim >= 4 && f.dim % 4 == 0 && f.dim <= 1024,
             "update: feature %d dim %d must be a multiple of 4 in [4,1024]", k, f.dim);
  HB_REQUIRE(f.nnz >= 0 && f.nnz <= INT32_MAX / 2 && f.nbags >= 0 && f.nbags <= INT32_MAX,
             "update: feature %d bad nnz/nbags", k);
  HB_REQUIRE(f.offsets != nullptr || f.nnz == f.nbags,
             "update: feature %d has no offsets, so nnz must equal nbags", k);
  HB_REQUIRE(f.rows >= 0 && f.rows < ((int64_t)1 << 32) - 2, "update: feature %d rows out of range", k);
  HB_REQUIRE(f.id_div >= 1, "update: feature %d id_div must be >= 1", k);
  HB_REQUIRE(f.grad_stride >= f.dim && f.grad_stride % 4 == 0,
             "update: feature %d grad_stride must be a multiple of 4 and >= dim", k);
  HB_REQUIRE(f.combiner >= HB_SUM && f.combiner <= HB_SQRTN, "update: feature %d bad combiner", k);
  if (f.nnz > 0 && (phases & kPhaseSort))
    HB_REQUIRE(f.ids != nullptr || (ex && ex->keys32 != nullptr), "update: feature %d null ids", k);
  if (f.nnz > 0 && (phases & kPhaseApply)) {
    HB_REQUIRE(f.grad && ((uintptr_t)f.grad & 15) == 0, "update: feature %d grad must be non-null and 16-byte aligned", k);
    if (!emit) {
      HB_REQUIRE(f.table && ((uintptr_t)f.table & 15) == 0, "update: feature %d table must be non-null and 16-byte aligned", k);
      if (opt->kind == HB_OPT_ADAGRAD)
        HB_REQUIRE(f.slot0 != nullptr, "update: feature %d Adagrad needs slot0 (accumulator)", k);
      if (opt->kind == HB_OPT_LAZY_ADAM)
        HB_REQUIRE(f.slot0 != nullptr && f.slot1 != nullptr, "update: feature %d LazyAdam needs slot0/slot1", k);
    }
  }
  return HB_OK;
}

// number of 9-bit digit positions a key space of `space` values needs; +2 keeps the
// two sentinels (invalid id, padding) strictly above every valid key in the covered
// bits, so they sort behind all real rows and never split a run
static int radix_passes(int64_t space) {
  const int bits = ilog2c(space + 2);
  int p = (bits + kRadixBits - 1) / kRadixBits;
  return p < 1 ? 1 : p;
}

size_t sparse_update_workspace_bytes(int n, const hbUpdateFeature* feats) {
  size_t o = 0, shared_total = 0;
  for (int c0 = 0; c0 < n; c0 += kMaxUpdFeats) {
    const int nc = (n - c0 < kMaxUpdFeats) ? n - c0 : kMaxUpdFeats;
    size_t items = 64, parts = 64;
    int max_dim = 4;
    for (int k = 0; k < nc; ++k) {
      const hbUpdateFeature& f = feats[c0 + k];
      o = upd_layout(f, o).end;
      items += long_items_of(f.nnz, f.dim);
      parts += long_parts_of(f.nnz, f.dim);
      if (f.dim > max_dim) max_dim = f.dim;
    }
    // one shared region per chunk: the hot-row queue lives from the sort phase to the apply
    shared_total += align_up(shared_layout(0, items, parts, max_dim).end, 256);
  }
  return align_up(o, 256) + shared_total + 256;
}

// Sort (row key, bag) pairs of every feature, find the runs and run the fused
// duplicate-sum + sink.  `id_div` is shared by the features of one call.
int sparse_update_run(int n, const hbUpdateFeature* feats, const hbOptimizer* opt, void* ws,
                      size_t ws_bytes, int32_t* d_status, cudaStream_t stream, const WaitSpec* wait,
                      const UpdExtra* extras, const EmitCtx* emit, int phases, UpdViews* views) {
  static const hbOptimizer kNoOpt = {HB_OPT_SGD, 0.f, 0.f, 0.f, 0.f, 0, 1};
  const bool is_emit = emit != nullptr;
  if ((!(phases & kPhaseApply) || is_emit) && opt == nullptr) opt = &kNoOpt;
  HB_REQUIRE(n >= 1 && feats && opt, "update: bad arguments");
  HB_REQUIRE(opt->kind >= HB_OPT_SGD && opt->kind <= HB_OPT_LAZY_ADAM, "update: bad optimizer kind %d", opt->kind);
  for (int k = 0; k < n; ++k) {
    int rc = validate_upd(k, feats[k], opt, phases, extras ? &extras[k] : nullptr, is_emit);
    if (rc != HB_OK) return rc;
    HB_REQUIRE(feats[k].id_div == feats[0].id_div, "update: all features of a call must share id_div");
    if (extras) HB_REQUIRE(extras[k].key_kind == extras[0].key_kind, "update: all features of a call must share the key kind");
  }
  const size_t need = sparse_update_workspace_bytes(n, feats);
  if (ws_bytes < need || (need > 0 && ws == nullptr)) {
    set_last_error("update: workspace %zu < required %zu bytes", ws_bytes, need);
    return HB_ERR_WORKSPACE;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);
  const int key_kind = extras ? extras[0].key_kind : 0;
  size_t feat_end = 0;
  for (int k = 0; k < n; ++k) feat_end = upd_layout(feats[k], feat_end).end;
  int rc = HB_OK;

  size_t off = 0;
  size_t shared_off = align_up(feat_end, 256);
  for (int c0 = 0; c0 < n; c0 += kMaxUpdFeats) {
    const int nc = (n - c0 < kMaxUpdFeats) ? n - c0 : kMaxUpdFeats;
    UpdLayout L[kMaxUpdFeats];
    int passes[kMaxUpdFeats];
    size_t o = off;
    size_t total_items = 64, total_parts = 64;
    int max_passes = 0, max_dim = 4;
    for (int k = 0; k < nc; ++k) {
      const hbUpdateFeature& f = feats[c0 + k];
      const UpdExtra* ex = extras ? &extras[c0 + k] : nullptr;
      L[k] = upd_layout(f, o);
      o = L[k].end;
      total_items += long_items_of(f.nnz, f.dim);
      total_parts += long_parts_of(f.nnz, f.dim);
      if (f.dim > max_dim) max_dim = f.dim;
      int64_t space = f.rows;
      if (ex && ex->key_kind == 1) space = (int64_t)feats[0].id_div << ex->lbits;
      passes[k] = radix_passes(space);
      HB_REQUIRE(passes[k] <= kMaxPasses, "update: feature %d key space needs more than %d radix passes", c0 + k, kMaxPasses);
      if (f.nnz > 0 && passes[k] > max_passes) max_passes = passes[k];
      if (views) {
        views[c0 + k].ukey = reinterpret_cast<const uint32_t*>(base + L[k].ukey);
        views[c0 + k].ustart = reinterpret_cast<const int32_t*>(base + L[k].ustart);
        views[c0 + k].counts = reinterpret_cast<const int32_t*>(base + L[k].counts);
      }
    }
    off = o;
    const SharedLayout S = shared_layout(shared_off, total_items, total_parts, max_dim);
    shared_off = align_up(S.end, 256);

    if (phases & kPhaseSort) {
      // 1. bag map for CSR features
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

      // 2. stable LSD radix sort (9-bit digits) + run detection: one cluster per feature
      //    (cluster_sort.cuh); it also zeroes the hot-row queue counters of the apply
      CsParams C;
      C.d_status = d_status;
      C.zero = reinterpret_cast<int32_t*>(base + S.zero_sort);
      C.zero_words = (int32_t)(S.zero_sort_bytes / 4);
      C.nfeats = nc;
      C.key_kind = key_kind;
      C.p = (key_kind == 1) ? (int32_t)feats[0].id_div : 1;
      C.div = feats[0].id_div;
      C.div_shift = ((C.div & (C.div - 1)) == 0 && C.div <= (1 << 30)) ? ilog2c(C.div) : -1;
      for (int k = 0; k < nc; ++k) {
        const hbUpdateFeature& f = feats[c0 + k];
        const UpdExtra* ex = extras ? &extras[c0 + k] : nullptr;
        CsFeat& F = C.f[k];
        F.in_keys = (key_kind == 2) ? (const void*)ex->keys32 : (const void*)f.ids;
        F.in_vals = f.offsets ? reinterpret_cast<int32_t*>(base + L[k].bagmap) : nullptr;
        // the requester-side sort carries POSITIONS (the inverse map needs them);
        // bags are looked up from positions afterwards
        if (ex && ex->inv != nullptr) F.in_vals = nullptr;
        F.keys[0] = reinterpret_cast<uint32_t*>(base + L[k].keysA);
        F.keys[1] = reinterpret_cast<uint32_t*>(base + L[k].keysB);
        F.vals[0] = reinterpret_cast<int32_t*>(base + L[k].valsA);
        F.vals[1] = reinterpret_cast<int32_t*>(base + L[k].valsB);
        F.hist = reinterpret_cast<uint32_t*>(base + L[k].cs_hist);
        F.tile_uniq = reinterpret_cast<int32_t*>(base + L[k].cs_uniq);
        F.ukey = reinterpret_cast<uint32_t*>(base + L[k].ukey);
        F.ustart = reinterpret_cast<int32_t*>(base + L[k].ustart);
        F.ubag4 = reinterpret_cast<int4*>(base + L[k].ubag4);
        F.counts = reinterpret_cast<int32_t*>(base + L[k].counts);
        F.inv = ex ? ex->inv : nullptr;
        F.owner_start1 = ex ? ex->owner_start1 : nullptr;
        F.n_dev = ex ? ex->n_dev : nullptr;
        F.n = (int32_t)f.nnz;
        F.passes = passes[k];
        F.lbits = ex ? ex->lbits : 0;
        F.key_limit = (uint32_t)f.rows;
      }
      {  // launch order: features with the most passes (and entries) first
        int idx[kCsMaxFeats];
        for (int k = 0; k < nc; ++k) idx[k] = k;
        std::stable_sort(idx, idx + nc, [&](int a, int b) {
          const int64_t wa = (int64_t)C.f[a].passes * cs_tiles(C.f[a].n), wb = (int64_t)C.f[b].passes * cs_tiles(C.f[b].n);
          return wa > wb;
        });
        for (int k = 0; k < nc; ++k) C.order[k] = (uint8_t)idx[k];
      }
      (void)max_passes;
      C.timing = cs_timing_buffer();
      if (key_kind == 1) rc = cluster_sort_launch<1>(C, stream, HB_K_SORT_PASS);
      else if (key_kind == 2) rc = cluster_sort_launch<2>(C, stream, HB_K_SORT_PASS);
      else rc = cluster_sort_launch<0>(C, stream, HB_K_SORT_PASS);
      if (rc != HB_OK) return rc;
    }

    // 4. fused duplicate-sum + sink: the short kernels of every vector class (they also
    //    queue the hot rows), then the long kernels
    if (!(phases & kPhaseApply)) continue;
    for (int which = 0; which < 2; ++which)
    for (int V = 1; V <= 8; V <<= 1) {
      UpdParams U;
      U.wait = wait ? *wait : WaitSpec{nullptr, 0, 0};
      if (is_emit) U.emit = *emit;
      else { for (int i = 0; i < kMaxWorld; ++i) U.emit.peers.p[i] = nullptr; U.emit.window_off = 0; U.emit.world = 1; }
      U.status = d_status;
      U.long_count = reinterpret_cast<int32_t*>(base + S.long_count);
      U.items = reinterpret_cast<LongItem*>(base + S.items);
      U.part = reinterpret_cast<float*>(base + S.part);
      U.tickets = reinterpret_cast<int32_t*>(base + S.tickets);
      U.item_cap = S.item_cap; U.part_cap = S.part_cap; U.part_stride = S.part_stride;
      U.ticket_idx = 2 + (V == 1 ? 0 : (V == 2 ? 1 : (V == 4 ? 2 : 3)));
      U.opt = opt->kind;
      U.fast = (opt->flags & HB_OPT_FLAG_FAST_MATH) ? 1 : 0;
      U.lr = opt->lr;
      U.beta1 = opt->beta1; U.beta2 = opt->beta2; U.eps = opt->eps;
      U.omb1 = 1.0f - opt->beta1; U.omb2 = 1.0f - opt->beta2;
      if (opt->kind == HB_OPT_LAZY_ADAM) {
        const double t = (double)(opt->step < 1 ? 1 : opt->step);
        U.lr = (float)((double)opt->lr * sqrt(1.0 - pow((double)opt->beta2, t)) /
                       (1.0 - pow((double)opt->beta1, t)));
      }
      bool any = false;
      U.nfeats = nc;  // queued pieces name features by their index in the chunk
      for (int k = 0; k < nc; ++k) {
        const hbUpdateFeature& f = feats[c0 + k];
        const UpdExtra* ex = extras ? &extras[c0 + k] : nullptr;
        const bool sel = L[k].V == V && f.nnz > 0;
        any = any || sel;
        UpdFeat& F = U.f[k];
        const bool inA = (passes[k] & 1) == 1;
        F.table = f.table; F.slot0 = f.slot0; F.slot1 = f.slot1;
        F.grad = f.grad; F.offsets = f.offsets;
        F.ukey = reinterpret_cast<uint32_t*>(base + L[k].ukey);
        F.ustart = reinterpret_cast<int32_t*>(base + L[k].ustart);
        F.counts = reinterpret_cast<int32_t*>(base + L[k].counts);
        F.ubag4 = reinterpret_cast<const int4*>(base + L[k].ubag4);
        F.vals = reinterpret_cast<int32_t*>(base + (inA ? L[k].valsA : L[k].valsB));
        F.pos2bag = (ex && ex->inv != nullptr && f.offsets != nullptr)
                        ? reinterpret_cast<const int32_t*>(base + L[k].bagmap) : nullptr;
        F.emit_send_off = ex ? ex->emit_send_off : nullptr;
        F.emit_remote_base = ex ? ex->emit_remote_base : nullptr;
        F.emit_off = ex ? ex->emit_off : 0;
        F.emit_cap = ex ? ex->emit_cap : 0;
        F.rows = f.rows; F.grad_stride = f.grad_stride;
        F.dim = f.dim; F.combiner = f.combiner;
        F.log2g = L[k].log2g;
        const int ng = 32 >> L[k].log2g;
        F.max_chunks = sel ? (int)((f.nnz + ng - 1) / ng) : 0;  // 0: not this launch's vector class
        F.piece = piece_rows(f.dim);
      }
      if (!any) continue;
      rc = launch_apply_v(V, is_emit, U, which, stream);
      if (rc != HB_OK) return rc;
      wait = nullptr;  // later launches are stream-ordered behind the first
    }
  }
  return HB_OK;
}

}  // namespace hb

extern "C" {

// debug only (not in the public header): copies the 64 stamps of the last sort to the host
int hbDebugSortTiming(unsigned long long* h_out) {
  if (hb::g_cs_timing == nullptr) return HB_ERR_INVALID;
  return cudaMemcpy(h_out, hb::g_cs_timing, 64 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

int hbGroupSparseUpdateWorkspaceBytes(int n, const hbUpdateFeature* feats, size_t* bytes) {
  using namespace hb;
  HB_REQUIRE(n >= 1 && feats && bytes, "hbGroupSparseUpdateWorkspaceBytes: bad argument");
  for (int k = 0; k < n; ++k)
    HB_REQUIRE(feats[k].nnz >= 0 && feats[k].nnz <= INT32_MAX / 2 && feats[k].dim >= 4 &&
                   feats[k].dim % 4 == 0 && feats[k].dim <= 1024,
               "hbGroupSparseUpdateWorkspaceBytes: feature %d bad nnz/dim", k);
  *bytes = sparse_update_workspace_bytes(n, feats);
  return HB_OK;
}

int hbGroupLookupBackwardUpdate(int n, const hbUpdateFeature* feats, const hbOptimizer* opt,
                                void* d_workspace, size_t workspace_bytes, int32_t* d_status,
                                hbStream stream) {
  return hb::sparse_update_run(n, feats, opt, d_workspace, workspace_bytes, d_status,
                               (cudaStream_t)stream, nullptr, nullptr, nullptr,
                               hb::kPhaseSort | hb::kPhaseApply, nullptr);
}

int hbGroupSparseSort(int n, const hbUpdateFeature* feats, void* d_workspace, size_t workspace_bytes,
                      int32_t* d_status, hbStream stream) {
  return hb::sparse_update_run(n, feats, nullptr, d_workspace, workspace_bytes, d_status,
                               (cudaStream_t)stream, nullptr, nullptr, nullptr, hb::kPhaseSort, nullptr);
}

int hbGroupSparseApply(int n, const hbUpdateFeature* feats, const hbOptimizer* opt, void* d_workspace,
                       size_t workspace_bytes, int32_t* d_status, hbStream stream) {
  return hb::sparse_update_run(n, feats, opt, d_workspace, workspace_bytes, d_status,
                               (cudaStream_t)stream, nullptr, nullptr, nullptr, hb::kPhaseApply, nullptr);
}

}  // extern "C"
