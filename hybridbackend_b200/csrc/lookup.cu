// lookup.cu -- K3: fused multi-table gather + segment pooling (forward).
//
// One launch covers every feature of the group.  A feature with row width
// dim = 4*G*V floats is served by sub-warp groups of G lanes (G a power of two
// <= 32), each lane owning V float4 columns of the row, so a warp issues
// 128-bit loads that cover whole rows (D=32: 4 rows per warp instruction, each
// row one full 128-byte line).  Accumulation is in registers, fp32, in bag
// order -- the order TF-1.15's SparseSegment* CPU kernel uses -- so results are
// bit-identical to the oracle (oracle/hb_oracle.c: hbo_embedding_lookup_sparse).
// Memory-level parallelism: every group keeps kBagsPerGroup independent bags
// (one-id-per-bag fast path) or 4 ids of one bag in flight.
//
// HBM traffic per pooled row (algorithmic): L*(8 + 4*dim) read + 4*dim written.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace hb {

constexpr int kLookupThreads = 256;
constexpr int kBagsPerGroup = 4;
constexpr int kMaxLookupFeats = 128;

struct LookupFeat {
  const float* table;
  const int64_t* ids;
  const int64_t* offsets;
  const int32_t* idx32;   // when set: row index of position p is idx32[p]
  float* out;
  int64_t rows;
  int64_t out_stride;
  int64_t id_div;
  int64_t nnz;
  int32_t nbags;
  int32_t dim;
  int32_t combiner;
  int32_t div_shift;  // log2(id_div) or -1
  int32_t cta_begin;
  int32_t log2g;
};

struct LookupParams {
  LookupFeat f[kMaxLookupFeats];
  WaitSpec wait;
  int32_t* status;
  int32_t nfeats;
  int32_t total_ctas;
};

__device__ __forceinline__ int find_feat(const LookupParams& P, int cta) {
  int lo = 0, hi = P.nfeats - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.f[mid].cta_begin <= cta) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <bool COH>
__device__ __forceinline__ float4 ld_row4(const float4* p) {
  if constexpr (COH) return *p;
  else return ld_nc_f4(p);
}

// row index of id position p
__device__ __forceinline__ int64_t row_of(const LookupFeat& F, int64_t p);

__device__ __forceinline__ int64_t local_row(int64_t id, const LookupFeat& F) {
  if (id < 0) return -1;  // never a valid row
  if (F.div_shift >= 0) return (int64_t)((uint64_t)id >> F.div_shift);
  return id / F.id_div;
}

__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z),
                     __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4_div(const float4& a, float c) {
  return make_float4(__fdiv_rn(a.x, c), __fdiv_rn(a.y, c), __fdiv_rn(a.z, c), __fdiv_rn(a.w, c));
}

__device__ __forceinline__ int64_t row_of(const LookupFeat& F, int64_t p) {
  if (F.idx32 != nullptr) return (int64_t)F.idx32[p];
  return local_row(ld_nc_i64(F.ids + p), F);
}

// Row indices of the kBagsPerGroup bags a group owns in chunk `cid` of a
// one-id-per-bag feature: >= 0 row, -1 out-of-range id, -2 no such bag.
__device__ __forceinline__ void load_rows(const LookupFeat& F, int cid, int64_t (&r)[kBagsPerGroup],
                                          bool& oob) {
  const int groups = kLookupThreads >> F.log2g;
  const int g = threadIdx.x >> F.log2g;
  const int bag0 = (cid - F.cta_begin) * groups * kBagsPerGroup;
  // all id loads first (independent, in flight together), then the row math
  int64_t raw[kBagsPerGroup];
#pragma unroll
  for (int u = 0; u < kBagsPerGroup; ++u) {
    const int b = bag0 + u * groups + g;
    const int bb = b < F.nbags ? b : 0;
    raw[u] = (F.idx32 != nullptr) ? (int64_t)F.idx32[bb] : ld_nc_i64(F.ids + bb);
  }
#pragma unroll
  for (int u = 0; u < kBagsPerGroup; ++u) {
    const int b = bag0 + u * groups + g;
    int64_t row = (F.idx32 != nullptr) ? raw[u] : local_row(raw[u], F);
    const bool bad = (uint64_t)row >= (uint64_t)F.rows;
    if (b >= F.nbags) row = -2;
    else if (bad) { oob = true; row = -1; }
    r[u] = row;
  }
}

// Persistent CTAs walk the chunk list (chunk = groups*kBagsPerGroup bags of one
// feature).  For one-id-per-bag features the row indices of the NEXT chunk are
// fetched while the rows of the current chunk are in flight, so every iteration
// exposes a single memory round trip with 4 x 128-bit loads per lane outstanding.
template <int V, bool COH>
__global__ void __launch_bounds__(kLookupThreads, (V == 1 ? 4 : 1))
lookup_fwd_kernel(const __grid_constant__ LookupParams P) {
  wait_spec(P.wait, P.status);
  bool oob = false;
  bool bad_off = false;
  int cid = blockIdx.x;
  if (cid >= P.total_ctas) return;
  int fi = find_feat(P, cid);
  int64_t r[kBagsPerGroup];
#pragma unroll
  for (int u = 0; u < kBagsPerGroup; ++u) r[u] = -2;
  if (P.f[fi].offsets == nullptr) load_rows(P.f[fi], cid, r, oob);

  while (true) {
    const LookupFeat& F = P.f[fi];
    const int log2g = F.log2g;
    const int groups = kLookupThreads >> log2g;
    const int g = threadIdx.x >> log2g;
    const int l = threadIdx.x & ((1 << log2g) - 1);
    const int dim = F.dim;
    const int nbags = F.nbags;
    const int bag0 = (cid - F.cta_begin) * groups * kBagsPerGroup;
    // column c of lane l, vector v:  (v * G + l) * 4
    int col[V];
    bool act[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      col[v] = ((v << log2g) + l) * 4;
      act[v] = col[v] < dim;
    }
    const int nid = cid + gridDim.x;
    const int nfi = (nid < P.total_ctas) ? find_feat(P, nid) : -1;
    int64_t rn[kBagsPerGroup];
#pragma unroll
    for (int u = 0; u < kBagsPerGroup; ++u) rn[u] = -2;

    if (F.offsets == nullptr) {
      // ---- one id per bag: pure gather ---------------------------------------
      float4 val[kBagsPerGroup][V];
#pragma unroll
      for (int u = 0; u < kBagsPerGroup; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) {
          val[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r[u] >= 0 && act[v])
            val[u][v] = ld_row4<COH>(reinterpret_cast<const float4*>(F.table + r[u] * dim + col[v]));
        }
      if (nfi >= 0 && P.f[nfi].offsets == nullptr) load_rows(P.f[nfi], nid, rn, oob);
#pragma unroll
      for (int u = 0; u < kBagsPerGroup; ++u) {
        const int b = bag0 + u * groups + g;
        if (r[u] == -2) continue;
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (act[v])
            *reinterpret_cast<float4*>(F.out + (int64_t)b * F.out_stride + col[v]) = val[u][v];
      }
    } else {
      // ---- CSR bags: sequential fp32 accumulation in bag order ----------------
#pragma unroll 1
      for (int u = 0; u < kBagsPerGroup; ++u) {
        const int b = bag0 + u * groups + g;
        if (b >= nbags) break;
        int64_t s = F.offsets[b];
        int64_t e = F.offsets[b + 1];
        if (e < s || s < 0 || e > F.nnz) { bad_off = true; e = s; }
        float4 acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t p = s; p < e; p += 4) {
          int64_t rr[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int64_t pp = (p + k < e) ? p + k : p;
            rr[k] = (F.idx32 != nullptr) ? (int64_t)F.idx32[pp] : ld_nc_i64(F.ids + pp);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (F.idx32 == nullptr) rr[k] = local_row(rr[k], F);
            const bool bad = (uint64_t)rr[k] >= (uint64_t)F.rows;
            if (p + k >= e) rr[k] = -1;
            else if (bad) { oob = true; rr[k] = -1; }
          }
          float4 x[4][V];
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int v = 0; v < V; ++v) {
              x[k][v] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (rr[k] >= 0 && act[v])
                x[k][v] = ld_row4<COH>(reinterpret_cast<const float4*>(F.table + rr[k] * dim + col[v]));
            }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (p + k < e)
#pragma unroll
              for (int v = 0; v < V; ++v) acc[v] = f4_add(acc[v], x[k][v]);
        }
        const int64_t cnt = e - s;
        if (cnt > 0 && F.combiner != HB_SUM) {
          const float c = (F.combiner == HB_MEAN) ? (float)cnt : __fsqrt_rn((float)cnt);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = f4_div(acc[v], c);
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (act[v])
            *reinterpret_cast<float4*>(F.out + (int64_t)b * F.out_stride + col[v]) = acc[v];
      }
      if (nfi >= 0 && P.f[nfi].offsets == nullptr) load_rows(P.f[nfi], nid, rn, oob);
    }
    if (nfi < 0) break;
    cid = nid;
    fi = nfi;
#pragma unroll
    for (int u = 0; u < kBagsPerGroup; ++u) r[u] = rn[u];
  }
  if (oob) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
  if (bad_off) raise_status(P.status, HB_STATUS_BAD_OFFSETS);
}


// ---- one id per bag: pure row gather through the copy engine --------------------------
// out[b, :] = table[row(ids[b]), :].  No register ever holds row data: a lane starts
// a 1-D bulk copy (cp.async.bulk, SASS UBLKCP) of each of its rows into the warp's
// shared-memory stage, the copies complete on the warp's mbarrier, and the staged
// rows leave again as bulk stores into the concatenated output.  Three stages per
// warp rotate, so the loads of the next block of rows are in flight while the
// current one drains: kLtWarps x kLtStages x kLtStageBytes = 192 KB of rows in
// flight per SM, against ~16 KB with 128-bit register loads.
constexpr int kLtWarps = 8;
constexpr int kLtStages = 3;
constexpr int kLtStageBytes = 8192;
constexpr int kLtMaxRows = 64;   // rows per stage at most (two per lane)

struct LtFeat {
  const float* table;
  const int64_t* ids;
  float* out;
  int64_t rows;
  int64_t out_stride;
  int64_t id_div;
  int32_t nbags;
  int32_t dim;
  int32_t div_shift;
  int32_t unit_begin;   // first work unit of this feature
  int32_t rows_per_unit;
};

struct LtParams {
  LtFeat f[kMaxLookupFeats];
  int32_t* status;
  int32_t nfeats;
  int32_t total_units;
};

__device__ __forceinline__ int lt_find(const LtParams& P, int unit) {
  int lo = 0, hi = P.nfeats - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.f[mid].unit_begin <= unit) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// start the copies of one unit into stage `buf`; returns through the barrier
__device__ __forceinline__ void lt_issue(const LtParams& P, int unit, float* buf, uint64_t* bar,
                                         unsigned lane, bool& oob) {
  const int fi = lt_find(P, unit);
  const LtFeat& F = P.f[fi];
  const uint32_t row_bytes = (uint32_t)F.dim * 4u;
  const int b0 = (unit - F.unit_begin) * F.rows_per_unit;
  int64_t idv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int slot = r * 32 + (int)lane;
    const int b = b0 + slot;
    idv[r] = (slot < F.rows_per_unit && b < F.nbags) ? ld_nc_i64(F.ids + b) : (int64_t)-1;
  }
  uint32_t bytes = 0;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int slot = r * 32 + (int)lane;
    const int b = b0 + slot;
    const bool in_unit = slot < F.rows_per_unit && b < F.nbags;
    int64_t row = -1;
    if (in_unit && idv[r] >= 0)
      row = F.div_shift >= 0 ? (int64_t)((uint64_t)idv[r] >> F.div_shift) : idv[r] / F.id_div;
    const bool ok = in_unit && (uint64_t)row < (uint64_t)F.rows;
    if (in_unit && !ok) {   // out-of-range id: the output row is zeros
      oob = true;
      float4* z = reinterpret_cast<float4*>(buf + (size_t)slot * F.dim);
      for (int c = 0; c < F.dim / 4; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (ok) bulk_g2s(buf + (size_t)slot * F.dim, F.table + row * F.dim, row_bytes, bar);
    bytes += __popc(__ballot_sync(0xffffffffu, ok)) * row_bytes;
  }
  if (lane == 0) mbar_arrive_expect_tx(bar, bytes);
}

__device__ __forceinline__ void lt_store(const LtParams& P, int unit, const float* buf, unsigned lane) {
  const int fi = lt_find(P, unit);
  const LtFeat& F = P.f[fi];
  const uint32_t row_bytes = (uint32_t)F.dim * 4u;
  const int b0 = (unit - F.unit_begin) * F.rows_per_unit;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int slot = r * 32 + (int)lane;
    const int b = b0 + slot;
    if (slot < F.rows_per_unit && b < F.nbags)
      bulk_s2g(F.out + (int64_t)b * F.out_stride, buf + (size_t)slot * F.dim, row_bytes);
  }
  bulk_commit();
}

__global__ void __launch_bounds__(kLtWarps * 32, 1)
lookup_rows_tma_kernel(const __grid_constant__ LtParams P) {
  extern __shared__ __align__(128) unsigned char s_lt[];
  __shared__ uint64_t s_bar[kLtWarps][kLtStages];
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int s = 0; s < kLtStages; ++s) mbar_init(&s_bar[warp][s], 1);
  mbar_init_fence();
  __syncthreads();
  float* stage0 = reinterpret_cast<float*>(s_lt + (size_t)warp * kLtStages * kLtStageBytes);
  const int stride = gridDim.x * kLtWarps;
  int unit = blockIdx.x * kLtWarps + warp;
  uint32_t parity = 0;   // bit s: parity of the next completion of stage s
  bool oob = false;
  // prologue: two units in flight
  int st = 0;
  if (unit < P.total_units) lt_issue(P, unit, stage0, &s_bar[warp][0], lane, oob);
  if (unit + stride < P.total_units)
    lt_issue(P, unit + stride, stage0 + kLtStageBytes / 4, &s_bar[warp][1], lane, oob);
  while (unit < P.total_units) {
    mbar_wait(&s_bar[warp][st], (parity >> st) & 1u);
    parity ^= 1u << st;
    fence_proxy_async();   // zero rows written with ordinary stores -> visible to the bulk stores
    lt_store(P, unit, stage0 + (size_t)st * (kLtStageBytes / 4), lane);
    // refill the stage drained in the PREVIOUS iteration: its bulk stores must have
    // finished reading shared memory (all store groups but the one just committed)
    const int ahead = unit + 2 * stride;
    const int st2 = (st + 2) % kLtStages;
    if (ahead < P.total_units) {
      bulk_wait_read<1>();
      __syncwarp();
      lt_issue(P, ahead, stage0 + (size_t)st2 * (kLtStageBytes / 4), &s_bar[warp][st2], lane, oob);
    }
    unit += stride;
    st = (st + 1) % kLtStages;
  }
  bulk_wait<0>();          // global writes complete before the kernel ends
  if (oob) raise_status(P.status, HB_STATUS_ID_OUT_OF_RANGE);
}

// Opt-in (HB_TMA_GATHER=1).  Measured on B200 (profiles/r2_fwd_sweep.md): the copy-engine
// gather loses to the register path at every row width of the target range -- 152 vs 79 us
// (D=32), 193 vs 134 us (D=64), 291 vs 256 us (D=128) for 1.7 M rows -- because one bulk
// copy per 128..512-byte row is bound by the per-SM bulk-copy issue rate, not by bytes.
static bool use_tma_gather() {
  static const bool on = [] {
    const char* e = getenv("HB_TMA_GATHER");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

static int launch_rows_tma(const LtParams& P, cudaStream_t stream, int kid) {
  if (P.total_units == 0) return HB_OK;
  const size_t smem = (size_t)kLtWarps * kLtStages * kLtStageBytes;
  HB_CUDA_OK(cudaFuncSetAttribute(lookup_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = device_sm_count();
  const int need = (P.total_units + kLtWarps - 1) / kLtWarps;
  if (grid > need) grid = need;
  KernelScope ks(kid, stream);
  lookup_rows_tma_kernel<<<grid, kLtWarps * 32, smem, stream>>>(P);
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

static int ilog2_ceil(int x) {
  int l = 0;
  while ((1 << l) < x) ++l;
  return l;
}

// Shape of the work decomposition for one feature: G lanes/row, V vectors/lane.
static void lookup_shape(int dim, int* log2g, int* v) {
  const int vecs = dim / 4;
  if (vecs <= 32) { *log2g = ilog2_ceil(vecs); *v = 1; return; }
  *log2g = 5;
  int vv = (vecs + 31) / 32;
  int p = 1;
  while (p < vv) p <<= 1;
  *v = p;
}

template <int V>
static int launch_lookup(const LookupParams& P, cudaStream_t stream, bool coherent, int kid) {
  if (P.total_ctas == 0) return HB_OK;
  {
    // persistent: exactly the co-resident number of CTAs walks all chunks
    int per_sm = 0;
    if (coherent)
      HB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lookup_fwd_kernel<V, true>, kLookupThreads, 0));
    else
      HB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lookup_fwd_kernel<V, false>, kLookupThreads, 0));
    const int maxg = device_sm_count() * (per_sm > 0 ? per_sm : 1);
    const int grid = P.total_ctas < maxg ? P.total_ctas : maxg;
    KernelScope ks(kid, stream);
    if (coherent) lookup_fwd_kernel<V, true><<<grid, kLookupThreads, 0, stream>>>(P);
    else lookup_fwd_kernel<V, false><<<grid, kLookupThreads, 0, stream>>>(P);
  }
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

static int validate_feature(int k, const hbLookupFeature& f, bool has_idx32) {
  HB_REQUIRE(f.dim >= 4 && f.dim % 4 == 0 && f.dim <= 1024,
             "lookup: feature %d dim %d must be a multiple of 4 in [4,1024]", k, f.dim);
  HB_REQUIRE(f.nbags >= 0 && f.nbags <= INT32_MAX, "lookup: feature %d bad nbags", k);
  HB_REQUIRE(f.rows >= 0, "lookup: feature %d negative rows", k);
  HB_REQUIRE(f.id_div >= 1, "lookup: feature %d id_div must be >= 1", k);
  HB_REQUIRE(f.out_stride >= f.dim && f.out_stride % 4 == 0,
             "lookup: feature %d out_stride %lld must be a multiple of 4 and >= dim", k,
             (long long)f.out_stride);
  HB_REQUIRE(f.combiner >= HB_SUM && f.combiner <= HB_SQRTN, "lookup: feature %d bad combiner", k);
  HB_REQUIRE(f.nnz >= 0 && (f.offsets != nullptr || f.nnz == f.nbags),
             "lookup: feature %d nnz %lld must be the id count (== nbags without offsets)", k, (long long)f.nnz);
  if (f.nbags > 0) {
    HB_REQUIRE(f.table && (f.ids || has_idx32) && f.out, "lookup: feature %d null pointer", k);
    HB_REQUIRE(((uintptr_t)f.table & 15) == 0 && ((uintptr_t)f.out & 15) == 0,
               "lookup: feature %d table/out must be 16-byte aligned", k);
  }
  return HB_OK;
}

int lookup_forward_run(int n, const hbLookupFeature* feats, const int32_t* const* idx32,
                       const WaitSpec* wait, bool coherent, int32_t* d_status,
                       cudaStream_t stream, int kernel_id) {
  HB_REQUIRE(n >= 1 && feats != nullptr, "hbGroupLookupForward: need n >= 1 features");
  for (int k = 0; k < n; ++k) {
    int rc = validate_feature(k, feats[k], idx32 != nullptr && idx32[k] != nullptr);
    if (rc != HB_OK) return rc;
  }
  bool waited = false;
  // one id per bag, local table, no peer-written data: the copy-engine gather
  std::vector<char> done(n, 0);
  if (use_tma_gather() && wait == nullptr && !coherent) {
    LtParams T;
    T.status = d_status;
    T.nfeats = 0;
    T.total_units = 0;
    auto flush_t = [&]() -> int {
      const int rc = T.nfeats > 0 ? launch_rows_tma(T, stream, kernel_id) : HB_OK;
      T.nfeats = 0;
      T.total_units = 0;
      return rc;
    };
    for (int k = 0; k < n; ++k) {
      const hbLookupFeature& f = feats[k];
      if (f.offsets != nullptr || (idx32 != nullptr && idx32[k] != nullptr) || f.nbags == 0) continue;
      if (f.dim * 4 > kLtStageBytes / 2 || ((uintptr_t)f.out & 15) || (f.out_stride & 3)) continue;
      LtFeat& F = T.f[T.nfeats++];
      F.table = f.table; F.ids = f.ids; F.out = f.out;
      F.rows = f.rows; F.out_stride = f.out_stride; F.id_div = f.id_div;
      F.nbags = (int32_t)f.nbags; F.dim = f.dim;
      F.div_shift = ((f.id_div & (f.id_div - 1)) == 0 && f.id_div <= (1 << 30)) ? ilog2_ceil((int)f.id_div) : -1;
      int rpu = kLtStageBytes / (f.dim * 4);
      if (rpu > kLtMaxRows) rpu = kLtMaxRows;
      F.rows_per_unit = rpu;
      F.unit_begin = T.total_units;
      T.total_units += (int)((f.nbags + rpu - 1) / rpu);
      done[k] = 1;
      if (T.nfeats == kMaxLookupFeats) {
        const int rc = flush_t();
        if (rc != HB_OK) return rc;
      }
    }
    const int rc = flush_t();
    if (rc != HB_OK) return rc;
  }
  // one launch per distinct V (1 for every dim <= 128), chunks of kMaxLookupFeats
  for (int V = 1; V <= 8; V <<= 1) {
    LookupParams P;
    P.status = d_status;
    P.nfeats = 0;
    P.total_ctas = 0;
    auto flush = [&]() -> int {
      int rc = HB_OK;
      if (P.nfeats > 0) {
        // only the first launch has to wait: later ones are stream-ordered after it
        P.wait = (wait != nullptr && !waited) ? *wait : WaitSpec{nullptr, 0, 0};
        waited = true;
        switch (V) {
          case 1: rc = launch_lookup<1>(P, stream, coherent, kernel_id); break;
          case 2: rc = launch_lookup<2>(P, stream, coherent, kernel_id); break;
          case 4: rc = launch_lookup<4>(P, stream, coherent, kernel_id); break;
          default: rc = launch_lookup<8>(P, stream, coherent, kernel_id); break;
        }
      }
      P.nfeats = 0;
      P.total_ctas = 0;
      return rc;
    };
    for (int k = 0; k < n; ++k) {
      const hbLookupFeature& f = feats[k];
      int log2g, v;
      lookup_shape(f.dim, &log2g, &v);
      if (v != V || f.nbags == 0 || done[k]) continue;
      LookupFeat& F = P.f[P.nfeats];
      F.table = f.table; F.ids = f.ids; F.offsets = f.offsets; F.out = f.out;
      F.idx32 = idx32 ? idx32[k] : nullptr;
      F.rows = f.rows; F.out_stride = f.out_stride; F.id_div = f.id_div; F.nnz = f.nnz;
      F.nbags = (int32_t)f.nbags; F.dim = f.dim; F.combiner = f.combiner;
      F.div_shift = ((f.id_div & (f.id_div - 1)) == 0) ? ilog2_ceil((int)f.id_div) : -1;
      if (f.id_div > (1 << 30)) F.div_shift = -1;
      F.cta_begin = P.total_ctas;
      F.log2g = log2g;
      const int groups = kLookupThreads >> log2g;
      const int per_cta = groups * kBagsPerGroup;
      P.total_ctas += (int)((f.nbags + per_cta - 1) / per_cta);
      P.nfeats++;
      if (P.nfeats == kMaxLookupFeats) {
        int rc = flush();
        if (rc != HB_OK) return rc;
      }
    }
    int rc = flush();
    if (rc != HB_OK) return rc;
  }
  return HB_OK;
}

}  // namespace hb

extern "C" int hbGroupLookupForward(int n, const hbLookupFeature* feats, int32_t* d_status,
                                    hbStream stream_) {
  return hb::lookup_forward_run(n, feats, nullptr, nullptr, false, d_status,
                                (cudaStream_t)stream_, HB_K_LOOKUP_FWD);
}

extern "C" int hbGroupLookupForwardHost(int n, const hbLookupFeature* feats, const void* h_in_block,
                                        void* d_in_block, size_t in_block_bytes,
                                        const void* d_out_block, void* h_out_block,
                                        size_t out_block_bytes, int32_t* d_status,
                                        hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(in_block_bytes == 0 || (h_in_block && d_in_block), "hbGroupLookupForwardHost: null input block");
  HB_REQUIRE(out_block_bytes == 0 || (h_out_block && d_out_block), "hbGroupLookupForwardHost: null output block");
  if (in_block_bytes > 0)
    HB_CUDA_OK(cudaMemcpyAsync(d_in_block, h_in_block, in_block_bytes, cudaMemcpyHostToDevice, stream));
  int rc = hbGroupLookupForward(n, feats, d_status, stream_);
  if (rc != HB_OK) return rc;
  if (out_block_bytes > 0)
    HB_CUDA_OK(cudaMemcpyAsync(h_out_block, d_out_block, out_block_bytes, cudaMemcpyDeviceToHost, stream));
  return HB_OK;
}
