// cluster_sort.cuh -- the whole "sort + run detection" phase of the sparse update as ONE
// launch: a thread-block cluster of 8 CTAs per feature.
//
// The per-digit kernels of bucket.cuh (global histogram, ticketed tiles, status-word
// look-back) pay a kernel boundary, a ticket and a look-back per digit: ~25 us per pass
// for the 65 536 entries a feature has per step, i.e. pure latency.  A feature's entries
// fit the tiles of one cluster, so here the passes are phases of one kernel separated
// by the hardware cluster barrier (barrier.cluster, ~1 us):
//   per digit   phase 1: every CTA ranks its tile(s) stably (warp match_any + running
//                        per-warp counters) and writes the tile's bin totals to hist[tile][bin]
//               barrier
//               phase 2: thread b sums hist[.][b] over the preceding tiles and over all
//                        tiles (bin bases), the tile is staged bin-sorted in shared
//                        memory and written out in contiguous runs
//               barrier
//   runs        every CTA scans its tile of the sorted keys: head flags, unique keys, run
//               starts, the first four values of every run, the inverse map (requester
//               side of the sharded path) -- tile prefixes again through one barrier.
// Same stable order as the reference CPU functor and as bucket.cuh; the data ping-pongs
// through global memory (L2 resident), read with ld.global.cg because other SMs wrote it.
// Features are independent: no grid-wide dependency, no tickets, no memset of scratch.
#pragma once
#include <cooperative_groups.h>

#include "bucket.cuh"

namespace hb {

namespace cg = cooperative_groups;

constexpr int kCsThreads = 512;
constexpr int kCsWarps = kCsThreads / 32;
constexpr int kCsItems = 16;
constexpr int kCsTile = kCsThreads * kCsItems;   // 8192
constexpr int kCsCluster = 8;
constexpr int kCsBits = 9;
constexpr int kCsBins = 1 << kCsBits;
constexpr int kCsMaxFeats = 96;
static_assert(kCsThreads == kCsBins, "thread b owns bin b in the histogram / base phases");

struct CsFeat {
  const void* in_keys;       // int64 ids (key kinds 0, 1) or uint32 keys (kind 2)
  const int32_t* in_vals;    // nullptr: value = input index
  uint32_t* keys[2];         // ping-pong; pass p writes keys[p & 1]
  int32_t* vals[2];
  uint32_t* hist;            // [tiles][kCsBins] scratch
  int32_t* tile_uniq;        // [tiles] scratch
  uint32_t* ukey;            // run outputs (see sparse_update.cu)
  int32_t* ustart;
  int4* ubag4;
  int32_t* counts;
  int32_t* inv;              // may be nullptr
  int32_t* owner_start1;     // may be nullptr
  const int32_t* n_dev;      // may be nullptr
  int32_t n;
  int32_t passes;
  int32_t lbits;
  uint32_t key_limit;
};

struct CsParams {
  CsFeat f[kCsMaxFeats];
  int32_t* d_status;
  int32_t* zero;             // words zeroed by the kernel (hot-row queue counters / tickets)
  int32_t zero_words;
  int32_t nfeats;
  int32_t key_kind;
  int32_t p;                 // composite keys: W
  int32_t div_shift;
  int64_t div;
  uint8_t order[kCsMaxFeats];   // cluster c sorts feature order[c]: most passes first, so the longest
                                // clusters start first and the short ones fill in behind them
  unsigned long long* timing;   // debug (HB_CS_TIMING=1): globaltimer stamps of feature 0, cluster rank 0
};

__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t* p) { return __ldcg(p); }

// block-wide exclusive scan of one int per thread (kCsThreads threads)
__device__ __forceinline__ int32_t cs_block_excl_scan(int32_t x, int32_t* s_warp /*[kCsWarps]*/, int32_t* total) {
  const unsigned lane = threadIdx.x & 31u;
  const int warp = threadIdx.x >> 5;
  int32_t incl = x;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= (unsigned)off) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int32_t wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kCsWarps; ++w) {
    const int32_t t = s_warp[w];
    if (w < warp) wbase += t;
    tot += t;
  }
  __syncthreads();
  if (total != nullptr) *total = tot;
  return wbase + incl - x;
}

// raw key of entry i (all loads of a tile are issued before the first conversion: the
// conversion branches, and a load behind a branch is a serialized round trip)
template <int KIND>
__device__ __forceinline__ int64_t cs_load(const CsFeat& F, int i) {
  if constexpr (KIND == 2) return (int64_t)reinterpret_cast<const uint32_t*>(F.in_keys)[i];
  else return ld_nc_i64(reinterpret_cast<const int64_t*>(F.in_keys) + i);
}

template <int KIND>
__device__ __forceinline__ uint32_t cs_conv(const CsParams& P, const CsFeat& F, int64_t v) {
  if constexpr (KIND == 2) {
    const uint32_t k = (uint32_t)v;
    if (k == 0xFFFFFFFFu) return k;
    return k >= F.key_limit ? 0xFFFFFFFEu : k;
  } else {
    if (v == INT64_MIN) return 0xFFFFFFFFu;  // padding entry: skipped silently
    if (v < 0) return 0xFFFFFFFEu;           // invalid id: skipped, raises the status word
    if constexpr (KIND == 0) {
      const uint64_t r = (P.div_shift >= 0) ? ((uint64_t)v >> P.div_shift) : (uint64_t)(v / P.div);
      return (r >= (uint64_t)F.key_limit) ? 0xFFFFFFFEu : (uint32_t)r;
    } else {
      uint64_t own, r;
      if (P.div_shift >= 0) { own = (uint64_t)v & (uint64_t)(P.p - 1); r = (uint64_t)v >> P.div_shift; }
      else { own = (uint64_t)v % (uint64_t)P.p; r = (uint64_t)v / (uint64_t)P.p; }
      if (r >= (uint64_t)F.key_limit) return 0xFFFFFFFEu;
      return (uint32_t)((own << F.lbits) | r);
    }
  }
}

// shared memory of one CTA (dynamic): staging of one tile + per-warp bin counters
struct CsSmem {
  uint32_t keys[kCsTile];
  int32_t vals[kCsTile];
  uint16_t wcnt[kCsWarps][kCsBins];   // per-warp counts -> exclusive prefix over warps
  int32_t gbase[kCsBins];
  int32_t lstart[kCsBins + 1];
  int32_t scan[kCsWarps];
};

template <int KIND>
__global__ void __launch_bounds__(kCsThreads, 2)
cluster_sort_runs_kernel(const __grid_constant__ CsParams P) {
  extern __shared__ __align__(16) unsigned char cs_raw[];
  CsSmem& S = *reinterpret_cast<CsSmem*>(cs_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int fi = P.order[blockIdx.x / kCsCluster];
  const CsFeat& F = P.f[fi];
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const int warp = tid >> 5;

  // the counters of the apply phase (stream-ordered behind this kernel)
  for (int i = blockIdx.x * kCsThreads + tid; i < P.zero_words; i += gridDim.x * kCsThreads) P.zero[i] = 0;

  int n = F.n;
  if (F.n_dev != nullptr) {
    const int d = *F.n_dev;
    n = d < 0 ? 0 : (d < F.n ? d : F.n);
  }
  const int T = (n + kCsTile - 1) / kCsTile;
  const bool single = T <= kCsCluster;   // at most one tile per CTA: ranks stay in registers
  // One tile in all (small batches, the owner side of the sharded path): rank 0 sorts it
  // entirely in shared memory -- no ping-pong through global memory, no fences, no cluster
  // barriers -- and the other CTAs of the cluster leave (nobody ever waits on a barrier).
  const bool local = T <= 1;
  if (local && crank != 0) return;

  int stamp_i = 0;
  auto stamp = [&]() {
    if (P.timing != nullptr && blockIdx.x == 0 && tid == 0 && stamp_i < 24) P.timing[stamp_i] = globaltimer_ns();
    ++stamp_i;
  };
  stamp();
  uint32_t key[kCsItems];
  uint32_t rnk2[kCsItems / 2];   // stable rank of the item among the warp's items of the same bin, two per word

  // load + convert the keys of tile t for pass p, rank them inside the warp slices
  auto rank_tile = [&](int t, int p) {
    stamp();
    for (int b = tid; b < kCsWarps * kCsBins; b += kCsThreads) (&S.wcnt[0][0])[b] = 0;
    const int wbase = t * kCsTile + warp * (kCsItems * 32);
    const int shift = p * kCsBits;
    if (p == 0) {
      int64_t raw[kCsItems];
#pragma unroll
      for (int j = 0; j < kCsItems; ++j) {
        const int i = wbase + j * 32 + (int)lane;
        raw[j] = (i < n) ? cs_load<KIND>(F, i) : INT64_MIN;
      }
#pragma unroll
      for (int j = 0; j < kCsItems; ++j) {
        const int i = wbase + j * 32 + (int)lane;
        key[j] = (i < n) ? cs_conv<KIND>(P, F, raw[j]) : 0xFFFFFFFFu;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kCsItems; ++j) {
        const int i = wbase + j * 32 + (int)lane;
        if (local) key[j] = (i < n) ? S.keys[i] : 0xFFFFFFFFu;   // staged by the previous pass
        else key[j] = (i < n) ? ldcg_u32(F.keys[(p - 1) & 1] + i) : 0xFFFFFFFFu;
      }
    }
    __syncthreads();
    stamp();
    uint16_t* wc = S.wcnt[warp];
    // (a) lanes holding the same bin, for all rounds up front: nine independent ballots per
    //     round (fixed latency; match.any takes one iteration per distinct value)
    uint32_t pk2[kCsItems / 2];   // per item 16 bits: lower peers | (peers - 1) << 6 | leader << 12 | valid << 13
    const bool full = wbase + kCsItems * 32 <= n;
#pragma unroll
    for (int j = 0; j < kCsItems; ++j) {
      const int i = wbase + j * 32 + (int)lane;
      const bool valid = full || i < n;
      const uint32_t b = (key[j] >> shift) & (kCsBins - 1);
      unsigned peers = 0xffffffffu;
      if (!full) {
        peers = __ballot_sync(0xffffffffu, valid);
        if (!valid) peers = ~peers;
      }
#pragma unroll
      for (int bit = 0; bit < kCsBits; ++bit) {
        const bool one = (b >> bit) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, one);
        peers &= one ? m : ~m;
      }
      const unsigned lower = peers & lanemask_lt();
      const uint32_t w16 = (uint32_t)__popc(lower) | ((uint32_t)(__popc(peers) - 1) << 6) |
                           ((lower == 0 ? 1u : 0u) << 12) | ((valid ? 1u : 0u) << 13);
      if (j & 1) pk2[j >> 1] |= w16 << 16; else pk2[j >> 1] = w16;
    }
    // (b) running per-warp counters: every lane reads its bin's counter (broadcast), the
    //     first lane of each group adds the group size
#pragma unroll
    for (int j = 0; j < kCsItems; ++j) {
      const uint32_t b = (key[j] >> shift) & (kCsBins - 1);
      const uint32_t w16 = (pk2[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
      const int base = wc[b];
      __syncwarp();
      if ((w16 >> 12) == 3u) wc[b] = (uint16_t)(base + 1 + (int)((w16 >> 6) & 63u));
      __syncwarp();
      const uint32_t rr = (uint32_t)(base + (int)(w16 & 63u));
      if (j & 1) rnk2[j >> 1] |= rr << 16; else rnk2[j >> 1] = rr;
    }
    __syncthreads();
  };

  for (int p = 0; p < F.passes; ++p) {
    const int shift = p * kCsBits;
    // ---- phase 1: tile totals per bin ---------------------------------------------------
    for (int t = crank; t < T; t += kCsCluster) {
      rank_tile(t, p);
      int tot = 0;
#pragma unroll
      for (int w = 0; w < kCsWarps; ++w) tot += S.wcnt[w][tid];
      F.hist[(size_t)t * kCsBins + tid] = (uint32_t)tot;
      if (!single) __syncthreads();
    }
    stamp();
    if (!local) __threadfence();
    stamp();
    if (!local) cluster.sync();
    stamp();
    // ---- phase 2: bases, staging, write-out ---------------------------------------------
    uint32_t* okeys = F.keys[p & 1];
    int32_t* ovals = F.vals[p & 1];
    for (int t = crank; t < T; t += kCsCluster) {
      int pre = 0, tot = 0;
      if (local) __syncthreads();   // hist row written by this CTA (read back through L2)
      for (int t0 = 0; t0 < T; t0 += 8) {
        uint32_t h[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) h[u] = (t0 + u < T) ? ldcg_u32(F.hist + (size_t)(t0 + u) * kCsBins + tid) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          tot += (int)h[u];
          if (t0 + u < t) pre += (int)h[u];
        }
      }
      if (!single) rank_tile(t, p);
      const int gb = cs_block_excl_scan(tot, S.scan, nullptr);
      S.gbase[tid] = gb + pre;
      // exclusive prefix over the warps of my bin, and the tile's total for it
      int run = 0;
#pragma unroll
      for (int w = 0; w < kCsWarps; ++w) {
        const int c = S.wcnt[w][tid];
        S.wcnt[w][tid] = (uint16_t)run;
        run += c;
      }
      int tile_n = 0;
      const int ls = cs_block_excl_scan(run, S.scan, &tile_n);
      S.lstart[tid] = ls;
      if (tid == 0) S.lstart[kCsBins] = tile_n;
      __syncthreads();
      // carried values: the input index in the first pass, else the previous pass's output
      const int wbase = t * kCsTile + warp * (kCsItems * 32);
      int32_t val[kCsItems];
#pragma unroll
      for (int j = 0; j < kCsItems; ++j) {
        const int i = wbase + j * 32 + (int)lane;
        if (p == 0) val[j] = (F.in_vals != nullptr && i < n) ? F.in_vals[i] : i;
        else if (local) val[j] = (i < n) ? S.vals[i] : 0;
        else val[j] = (i < n) ? __ldcg(F.vals[(p - 1) & 1] + i) : 0;
      }
      if (local) __syncthreads();   // every thread holds its keys / values before the stage is rewritten
#pragma unroll
      for (int j = 0; j < kCsItems; ++j) {
        const int i = wbase + j * 32 + (int)lane;
        if (i < n) {
          const int b = (int)((key[j] >> shift) & (kCsBins - 1));
          const int lpos = S.lstart[b] + (int)S.wcnt[warp][b] + (int)((rnk2[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu);
          S.keys[lpos] = key[j];
          S.vals[lpos] = val[j];
        }
      }
      __syncthreads();
      if (!local || p == F.passes - 1)
      for (int k = tid; k < tile_n; k += kCsThreads) {
        const uint32_t kk = S.keys[k];
        const int b = (int)((kk >> shift) & (kCsBins - 1));
        const int g = S.gbase[b] + (k - S.lstart[b]);
        okeys[g] = kk;
        ovals[g] = S.vals[k];
      }
      __syncthreads();
    }
    stamp();
    if (!local) __threadfence();
    stamp();
    if (!local) cluster.sync();
    stamp();
  }

  // ---- runs -------------------------------------------------------------------------------
  const uint32_t* skeys = F.keys[(F.passes - 1) & 1];
  const int32_t* svals = F.vals[(F.passes - 1) & 1];
  bool oob = false;
  uint32_t k[kCsItems + 2];
  uint32_t heads = 0;
  auto scan_tile = [&](int t) -> int {   // head flags of my 16 consecutive entries; returns their count
    const int i0 = t * kCsTile + tid * kCsItems;
    if (i0 + kCsItems <= n) {
      const uint4* pk = reinterpret_cast<const uint4*>(skeys + i0);
#pragma unroll
      for (int q = 0; q < kCsItems / 4; ++q) {
        const uint4 v = __ldcg(pk + q);
        k[1 + 4 * q] = v.x; k[2 + 4 * q] = v.y; k[3 + 4 * q] = v.z; k[4 + 4 * q] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kCsItems; ++j) k[j + 1] = (i0 + j < n) ? ldcg_u32(skeys + i0 + j) : 0xFFFFFFFFu;
    }
    k[0] = (i0 > 0 && i0 - 1 < n) ? ldcg_u32(skeys + i0 - 1) : 0xFFFFFFFFu;
    k[kCsItems + 1] = (i0 + kCsItems < n) ? ldcg_u32(skeys + i0 + kCsItems) : 0xFFFFFFFFu;
    heads = 0;
#pragma unroll
    for (int j = 0; j < kCsItems; ++j) {
      const int i = i0 + j;
      const bool valid = i < n && k[j + 1] < 0xFFFFFFFEu;
      if (i < n && k[j + 1] == 0xFFFFFFFEu) oob = true;
      if (valid && (i == 0 || k[j + 1] != k[j])) heads |= 1u << j;
    }
    return __popc(heads);
  };
  int excl = 0;
  for (int t = crank; t < T; t += kCsCluster) {
    int tile_total = 0;
    excl = cs_block_excl_scan(scan_tile(t), S.scan, &tile_total);
    if (tid == 0) F.tile_uniq[t] = tile_total;
  }
  stamp();
  if (!local) {
    __threadfence();
    cluster.sync();
  } else {
    __syncthreads();
  }
  stamp();
  for (int t = crank; t < T; t += kCsCluster) {
    if (!single) excl = cs_block_excl_scan(scan_tile(t), S.scan, nullptr);
    int pre = 0;
#pragma unroll 8
    for (int tp = 0; tp < t; ++tp) pre += __ldcg(F.tile_uniq + tp);
    const int i0 = t * kCsTile + tid * kCsItems;
    // values of my entries and of the three behind them (first four values of a run)
    int32_t sv[kCsItems + 3];
    if (i0 + kCsItems <= n) {
      const int4* pv = reinterpret_cast<const int4*>(svals + i0);
#pragma unroll
      for (int q = 0; q < kCsItems / 4; ++q) {
        const int4 v = __ldcg(pv + q);
        sv[4 * q] = v.x; sv[4 * q + 1] = v.y; sv[4 * q + 2] = v.z; sv[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kCsItems; ++j) sv[j] = (i0 + j < n) ? __ldcg(svals + i0 + j) : 0;
    }
#pragma unroll
    for (int j = kCsItems; j < kCsItems + 3; ++j) sv[j] = (i0 + j < n) ? __ldcg(svals + i0 + j) : 0;
    int u = pre + excl;   // index the next head of this thread gets
#pragma unroll
    for (int j = 0; j < kCsItems; ++j) {
      const int i = i0 + j;
      if (i >= n) break;
      const bool valid = k[j + 1] < 0xFFFFFFFEu;
      if ((heads >> j) & 1u) {
        F.ukey[u] = k[j + 1];
        F.ustart[u] = i;
        // first four values of the run (entries past its end are never used)
        F.ubag4[u] = make_int4(sv[j], sv[j + 1], sv[j + 2], sv[j + 3]);
        if (F.owner_start1 != nullptr) {
          const uint32_t own = k[j + 1] >> F.lbits;
          if (i == 0 || (k[j] >> F.lbits) != own) F.owner_start1[own] = u + 1;
        }
        ++u;
      }
      if (F.inv != nullptr) F.inv[sv[j]] = valid ? u - 1 : -1;
      if (valid && (i + 1 >= n || k[j + 2] >= 0xFFFFFFFEu)) {  // last valid entry of the feature
        F.ustart[u] = i + 1;
        F.counts[0] = u;
        F.counts[1] = i + 1;
      }
    }
  }
  if (crank == 0 && tid == 0 && (n == 0 || ldcg_u32(skeys) >= 0xFFFFFFFEu)) {  // no valid entry at all
    F.ustart[0] = 0;
    F.counts[0] = 0;
    F.counts[1] = 0;
  }
  stamp();
  if (P.timing != nullptr && crank == 0 && tid == 0 && fi < 40) P.timing[24 + fi] = globaltimer_ns();
  if (oob) raise_status(P.d_status, HB_STATUS_ID_OUT_OF_RANGE);
}

static inline int cs_tiles(int64_t n) { return (int)((n + kCsTile - 1) / kCsTile); }

template <int KIND>
static int cluster_sort_launch(const CsParams& P, cudaStream_t stream, int kid) {
  const size_t smem = sizeof(CsSmem);
  HB_CUDA_OK(cudaFuncSetAttribute(cluster_sort_runs_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P.nfeats * kCsCluster), 1, 1);
  cfg.blockDim = dim3(kCsThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCsCluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  KernelScope ks(kid, stream);
  HB_CUDA_OK(cudaLaunchKernelEx(&cfg, cluster_sort_runs_kernel<KIND>, P));
  return HB_OK;
}

}  // namespace hb
