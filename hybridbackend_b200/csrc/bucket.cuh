// bucket.cuh -- stable multi-segment bucket scatter for sm_100a.
//
// One mechanism serves two users:
//   * K1 partition (HbPartitionByModulo[N] / dual modulo): bin = shard(id),
//     emits permuted ids, per-bin sizes and the inverse permutation;
//   * the LSD radix sort that groups row ids before the fused sparse update
//     (bin = 9-bit digit, carries the original position as value).
// Three launches cover ALL segments (features) at once:
//   count   : per (segment, tile) histogram            -> counts[seg][bin][tile]
//   scan    : per segment exclusive scan in (bin, tile) order (+ sizes[bin])
//   scatter : per tile stable rank (warp match_any + per-warp running counters),
//             tile staged bin-sorted in shared memory so that global writes are
//             contiguous runs per bin (coalesced), inverse written by input index.
// Stability: warp w owns a contiguous slice of the tile and walks it in rounds
// of 32 consecutive items; rank = (# earlier warps' items of the bin) +
// (running count in this warp) + (# lower lanes with the same bin this round).
// Hence output order inside a bin == input order, i.e. bit-identical to the
// reference CPU counting sort (partition_by_modulo_functors.cc:48-69).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hb {

constexpr int kBucketThreads = 256;
constexpr int kBucketWarps = kBucketThreads / 32;
constexpr int kBucketItems = 8;
constexpr int kBucketTile = kBucketThreads * kBucketItems;  // 2048
constexpr int kMaxSegs = 128;
constexpr int kMaxBins = 512;

struct BucketSeg {
  const void* in_keys;
  const int32_t* in_vals;  // nullptr: value = input index
  void* out_keys;          // may be nullptr (keys not needed downstream)
  int32_t* out_vals;       // may be nullptr
  int32_t* out_inv;        // may be nullptr: out_inv[i] = output position of i
  int32_t* out_sizes;      // may be nullptr: [nbins]
  int32_t n;
  int32_t tile_begin;      // global tile index of this segment's first tile
  int32_t shift;           // radix digit shift
  uint32_t key_limit;      // radix-first: keys >= key_limit become 0xFFFFFFFF (0: off)
};

struct BucketParams {
  BucketSeg seg[kMaxSegs];
  int32_t* counts;     // [total_tiles * nbins]; seg s at tile_begin[s]*nbins
  int32_t nsegs;
  int32_t nbins;
  int32_t total_tiles;
  int32_t p;           // num_partitions (modulo modes)
  int32_t m;           // modulus (dual modulo)
  int32_t pow2_mask;   // p-1 if p is a power of two else -1
  int64_t div;         // radix-first: key = id / div
  int32_t div_shift;   // log2(div) if power of two else -1
  int32_t pad;
};

// ---- bin traits -----------------------------------------------------------
template <typename T>
__device__ __forceinline__ int floor_mod(T v, int p) {
  // (v % p + p) % p of the reference, in T's arithmetic.
  if constexpr (std::is_signed<T>::value) {
    T r = v % (T)p;
    if (r < 0) r += (T)p;
    return (int)r;
  } else {
    return (int)(v % (T)p);
  }
}

template <typename T>
struct ModuloTraits {  // partition_by_modulo_functors.cc:56-57
  using In = T;
  using Out = T;
  static __device__ __forceinline__ Out conv(In v, const BucketParams&, const BucketSeg&) { return v; }
  static __device__ __forceinline__ int bin(In v, const BucketParams& P, const BucketSeg&) {
    if (P.pow2_mask >= 0) return (int)(v & (T)P.pow2_mask);
    return floor_mod<T>(v, P.p);
  }
};

template <typename T, int STAGE>
struct DualModuloTraits {  // partition_by_dual_modulo_functors.cc:37-49,:66-71
  using In = T;
  using Out = T;
  static __device__ __forceinline__ Out conv(In v, const BucketParams&, const BucketSeg&) { return v; }
  static __device__ __forceinline__ int bin(In v, const BucketParams& P, const BucketSeg&) {
    const int pre = floor_mod<T>(v, P.p * P.m);
    return STAGE == 1 ? pre % P.p : pre / P.m;
  }
};

struct RadixFirstTraits {  // int64 global id -> uint32 local row, first digit
  using In = int64_t;
  using Out = uint32_t;
  static __device__ __forceinline__ Out conv(In v, const BucketParams& P, const BucketSeg& sg) {
    if (v == INT64_MIN) return 0xFFFFFFFFu;  // padding entry: skipped silently
    if (v < 0) return 0xFFFFFFFEu;           // invalid id: skipped, raises the status word
    const uint64_t r = (P.div_shift >= 0) ? ((uint64_t)v >> P.div_shift) : (uint64_t)(v / P.div);
    return (r >= (uint64_t)sg.key_limit) ? 0xFFFFFFFEu : (uint32_t)r;
  }
  static __device__ __forceinline__ int bin(In v, const BucketParams& P, const BucketSeg& sg) {
    return (int)((conv(v, P, sg) >> sg.shift) & (uint32_t)(P.nbins - 1));
  }
};

struct RadixNextTraits {
  using In = uint32_t;
  using Out = uint32_t;
  static __device__ __forceinline__ Out conv(In v, const BucketParams&, const BucketSeg&) { return v; }
  static __device__ __forceinline__ int bin(In v, const BucketParams& P, const BucketSeg& sg) {
    return (int)((v >> sg.shift) & (uint32_t)(P.nbins - 1));
  }
};

__device__ __forceinline__ int find_seg(const BucketParams& P, int tile) {
  int lo = 0, hi = P.nsegs - 1;
  while (lo < hi) {  // last seg with tile_begin <= tile (skips empty segs)
    const int mid = (lo + hi + 1) >> 1;
    if (P.seg[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---- count ----------------------------------------------------------------
template <typename Tr>
__global__ void __launch_bounds__(kBucketThreads)
bucket_count_kernel(const __grid_constant__ BucketParams P) {
  extern __shared__ int32_t s_hist[];  // [nbins]
  using In = typename Tr::In;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const int s = find_seg(P, tile);
    const BucketSeg& sg = P.seg[s];
    const int t = tile - sg.tile_begin;
    const int ntiles = (sg.n + kBucketTile - 1) / kBucketTile;
    for (int b = threadIdx.x; b < P.nbins; b += kBucketThreads) s_hist[b] = 0;
    __syncthreads();
    const In* in = reinterpret_cast<const In*>(sg.in_keys);
    const int base = t * kBucketTile;
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = base + j * kBucketThreads + threadIdx.x;
      const bool valid = i < sg.n;
      int b = -1;
      if (valid) b = Tr::bin(in[i], P, sg);
      // warp-aggregate: one shared atomic per distinct bin per warp
      const unsigned peers = __match_any_sync(0xffffffffu, b);
      if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&s_hist[b], __popc(peers));
    }
    __syncthreads();
    int32_t* cnt = P.counts + (size_t)sg.tile_begin * P.nbins;
    for (int b = threadIdx.x; b < P.nbins; b += kBucketThreads)
      cnt[(size_t)b * ntiles + t] = s_hist[b];
    __syncthreads();
  }
}

// ---- scan -----------------------------------------------------------------
// One CTA per segment: exclusive scan of counts in (bin-major, tile-minor) order.
static __global__ void __launch_bounds__(kBucketThreads)
bucket_scan_kernel(const __grid_constant__ BucketParams P) {
  __shared__ int32_t s_part[kBucketThreads];
  __shared__ int32_t s_carry;
  const int s = blockIdx.x;
  const BucketSeg& sg = P.seg[s];
  const int ntiles = (sg.n + kBucketTile - 1) / kBucketTile;
  const int total = ntiles * P.nbins;
  int32_t* cnt = P.counts + (size_t)sg.tile_begin * P.nbins;
  if (sg.out_sizes != nullptr && ntiles == 0) {
    for (int b = threadIdx.x; b < P.nbins; b += kBucketThreads) sg.out_sizes[b] = 0;
    return;
  }
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  // chunks of kBucketThreads * per elements; each thread scans `per` contiguous
  const int per = 16;
  const int chunk = kBucketThreads * per;
  for (int c0 = 0; c0 < total; c0 += chunk) {
    int32_t v[per];
    int32_t sum = 0;
    const int b0 = c0 + threadIdx.x * per;
#pragma unroll
    for (int k = 0; k < per; ++k) {
      v[k] = (b0 + k < total) ? cnt[b0 + k] : 0;
      sum += v[k];
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 256 partials
    for (int off = 1; off < kBucketThreads; off <<= 1) {
      int32_t add = (threadIdx.x >= off) ? s_part[threadIdx.x - off] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    int32_t run = s_carry + s_part[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < per; ++k) {
      if (b0 + k < total) cnt[b0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == kBucketThreads - 1) s_carry = run;
    __syncthreads();
  }
  if (sg.out_sizes != nullptr) {
    // size[b] = start(b+1) - start(b); start(b) = scanned cnt[b*ntiles]
    for (int b = threadIdx.x; b < P.nbins; b += kBucketThreads) {
      const int32_t lo = cnt[(size_t)b * ntiles];
      const int32_t hi = (b + 1 < P.nbins) ? cnt[(size_t)(b + 1) * ntiles] : sg.n;
      sg.out_sizes[b] = hi - lo;
    }
  }
}

// ---- scatter --------------------------------------------------------------
// dynamic smem layout (see bucket_scatter_smem_bytes)
template <typename Tr>
__global__ void __launch_bounds__(kBucketThreads)
bucket_scatter_kernel(const __grid_constant__ BucketParams P) {
  using In = typename Tr::In;
  using Out = typename Tr::Out;
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int nb = P.nbins;
  Out* s_keys = reinterpret_cast<Out*>(s_raw);                                  // [tile]
  int32_t* s_vals = reinterpret_cast<int32_t*>(s_raw + sizeof(Out) * kBucketTile);  // [tile]
  uint16_t* s_bin = reinterpret_cast<uint16_t*>(s_vals + kBucketTile);          // [tile]
  int32_t* s_wcnt = reinterpret_cast<int32_t*>(s_bin + kBucketTile);            // [warps][nb]
  int32_t* s_gbase = s_wcnt + kBucketWarps * nb;                                // [nb]
  int32_t* s_lstart = s_gbase + nb;                                             // [nb+1]
  __shared__ int32_t s_scan[kBucketThreads];

  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;

  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const int s = find_seg(P, tile);
    const BucketSeg& sg = P.seg[s];
    const int t = tile - sg.tile_begin;
    const int ntiles = (sg.n + kBucketTile - 1) / kBucketTile;
    const In* in = reinterpret_cast<const In*>(sg.in_keys);
    const int32_t* cnt = P.counts + (size_t)sg.tile_begin * nb;

    for (int b = threadIdx.x; b < kBucketWarps * nb; b += kBucketThreads) s_wcnt[b] = 0;
    for (int b = threadIdx.x; b < nb; b += kBucketThreads)
      s_gbase[b] = cnt[(size_t)b * ntiles + t];
    __syncthreads();

    // 1. per-warp stable ranks over the warp's contiguous slice
    const int wbase = t * kBucketTile + warp * (kBucketItems * 32);
    In key[kBucketItems];
    int32_t rank[kBucketItems];
    int bin[kBucketItems];
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = wbase + j * 32 + lane;
      key[j] = (i < sg.n) ? in[i] : In(0);
    }
    int32_t* wc = s_wcnt + warp * nb;
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = wbase + j * 32 + lane;
      const bool valid = i < sg.n;
      bin[j] = valid ? Tr::bin(key[j], P, sg) : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, bin[j]);
      const int leader = __ffs(peers) - 1;
      int32_t base = 0;
      if (valid && lane == (unsigned)leader) {
        base = wc[bin[j]];
        wc[bin[j]] = base + __popc(peers);
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      rank[j] = base + __popc(peers & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();

    // 2. per bin: exclusive prefix over warps (in place) and tile totals
    int32_t mytot[(kMaxBins + kBucketThreads - 1) / kBucketThreads];
    int32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < (kMaxBins + kBucketThreads - 1) / kBucketThreads; ++k) {
      const int b = threadIdx.x * ((kMaxBins + kBucketThreads - 1) / kBucketThreads) + k;
      mytot[k] = 0;
      if (b < nb) {
        int32_t run = 0;
#pragma unroll
        for (int w = 0; w < kBucketWarps; ++w) {
          const int32_t c = s_wcnt[w * nb + b];
          s_wcnt[w * nb + b] = run;
          run += c;
        }
        mytot[k] = run;
      }
      tsum += mytot[k];
    }
    // block exclusive scan of per-thread sums -> local bin starts
    s_scan[threadIdx.x] = tsum;
    __syncthreads();
    for (int off = 1; off < kBucketThreads; off <<= 1) {
      int32_t add = (threadIdx.x >= off) ? s_scan[threadIdx.x - off] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += add;
      __syncthreads();
    }
    {
      int32_t run = s_scan[threadIdx.x] - tsum;
#pragma unroll
      for (int k = 0; k < (kMaxBins + kBucketThreads - 1) / kBucketThreads; ++k) {
        const int b = threadIdx.x * ((kMaxBins + kBucketThreads - 1) / kBucketThreads) + k;
        if (b < nb) s_lstart[b] = run;
        run += mytot[k];
      }
    }
    __syncthreads();

    // 3. stage the tile bin-sorted in smem; write the inverse by input index
    const int tile_n = min(kBucketTile, sg.n - t * kBucketTile);
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = wbase + j * 32 + lane;
      if (i < sg.n) {
        const int b = bin[j];
        const int32_t in_bin = s_wcnt[warp * nb + b] + rank[j];
        const int32_t lpos = s_lstart[b] + in_bin;
        s_keys[lpos] = Tr::conv(key[j], P, sg);
        s_bin[lpos] = (uint16_t)b;
        if (sg.out_vals != nullptr)
          s_vals[lpos] = (sg.in_vals != nullptr) ? sg.in_vals[i] : i;
        if (sg.out_inv != nullptr) sg.out_inv[i] = s_gbase[b] + in_bin;
      }
    }
    __syncthreads();

    // 4. coalesced write-out: consecutive smem slots of one bin are consecutive
    //    in global memory
    Out* okeys = reinterpret_cast<Out*>(sg.out_keys);
    for (int k = threadIdx.x; k < tile_n; k += kBucketThreads) {
      const int b = s_bin[k];
      const int32_t g = s_gbase[b] + (k - s_lstart[b]);
      if (okeys != nullptr) okeys[g] = s_keys[k];
      if (sg.out_vals != nullptr) sg.out_vals[g] = s_vals[k];
    }
    __syncthreads();
  }
}

template <typename Out>
static inline size_t bucket_scatter_smem_bytes(int nbins) {
  return sizeof(Out) * kBucketTile + sizeof(int32_t) * kBucketTile +
         sizeof(uint16_t) * kBucketTile +
         sizeof(int32_t) * ((size_t)kBucketWarps * nbins + 2 * (size_t)nbins + 1);
}

static inline int bucket_tiles(int64_t n) {
  return (int)((n + kBucketTile - 1) / kBucketTile);
}

// Launch count+scan+scatter for a prepared BucketParams (total_tiles may be 0).
template <typename Tr>
int bucket_pass_launch(const BucketParams& P, cudaStream_t stream, int kid_base = HB_K_PART_COUNT) {
  using Out = typename Tr::Out;
  const int max_grid = device_sm_count() * 8;
  if (P.total_tiles > 0) {
    const int grid = P.total_tiles < max_grid ? P.total_tiles : max_grid;
    KernelScope ks(kid_base, stream);
    bucket_count_kernel<Tr><<<grid, kBucketThreads, sizeof(int32_t) * P.nbins, stream>>>(P);
    HB_CUDA_OK(cudaGetLastError());
  }
  {
    KernelScope ks(kid_base + 1, stream);
    bucket_scan_kernel<<<P.nsegs, kBucketThreads, 0, stream>>>(P);
  }
  HB_CUDA_OK(cudaGetLastError());
  if (P.total_tiles > 0) {
    const int grid = P.total_tiles < max_grid ? P.total_tiles : max_grid;
    const size_t smem = bucket_scatter_smem_bytes<Out>(P.nbins);
    if (smem > 48 * 1024)
      HB_CUDA_OK(cudaFuncSetAttribute(bucket_scatter_kernel<Tr>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    KernelScope ks(kid_base + 2, stream);
    bucket_scatter_kernel<Tr><<<grid, kBucketThreads, smem, stream>>>(P);
    HB_CUDA_OK(cudaGetLastError());
  }
  return HB_OK;
}

}  // namespace hb
