// bucket.cuh -- stable multi-segment bucket scatter for sm_100a.
//
// K1 partition (HbPartitionByModulo[N] / dual modulo): bin = shard(id), emits the
// permuted ids, per-bin sizes and the inverse permutation.  (Round 1 also ran the
// backward's LSD radix sort through this mechanism, one launch per digit; that job
// moved to the one-launch cluster kernel of cluster_sort.cuh.)
// Launches (ALL segments = features at once):
//   memset  : zero histograms, tile status words and tickets (one node)
//   hist    : global per-(segment, pass, bin) histograms of EVERY digit position in
//             one read of the input (onesweep-style up-front histogram)
//   pass    : one kernel per digit position.  A CTA takes a tile by ticket (so a
//             tile only ever waits for tiles that started before it), ranks its
//             items stably (warp match_any + per-warp running counters), publishes
//             its per-bin counts as self-flagged 32-bit status words, sums the
//             counts of ALL preceding tiles of its segment with independent,
//             unrolled loads (no serial look-back chain), adds the bin bases from
//             the global histogram and scatters.  The tile is staged bin-sorted in
//             shared memory so global writes are contiguous runs per bin
//             (coalesced); the inverse permutation is written by input index.
// Stability: warp w owns a contiguous slice of the tile and walks it in rounds
// of 32 consecutive items; rank = (# items of the bin in earlier tiles) +
// (# in earlier warps) + (running count in this warp) + (# lower lanes with the
// same bin this round).  Output order inside a bin == input order, i.e.
// bit-identical to the reference CPU counting sort
// (partition_by_modulo_functors.cc:48-69).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hb {

constexpr int kBucketThreads = 256;
constexpr int kBucketWarps = kBucketThreads / 32;
constexpr int kBucketItems = 16;
constexpr int kBucketTile = kBucketThreads * kBucketItems;  // 4096
constexpr int kMaxSegs = 128;
constexpr int kMaxBins = 512;
constexpr int kMaxPasses = 4;
constexpr uint32_t kReady = 0x80000000u;

struct BucketSeg {
  const void* in_keys;
  const int32_t* in_vals;  // nullptr: value = input index
  void* out_keys;          // may be nullptr (keys not needed downstream)
  int32_t* out_vals;       // may be nullptr
  int32_t* out_inv;        // may be nullptr: out_inv[i] = output position of i
  int32_t* out_sizes;      // may be nullptr: [nbins]
  const int32_t* n_dev;    // may be nullptr: actual length on the device (<= n)
  int32_t n;               // length, or its static upper bound when n_dev is set
  int32_t tile_begin;      // global tile index of this segment's first tile
  int32_t shift;           // radix digit shift of THIS pass
  uint32_t key_limit;      // radix-first: keys >= key_limit are invalid (0: off)
  int32_t hist_slot;       // histogram row of this segment: hist[(hist_slot*npass + pass)*nbins]
  int32_t passes;          // number of digit positions this segment needs
  int32_t lbits;           // composite keys: bits of the local row (owner sits above them)
};

struct BucketParams {
  BucketSeg seg[kMaxSegs];
  uint32_t* hist;      // [hist_slots][npass][nbins]
  uint32_t* status;    // this pass: [total_tiles][nbins]
  uint32_t* ticket;    // this pass: dynamic tile counter
  int32_t nsegs;
  int32_t nbins;
  int32_t total_tiles;
  int32_t npass;       // histogram rows per segment
  int32_t pass;        // digit position of this launch
  int32_t digit_bits;  // radix: log2(nbins)
  int32_t p;           // num_partitions (modulo modes)
  int32_t m;           // modulus (dual modulo)
  int32_t pow2_mask;   // p-1 if p is a power of two else -1
  int32_t div_shift;   // log2(div) if power of two else -1
  int64_t div;         // radix-first: key = id / div
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
  static constexpr bool kRadix = false;
  static __device__ __forceinline__ Out conv(In v, const BucketParams&, const BucketSeg&) { return v; }
  static __device__ __forceinline__ int bin(In v, const BucketParams& P, const BucketSeg&, int) {
    if (P.pow2_mask >= 0) return (int)(v & (T)P.pow2_mask);
    return floor_mod<T>(v, P.p);
  }
};

template <typename T, int STAGE>
struct DualModuloTraits {  // partition_by_dual_modulo_functors.cc:37-49,:66-71
  using In = T;
  using Out = T;
  static constexpr bool kRadix = false;
  static __device__ __forceinline__ Out conv(In v, const BucketParams&, const BucketSeg&) { return v; }
  static __device__ __forceinline__ int bin(In v, const BucketParams& P, const BucketSeg&, int) {
    const int pre = floor_mod<T>(v, P.p * P.m);
    return STAGE == 1 ? pre % P.p : pre / P.m;
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

// effective length of a segment (device-side count clamps the static bound)
__device__ __forceinline__ int seg_len(const BucketSeg& sg) {
  if (sg.n_dev == nullptr) return sg.n;
  const int d = *sg.n_dev;
  return d < 0 ? 0 : (d < sg.n ? d : sg.n);
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Block-wide exclusive scan of one int per thread (256 threads): warp shuffles +
// one smem hop; returns the exclusive prefix, *total gets the block sum.
__device__ __forceinline__ int32_t block_excl_scan(int32_t x, int32_t* s_warp /*[8]*/, int32_t* total) {
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
  for (int w = 0; w < kBucketWarps; ++w) {
    const int32_t t = s_warp[w];
    if (w < warp) wbase += t;
    tot += t;
  }
  __syncthreads();
  if (total != nullptr) *total = tot;
  return wbase + incl - x;
}

// ---- hist -------------------------------------------------------------------
// Histograms of every digit position of every segment in one read of the input.
// digit of an already converted key (radix traits) / bin of a raw id (modulo traits)
template <typename Tr>
__device__ __forceinline__ int bin_of(typename Tr::In raw, typename Tr::Out key, const BucketParams& P,
                                      const BucketSeg& sg, int shift) {
  if constexpr (Tr::kRadix) return (int)((key >> shift) & (uint32_t)(P.nbins - 1));
  else return Tr::bin(raw, P, sg, shift);
}

template <typename Tr>
__global__ void __launch_bounds__(kBucketThreads)
bucket_hist_kernel(const __grid_constant__ BucketParams P) {
  extern __shared__ uint32_t s_hist[];  // [npass][nbins]
  using In = typename Tr::In;
  using Out = typename Tr::Out;
  const int nb = P.nbins;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const int s = find_seg(P, tile);
    const BucketSeg& sg = P.seg[s];
    const int t = tile - sg.tile_begin;
    const int np = Tr::kRadix ? sg.passes : 1;
    const int n = seg_len(sg);
    const int base = t * kBucketTile;
    if (base >= n) continue;  // uniform per CTA
    for (int b = threadIdx.x; b < np * nb; b += kBucketThreads) s_hist[b] = 0;
    const In* in = reinterpret_cast<const In*>(sg.in_keys);
    In v[kBucketItems];
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {  // all loads in flight together
      const int i = base + j * kBucketThreads + threadIdx.x;
      v[j] = (i < n) ? in[i] : In(0);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = base + j * kBucketThreads + threadIdx.x;
      if (i < n) {
        const Out key = Tr::conv(v[j], P, sg);  // converted once, every digit taken from it
        for (int p = 0; p < np; ++p)
          atomicAdd(&s_hist[p * nb + bin_of<Tr>(v[j], key, P, sg, p * P.digit_bits)], 1u);
      }
    }
    __syncthreads();
    uint32_t* gh = P.hist + (size_t)sg.hist_slot * P.npass * nb;
    for (int b = threadIdx.x; b < np * nb; b += kBucketThreads)
      if (s_hist[b]) atomicAdd(&gh[b], s_hist[b]);
    __syncthreads();
  }
  // sizes of empty segments (no tile will ever write them)
  if (blockIdx.x == 0) {
    for (int s = 0; s < P.nsegs; ++s) {
      const BucketSeg& sg = P.seg[s];
      if (seg_len(sg) == 0 && sg.out_sizes != nullptr)
        for (int b = threadIdx.x; b < nb; b += kBucketThreads) sg.out_sizes[b] = 0;
    }
  }
}

// ---- pass -------------------------------------------------------------------
template <typename Tr>
__global__ void __launch_bounds__(kBucketThreads, 3)
bucket_pass_kernel(const __grid_constant__ BucketParams P) {
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
  __shared__ int s_tile;
  constexpr int kBinsPerThread = (kMaxBins + kBucketThreads - 1) / kBucketThreads;  // 2

  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;

  // ticket: tiles are started in global order, so a tile only waits for tiles
  // that already run (or finished) -- no deadlock whatever the grid size
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(P.ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= P.total_tiles) return;
  const int s = find_seg(P, tile);
  const BucketSeg& sg = P.seg[s];
  const int t = tile - sg.tile_begin;
  const In* in = reinterpret_cast<const In*>(sg.in_keys);
  const int n = seg_len(sg);
  if (t * kBucketTile >= n) return;  // tile beyond the device-side length (uniform)

  for (int b = threadIdx.x; b < kBucketWarps * nb; b += kBucketThreads) s_wcnt[b] = 0;
  __syncthreads();

  // 1. per-warp stable ranks over the warp's contiguous slice.  Everything the tile
  //    needs from global memory is requested up front (keys, carried values, the
  //    histogram row): one exposed round trip instead of three.
  const int wbase = t * kBucketTile + warp * (kBucketItems * 32);
  Out key[kBucketItems];
  int32_t val[kBucketItems];
  uint32_t rb[kBucketItems];  // stable rank inside the warp slice (low 16 bits) | bin (high 16; kNoBin: none)
  constexpr uint32_t kNoBin = 0xFFFFu;
  const bool carry_vals = sg.out_vals != nullptr && sg.in_vals != nullptr;
  {
    In raw[kBucketItems];
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = wbase + j * 32 + lane;
      raw[j] = (i < n) ? in[i] : In(0);
    }
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = wbase + j * 32 + lane;
      val[j] = (carry_vals && i < n) ? sg.in_vals[i] : i;
    }
#pragma unroll
    for (int j = 0; j < kBucketItems; ++j) {
      const int i = wbase + j * 32 + lane;
      key[j] = Tr::conv(raw[j], P, sg);
      rb[j] = ((i < n) ? (uint32_t)bin_of<Tr>(raw[j], key[j], P, sg, sg.shift) : kNoBin) << 16;
    }
  }
  const uint32_t* gh = P.hist + ((size_t)sg.hist_slot * P.npass + P.pass) * nb;
  int32_t hpre[kBinsPerThread];
#pragma unroll
  for (int k = 0; k < kBinsPerThread; ++k) {
    const int b = threadIdx.x * kBinsPerThread + k;
    hpre[k] = (b < nb) ? (int32_t)gh[b] : 0;
  }
  int32_t* wc = s_wcnt + warp * nb;
#pragma unroll
  for (int j = 0; j < kBucketItems; ++j) {
    const int i = wbase + j * 32 + lane;
    const bool valid = i < n;
    const uint32_t b = rb[j] >> 16;
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    int32_t base = 0;
    if (valid && lane == (unsigned)leader) {
      base = wc[b];
      wc[b] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rb[j] |= (uint32_t)(base + __popc(peers & lanemask_lt()));
    __syncwarp();
  }
  __syncthreads();

  // 2. per bin: exclusive prefix over warps (in place), tile totals, publish
  uint32_t* my_status = P.status + (size_t)tile * nb;
  int32_t mytot[kBinsPerThread];
#pragma unroll
  for (int k = 0; k < kBinsPerThread; ++k) {
    const int b = k * kBucketThreads + threadIdx.x;  // coalesced status access
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
      *reinterpret_cast<volatile uint32_t*>(&my_status[b]) = (uint32_t)run | kReady;
    }
  }

  // 3. bin bases from the global histogram (exclusive scan over bins) ...
  {
    // thread owns bins {2*tid, 2*tid+1} for the scan (contiguous), totals from hist
    int32_t h[kBinsPerThread];
    int32_t hsum = 0;
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
      h[k] = hpre[k];
      hsum += h[k];
    }
    int32_t run = block_excl_scan(hsum, s_scan, nullptr);
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
      const int b = threadIdx.x * kBinsPerThread + k;
      if (b < nb) s_gbase[b] = run;
      run += h[k];
    }
    if (t == 0 && sg.out_sizes != nullptr) {
#pragma unroll
      for (int k = 0; k < kBinsPerThread; ++k) {
        const int b = threadIdx.x * kBinsPerThread + k;
        if (b < nb) sg.out_sizes[b] = h[k];
      }
    }
  }
  __syncthreads();
  // ... plus the counts of all preceding tiles of this segment: independent loads
  // (both bins of the thread, 8 tiles each = 16 in flight), spinning only on words
  // not yet published
  {
    const uint32_t* st = P.status + (size_t)sg.tile_begin * nb;
    int32_t pre[kBinsPerThread];
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) pre[k] = 0;
    constexpr int kLook = 4;
    for (int tp = 0; tp < t; tp += kLook) {
      uint32_t v[kBinsPerThread][kLook];
#pragma unroll
      for (int k = 0; k < kBinsPerThread; ++k) {
        const int b = k * kBucketThreads + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kLook; ++u)
          v[k][u] = (b < nb && tp + u < t) ? ld_volatile_u32(st + (size_t)(tp + u) * nb + b) : kReady;
      }
#pragma unroll
      for (int k = 0; k < kBinsPerThread; ++k) {
        const int b = k * kBucketThreads + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kLook; ++u) {
          while (!(v[k][u] & kReady)) v[k][u] = ld_volatile_u32(st + (size_t)(tp + u) * nb + b);
          pre[k] += (int32_t)(v[k][u] & ~kReady);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
      const int b = k * kBucketThreads + threadIdx.x;
      if (b < nb) s_gbase[b] += pre[k];
    }
  }
  // local bin starts inside the tile (exclusive scan of tile totals, bin order)
  {
    // mytot[k] belongs to bin k*256+tid; transpose through smem to contiguous order
    int32_t* s_tot = s_lstart;  // scratch [nb], overwritten with the starts below
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
      const int b = k * kBucketThreads + threadIdx.x;
      if (b < nb) s_tot[b] = mytot[k];
    }
    __syncthreads();
    int32_t tt[kBinsPerThread];
    int32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
      const int b = threadIdx.x * kBinsPerThread + k;
      tt[k] = (b < nb) ? s_tot[b] : 0;
      tsum += tt[k];
    }
    int32_t run = block_excl_scan(tsum, s_scan, nullptr);
#pragma unroll
    for (int k = 0; k < kBinsPerThread; ++k) {
      const int b = threadIdx.x * kBinsPerThread + k;
      if (b < nb) s_lstart[b] = run;
      run += tt[k];
    }
  }
  __syncthreads();

  // 4. stage the tile bin-sorted in smem; write the inverse by input index
  const int tile_n = min(kBucketTile, n - t * kBucketTile);
#pragma unroll
  for (int j = 0; j < kBucketItems; ++j) {
    const int i = wbase + j * 32 + lane;
    if (i < n) {
      const int b = (int)(rb[j] >> 16);
      const int32_t in_bin = s_wcnt[warp * nb + b] + (int32_t)(rb[j] & 0xFFFFu);
      const int32_t lpos = s_lstart[b] + in_bin;
      s_keys[lpos] = key[j];
      s_bin[lpos] = (uint16_t)b;
      if (sg.out_vals != nullptr) s_vals[lpos] = val[j];
      if (sg.out_inv != nullptr) sg.out_inv[i] = s_gbase[b] + in_bin;
    }
  }
  __syncthreads();

  // 5. coalesced write-out: consecutive smem slots of one bin are consecutive
  //    in global memory
  Out* okeys = reinterpret_cast<Out*>(sg.out_keys);
  for (int k = threadIdx.x; k < tile_n; k += kBucketThreads) {
    const int b = s_bin[k];
    const int32_t g = s_gbase[b] + (k - s_lstart[b]);
    if (okeys != nullptr) okeys[g] = s_keys[k];
    if (sg.out_vals != nullptr) sg.out_vals[g] = s_vals[k];
  }
}

template <typename Out>
static inline size_t bucket_pass_smem_bytes(int nbins) {
  return sizeof(Out) * kBucketTile + sizeof(int32_t) * kBucketTile +
         sizeof(uint16_t) * kBucketTile +
         sizeof(int32_t) * ((size_t)kBucketWarps * nbins + 2 * (size_t)nbins + 1);
}

static inline int bucket_tiles(int64_t n) {
  return (int)((n + kBucketTile - 1) / kBucketTile);
}

// Scratch (uint32 words) a group of segments needs for `npass` digit positions:
// histograms + per-pass tile status + per-pass tickets.  Must be zeroed (one
// cudaMemsetAsync) before the hist kernel.
static inline size_t bucket_scratch_words(int nsegs, size_t total_tiles, int nbins, int npass) {
  return (size_t)nsegs * npass * nbins + (size_t)npass * total_tiles * nbins + 64;
}

struct BucketScratch {
  uint32_t* hist;
  uint32_t* status[kMaxPasses];
  uint32_t* ticket[kMaxPasses];
};

static inline BucketScratch bucket_scratch_carve(uint32_t* base, int nsegs, size_t total_tiles,
                                                 int nbins, int npass) {
  BucketScratch s;
  s.hist = base;
  uint32_t* p = base + (size_t)nsegs * npass * nbins;
  for (int i = 0; i < kMaxPasses; ++i)
    s.status[i] = (i < npass) ? p + (size_t)i * total_tiles * nbins : nullptr;
  uint32_t* tk = p + (size_t)npass * total_tiles * nbins;
  for (int i = 0; i < kMaxPasses; ++i) s.ticket[i] = tk + i * 16;
  return s;
}

template <typename Tr>
int bucket_hist_launch(const BucketParams& P, cudaStream_t stream, int kid) {
  const int max_grid = device_sm_count() * 8;
  const int grid = P.total_tiles < 1 ? 1 : (P.total_tiles < max_grid ? P.total_tiles : max_grid);
  KernelScope ks(kid, stream);
  bucket_hist_kernel<Tr><<<grid, kBucketThreads, sizeof(uint32_t) * P.npass * P.nbins, stream>>>(P);
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

template <typename Tr>
int bucket_pass_launch(const BucketParams& P, cudaStream_t stream, int kid) {
  using Out = typename Tr::Out;
  if (P.total_tiles == 0) return HB_OK;
  const size_t smem = bucket_pass_smem_bytes<Out>(P.nbins);
  if (smem > 48 * 1024)
    HB_CUDA_OK(cudaFuncSetAttribute(bucket_pass_kernel<Tr>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KernelScope ks(kid, stream);
  bucket_pass_kernel<Tr><<<P.total_tiles, kBucketThreads, smem, stream>>>(P);
  HB_CUDA_OK(cudaGetLastError());
  return HB_OK;
}

}  // namespace hb
