// transfer.cu -- fused multi-tensor host -> device staging of the step's sparse features.
// Replaces HbH2DTransferN (ops/transfer/transfer_functors.cu.cc:38-238; fed by the
// prefetch pipeline, data/prefetch/prefetch.cc:41-481): N host tensors (one per sparse
// feature: ids, offsets) reach N device tensors with ONE kernel launch that reads the
// PINNED host buffers over PCIe directly (zero copy: the SMs issue 128-bit loads on the
// mapped host pointers), instead of N cudaMemcpyAsync calls of a few hundred KB each.
// Pageable inputs cannot be read by a kernel: they go through cudaMemcpyAsync one by one,
// as in the reference (its large unpinned branch, :96-110).
#include <vector>

#include "common.cuh"

namespace hb {

constexpr int kMaxTransfer = 256;
constexpr uint64_t kTransferChunk = 16384;   // bytes per work unit

struct TransferSeg {
  const unsigned char* src;   // device-visible alias of the pinned host buffer
  unsigned char* dst;
  uint64_t bytes;
  uint64_t chunk_begin;
};
struct TransferParams {
  TransferSeg seg[kMaxTransfer];
  uint64_t total_chunks;
  int32_t n;
};

__global__ void __launch_bounds__(256) h2d_transfer_kernel(const __grid_constant__ TransferParams P) {
  for (uint64_t c = blockIdx.x; c < P.total_chunks; c += gridDim.x) {
    int lo = 0, hi = P.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (P.seg[mid].chunk_begin <= c) lo = mid; else hi = mid - 1;
    }
    const TransferSeg& S = P.seg[lo];
    const uint64_t o = (c - S.chunk_begin) * kTransferChunk;
    const uint64_t len = (S.bytes - o < kTransferChunk) ? S.bytes - o : kTransferChunk;
    const unsigned char* src = S.src + o;
    unsigned char* dst = S.dst + o;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
      const uint64_t nv = len >> 4;
      const int4* s4 = reinterpret_cast<const int4*>(src);
      int4* d4 = reinterpret_cast<int4*>(dst);
      // four independent 128-bit host reads per thread in flight: PCIe round trips are long
      uint64_t i = threadIdx.x;
      for (; i + 3 * 256 < nv; i += 4 * 256) {
        const int4 a = s4[i], b = s4[i + 256], c4 = s4[i + 512], d = s4[i + 768];
        d4[i] = a; d4[i + 256] = b; d4[i + 512] = c4; d4[i + 768] = d;
      }
      for (; i < nv; i += 256) d4[i] = s4[i];
      for (uint64_t j = (nv << 4) + threadIdx.x; j < len; j += 256) dst[j] = src[j];
    } else {
      for (uint64_t j = threadIdx.x; j < len; j += 256) dst[j] = src[j];
    }
  }
}

}  // namespace hb

extern "C" int hbH2DTransferN(int n, const void* const* h_inputs, void* const* d_outputs, const int64_t* bytes,
                              hbStream stream_) {
  using namespace hb;
  cudaStream_t stream = (cudaStream_t)stream_;
  HB_REQUIRE(n >= 0 && (n == 0 || (h_inputs && d_outputs && bytes)), "hbH2DTransferN: bad arguments");
  TransferParams P;
  P.n = 0;
  P.total_chunks = 0;
  auto flush = [&]() -> int {
    if (P.n > 0) {
      const int maxg = device_sm_count() * 4;
      const int grid = P.total_chunks < (uint64_t)maxg ? (int)P.total_chunks : maxg;
      KernelScope ks(HB_K_H2D_STAGE, stream);
      h2d_transfer_kernel<<<grid, 256, 0, stream>>>(P);
      HB_CUDA_OK(cudaGetLastError());
    }
    P.n = 0;
    P.total_chunks = 0;
    return HB_OK;
  };
  for (int k = 0; k < n; ++k) {
    HB_REQUIRE(bytes[k] >= 0, "hbH2DTransferN: negative size for tensor %d", k);
    if (bytes[k] == 0) continue;
    HB_REQUIRE(h_inputs[k] && d_outputs[k], "hbH2DTransferN: null pointer for tensor %d", k);
    // pinned (page-locked) host memory has a device alias the SMs can read; pageable
    // memory is reported as unregistered (interior pointers of an allocation are fine)
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, h_inputs[k]) == cudaSuccess &&
                        attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
    void* alias = pinned ? attr.devicePointer : nullptr;
    if (!pinned) {
      (void)cudaGetLastError();   // a failed query must not leak into the next CUDA call's status
      HB_CUDA_OK(cudaMemcpyAsync(d_outputs[k], h_inputs[k], (size_t)bytes[k], cudaMemcpyHostToDevice, stream));
      continue;
    }
    TransferSeg& S = P.seg[P.n++];
    S.src = reinterpret_cast<const unsigned char*>(alias);
    S.dst = reinterpret_cast<unsigned char*>(d_outputs[k]);
    S.bytes = (uint64_t)bytes[k];
    S.chunk_begin = P.total_chunks;
    P.total_chunks += (S.bytes + kTransferChunk - 1) / kTransferChunk;
    if (P.n == kMaxTransfer) {
      const int rc = flush();
      if (rc != HB_OK) return rc;
    }
  }
  return flush();
}
