// cast.cu -- fp32 <-> fp16 wire casts for N tensors in one launch.
// Replaces functor::Cast / CastN (tensorflow/common/cast.cu.cc:37-495).
#include <cuda_fp16.h>

#include "common.cuh"

namespace hb {

constexpr int kMaxCast = 128;
struct CastParams {
  const void* in[kMaxCast];
  void* out[kMaxCast];
  int64_t count[kMaxCast];
  int32_t cta_begin[kMaxCast];
  int32_t n;
  int32_t to_half;
};
constexpr int kCastPerCta = 256 * 8;

__global__ void __launch_bounds__(256) cast_n_kernel(const __grid_constant__ CastParams P) {
  int lo = 0, hi = P.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.cta_begin[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const int64_t base = (int64_t)(blockIdx.x - P.cta_begin[lo]) * kCastPerCta;
  const int64_t n = P.count[lo];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t i = base + j * 256 + threadIdx.x;
    if (i >= n) break;
    if (P.to_half)
      reinterpret_cast<__half*>(P.out[lo])[i] = __float2half(reinterpret_cast<const float*>(P.in[lo])[i]);
    else
      reinterpret_cast<float*>(P.out[lo])[i] = __half2float(reinterpret_cast<const __half*>(P.in[lo])[i]);
  }
}

}  // namespace hb

extern "C" int hbCastN(int n, const void* const* d_inputs, void* const* d_outputs,
                       const int64_t* counts, int from_dtype, int to_dtype, hbStream stream) {
  using namespace hb;
  HB_REQUIRE(n >= 1 && d_inputs && d_outputs && counts, "hbCastN: bad argument");
  HB_REQUIRE((from_dtype == HB_F32 && to_dtype == HB_F16) || (from_dtype == HB_F16 && to_dtype == HB_F32),
             "hbCastN: only float<->half supported (got %d -> %d)", from_dtype, to_dtype);
  for (int c0 = 0; c0 < n; c0 += kMaxCast) {
    CastParams P;
    P.n = 0;
    P.to_half = to_dtype == HB_F16;
    int ctas = 0;
    for (int k = c0; k < n && k < c0 + kMaxCast; ++k) {
      HB_REQUIRE(counts[k] >= 0, "hbCastN: negative count");
      if (counts[k] == 0) continue;
      P.in[P.n] = d_inputs[k];
      P.out[P.n] = d_outputs[k];
      P.count[P.n] = counts[k];
      P.cta_begin[P.n] = ctas;
      ctas += (int)((counts[k] + kCastPerCta - 1) / kCastPerCta);
      P.n++;
    }
    if (ctas > 0) {
      KernelScope ks(HB_K_CAST, (cudaStream_t)stream);
      cast_n_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(P);
      HB_CUDA_OK(cudaGetLastError());
    }
  }
  return HB_OK;
}
