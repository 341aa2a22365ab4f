// gather_peak.cu -- calibration microbenchmark (not part of the library): what does a
// B200 sustain on RANDOM 128/256/512-byte row reads, and on the read-modify-write of
// such rows, at maximum memory-level parallelism?  The streaming-copy peak of
// MEASURED_PEAKS.json is the denominator of every roofline fraction this repo reports;
// this number says how much of the gap is the access pattern rather than the kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_peak tools/gather_peak.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// G lanes per row (G*16 bytes), R rows in flight per group
template <int R, bool RMW>
__global__ void __launch_bounds__(256) gather_kernel(float* table, const int64_t* rows, int n, int log2g,
                                                     float* sink) {
  const int G = 1 << log2g;
  const int groups = (gridDim.x * 256) >> log2g;
  const int g = (blockIdx.x * 256 + threadIdx.x) >> log2g;
  const int l = threadIdx.x & (G - 1);
  float4 acc = make_float4(0, 0, 0, 0);
  for (int i0 = g * R; i0 < n; i0 += groups * R) {
    int64_t r[R];
#pragma unroll
    for (int u = 0; u < R; ++u) r[u] = (i0 + u < n) ? rows[i0 + u] : -1;
    float4 v[R];
#pragma unroll
    for (int u = 0; u < R; ++u)
      v[u] = r[u] >= 0 ? *reinterpret_cast<const float4*>(table + r[u] * (G * 4) + l * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (RMW) {
        if (r[u] >= 0) {
          v[u].x += 1.0f; v[u].y += 1.0f; v[u].z += 1.0f; v[u].w += 1.0f;
          *reinterpret_cast<float4*>(table + r[u] * (G * 4) + l * 4) = v[u];
        }
      } else {
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
  }
  if (!RMW && acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <int R, bool RMW>
static void run(const char* name, float* table, const int64_t* d_rows, int n, int log2g, float* sink, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 2; ++w) gather_kernel<R, RMW><<<grid, 256>>>(table, d_rows, n, log2g, sink);
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int w = 0; w < reps; ++w) gather_kernel<R, RMW><<<grid, 256>>>(table, d_rows, n, log2g, sink);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double bytes = (double)n * (16 << log2g) * (RMW ? 2 : 1);
  printf("%-28s row %4d B  R=%2d ctas/sm=%d : %8.1f us  %7.1f GB/s\n", name, 16 << log2g, R, ctas_per_sm, ms * 1e3,
         bytes / ms / 1e6);
}

int main() {
  const size_t table_bytes = (size_t)16 << 30;   // 16 GB: far beyond the 126 MB L2
  float* table;
  CK(cudaMalloc(&table, table_bytes));
  CK(cudaMemset(table, 0, table_bytes));
  float* sink;
  CK(cudaMalloc(&sink, 256));
  const int n = 1 << 22;  // 4 M random rows per launch
  for (int log2g = 3; log2g <= 5; ++log2g) {
    const size_t nrows = table_bytes / (16 << log2g);
    std::vector<int64_t> h(n);
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int64_t)(s % nrows); }
    int64_t* d_rows;
    CK(cudaMalloc(&d_rows, n * 8));
    CK(cudaMemcpy(d_rows, h.data(), n * 8, cudaMemcpyHostToDevice));
    run<4, false>("random row read", table, d_rows, n, log2g, sink, 8);
    run<8, false>("random row read", table, d_rows, n, log2g, sink, 4);
    run<8, false>("random row read", table, d_rows, n, log2g, sink, 8);
    run<16, false>("random row read", table, d_rows, n, log2g, sink, 4);
    run<4, true>("random row read-modify-write", table, d_rows, n, log2g, sink, 8);
    run<8, true>("random row read-modify-write", table, d_rows, n, log2g, sink, 8);
    CK(cudaFree(d_rows));
  }
  // streaming copy for reference (1 GiB each way)
  {
    const size_t nvec = ((size_t)1 << 30) / 16;
    float4* a = reinterpret_cast<float4*>(table);
    float4* b = a + nvec * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    copy_kernel<<<148 * 8, 256>>>(a, b, nvec);
    CK(cudaEventRecord(e0));
    for (int w = 0; w < 5; ++w) copy_kernel<<<148 * 8, 256>>>(a, b, nvec);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("streaming copy 1 GiB: %.1f GB/s (read + write)\n", 2.0 * nvec * 16 / (ms / 5) / 1e6);
  }
  return 0;
}
