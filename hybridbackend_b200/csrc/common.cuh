// common.cuh -- shared device/host helpers for the hb_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hb_b200.h"

namespace hb {

// ---- host-side error plumbing (C-ABI returns int, message via TLS) --------
void set_last_error(const char* fmt, ...);

#define HB_CUDA_OK(expr)                                                     \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      ::hb::set_last_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,    \
                           cudaGetErrorName(_e), cudaGetErrorString(_e));    \
      return HB_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

#define HB_REQUIRE(cond, ...)                                                \
  do {                                                                       \
    if (!(cond)) {                                                           \
      ::hb::set_last_error(__VA_ARGS__);                                     \
      return HB_ERR_INVALID;                                                 \
    }                                                                        \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int device_sm_count();

// Every kernel launch of the library goes through a KernelScope: it counts the
// launch (hbGetLaunchCount) and, when profiling is enabled (hbProfileEnable),
// brackets it with CUDA events on the launching stream (hbProfileGet).
struct KernelScope {
  int id;
  cudaStream_t stream;
  void* rec;
  KernelScope(int id_, cudaStream_t s);
  ~KernelScope();
};

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// 128-bit streaming loads/stores (read-once data: bypass L1 allocation).
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
  float4 r;
  // not volatile: a pure load of read-only data, free to be hoisted and batched
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_f4(const float4* p) { return *p; }
__device__ __forceinline__ void st_na_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int64_t ld_nc_i64(const int64_t* p) {
  int64_t r;
  asm("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}

// system-scope release/acquire for cross-GPU flags.
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Cross-GPU dependency of a kernel: spin until flags[0..n) (epoch-valued words
// in LOCAL memory, written by peers with st.release.sys) reach `epoch`.
struct WaitSpec {
  const uint32_t* flags;
  uint32_t epoch;
  int32_t n;
};

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Bounded spin on an epoch flag: a peer that never arrives (crashed rank, ranks
// issuing different op sequences) must not hang the GPU forever -- after
// kSpinTimeoutNs the waiter gives up and raises HB_STATUS_PEER_TIMEOUT (the
// reference relies on an NCCL watchdog thread for the same failure,
// nccl/nccl_create.cc:104-117).  Returns false on timeout.
constexpr uint64_t kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ bool spin_until(const uint32_t* flag, uint32_t epoch) {
  if ((int32_t)(ld_acquire_sys_u32(flag) - epoch) >= 0) return true;
  const uint64_t t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys_u32(flag) - epoch) < 0) {
    if (globaltimer_ns() - t0 > kSpinTimeoutNs) return false;
  }
  return true;
}

__device__ __forceinline__ void wait_spec(const WaitSpec& w, int32_t* status = nullptr) {
  if (w.flags != nullptr) {
    if ((int)threadIdx.x < w.n) {
      if (!spin_until(w.flags + threadIdx.x, w.epoch) && status != nullptr)
        atomicOr(status, HB_STATUS_PEER_TIMEOUT);
    }
    __syncthreads();
  }
}

// sticky device status word bits (hbStatusWord)
__device__ __forceinline__ void raise_status(int32_t* status, int32_t bit) {
  if (status != nullptr) atomicOr(status, bit);
}

// ---- internal entry points shared between translation units ----------------
// lookup.cu: idx32[k] != nullptr -> rows of feature k are addressed by int32
// indices (stitch of the sharded path) instead of int64 ids; `coherent` uses
// plain ld.global (data written by peers) instead of ld.global.nc.
int lookup_forward_run(int n, const hbLookupFeature* feats, const int32_t* const* idx32,
                       const WaitSpec* wait, bool coherent, int32_t* d_status,
                       cudaStream_t stream, int kernel_id);
}  // namespace hb
