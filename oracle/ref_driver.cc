// C entry points over the reference's own CPU functors, compiled unmodified from
// /root/reference by oracle/build_ref.sh (outputs only into oracle/_ref/).
// Test infrastructure: used to pin oracle/hb_oracle.c and, when present, as the
// "reference" CPU baseline of the partition stage.
#include <cstdint>
#include <tensorflow/core/framework/tensor.h>
#include "hybridbackend/tensorflow/distribute/partition/dual_modulo_functors.h"
#include "hybridbackend/tensorflow/distribute/partition/modulo_functors.h"

#define __host__
#define __device__
#include "hybridbackend/common/murmur3.cu.h"

namespace tensorflow {
namespace hybridbackend {
namespace functor {
struct ComputeShardAtStageOne;
struct ComputeShardAtStageTwo;
}  // namespace functor
}  // namespace hybridbackend
}  // namespace tensorflow

using namespace tensorflow;
using CPUDevice = Eigen::ThreadPoolDevice;
namespace hbf = tensorflow::hybridbackend::functor;

template <typename T>
static int run_modulo(const void* in, int32_t n, int32_t p, void* out,
                      int32_t* sizes, int32_t* idx) {
  Tensor ti(const_cast<void*>(in), n), to(out, n), ts(sizes, p), tx(idx, n);
  OpKernelContext ctx;
  hbf::PartitionByModulo<CPUDevice, T>()(p, ti, &to, &ts, &tx, &ctx);
  return ctx.status.ok() ? 0 : 1;
}

template <typename T, typename Stage>
static int run_dual(const void* in, int32_t n, int32_t p, int32_t m, void* out,
                    int32_t* sizes, int32_t* idx) {
  Tensor ti(const_cast<void*>(in), n), to(out, n), ts(sizes, p), tx(idx, n);
  OpKernelContext ctx;
  hbf::PartitionByDualModulo<CPUDevice, T, Stage>()(p, m, ti, &to, &ts, &tx, &ctx);
  return ctx.status.ok() ? 0 : 1;
}

extern "C" {
// dtype codes: 0 int32, 1 int64, 2 uint32, 3 uint64 (as oracle/hb_oracle.h)
int hbref_partition_by_modulo(int dtype, const void* in, int32_t n, int32_t p,
                              void* out, int32_t* sizes, int32_t* idx) {
  switch (dtype) {
    case 0: return run_modulo<int32>(in, n, p, out, sizes, idx);
    case 1: return run_modulo<int64>(in, n, p, out, sizes, idx);
    case 2: return run_modulo<uint32>(in, n, p, out, sizes, idx);
    case 3: return run_modulo<uint64>(in, n, p, out, sizes, idx);
  }
  return 1;
}

int hbref_partition_by_dual_modulo(int dtype, int stage, const void* in,
                                   int32_t n, int32_t p, int32_t m, void* out,
                                   int32_t* sizes, int32_t* idx) {
#define HBREF_DUAL(T)                                                         \
  return stage == 1                                                           \
             ? run_dual<T, hbf::ComputeShardAtStageOne>(in, n, p, m, out, sizes, idx) \
             : run_dual<T, hbf::ComputeShardAtStageTwo>(in, n, p, m, out, sizes, idx)
  switch (dtype) {
    case 0: HBREF_DUAL(int32);
    case 1: HBREF_DUAL(int64);
    case 2: HBREF_DUAL(uint32);
    case 3: HBREF_DUAL(uint64);
  }
  return 1;
}

uint32_t hbref_murmur3_hash32_i64(long long key) { return murmur3_hash32<long long>(key); }
}
