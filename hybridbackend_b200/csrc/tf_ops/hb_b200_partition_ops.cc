/* hb_b200_partition_ops.cc -- TensorFlow 1.15 OpKernel shim over libhb_b200.so.
 *
 * NOT BUILT OR TESTED IN THIS REPOSITORY'S ENVIRONMENT (no TensorFlow headers,
 * Python 3.12): it is the binding a HybridBackend maintainer adds, written against
 * the reference's op definitions, which stay where they are:
 *   REGISTER_OP("HbPartitionByModulo")   distribute/partition/partition_by_modulo_ops.cc:46-60
 *   REGISTER_OP("HbPartitionByModuloN")  ...:124-143
 *   REGISTER_OP("HbPartitionByDualModuloStage{One,Two}[N]")  partition_by_dual_modulo_ops.cc:46-53,...
 * Only the GPU REGISTER_KERNEL_BUILDER lines of those files are replaced by the
 * ones below.  Build (where TF 1.15 is installed):
 *   g++ -std=c++11 -shared -fPIC hb_b200_partition_ops.cc -I<repo>/include \
 *       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))') \
 *       -L<repo>/hybridbackend_b200/lib -lhb_b200 \
 *       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
 */
#if HB_B200_WITH_TENSORFLOW

#define EIGEN_USE_GPU
#include <vector>

#include <tensorflow/core/framework/op_kernel.h>
#include <tensorflow/core/framework/register_types.h>
#include <tensorflow/core/framework/tensor.h>

#include "hb_b200.h"

namespace tensorflow {
namespace hybridbackend {

using GPUDevice = Eigen::GpuDevice;

template <typename T> struct HbDType;
template <> struct HbDType<int32> { static constexpr int v = HB_I32; };
template <> struct HbDType<int64> { static constexpr int v = HB_I64; };
template <> struct HbDType<uint32> { static constexpr int v = HB_U32; };
template <> struct HbDType<uint64> { static constexpr int v = HB_U64; };

// stage 0: modulo; 1 / 2: dual modulo stage one / two.  N == 0: single-input op.
template <typename T, int STAGE, bool PACKED>
class HbB200PartitionOp : public OpKernel {
 public:
  explicit HbB200PartitionOp(OpKernelConstruction* ctx) : OpKernel(ctx), modulus_(1) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("num_partitions", &num_partitions_));
    if (STAGE != 0) OP_REQUIRES_OK(ctx, ctx->GetAttr("modulus", &modulus_));
  }

  void Compute(OpKernelContext* ctx) override {
    std::vector<const Tensor*> inputs;
    if (PACKED) {
      OpInputList list;
      OP_REQUIRES_OK(ctx, ctx->input_list("inputs", &list));
      for (int i = 0; i < list.size(); ++i) inputs.push_back(&list[i]);
    } else {
      inputs.push_back(&ctx->input(0));
    }
    const int n = static_cast<int>(inputs.size());
    std::vector<const void*> in(n);
    std::vector<void*> out(n);
    std::vector<int32*> sz(n), ix(n);
    std::vector<int32> lens(n);
    for (int i = 0; i < n; ++i) {
      OP_REQUIRES(ctx, TensorShapeUtils::IsVector(inputs[i]->shape()),
                  errors::InvalidArgument("partition_by_modulo expects a 1D vector."));
      lens[i] = static_cast<int32>(inputs[i]->NumElements());
      // output index layout of the packed op: [outputs x N, sizes x N, indices x N]
      Tensor *o = nullptr, *s = nullptr, *x = nullptr;
      OP_REQUIRES_OK(ctx, ctx->allocate_output(i, {lens[i]}, &o));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(n + i, {num_partitions_}, &s));
      OP_REQUIRES_OK(ctx, ctx->allocate_output(2 * n + i, {lens[i]}, &x));
      in[i] = inputs[i]->flat<T>().data();
      out[i] = o->flat<T>().data();
      sz[i] = s->flat<int32>().data();
      ix[i] = x->flat<int32>().data();
    }
    size_t ws_bytes = 0;
    OP_REQUIRES(ctx, hbPartitionWorkspaceBytes(n, lens.data(), num_partitions_, &ws_bytes) == HB_OK,
                errors::Internal(hbGetLastErrorString()));
    Tensor ws;  // scratch: TF keeps it alive until the stream has passed this op
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT8, {static_cast<int64>(ws_bytes)}, &ws));
    auto stream = ctx->eigen_device<GPUDevice>().stream();
    int rc;
    if (STAGE == 0) {
      rc = hbPartitionByModuloN(HbDType<T>::v, n, in.data(), lens.data(), num_partitions_, out.data(),
                                sz.data(), ix.data(), ws.flat<int8>().data(), ws_bytes, stream);
    } else {
      rc = hbPartitionByDualModuloN(HbDType<T>::v, STAGE, n, in.data(), lens.data(), num_partitions_,
                                    modulus_, out.data(), sz.data(), ix.data(),
                                    ws.flat<int8>().data(), ws_bytes, stream);
    }
    OP_REQUIRES(ctx, rc == HB_OK, errors::Internal(hbGetLastErrorString()));
  }

 private:
  int32 num_partitions_;
  int32 modulus_;
};

#define HB_B200_REGISTER(NAME, T, STAGE, PACKED)                                        \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(DEVICE_GPU).TypeConstraint<T>("T"),         \
                          HbB200PartitionOp<T, STAGE, PACKED>)
#define HB_B200_REGISTER_ALL(T)                                           \
  HB_B200_REGISTER("HbPartitionByModulo", T, 0, false);                   \
  HB_B200_REGISTER("HbPartitionByModuloN", T, 0, true);                   \
  HB_B200_REGISTER("HbPartitionByDualModuloStageOne", T, 1, false);       \
  HB_B200_REGISTER("HbPartitionByDualModuloStageOneN", T, 1, true);       \
  HB_B200_REGISTER("HbPartitionByDualModuloStageTwo", T, 2, false);       \
  HB_B200_REGISTER("HbPartitionByDualModuloStageTwoN", T, 2, true)
HB_B200_REGISTER_ALL(int32);
HB_B200_REGISTER_ALL(int64);
HB_B200_REGISTER_ALL(uint32);
HB_B200_REGISTER_ALL(uint64);

}  // namespace hybridbackend
}  // namespace tensorflow

#endif  // HB_B200_WITH_TENSORFLOW
