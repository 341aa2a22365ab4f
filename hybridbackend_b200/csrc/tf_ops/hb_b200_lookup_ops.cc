/* hb_b200_lookup_ops.cc -- TensorFlow 1.15 OpKernel shim of HbLookup (the slab-hash cache
 * probe of the reference's EmbeddingService path) over hbCacheLookup.
 *   REGISTER_OP("HbLookup")  embedding/lookup_ops.cc:38-58: keys_cache:T, keys:T ->
 *     hit_keys_indices:int32, hit_cache_indices:T, miss_keys_indices:int32, miss_keys:T;
 *     attr cache_slab_size; int64 GPU kernel only (:141-145).
 * The reference kernel learns the miss count on the host to shape its four outputs
 * (lookup_ops.cc:94-137); so does this shim (one 8-byte D2H + stream wait).
 * Type-checked by tests/test_tf_shims.py against oracle/tf_shim_stub; not linked here.
 */
#if HB_B200_WITH_TENSORFLOW

#include <cuda_runtime.h>

#include <tensorflow/core/framework/op_kernel.h>
#include <tensorflow/core/framework/tensor.h>

#include "hb_b200.h"

namespace tensorflow {
namespace hybridbackend {

class HbB200LookupOp : public OpKernel {
 public:
  explicit HbB200LookupOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("cache_slab_size", &cache_slab_size_));
    OP_REQUIRES(ctx, cache_slab_size_ == 32,
                errors::InvalidArgument("hb_b200 probes 32-key slabs (one warp per slab)"));
  }

  void Compute(OpKernelContext* ctx) override {
    const Tensor& keys_cache = ctx->input(0);
    const Tensor& keys = ctx->input(1);
    const int32 n = static_cast<int32>(keys.NumElements());
    const int64 slabs = keys_cache.NumElements() / cache_slab_size_;
    cudaStream_t stream = static_cast<cudaStream_t>(ctx->eigen_device<Eigen::GpuDevice>().stream());
    // hits from the front, misses from the back of two n-sized scratch vectors
    Tensor idx, val, counts, h_counts;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT32, TensorShape({n}), &idx));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT64, TensorShape({n}), &val));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT32, TensorShape({2}), &counts));
    AllocatorAttributes host;
    host.set_on_host(true);
    host.set_gpu_compatible(true);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT32, TensorShape({2}), &h_counts, host));
    OP_REQUIRES(ctx,
                hbCacheLookup(reinterpret_cast<const int64_t*>(keys_cache.flat<int64>().data()), slabs,
                              reinterpret_cast<const int64_t*>(keys.flat<int64>().data()), n,
                              idx.flat<int32>().data(), reinterpret_cast<int64_t*>(val.flat<int64>().data()),
                              counts.flat<int32>().data(), stream) == HB_OK,
                errors::Internal(hbGetLastErrorString()));
    OP_REQUIRES(ctx,
                cudaMemcpyAsync(h_counts.flat<int32>().data(), counts.flat<int32>().data(), 8,
                                cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
                    cudaStreamSynchronize(stream) == cudaSuccess,
                errors::Internal("hb_b200: reading the miss count failed"));
    const int32 miss = h_counts.flat<int32>()(0), hit = h_counts.flat<int32>()(1);
    Tensor *hit_idx = nullptr, *hit_cache = nullptr, *miss_idx = nullptr, *miss_keys = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({hit}), &hit_idx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({hit}), &hit_cache));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({miss}), &miss_idx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({miss}), &miss_keys));
    const cudaMemcpyKind d2d = cudaMemcpyDeviceToDevice;
    OP_REQUIRES(ctx,
                cudaMemcpyAsync(hit_idx->flat<int32>().data(), idx.flat<int32>().data(), 4 * size_t(hit), d2d, stream) == cudaSuccess &&
                    cudaMemcpyAsync(hit_cache->flat<int64>().data(), val.flat<int64>().data(), 8 * size_t(hit), d2d, stream) == cudaSuccess &&
                    cudaMemcpyAsync(miss_idx->flat<int32>().data(), idx.flat<int32>().data() + (n - miss), 4 * size_t(miss), d2d, stream) == cudaSuccess &&
                    cudaMemcpyAsync(miss_keys->flat<int64>().data(), val.flat<int64>().data() + (n - miss), 8 * size_t(miss), d2d, stream) == cudaSuccess,
                errors::Internal("hb_b200: splitting hits and misses failed"));
  }

 private:
  int64 cache_slab_size_;
};

REGISTER_KERNEL_BUILDER(Name("HbLookup").Device(DEVICE_GPU).TypeConstraint<int64>("T"), HbB200LookupOp);

}  // namespace hybridbackend
}  // namespace tensorflow

#endif  // HB_B200_WITH_TENSORFLOW
