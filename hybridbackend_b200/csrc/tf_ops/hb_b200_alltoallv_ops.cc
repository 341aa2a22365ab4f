/* hb_b200_alltoallv_ops.cc -- TensorFlow 1.15 AsyncOpKernel shim for HbNcclAlltoallvN
 * over libhb_b200.so's NVSwitch peer-memory communicator.
 *
 * NOT BUILT OR TESTED IN THIS REPOSITORY'S ENVIRONMENT (no TensorFlow headers).  It
 * shows where each C-ABI call lands inside the reference's op
 *   REGISTER_OP("HbNcclAlltoallvN")  distribute/nccl/nccl_alltoallv.cc:359-387
 * whose kernel NcclAlltoallvNOp::CollectiveComputeAsync is nccl_alltoallv.cc:418-564.
 * The communicator resource below replaces NcclCollective
 * (distribute/nccl/nccl_collective.cc:40-65, :434-465): its Create() receives the
 * `[W, 16] int64` all-gathered tokens instead of a broadcast NCCL id.
 */
#if HB_B200_WITH_TENSORFLOW

#include <vector>

#include <tensorflow/core/framework/op_kernel.h>
#include <tensorflow/core/framework/resource_mgr.h>
#include <tensorflow/core/framework/tensor.h>

#include "hb_b200.h"

namespace tensorflow {
namespace hybridbackend {

// Session resource owning the hbComm (one per process/GPU).
class HbB200Collective : public ResourceBase {
 public:
  HbB200Collective() : comm_(nullptr) {}
  ~HbB200Collective() override { hbCommDestroy(comm_); }
  Status Create(int rank, int world, int local, size_t window_bytes, unsigned char* token_out) {
    if (hbCommCreate(rank, world, local, window_bytes, &comm_, token_out) != HB_OK)
      return errors::Internal(hbGetLastErrorString());
    return Status::OK();
  }
  Status Connect(const unsigned char* all_tokens) {
    if (hbCommConnect(comm_, all_tokens) != HB_OK) return errors::Internal(hbGetLastErrorString());
    return Status::OK();
  }
  hbComm* comm() const { return comm_; }
  string DebugString() const override { return "HbB200Collective"; }

 private:
  hbComm* comm_;
};

template <typename DTYPE>
class HbB200AlltoallvNOp : public AsyncOpKernel {
 public:
  explicit HbB200AlltoallvNOp(OpKernelConstruction* ctx) : AsyncOpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("N", &N_));
    std::vector<PartialTensorShape> common_shape;
    OP_REQUIRES_OK(ctx, ctx->GetAttr("common_shape", &common_shape));
    for (int k = 0; k < N_; ++k) {
      int64 c = 1;
      for (int d = 0; d < common_shape[k].dims(); ++d) c *= common_shape[k].dim_size(d);
      common_sizes_.push_back(c);        // nccl_alltoallv.cc:408-415
      common_shapes_.push_back(common_shape[k]);
    }
  }

  void ComputeAsync(OpKernelContext* ctx, DoneCallback done) override {
    HbB200Collective* coll = nullptr;
    OP_REQUIRES_OK_ASYNC(ctx, LookupResource(ctx, HandleFromInput(ctx, 0), &coll), done);
    core::ScopedUnref unref(coll);
    OpInputList n_input, n_input_sizes;
    OP_REQUIRES_OK_ASYNC(ctx, ctx->input_list("n_input", &n_input), done);
    OP_REQUIRES_OK_ASYNC(ctx, ctx->input_list("n_input_sizes", &n_input_sizes), done);
    const int W = hbCommWorldSize(coll->comm());
    auto stream = ctx->eigen_device<Eigen::GpuDevice>().stream();

    // phase 1: sizes (replaces the D2H of input sizes + NCCL AlltoallN + D2H, :497-533)
    std::vector<const int32*> d_send(N_);
    std::vector<int32*> d_recv(N_);
    for (int k = 0; k < N_; ++k) {
      OP_REQUIRES_ASYNC(ctx, n_input_sizes[k].NumElements() == W,
                        errors::InvalidArgument("n_input_sizes must have one entry per rank"), done);
      d_send[k] = n_input_sizes[k].flat<int32>().data();
      Tensor* out_sizes = nullptr;
      OP_REQUIRES_OK_ASYNC(ctx, ctx->allocate_output(N_ + k, {W}, &out_sizes), done);
      d_recv[k] = out_sizes->flat<int32>().data();
    }
    AllocatorAttributes host_attrs;
    host_attrs.set_on_host(true);
    host_attrs.set_gpu_compatible(true);   // pinned, as the reference (:420-422)
    Tensor h_sizes;
    OP_REQUIRES_OK_ASYNC(ctx, ctx->allocate_temp(DT_INT32, {N_ * W}, &h_sizes, host_attrs), done);
    OP_REQUIRES_ASYNC(ctx,
                      hbAlltoallvNSizes(coll->comm(), N_, d_send.data(), d_recv.data(),
                                        h_sizes.flat<int32>().data(), stream) == HB_OK,
                      errors::Internal(hbGetLastErrorString()), done);
    // the output shapes are data dependent: block like the reference (:533)
    OP_REQUIRES_ASYNC(ctx, cudaStreamSynchronize(stream) == cudaSuccess,
                      errors::Internal("stream synchronize failed"), done);

    // phase 2: payload (replaces NcclCollective::AlltoallvN, nccl_collective.cc:290-336)
    std::vector<const void*> d_in(N_);
    std::vector<void*> d_out(N_);
    std::vector<int32> elem_bytes(N_, static_cast<int32>(sizeof(DTYPE)));
    for (int k = 0; k < N_; ++k) {
      int64 total = 0;
      for (int q = 0; q < W; ++q) total += h_sizes.flat<int32>()(k * W + q);
      TensorShape shape({total});
      shape.AppendShape(TensorShape(common_shapes_[k].dim_sizes()));
      Tensor* out = nullptr;
      OP_REQUIRES_OK_ASYNC(ctx, ctx->allocate_output(k, shape, &out), done);   // :534-553
      d_in[k] = n_input[k].flat<DTYPE>().data();
      d_out[k] = out->flat<DTYPE>().data();
    }
    OP_REQUIRES_ASYNC(ctx,
                      hbAlltoallvN(coll->comm(), N_, d_in.data(), common_sizes_.data(),
                                   elem_bytes.data(), d_out.data(), /*d_status=*/nullptr,
                                   stream) == HB_OK,
                      errors::Internal(hbGetLastErrorString()), done);
    done();
  }

 private:
  int64 N_;
  std::vector<int64> common_sizes_;
  std::vector<PartialTensorShape> common_shapes_;
};

#define HB_B200_REGISTER_A2AV(T)                                                        \
  REGISTER_KERNEL_BUILDER(Name("HbNcclAlltoallvN").Device(DEVICE_GPU)                   \
                              .TypeConstraint<T>("dtype").TypeConstraint<float>("wire_dtype") \
                              .HostMemory("handle"),                                    \
                          HbB200AlltoallvNOp<T>)
HB_B200_REGISTER_A2AV(int32);
HB_B200_REGISTER_A2AV(int64);
HB_B200_REGISTER_A2AV(float);
HB_B200_REGISTER_A2AV(double);
HB_B200_REGISTER_A2AV(Eigen::half);

}  // namespace hybridbackend
}  // namespace tensorflow

#endif  // HB_B200_WITH_TENSORFLOW
