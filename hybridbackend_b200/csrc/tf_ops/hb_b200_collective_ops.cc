/* hb_b200_collective_ops.cc -- TensorFlow 1.15 OpKernel shims of the communicator and
 * the all-to-all family over libhb_b200.so's NVSwitch peer-memory communicator.
 *
 * TensorFlow is not in this repository's image: the file is TYPE-CHECKED against a
 * declaration-level stand-in of the TF-1.15 API (tests/test_tf_shims.py,
 * oracle/tf_shim_stub) but not linked or run here.  The reference's REGISTER_OP
 * definitions stay where they are; these kernels replace its REGISTER_KERNEL_BUILDER
 * lines:
 *   HbNcclCollectiveHandleOp / HbIsNcclCollectiveInitialized   nccl_create.cc:32-43,:65-72
 *   HbCreateNcclCollective(handle, id; world_size, local_size, rank, shared_name)  :45-62,:74-134
 *   HbGetNcclId -> id int64[16]                                 nccl_get_id.cc:35-70
 *   HbNcclAlltoall / HbNcclAlltoallN                            nccl_alltoall.cc:169-175,:242-249
 *   HbNcclAlltoallv / HbNcclAlltoallvN                          nccl_alltoallv.cc:200-223,:359-387
 * Threading follows the reference (common/stream.cc:83-130): the collective runs on
 * the communicator's OWN stream, fenced against the TF compute stream by events both
 * ways, and the host wait for the receive sizes happens on the communicator's worker
 * thread, never on the executor thread.
 */
#if HB_B200_WITH_TENSORFLOW

#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include <tensorflow/core/framework/op_kernel.h>
#include <tensorflow/core/framework/resource_mgr.h>
#include <tensorflow/core/framework/tensor.h>
#include <tensorflow/core/lib/core/threadpool.h>

#include "hb_b200.h"

namespace tensorflow {
namespace hybridbackend {

static const int64 kIdElements = HB_COMM_TOKEN_BYTES / sizeof(int64);   // 16, as the NCCL id
static const size_t kDefaultWindowBytes = size_t(1) << 30;

// Session resource owning the hbComm (one per process/GPU), its stream, the two fence
// events and a one-thread pool (the reference uses three threads for the same job).
class HbB200Collective : public ResourceBase {
 public:
  HbB200Collective() : comm_(nullptr), stream_(nullptr), in_(nullptr), out_(nullptr), pool_(nullptr) {}
  ~HbB200Collective() override {
    hbCommDestroy(comm_);
    if (in_) cudaEventDestroy(in_);
    if (out_) cudaEventDestroy(out_);
    if (stream_) cudaStreamDestroy(stream_);
  }
  Status Create(const unsigned char* id, int rank, int world, int local, thread::ThreadPool* pool) {
    pool_ = pool;
    if (cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&in_, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&out_, cudaEventDisableTiming) != cudaSuccess)
      return errors::Internal("hb_b200: cannot create the communicator stream");
    if (hbCommCreateFromId(id, rank, world, local, kDefaultWindowBytes, &comm_) != HB_OK)
      return errors::Internal(hbGetLastErrorString());
    return Status::OK();
  }
  hbComm* comm() const { return comm_; }
  cudaStream_t stream() const { return stream_; }
  // comm stream waits for what the compute stream has enqueued so far (stream.cc:83-101)
  Status WaitCompute(cudaStream_t compute) {
    if (cudaEventRecord(in_, compute) != cudaSuccess || cudaStreamWaitEvent(stream_, in_, 0) != cudaSuccess)
      return errors::Internal("hb_b200: event fence (compute -> comm) failed");
    return Status::OK();
  }
  // compute stream waits for the collective (stream.cc:103-130); outputs are then safe
  // to consume by ops the executor enqueues after done()
  Status BlockCompute(cudaStream_t compute) {
    if (cudaEventRecord(out_, stream_) != cudaSuccess || cudaStreamWaitEvent(compute, out_, 0) != cudaSuccess)
      return errors::Internal("hb_b200: event fence (comm -> compute) failed");
    return Status::OK();
  }
  void Schedule(std::function<void()> fn) { pool_->Schedule(std::move(fn)); }
  string DebugString() const override { return "HbB200Collective"; }

 private:
  hbComm* comm_;
  cudaStream_t stream_;
  cudaEvent_t in_, out_;
  thread::ThreadPool* pool_;
};

// HbGetNcclId: the 128-byte id is a rendezvous name (hbGetUniqueId), not an NCCL id; the
// Python side broadcasts it unchanged (distribute/collective.py:108-115).
class HbB200GetIdOp : public OpKernel {
 public:
  explicit HbB200GetIdOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    Tensor* id = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({kIdElements}), &id));
    OP_REQUIRES(ctx, hbGetUniqueId(reinterpret_cast<unsigned char*>(id->flat<int64>().data())) == HB_OK,
                errors::Internal(hbGetLastErrorString()));
  }
};
REGISTER_KERNEL_BUILDER(Name("HbGetNcclId").Device(DEVICE_GPU).HostMemory("id"), HbB200GetIdOp);
REGISTER_KERNEL_BUILDER(Name("HbGetNcclId").Device(DEVICE_CPU), HbB200GetIdOp);

class HbB200CreateCollectiveOp : public AsyncOpKernel {
 public:
  explicit HbB200CreateCollectiveOp(OpKernelConstruction* ctx) : AsyncOpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("world_size", &world_size_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("local_size", &local_size_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rank", &rank_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("shared_name", &shared_name_));
  }
  void ComputeAsync(OpKernelContext* ctx, DoneCallback done) override {
    const Tensor* id = nullptr;
    OP_REQUIRES_OK_ASYNC(ctx, ctx->input("id", &id), done);
    OP_REQUIRES_ASYNC(ctx, id->NumElements() == kIdElements,
                      errors::InvalidArgument("id must be the int64[16] HbGetNcclId produced"), done);
    HbB200Collective* coll = new HbB200Collective();
    // blocks until every rank arrived: like ncclCommInitRank in the reference (nccl_create.cc:101)
    Status s = coll->Create(reinterpret_cast<const unsigned char*>(id->flat<int64>().data()), rank_,
                            world_size_, local_size_, WorkerPool());
    if (s.ok()) s = CreateResource(ctx, HandleFromInput(ctx, 0), coll);
    OP_REQUIRES_OK_ASYNC(ctx, s, done);
    done();
  }

 private:
  static thread::ThreadPool* WorkerPool();   // process-wide, created on first use
  int world_size_, local_size_, rank_;
  string shared_name_;
};
thread::ThreadPool* HbB200CreateCollectiveOp::WorkerPool() {
  static thread::ThreadPool* pool = new thread::ThreadPool(Env::Default(), "hb_b200_comm", 1);
  return pool;
}
REGISTER_KERNEL_BUILDER(Name("HbCreateNcclCollective").Device(DEVICE_GPU).HostMemory("id"),
                        HbB200CreateCollectiveOp);

// ---- HbNcclAlltoallv / HbNcclAlltoallvN ---------------------------------------------------
// PACKED = false: inputs (handle, input, input_sizes) -> (output, output_sizes)
// PACKED = true : inputs (handle, n_input x N, n_input_sizes x N) -> (n_output x N, n_output_sizes x N)
template <typename DTYPE, bool PACKED>
class HbB200AlltoallvOp : public AsyncOpKernel {
 public:
  explicit HbB200AlltoallvOp(OpKernelConstruction* ctx) : AsyncOpKernel(ctx), N_(1) {
    if (PACKED) {
      OP_REQUIRES_OK(ctx, ctx->GetAttr("N", &N_));
      std::vector<PartialTensorShape> shapes;
      OP_REQUIRES_OK(ctx, ctx->GetAttr("common_shape", &shapes));
      common_shapes_ = shapes;
    } else {
      PartialTensorShape shape;
      OP_REQUIRES_OK(ctx, ctx->GetAttr("common_shape", &shape));
      common_shapes_.push_back(shape);
    }
    for (size_t k = 0; k < common_shapes_.size(); ++k) {
      int64 c = 1;   // nccl_alltoallv.cc:408-415
      for (int d = 0; d < common_shapes_[k].dims(); ++d) c *= common_shapes_[k].dim_size(d);
      common_sizes_.push_back(c);
    }
  }

  void ComputeAsync(OpKernelContext* ctx, DoneCallback done) override {
    HbB200Collective* coll = nullptr;
    OP_REQUIRES_OK_ASYNC(ctx, LookupResource(ctx, HandleFromInput(ctx, 0), &coll), done);
    std::vector<const Tensor*> inputs(N_), sizes(N_);
    if (PACKED) {
      OpInputList n_input, n_input_sizes;
      OP_REQUIRES_OK_ASYNC(ctx, ctx->input_list("n_input", &n_input), done);
      OP_REQUIRES_OK_ASYNC(ctx, ctx->input_list("n_input_sizes", &n_input_sizes), done);
      for (int k = 0; k < N_; ++k) { inputs[k] = &n_input[k]; sizes[k] = &n_input_sizes[k]; }
    } else {
      inputs[0] = &ctx->input(1);
      sizes[0] = &ctx->input(2);
    }
    const int W = hbCommWorldSize(coll->comm());
    const int N = static_cast<int>(N_);
    cudaStream_t compute = static_cast<cudaStream_t>(ctx->eigen_device<Eigen::GpuDevice>().stream());

    // phase 1: sizes (replaces the D2H of the input sizes + NCCL AlltoallN + D2H, :497-533)
    std::vector<const int32*> d_send(N);
    std::vector<int32*> d_recv(N);
    for (int k = 0; k < N; ++k) {
      OP_REQUIRES_ASYNC(ctx, sizes[k]->NumElements() == W,
                        errors::InvalidArgument("input sizes must have one entry per rank"), done);
      d_send[k] = sizes[k]->flat<int32>().data();
      Tensor* out_sizes = nullptr;
      OP_REQUIRES_OK_ASYNC(ctx, ctx->allocate_output(N + k, TensorShape({W}), &out_sizes), done);
      d_recv[k] = out_sizes->flat<int32>().data();
    }
    AllocatorAttributes host_attrs;
    host_attrs.set_on_host(true);
    host_attrs.set_gpu_compatible(true);   // pinned, as the reference (:420-422)
    Tensor* h_sizes = new Tensor();
    Status s = ctx->allocate_temp(DT_INT32, TensorShape({N * W}), h_sizes, host_attrs);
    if (s.ok()) s = coll->WaitCompute(compute);
    if (s.ok() && hbAlltoallvNSizes(coll->comm(), N, d_send.data(), d_recv.data(),
                                    h_sizes->flat<int32>().data(), coll->stream()) != HB_OK)
      s = errors::Internal(hbGetLastErrorString());
    if (!s.ok()) {
      delete h_sizes;
      coll->Unref();
      OP_REQUIRES_OK_ASYNC(ctx, s, done);
    }
    // the output shapes are data dependent: wait for the sizes on the communicator's
    // worker thread (the reference blocks its comm thread the same way, :533)
    coll->Schedule([this, ctx, done, coll, inputs, h_sizes, compute, W, N]() {
      core::ScopedUnref unref(coll);
      Status st = cudaStreamSynchronize(coll->stream()) == cudaSuccess
                      ? Status::OK() : errors::Internal("hb_b200: stream synchronize failed");
      // phase 2: payload (replaces NcclCollective::AlltoallvN, nccl_collective.cc:290-336)
      std::vector<const void*> d_in(N);
      std::vector<void*> d_out(N);
      std::vector<int32> elem_bytes(N, static_cast<int32>(sizeof(DTYPE)));
      for (int k = 0; k < N && st.ok(); ++k) {
        int64 total = 0;
        for (int q = 0; q < W; ++q) total += h_sizes->flat<int32>()(k * W + q);
        TensorShape shape({total});
        shape.AppendShape(TensorShape(common_shapes_[k].dim_sizes()));
        Tensor* out = nullptr;
        st = ctx->allocate_output(k, shape, &out);   // :534-553
        if (!st.ok()) break;
        d_in[k] = inputs[k]->flat<DTYPE>().data();
        d_out[k] = out->flat<DTYPE>().data();
      }
      delete h_sizes;
      if (st.ok() && hbAlltoallvN(coll->comm(), N, d_in.data(), common_sizes_.data(), elem_bytes.data(),
                                  d_out.data(), /*d_status=*/nullptr, coll->stream()) != HB_OK)
        st = errors::Internal(hbGetLastErrorString());
      if (st.ok()) st = coll->BlockCompute(compute);
      OP_REQUIRES_OK_ASYNC(ctx, st, done);
      done();
    });
  }

 private:
  int64 N_;
  std::vector<int64_t> common_sizes_;   // int64_t: the C-ABI's type (TF's int64 is long long)
  std::vector<PartialTensorShape> common_shapes_;
};

// ---- HbNcclAlltoall / HbNcclAlltoallN: equal split of dim 0 -----------------------------
// (nccl_alltoall.cc:177-230, :251-330): static sizes, so no size exchange on the wire is
// needed for the shapes; the kernels still take them from the device size matrix.
template <typename DTYPE, bool PACKED>
class HbB200AlltoallOp : public AsyncOpKernel {
 public:
  explicit HbB200AlltoallOp(OpKernelConstruction* ctx) : AsyncOpKernel(ctx), N_(1) {
    if (PACKED) OP_REQUIRES_OK(ctx, ctx->GetAttr("N", &N_));
  }
  void ComputeAsync(OpKernelContext* ctx, DoneCallback done) override {
    HbB200Collective* coll = nullptr;
    OP_REQUIRES_OK_ASYNC(ctx, LookupResource(ctx, HandleFromInput(ctx, 0), &coll), done);
    core::ScopedUnref unref(coll);
    const int W = hbCommWorldSize(coll->comm());
    const int N = static_cast<int>(N_);
    cudaStream_t compute = static_cast<cudaStream_t>(ctx->eigen_device<Eigen::GpuDevice>().stream());
    std::vector<const Tensor*> inputs(N);
    OpInputList n_input;
    if (PACKED) {
      OP_REQUIRES_OK_ASYNC(ctx, ctx->input_list("n_input", &n_input), done);
      for (int k = 0; k < N; ++k) inputs[k] = &n_input[k];
    } else {
      inputs[0] = &ctx->input(1);
    }
    // sizes[k][r] = rows / W for every r; one small pinned->device copy feeds the size kernel
    Tensor d_sizes;
    OP_REQUIRES_OK_ASYNC(ctx, ctx->allocate_temp(DT_INT32, TensorShape({N * W}), &d_sizes), done);
    std::vector<int32> h(N * W);
    std::vector<const void*> d_in(N);
    std::vector<void*> d_out(N);
    std::vector<int64_t> common(N);
    std::vector<int32> elem_bytes(N, static_cast<int32>(sizeof(DTYPE)));
    std::vector<const int32*> d_send(N);
    for (int k = 0; k < N; ++k) {
      const int64 rows = inputs[k]->dim_size(0);
      OP_REQUIRES_ASYNC(ctx, rows % W == 0,
                        errors::InvalidArgument("alltoall: dim 0 must be divisible by the world size"), done);
      for (int r = 0; r < W; ++r) h[k * W + r] = static_cast<int32>(rows / W);
      common[k] = rows > 0 ? inputs[k]->NumElements() / rows : 1;
      Tensor* out = nullptr;
      OP_REQUIRES_OK_ASYNC(ctx, ctx->allocate_output(k, inputs[k]->shape(), &out), done);
      d_in[k] = inputs[k]->flat<DTYPE>().data();
      d_out[k] = out->flat<DTYPE>().data();
      d_send[k] = d_sizes.flat<int32>().data() + k * W;
    }
    OP_REQUIRES_ASYNC(ctx,
                      cudaMemcpyAsync(d_sizes.flat<int32>().data(), h.data(), sizeof(int32) * N * W,
                                      cudaMemcpyHostToDevice, compute) == cudaSuccess &&
                          cudaStreamSynchronize(compute) == cudaSuccess,   // h is a stack vector
                      errors::Internal("hb_b200: staging the split sizes failed"), done);
    OP_REQUIRES_OK_ASYNC(ctx, coll->WaitCompute(compute), done);
    OP_REQUIRES_ASYNC(ctx,
                      hbAlltoallvNSizes(coll->comm(), N, d_send.data(), nullptr, nullptr, coll->stream()) == HB_OK &&
                          hbAlltoallvN(coll->comm(), N, d_in.data(), common.data(), elem_bytes.data(),
                                       d_out.data(), nullptr, coll->stream()) == HB_OK,
                      errors::Internal(hbGetLastErrorString()), done);
    OP_REQUIRES_OK_ASYNC(ctx, coll->BlockCompute(compute), done);
    done();
  }

 private:
  int64 N_;
};

#define HB_B200_REGISTER_A2A(T)                                                                      \
  REGISTER_KERNEL_BUILDER(Name("HbNcclAlltoallv").Device(DEVICE_GPU).TypeConstraint<T>("dtype")      \
                              .TypeConstraint<float>("wire_dtype").HostMemory("handle"),             \
                          HbB200AlltoallvOp<T, false>);                                              \
  REGISTER_KERNEL_BUILDER(Name("HbNcclAlltoallvN").Device(DEVICE_GPU).TypeConstraint<T>("dtype")     \
                              .TypeConstraint<float>("wire_dtype").HostMemory("handle"),             \
                          HbB200AlltoallvOp<T, true>);                                               \
  REGISTER_KERNEL_BUILDER(Name("HbNcclAlltoall").Device(DEVICE_GPU).TypeConstraint<T>("dtype")       \
                              .TypeConstraint<float>("wire_dtype").HostMemory("handle"),             \
                          HbB200AlltoallOp<T, false>);                                               \
  REGISTER_KERNEL_BUILDER(Name("HbNcclAlltoallN").Device(DEVICE_GPU).TypeConstraint<T>("dtype")      \
                              .TypeConstraint<float>("wire_dtype").HostMemory("handle"),             \
                          HbB200AlltoallOp<T, true>)
HB_B200_REGISTER_A2A(int8);
HB_B200_REGISTER_A2A(uint8);
HB_B200_REGISTER_A2A(int32);
HB_B200_REGISTER_A2A(uint32);
HB_B200_REGISTER_A2A(int64);
HB_B200_REGISTER_A2A(uint64);
HB_B200_REGISTER_A2A(float);
HB_B200_REGISTER_A2A(double);
HB_B200_REGISTER_A2A(Eigen::half);

}  // namespace hybridbackend
}  // namespace tensorflow

#endif  // HB_B200_WITH_TENSORFLOW
