// TEST INFRASTRUCTURE ONLY -- a declaration-level stand-in for the slice of the
// TensorFlow 1.15 C++ API that the OpKernel shims under hybridbackend_b200/csrc/tf_ops
// use, so that tests/test_tf_shims.py can type-check them (g++ -fsyntax-only) in an
// image without TensorFlow.  Signatures follow tensorflow/core/framework/{op_kernel.h,
// tensor.h,tensor_shape.h,resource_mgr.h,types.h} of r1.15; nothing here runs.
#ifndef HB_TF_SHIM_STUB_OP_KERNEL_H_
#define HB_TF_SHIM_STUB_OP_KERNEL_H_
#include <cstddef>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <string>
#include <typeinfo>
#include <vector>

namespace Eigen {
struct half { uint16_t x; };
struct ThreadPoolDevice {};
struct GpuDevice {
  void* stream() const;   // cudaStream_t
};
}  // namespace Eigen

namespace tensorflow {
typedef int8_t int8;
typedef uint8_t uint8;
typedef int32_t int32;
typedef long long int64;
typedef uint32_t uint32;
typedef unsigned long long uint64;
using std::string;

enum DataType { DT_INVALID = 0, DT_FLOAT = 1, DT_DOUBLE = 2, DT_INT32 = 3, DT_UINT8 = 4, DT_INT8 = 6,
                DT_INT64 = 9, DT_HALF = 19, DT_RESOURCE = 20, DT_UINT32 = 22, DT_UINT64 = 23 };
extern const char* const DEVICE_CPU;
extern const char* const DEVICE_GPU;

class Status {
 public:
  Status();
  static Status OK();
  bool ok() const;
  const string& error_message() const;
};
namespace errors {
Status Internal(const char* a, const char* b = "", const char* c = "");
Status Internal(const string& a);
Status InvalidArgument(const char* a, const char* b = "", const char* c = "");
Status Unimplemented(const char* a);
Status FailedPrecondition(const char* a);
}  // namespace errors

class TensorShape {
 public:
  TensorShape();
  TensorShape(std::initializer_list<int64> dims);
  explicit TensorShape(const std::vector<int64>& dims);
  int dims() const;
  int64 dim_size(int d) const;
  int64 num_elements() const;
  void AddDim(int64 size);
  void AppendShape(const TensorShape& s);
};
class PartialTensorShape {
 public:
  int dims() const;
  int64 dim_size(int d) const;
  std::vector<int64> dim_sizes() const;
  bool IsFullyDefined() const;
};
struct TensorShapeUtils {
  static bool IsScalar(const TensorShape& s);
  static bool IsVector(const TensorShape& s);
  static bool IsMatrix(const TensorShape& s);
};

class Tensor {
 public:
  Tensor();
  template <typename T>
  struct Flat {
    T* data() const;
    T& operator()(int64 i) const;
    int64 size() const;
  };
  const TensorShape& shape() const;
  int64 NumElements() const;
  int64 dim_size(int d) const;
  size_t TotalBytes() const;
  DataType dtype() const;
  template <typename T> Flat<T> flat();
  template <typename T> Flat<const T> flat() const;
  template <typename T> Flat<T> vec();
  template <typename T> Flat<const T> vec() const;
  template <typename T> Flat<T> scalar();
  template <typename T> Flat<const T> scalar() const;
};
class TensorReference {
 public:
  explicit TensorReference(const Tensor& t);
  void Unref() const;
};

class OpInputList {
 public:
  int size() const;
  const Tensor& operator[](int i) const;
};
class OpOutputList {
 public:
  int size() const;
  Status allocate(int i, const TensorShape& shape, Tensor** output);
};

struct AllocatorAttributes {
  void set_on_host(bool v);
  void set_gpu_compatible(bool v);
};

namespace core {
class RefCounted {
 public:
  void Ref() const;
  bool Unref() const;
 protected:
  virtual ~RefCounted();
};
class ScopedUnref {
 public:
  explicit ScopedUnref(const RefCounted* o);
  ~ScopedUnref();
};
}  // namespace core

class ResourceBase : public core::RefCounted {
 public:
  virtual string DebugString() const = 0;
};
class ResourceHandle {};
class ResourceMgr {
 public:
  template <typename T> Status Create(const string& container, const string& name, T* resource);
  template <typename T> Status Lookup(const string& container, const string& name, T** resource) const;
  template <typename T> Status LookupOrCreate(const string& container, const string& name, T** resource,
                                              std::function<Status(T**)> creator);
  const string& default_container() const;
};

class Env {
 public:
  static Env* Default();
};

namespace thread {
class ThreadPool {
 public:
  ThreadPool(Env* env, const string& name, int num_threads);
  void Schedule(std::function<void()> fn);
};
}  // namespace thread

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char* name, T* value) const;
  void SetStatus(const Status& s);
  void CtxFailure(const Status& s);
  void CtxFailureWithWarning(const Status& s);
};

class OpKernelContext {
 public:
  const Tensor& input(int index);
  Status input(const char* name, const Tensor** tensor);
  Status input_list(const char* name, OpInputList* list);
  Status output_list(const char* name, OpOutputList* list);
  int num_inputs() const;
  Status allocate_output(int index, const TensorShape& shape, Tensor** tensor);
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp);
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp, AllocatorAttributes attr);
  template <typename Device> const Device& eigen_device() const;
  ResourceMgr* resource_manager() const;
  void SetStatus(const Status& s);
  void CtxFailure(const Status& s);
  void CtxFailureWithWarning(const Status& s);
  const Status& status() const;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction* ctx);
  virtual ~OpKernel();
  virtual void Compute(OpKernelContext* ctx) = 0;
};
class AsyncOpKernel : public OpKernel {
 public:
  typedef std::function<void()> DoneCallback;
  explicit AsyncOpKernel(OpKernelConstruction* ctx);
  virtual void ComputeAsync(OpKernelContext* ctx, DoneCallback done) = 0;
  void Compute(OpKernelContext* ctx) override;
};

const ResourceHandle& HandleFromInput(OpKernelContext* ctx, int input);
template <typename T> Status LookupResource(OpKernelContext* ctx, const ResourceHandle& h, T** value);
template <typename T> Status CreateResource(OpKernelContext* ctx, const ResourceHandle& h, T* value);

// REGISTER_KERNEL_BUILDER(Name("Op").Device(DEVICE_GPU).TypeConstraint<T>("T"), Class)
struct KernelDefBuilder {
  KernelDefBuilder& Device(const char* d);
  template <typename T> KernelDefBuilder& TypeConstraint(const char* attr);
  KernelDefBuilder& HostMemory(const char* arg);
  KernelDefBuilder& Priority(int p);
};
KernelDefBuilder Name(const char* op);
// constructing K instantiates its virtual members (Compute / ComputeAsync), so kernels that
// are class templates are type-checked too
template <typename K> struct KernelRegistrarStub {
  explicit KernelRegistrarStub(const KernelDefBuilder&) {
    OpKernel* (*factory)(OpKernelConstruction*) = [](OpKernelConstruction* c) -> OpKernel* { return new K(c); };
    (void)factory;
  }
};
#define HB_STUB_CAT_(a, b) a##b
#define HB_STUB_CAT(a, b) HB_STUB_CAT_(a, b)
#define REGISTER_KERNEL_BUILDER(kernel_builder, ...) \
  static ::tensorflow::KernelRegistrarStub<__VA_ARGS__> HB_STUB_CAT(hb_stub_registrar_, __COUNTER__)(kernel_builder)

#define OP_REQUIRES(CTX, EXP, STATUS)          \
  do {                                         \
    if (!(EXP)) { (CTX)->CtxFailure((STATUS)); return; } \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)               \
  do {                                         \
    ::tensorflow::Status _s(__VA_ARGS__);      \
    if (!_s.ok()) { (CTX)->CtxFailureWithWarning(_s); return; } \
  } while (0)
#define OP_REQUIRES_ASYNC(CTX, EXP, STATUS, CALLBACK) \
  do {                                         \
    if (!(EXP)) { (CTX)->CtxFailure((STATUS)); (CALLBACK)(); return; } \
  } while (0)
#define OP_REQUIRES_OK_ASYNC(CTX, STATUS, CALLBACK) \
  do {                                         \
    ::tensorflow::Status _s(STATUS);           \
    if (!_s.ok()) { (CTX)->CtxFailureWithWarning(_s); (CALLBACK)(); return; } \
  } while (0)

}  // namespace tensorflow
#endif  // HB_TF_SHIM_STUB_OP_KERNEL_H_
