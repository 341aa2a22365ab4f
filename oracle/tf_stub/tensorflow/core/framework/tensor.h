// Minimal stand-in for the four TensorFlow headers that the reference's CPU
// partition functors include (partition_by_modulo_functors.cc:22-25).  It lets
// oracle/build_ref.sh compile those reference sources UNMODIFIED, from where
// they lie under /root/reference, without TensorFlow.  Test infrastructure only.
#ifndef HB_ORACLE_TF_STUB_TENSOR_H_
#define HB_ORACLE_TF_STUB_TENSOR_H_
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace Eigen {
struct ThreadPoolDevice {};
struct GpuDevice {};
}  // namespace Eigen

namespace tensorflow {
typedef int32_t int32;
typedef long long int64;
typedef uint32_t uint32;
typedef unsigned long long uint64;
using std::string;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& m) : ok_(false), msg_(m) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }
 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
inline Status Unimplemented(const char* m) { return Status(m); }
inline Status InvalidArgument(const char* m) { return Status(m); }
}  // namespace errors

// A Tensor here is a non-owning (pointer, element count) view.
class Tensor {
 public:
  Tensor() : data_(nullptr), n_(0) {}
  Tensor(void* data, int64 n) : data_(data), n_(n) {}
  int64 NumElements() const { return n_; }
  template <typename T>
  struct Flat {
    T* p;
    T* data() const { return p; }
  };
  template <typename T>
  Flat<T> flat() { return Flat<T>{static_cast<T*>(data_)}; }
  template <typename T>
  Flat<const T> flat() const { return Flat<const T>{static_cast<const T*>(data_)}; }
 private:
  void* data_;
  int64 n_;
};

class OpKernelContext {
 public:
  Status status;
};

#define OP_REQUIRES_OK(CTX, ...)            \
  do {                                      \
    ::tensorflow::Status _s(__VA_ARGS__);   \
    if (!_s.ok()) { (CTX)->status = _s; return; } \
  } while (0)

}  // namespace tensorflow
#endif  // HB_ORACLE_TF_STUB_TENSOR_H_
