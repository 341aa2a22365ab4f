#include "tensorflow/core/framework/op_kernel.h"
#include <mutex>
namespace tensorflow {
typedef std::mutex mutex;
typedef std::lock_guard<std::mutex> mutex_lock;
}
