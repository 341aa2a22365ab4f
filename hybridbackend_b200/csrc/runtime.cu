// runtime.cu -- error plumbing, build info, device queries for libhb_b200.so.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace hb {

static thread_local char g_last_error[1024] = "";

const char* get_last_error() { return g_last_error; }

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_profile{0};
struct ProfRec { int id; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec*> g_recs;

KernelScope::KernelScope(int id_, cudaStream_t s) : id(id_), stream(s), rec(nullptr) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_profile.load(std::memory_order_relaxed)) {
    ProfRec* r = new ProfRec();
    r->id = id_;
    if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
    cudaEventRecord(r->a, s);
    rec = r;
  }
}

KernelScope::~KernelScope() {
  if (rec != nullptr) {
    ProfRec* r = reinterpret_cast<ProfRec*>(rec);
    cudaEventRecord(r->b, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_recs.push_back(r);
  }
}

static const char* kKernelNames[HB_K_COUNT] = {
    "", "partition_hist", "partition_pass", "", "sort_hist", "sort_pass",
    "", "bag_of_position", "lookup_fwd", "sparse_update", "sparse_update_fixup",
    "cast_n", "cache_lookup", "barrier", "a2a_sizes", "a2a_tables", "a2a_push", "a2a_copyout",
    "sharded_exchange", "sharded_push_ids", "sharded_owner_gather", "sharded_stitch_pool",
    "sharded_push_grads", "sharded_pad", "allreduce_push", "allreduce_reduce", "update_runs",
    "sparse_update_long", "sharded_publish", "sharded_unique", "h2d_stage", ""};

}  // namespace hb

extern "C" {
int64_t hbGetLaunchCount(void) { return hb::g_launches.load(); }
int hbProfileEnable(int on) { hb::g_profile.store(on ? 1 : 0); return HB_OK; }
int hbProfileReset(void) {
  std::lock_guard<std::mutex> lk(hb::g_prof_mu);
  for (auto* r : hb::g_recs) { cudaEventDestroy(r->a); cudaEventDestroy(r->b); delete r; }
  hb::g_recs.clear();
  return HB_OK;
}
int hbProfileGet(int kernel_id, double* total_ms, int64_t* launches) {
  std::lock_guard<std::mutex> lk(hb::g_prof_mu);
  double t = 0;
  int64_t n = 0;
  for (auto* r : hb::g_recs) {
    if (r->id != kernel_id) continue;
    if (cudaEventSynchronize(r->b) != cudaSuccess) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) { t += ms; ++n; }
  }
  if (total_ms) *total_ms = t;
  if (launches) *launches = n;
  return HB_OK;
}
const char* hbKernelName(int id) { return (id > 0 && id < HB_K_COUNT) ? hb::kKernelNames[id] : ""; }
const char* hbGetLastErrorString(void) { return hb::g_last_error; }
int hbGetVersion(void) { return 100; }
const char* hbGetBuildInfo(void) { return "hb_b200 0.1.0; CUDA sm_100a; " __DATE__; }
}
