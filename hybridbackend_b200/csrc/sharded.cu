// sharded.cu -- placeholder while the fused sharded path is being written.
#include "comm.cuh"
struct hbShardedPlan { int unused; };
extern "C" {
int hbShardedPlanCreate(hbComm*, int, const int64_t*, const int32_t*, double, hbShardedPlan**) {
  hb::set_last_error("sharded plan: not built yet"); return HB_ERR_INVALID; }
int hbShardedPlanDestroy(hbShardedPlan*) { return HB_OK; }
size_t hbShardedPlanWindowBytes(int, int, const int64_t*, const int32_t*, double) { return 0; }
int hbShardedLookupForward(hbShardedPlan*, const hbShardedFeature*, int32_t*, hbStream) {
  hb::set_last_error("sharded plan: not built yet"); return HB_ERR_INVALID; }
int hbShardedLookupBackwardUpdate(hbShardedPlan*, const hbShardedFeature*, const hbOptimizer*, int32_t*, hbStream) {
  hb::set_last_error("sharded plan: not built yet"); return HB_ERR_INVALID; }
}
