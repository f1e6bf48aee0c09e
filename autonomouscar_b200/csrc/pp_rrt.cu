#include "pp_common.cuh"
namespace pp {
int rrt_plan_batch(pp_handle*, int, const pp_query*, pp_result*) { set_last_error("pp_rrt_plan: not built yet"); return PP_ERR_UNSUPPORTED; }
void rrt_free(pp_handle*) {}
}
