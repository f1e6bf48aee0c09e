#include "pp_common.cuh"
namespace pp {
int plan_batch(pp_handle*, int, const pp_query*, pp_result*) { set_last_error("pp_plan: not built yet"); return PP_ERR_UNSUPPORTED; }
void plan_free(pp_handle*) {}
}
