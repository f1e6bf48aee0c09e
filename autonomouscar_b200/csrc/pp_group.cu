#include "pp_common.cuh"
extern "C" {
int pp_group_create(const pp_params*, int, const int*, pp_group**) { pp::set_last_error("pp_group: not built yet"); return PP_ERR_UNSUPPORTED; }
int pp_group_destroy(pp_group*) { return PP_OK; }
int pp_group_size(const pp_group*, int*) { return PP_ERR_UNSUPPORTED; }
int pp_group_set_grid(pp_group*, const int8_t*, int32_t, int32_t, double, int) { return PP_ERR_UNSUPPORTED; }
int pp_group_collide_batch(pp_group*, int64_t, const float*, uint8_t*) { return PP_ERR_UNSUPPORTED; }
int pp_group_plan_batch(pp_group*, int32_t, const pp_query*, pp_result*) { return PP_ERR_UNSUPPORTED; }
}
