// smj_dist.cu -- multi-GPU key-range partitioned path (placeholder until the exchange step lands).
#include "smj_internal.h"

bool smj_dist_active(void) { return false; }
int smj_dist_shutdown(void) { return SMJ_OK; }

int smj_run_multi(const smj_config_t *, const smj_table_t *, const smj_table_t *, smj_table_t *, smj_stats_t *)
{
    return smj_set_error(SMJ_EINVAL, "multi-GPU smj_run is not built into this library yet");
}
extern "C" int smj_dist_unique_id(void *) { return smj_set_error(SMJ_ENCCL, "multi-GPU support not built"); }
extern "C" int smj_init_dist(const smj_config_t *, int, int, int, const void *)
{
    return smj_set_error(SMJ_ENCCL, "multi-GPU support not built");
}
