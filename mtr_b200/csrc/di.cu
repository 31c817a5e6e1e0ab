// di.cu -- K1/K2: directional index (placeholder until the kernels land; fails loudly, never falls back).
#include "mtr_internal.h"
void di_state_free(mtr_ctx *) {}
extern "C" int mtr_di_run(mtr_ctx *ctx, int, const uint16_t *, const int64_t *, const int64_t *, double *, int32_t *, int32_t *)
{
    if (!ctx) return MTR_EINVAL;
    mtr_set_error(ctx, "mtr_di_run: not built yet");
    return MTR_EINVAL;
}
