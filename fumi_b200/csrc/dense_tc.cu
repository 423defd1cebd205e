// tcgen05 (5th-gen tensor core) 3xTF32 path of the dense layers -- filled in after the fp32 path is
// parity-green.  Until then precision=1 is refused loudly rather than silently
// falling back.
#include <cstdint>

#include "../../include/fumi_b200.h"
#include "common.cuh"

int fumi_linear_fwd_tc(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*) {
    fumi_set_error("fumi_linear_fwd: precision=1 (tcgen05 3xTF32) is not built yet");
    return FUMI_ERR_UNSUPPORTED;
}
int fumi_linear_wgrad_tc(const float*, const float*, float*, int64_t, int64_t, int64_t, int32_t, void*) {
    fumi_set_error("fumi_linear_wgrad: precision=1 (tcgen05 3xTF32) is not built yet");
    return FUMI_ERR_UNSUPPORTED;
}
