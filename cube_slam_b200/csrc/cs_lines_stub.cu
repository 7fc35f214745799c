/* placeholder until the line detectors land: keeps the ABI complete */
#include "cs_internal.h"
extern "C" {
int cs_detect_lines(cs_ctx *, const uint8_t *, int, int, int, int, const cs_line_params *, float *, int32_t *) { return CS_ERR_UNSUPPORTED; }
int cs_detect_lines_batch(cs_ctx *, const uint8_t *, int, int, int, int, int, const cs_line_params *, float *, int32_t, int32_t *)
{
    return CS_ERR_UNSUPPORTED;
}
}
