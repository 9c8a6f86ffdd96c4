// Placeholder entry points (implemented in a later milestone).
#include "common.cuh"
extern "C" {
int dgcn_genconv_aggregate_backward(const float*, const float*, int64_t, int64_t, int64_t, const int32_t*,
                                    const int32_t*, const int32_t*, const float*, const dgcn_genconv_params*, int32_t,
                                    const float*, float*, float*, float*, float*, dgcn_stream_t) {
  return DGCN_ERR_UNSUPPORTED;
}
}
