// Placeholder entry points (implemented in a later milestone).
#include "common.cuh"
extern "C" {
size_t dgcn_graph_conv_backward_workspace_bytes(int32_t, int64_t, int64_t, int64_t, int64_t, int64_t) { return 0; }
int dgcn_graph_conv_backward(int32_t, const float*, int64_t, int64_t, int64_t, int64_t, int64_t, const int32_t*,
                             int64_t, const dgcn_basic_conv*, int64_t, const float*, float*, float*, float*, float*,
                             float*, float*, void*, size_t, dgcn_stream_t) {
  return DGCN_ERR_UNSUPPORTED;
}
int dgcn_genconv_aggregate_backward(const float*, const float*, int64_t, int64_t, int64_t, const int32_t*,
                                    const int32_t*, const int32_t*, const float*, const dgcn_genconv_params*, int32_t,
                                    const float*, float*, float*, float*, float*, dgcn_stream_t) {
  return DGCN_ERR_UNSUPPORTED;
}
}
