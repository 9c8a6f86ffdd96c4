// Status strings, version and per-thread CUDA error text for the dgcn C ABI.
#include <stdio.h>
#include "common.cuh"

namespace dgcn {
static thread_local char g_last_error[512] = "";
void set_last_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file,
           line);
}
}  // namespace dgcn

extern "C" {
int dgcn_version(void) { return 100; }
const char* dgcn_status_string(int status) {
  switch (status) {
    case DGCN_OK: return "ok";
    case DGCN_ERR_BAD_ARG: return "bad argument (null pointer or inconsistent size)";
    case DGCN_ERR_UNSUPPORTED: return "request outside what the sm_100a kernels cover";
    case DGCN_ERR_WORKSPACE: return "workspace too small";
    case DGCN_ERR_CUDA: return "CUDA launch failed";
    default: return "unknown status";
  }
}
const char* dgcn_last_cuda_error(void) { return dgcn::g_last_error; }
}
