// api.cu -- error reporting, version and launch accounting of libdbb200.so
#include "common.cuh"
#include <string.h>

namespace dbb {
thread_local char g_last_error[512] = "";
uint64_t g_launch_count = 0;

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return DBB_ECUDA;
}
int set_error(int code, const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}
}  // namespace dbb

extern "C" int dbb_version(void) { return DBB_VERSION; }

extern "C" const char* dbb_strerror(int code) {
  switch (code) {
    case DBB_OK: return "ok";
    case DBB_EINVAL: return "invalid argument or shape";
    case DBB_EALIGN: return "pointer not 16-byte aligned";
    case DBB_EWORKSPACE: return "workspace too small";
    case DBB_ECUDA: return "CUDA error";
    case DBB_EUNSUPPORTED: return "configuration not supported";
    default: return "unknown error";
  }
}
extern "C" const char* dbb_last_cuda_error(void) { return dbb::g_last_error; }
extern "C" uint64_t dbb_launch_count(void) { return dbb::g_launch_count; }
