// C-ABI glue of libagenda_b200.so: error reporting, device probe, and the cross-attention dispatcher.
#include "common.cuh"

namespace agenda {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

struct TokenList;
int attn_common_checks(const char* who, const void* q, const void* k, const void* v, void* out, int dtype, int B,
                       int H, int N, int M, int d);

}  // namespace agenda

using namespace agenda;

extern "C" int agenda_version(void) { return 1000; }

extern "C" const char* agenda_last_error(void) { return last_error_buf(); }

extern "C" int agenda_device_ok(void) {
  int dev = 0, major = 0;
  AGENDA_CUDA(cudaGetDevice(&dev));
  AGENDA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}
