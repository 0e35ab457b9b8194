// ABI bookkeeping: version, thread-local error string, launch counter.
#include "common.cuh"

namespace dagnn {

std::atomic<int64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace dagnn

extern "C" {

int dagnn_abi_version(void) { return DAGNN_ABI_VERSION; }
const char* dagnn_last_error(void) { return dagnn::err_buf(); }
int64_t dagnn_launch_count(void) { return dagnn::g_launches.load(); }

}  // extern "C"
