// ABI version and thread-local last-error text.
#include <stdarg.h>

#include "common.cuh"

namespace ppg {
static thread_local char g_last_error[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
}  // namespace ppg

extern "C" int ppg_abi_version(void) { return PPG_ABI_VERSION; }
extern "C" const char* ppg_last_error(void) { return ppg::g_last_error; }
extern "C" unsigned long long ppg_launch_count(void) { return ppg::g_launch_count; }

#ifdef PPG_SORT_TRACE
namespace ppg { __device__ unsigned long long* g_sort_trace = nullptr; }
extern "C" int ppg_debug_set_sort_trace(unsigned long long* device_buffer) {
  return cudaMemcpyToSymbol(ppg::g_sort_trace, &device_buffer, sizeof(device_buffer)) == cudaSuccess ? 0 : 2;
}
#endif
