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

namespace ppg { PassProfile g_pass_profile; }

extern "C" int ppg_profile_begin(void) {
  ppg::PassProfile& p = ppg::g_pass_profile;
  if (!p.created) {
    for (int i = 0; i < ppg::PassProfile::kCapacity; ++i) {
      PPG_CUDA_TRY(cudaEventCreate(&p.start[i]));
      PPG_CUDA_TRY(cudaEventCreate(&p.stop[i]));
    }
    p.created = true;
  }
  p.count = 0;
  p.enabled = true;
  return PPG_OK;
}

extern "C" int ppg_profile_end(float* h_ms, int64_t* h_items, int* h_bytes_per_item, int* h_kind, int capacity, int* h_count) {
  ppg::PassProfile& p = ppg::g_pass_profile;
  p.enabled = false;
  PPG_CUDA_TRY(cudaDeviceSynchronize());
  const int n = p.count < capacity ? p.count : capacity;
  for (int i = 0; i < n; ++i) {
    PPG_CUDA_TRY(cudaEventElapsedTime(&h_ms[i], p.start[i], p.stop[i]));
    h_items[i] = p.items[i];
    h_bytes_per_item[i] = p.bytes_per_item[i];
    if (h_kind != nullptr) h_kind[i] = p.kind[i];
  }
  *h_count = n;
  p.count = 0;
  return PPG_OK;
}

extern "C" int ppg_abi_version(void) { return PPG_ABI_VERSION; }
extern "C" const char* ppg_last_error(void) { return ppg::g_last_error; }
extern "C" unsigned long long ppg_launch_count(void) { return ppg::g_launch_count; }

#ifdef PPG_SORT_TRACE
namespace ppg { __device__ unsigned long long* g_sort_trace = nullptr; }
extern "C" int ppg_debug_set_sort_trace(unsigned long long* device_buffer) {
  return cudaMemcpyToSymbol(ppg::g_sort_trace, &device_buffer, sizeof(device_buffer)) == cudaSuccess ? 0 : 2;
}
#endif
