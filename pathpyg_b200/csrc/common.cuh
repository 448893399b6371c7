// Shared device/host helpers for the pathpyg_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pathpyg_b200.h"

namespace ppg {

constexpr int kNumSMsB200 = 148;
constexpr unsigned kFullMask = 0xffffffffu;

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

#define PPG_CUDA_TRY(expr)                                                               \
  do {                                                                                   \
    cudaError_t ppg_err_ = (expr);                                                       \
    if (ppg_err_ != cudaSuccess) {                                                       \
      ::ppg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                     \
                       cudaGetErrorString(ppg_err_));                                    \
      return PPG_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// after every kernel launch: surface launch errors and count the launch (bench.py reports the count)
extern unsigned long long g_launch_count;
#define PPG_LAUNCHED()                                                                   \
  do {                                                                                   \
    PPG_CUDA_TRY(cudaGetLastError());                                                    \
    ++::ppg::g_launch_count;                                                             \
  } while (0)

// opt-in timing of the hot kernels (radix digit passes, chain tiles, owner merges, fused GCN layers) with CUDA events on the launching stream (ppg_profile_begin / _end):
// the benchmark reads the durations of the passes that ran INSIDE a real step instead of probing a synthetic sort
struct PassProfile {
  static constexpr int kCapacity = 512;
  bool enabled = false;
  int count = 0;
  cudaEvent_t start[kCapacity], stop[kCapacity];
  long long items[kCapacity];
  int bytes_per_item[kCapacity];
  int kind[kCapacity];
  bool created = false;
};
extern PassProfile g_pass_profile;
inline void profile_pass_begin(cudaStream_t stream) {
  PassProfile& p = g_pass_profile;
  if (p.enabled && p.count < PassProfile::kCapacity) cudaEventRecord(p.start[p.count], stream);
}
inline void profile_pass_end(cudaStream_t stream, long long items, int bytes_per_item, int kind = PPG_PROFILE_DIGIT_PASS) {
  PassProfile& p = g_pass_profile;
  if (p.enabled && p.count < PassProfile::kCapacity) {
    cudaEventRecord(p.stop[p.count], stream);
    p.items[p.count] = items;
    p.bytes_per_item[p.count] = bytes_per_item;
    p.kind[p.count] = kind;
    ++p.count;
  }
}

#define PPG_REQUIRE(cond, code, ...)                                                     \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::ppg::set_error(__VA_ARGS__);                                                     \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)

#define PPG_TRY(expr)                                                                    \
  do {                                                                                   \
    int ppg_rc_ = (expr);                                                                \
    if (ppg_rc_ != PPG_OK) return ppg_rc_;                                               \
  } while (0)

// ---------------------------------------------------------------- misc host/device
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// number of bits needed to represent values in [0, max_value]
__host__ __device__ inline int bits_for(uint64_t max_value) {
  int b = 0;
  while (max_value) { ++b; max_value >>= 1; }
  return b;
}

// Bump allocator over a caller-provided workspace. With base == nullptr it only measures.
struct Workspace {
  char* base;
  size_t capacity;
  size_t used = 0;
  Workspace(void* b, size_t cap) : base(static_cast<char*>(b)), capacity(cap) {}
  template <typename T>
  T* take(size_t count) {
    used = align_up(used, 256);
    T* p = base ? reinterpret_cast<T*>(base + used) : nullptr;
    used += count * sizeof(T);
    return p;
  }
  bool fits() const { return base == nullptr || used <= capacity; }
};

// ---------------------------------------------------------------- decoupled look-back state words
// One 64-bit word carries an 8-bit code and a 56-bit value, so flag and value are published by a
// single relaxed store (no fence needed). Codes are unique per (call, pass): a word written by an
// earlier pass never looks valid to a later one, which lets one memset serve a multi-pass sort.
constexpr unsigned long long kStateValueMask = (1ull << 56) - 1;

__device__ __forceinline__ void state_store(unsigned long long* p, unsigned code, unsigned long long value) {
  unsigned long long w = (static_cast<unsigned long long>(code) << 56) | (value & kStateValueMask);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long state_load(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

// 32-bit variant for per-tile digit counts (<= 4096): code << 24 | count
__device__ __forceinline__ void tile_state_store(uint32_t* p, unsigned code, uint32_t value) {
  const uint32_t w = (code << 24) | (value & 0xffffffu);
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t tile_state_load(const uint32_t* p) {
  uint32_t w;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(p) : "memory");
  return w;
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T x) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T y = __shfl_up_sync(kFullMask, x, d);
    if (lane_id() >= static_cast<unsigned>(d)) x += y;
  }
  return x;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(kFullMask, x, d);
  return x;
}

// streaming (read-once / write-once) accesses: keep them out of L1
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v) { __stcs(p, v); }

// host copy of a device scalar through the stream (the one read-back of a count->allocate->fill op)
template <typename T>
inline int read_back(T* host_dst, const T* dev_src, cudaStream_t stream) {
  PPG_CUDA_TRY(cudaMemcpyAsync(host_dst, dev_src, sizeof(T), cudaMemcpyDeviceToHost, stream));
  PPG_CUDA_TRY(cudaStreamSynchronize(stream));
  return PPG_OK;
}

inline int grid_for(int64_t work_items, int per_block, int max_blocks = kNumSMsB200 * 16) {
  int64_t g = ceil_div(work_items, per_block);
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

// development aid: per-tile phase time stamps (globaltimer, ns) of the onesweep digit passes and of the fused
// GCN layer; enabled with -DPPG_SORT_TRACE (make trace), read with scripts/sort_trace.py / scripts/gcn_trace.py
#ifdef PPG_SORT_TRACE
extern __device__ unsigned long long* g_sort_trace;  // [tiles][8]
__device__ __forceinline__ void sort_trace(unsigned tile, int slot) {
  if (threadIdx.x == 0 && g_sort_trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_sort_trace[static_cast<size_t>(tile) * 8 + slot] = t;
  }
}
__device__ __forceinline__ void sort_trace_if(bool who, unsigned tile, int slot) {
  if (who && g_sort_trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_sort_trace[static_cast<size_t>(tile) * 8 + slot] = t;
  }
}
#define PPG_TRACE(tile, slot) sort_trace(tile, slot)
#define PPG_TRACE_IF(who, tile, slot) sort_trace_if(who, tile, slot)
#else
#define PPG_TRACE(tile, slot)
#define PPG_TRACE_IF(who, tile, slot)
#endif

}  // namespace ppg
