// Consumers of the built De Bruijn layers: weighted degrees and transition
// probabilities (core/graph.py:486-533), walk counts for the degrees of freedom
// (core/multi_order_model.py:243-312) and the log-likelihood sums (:314-409).
//
// All of it is streaming / segment work over arrays the lift already left on the device:
//   sorted_ids_ptr   : CSR pointer of a SORTED id column by one binary search per node (no atomics, no scan)
//   segment_sum      : per-segment fp32 sum in slot order (== the order of a sequential scatter_add);
//                      segments longer than kSerialRun go to a list and are reduced by one CTA each with a
//                      fixed-shape fp64 tree (deterministic, no floating-point atomics)
//   edge_ratio       : out[e] = w[e] / denom[ids[e]]
//   walk_step        : c_k[row e] += c_{k-1}[col e] (u64 integer atomics: order-independent), with the grand
//                      total and the number of non-zero rows reduced per CTA
//   weighted_log_sum : sum_i f[i] * logf(p[j(i)]) -- terms in fp32 as the reference forms them, accumulated in
//                      fp64 through per-CTA partials + one final CTA (fixed order)
#include "common.cuh"

namespace ppg {

constexpr int kSerialRun = 64;
constexpr int kSelBlock = 256;

// ------------------------------------------------------------------ CSR pointer of a sorted id column
__global__ void __launch_bounds__(kSelBlock)
sorted_ids_ptr_kernel(const int64_t* __restrict__ ids, int64_t E, int64_t n, int32_t* __restrict__ ptr) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v <= n; v += stride) {
    int64_t lo = 0, hi = E;  // first slot with ids[slot] >= v
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(ids + mid) < v) lo = mid + 1; else hi = mid;
    }
    ptr[v] = static_cast<int32_t>(lo);
  }
}

// ------------------------------------------------------------------ segment sums
struct LongRun {
  int32_t node, begin, end, pad;
};

__global__ void __launch_bounds__(kSelBlock)
segment_sum_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm, const float* __restrict__ w,
                   int64_t n, float* __restrict__ out, unsigned* __restrict__ long_count, LongRun* __restrict__ long_runs) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < n; v += stride) {
    const int32_t a = ptr[v], b = ptr[v + 1];
    if (w == nullptr) {
      out[v] = static_cast<float>(b - a);
    } else if (b - a <= kSerialRun) {
      float s = 0.f;
      for (int32_t i = a; i < b; ++i) s += w[perm ? perm[i] : i];
      out[v] = s;
    } else {
      const unsigned slot = atomicAdd(long_count, 1u);
      long_runs[slot] = LongRun{static_cast<int32_t>(v), a, b, 0};
    }
  }
}

__device__ __forceinline__ double block_sum_f64(double x, double* s_part) {
  // fixed-shape tree: xor-shuffle inside a warp, then warp 0 over the warp partials
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(kFullMask, x, d);
  if (lane_id() == 0) s_part[threadIdx.x >> 5] = x;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(kFullMask, t, d);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(kSelBlock)
segment_sum_long_kernel(const int32_t* __restrict__ perm, const float* __restrict__ w, const unsigned* __restrict__ long_count,
                        const LongRun* __restrict__ long_runs, float* __restrict__ out) {
  __shared__ double s_part[kSelBlock / 32];
  const unsigned count = *long_count;
  for (unsigned r = blockIdx.x; r < count; r += gridDim.x) {
    const LongRun run = long_runs[r];
    double acc = 0.0;
    for (int32_t i = run.begin + threadIdx.x; i < run.end; i += kSelBlock) acc += static_cast<double>(w[perm ? perm[i] : i]);
    const double total = block_sum_f64(acc, s_part);
    if (threadIdx.x == 0) out[run.node] = static_cast<float>(total);
  }
}

__global__ void __launch_bounds__(kSelBlock)
edge_ratio_kernel(const int64_t* __restrict__ ids, const float* __restrict__ w, const float* __restrict__ denom, int64_t E,
                  float* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < E; e += stride) {
    const float num = w ? ld_stream(w + e) : 1.f;
    st_stream(out + e, num / __ldg(denom + ld_stream(ids + e)));
  }
}

// ------------------------------------------------------------------ walk counts
__global__ void __launch_bounds__(kSelBlock)
fill_u64_kernel(unsigned long long* __restrict__ p, int64_t n, unsigned long long v) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// next[row e] += prev[col e]; totals[0] += sum over edges (== number of walks one step longer)
__global__ void __launch_bounds__(kSelBlock)
walk_step_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t E, int64_t n,
                 const unsigned long long* __restrict__ prev, unsigned long long* __restrict__ next,
                 unsigned long long* __restrict__ total, unsigned long long* __restrict__ status) {
  __shared__ unsigned long long s_part[kSelBlock / 32];
  unsigned long long acc = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t r = ld_stream(row + e), c = ld_stream(col + e);
    if (r < 0 || r >= n || c < 0 || c >= n) {
      atomicOr(status, 1ull);
      continue;
    }
    const unsigned long long x = prev[c];
    if (x) {
      atomicAdd(next + r, x);
      acc += x;
    }
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kSelBlock / 32; ++w) t += s_part[w];
    if (t) atomicAdd(total, t);
  }
}

// sources[0] += #{v : cur[v] > 0}; clears `clear` (the buffer the next step accumulates into)
__global__ void __launch_bounds__(kSelBlock)
walk_sources_kernel(const unsigned long long* __restrict__ cur, unsigned long long* __restrict__ clear, int64_t n,
                    unsigned long long* __restrict__ sources) {
  __shared__ unsigned s_part[kSelBlock / 32];
  unsigned acc = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < n; v += stride) {
    acc += cur[v] != 0;
    clear[v] = 0;
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < kSelBlock / 32; ++w) t += s_part[w];
    if (t) atomicAdd(sources, static_cast<unsigned long long>(t));
  }
}

// ------------------------------------------------------------------ sum_i f[i] * log(p[j(i)])
__global__ void __launch_bounds__(kSelBlock)
weighted_log_partial_kernel(const float* __restrict__ freq, const float* __restrict__ prob, const int64_t* __restrict__ idx,
                            const int64_t* __restrict__ idx2, int64_t n, int64_t prob_len, int64_t idx2_len,
                            double* __restrict__ partial, unsigned long long* __restrict__ status) {
  __shared__ double s_part[kSelBlock / 32];
  double acc = 0.0;
  // contiguous chunk per CTA, strided inside: the summation tree depends on (n, grid) only
  const int64_t per = ceil_div(n, gridDim.x);
  const int64_t lo = per * blockIdx.x;
  const int64_t hi = lo + per < n ? lo + per : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kSelBlock) {
    int64_t j = i;
    if (idx) j = idx[j];
    if (idx2) {
      if (j < 0 || j >= idx2_len) { atomicOr(status, 1ull); continue; }
      j = idx2[j];
    }
    if (j < 0 || j >= prob_len) { atomicOr(status, 1ull); continue; }
    const float term = freq[i] * logf(prob[j]);  // fp32 product of fp32 factors, as torch.mul(frequencies, torch.log(.))
    acc += static_cast<double>(term);
  }
  const double t = block_sum_f64(acc, s_part);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kSelBlock)
sum_partials_kernel(const double* __restrict__ partial, int count, double* __restrict__ out) {
  __shared__ double s_part[kSelBlock / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += kSelBlock) acc += partial[i];
  const double t = block_sum_f64(acc, s_part);
  if (threadIdx.x == 0) *out = t;
}

constexpr int kLogSumGrid = kNumSMsB200 * 4;

}  // namespace ppg

using namespace ppg;

extern "C" int ppg_sorted_ids_ptr(const int64_t* sorted_ids, int64_t num_ids, int64_t num_nodes, int32_t* out_ptr,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(num_ids >= 0 && num_ids < (1ll << 31) && num_nodes >= 0 && num_nodes < (1ll << 31) - 1, PPG_ERR_INVALID,
              "sorted_ids_ptr: sizes (%lld ids, %lld nodes) exceed the 2^31 limit", (long long)num_ids, (long long)num_nodes);
  sorted_ids_ptr_kernel<<<grid_for(num_nodes + 1, kSelBlock), kSelBlock, 0, stream>>>(sorted_ids, num_ids, num_nodes, out_ptr);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" size_t ppg_segment_sum_workspace_bytes(int64_t num_slots) {
  Workspace ws(nullptr, 0);
  ws.take<unsigned long long>(1);
  ws.take<LongRun>(static_cast<size_t>(num_slots / kSerialRun + 1));
  return ws.used + 256;
}

extern "C" int ppg_segment_sum(const int32_t* ptr, const int32_t* perm, const float* weights, int64_t num_segments,
                               int64_t num_slots, void* workspace, size_t workspace_bytes, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_segments == 0) return PPG_OK;
  Workspace ws(workspace, workspace_bytes);
  unsigned* long_count = reinterpret_cast<unsigned*>(ws.take<unsigned long long>(1));
  LongRun* long_runs = ws.take<LongRun>(static_cast<size_t>(num_slots / kSerialRun + 1));
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "segment_sum: workspace of %zu bytes is too small (%zu needed)", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(long_count, 0, sizeof(unsigned long long), stream));
  segment_sum_kernel<<<grid_for(num_segments, kSelBlock), kSelBlock, 0, stream>>>(ptr, perm, weights, num_segments, out,
                                                                                   long_count, long_runs);
  PPG_LAUNCHED();
  if (weights != nullptr && num_slots > kSerialRun) {
    segment_sum_long_kernel<<<kNumSMsB200 * 2, kSelBlock, 0, stream>>>(perm, weights, long_count, long_runs, out);
    PPG_LAUNCHED();
  }
  return PPG_OK;
}

extern "C" int ppg_edge_ratio(const int64_t* ids, const float* weights, const float* denom, int64_t num_edges, float* out,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_edges == 0) return PPG_OK;
  edge_ratio_kernel<<<grid_for(num_edges, kSelBlock * 4), kSelBlock, 0, stream>>>(ids, weights, denom, num_edges, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" size_t ppg_walk_counts_workspace_bytes(int64_t num_nodes, int max_len) {
  Workspace ws(nullptr, 0);
  ws.take<unsigned long long>(2 * static_cast<size_t>(max_len > 0 ? max_len : 1) + 1);
  ws.take<unsigned long long>(static_cast<size_t>(num_nodes));
  ws.take<unsigned long long>(static_cast<size_t>(num_nodes));
  return ws.used + 256;
}

extern "C" int ppg_walk_counts(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int max_len, void* workspace,
                               size_t workspace_bytes, int64_t* h_num_walks, int64_t* h_num_sources, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(max_len >= 1 && max_len <= 64, PPG_ERR_INVALID, "walk_counts: max_len %d outside [1, 64]", max_len);
  Workspace ws(workspace, workspace_bytes);
  unsigned long long* words = ws.take<unsigned long long>(2 * static_cast<size_t>(max_len) + 1);  // totals, sources, status
  unsigned long long* a = ws.take<unsigned long long>(static_cast<size_t>(num_nodes));
  unsigned long long* b = ws.take<unsigned long long>(static_cast<size_t>(num_nodes));
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "walk_counts: workspace of %zu bytes is too small (%zu needed)", workspace_bytes, ws.used);
  unsigned long long* totals = words;
  unsigned long long* sources = words + max_len;
  unsigned long long* status = words + 2 * max_len;
  PPG_CUDA_TRY(cudaMemsetAsync(words, 0, (2 * static_cast<size_t>(max_len) + 1) * sizeof(unsigned long long), stream));
  if (num_nodes > 0 && num_edges > 0) {
    const int gn = grid_for(num_nodes, kSelBlock * 4);
    const int ge = grid_for(num_edges, kSelBlock * 4);
    fill_u64_kernel<<<gn, kSelBlock, 0, stream>>>(a, num_nodes, 1ull);
    PPG_LAUNCHED();
    PPG_CUDA_TRY(cudaMemsetAsync(b, 0, static_cast<size_t>(num_nodes) * sizeof(unsigned long long), stream));
    for (int k = 0; k < max_len; ++k) {
      walk_step_kernel<<<ge, kSelBlock, 0, stream>>>(edge_index, edge_index + num_edges, num_edges, num_nodes, a, b,
                                                     totals + k, status);
      PPG_LAUNCHED();
      walk_sources_kernel<<<gn, kSelBlock, 0, stream>>>(b, a, num_nodes, sources + k);
      PPG_LAUNCHED();
      unsigned long long* t = a; a = b; b = t;
    }
  }
  unsigned long long h[2 * 64 + 1];
  PPG_CUDA_TRY(cudaMemcpyAsync(h, words, (2 * static_cast<size_t>(max_len) + 1) * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, stream));
  PPG_CUDA_TRY(cudaStreamSynchronize(stream));
  PPG_REQUIRE(h[2 * max_len] == 0, PPG_ERR_INVALID, "walk_counts: node id outside [0, %lld)", (long long)num_nodes);
  for (int k = 0; k < max_len; ++k) {
    h_num_walks[k] = static_cast<int64_t>(h[k]);
    h_num_sources[k] = static_cast<int64_t>(h[max_len + k]);
  }
  return PPG_OK;
}

extern "C" size_t ppg_weighted_log_sum_workspace_bytes(void) {
  return (static_cast<size_t>(kLogSumGrid) + 2) * sizeof(double) + 512;
}

extern "C" int ppg_weighted_log_sum(const float* freq, const float* prob, const int64_t* idx, const int64_t* idx2, int64_t n,
                                    int64_t prob_len, int64_t idx2_len, void* workspace, size_t workspace_bytes,
                                    double* h_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  *h_out = 0.0;
  if (n == 0) return PPG_OK;
  Workspace ws(workspace, workspace_bytes);
  unsigned long long* status = ws.take<unsigned long long>(1);
  double* result = ws.take<double>(1);
  double* partial = ws.take<double>(kLogSumGrid);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "weighted_log_sum: workspace of %zu bytes is too small (%zu needed)", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(unsigned long long), stream));
  const int grid = grid_for(n, kSelBlock * 4, kLogSumGrid);
  weighted_log_partial_kernel<<<grid, kSelBlock, 0, stream>>>(freq, prob, idx, idx2, n, prob_len, idx2_len, partial, status);
  PPG_LAUNCHED();
  sum_partials_kernel<<<1, kSelBlock, 0, stream>>>(partial, grid, result);
  PPG_LAUNCHED();
  unsigned long long h_status = 0;
  PPG_CUDA_TRY(cudaMemcpyAsync(h_out, result, sizeof(double), cudaMemcpyDeviceToHost, stream));
  PPG_CUDA_TRY(cudaMemcpyAsync(&h_status, status, sizeof(h_status), cudaMemcpyDeviceToHost, stream));
  PPG_CUDA_TRY(cudaStreamSynchronize(stream));
  PPG_REQUIRE(h_status == 0, PPG_ERR_INVALID, "weighted_log_sum: index out of range");
  return PPG_OK;
}
