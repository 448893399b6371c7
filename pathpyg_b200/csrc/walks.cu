// Per-walk bookkeeping of a7 / a9 (MultiOrderModel.from_path_data, the log-likelihoods, PathData.append_walks):
// what the reference writes as torch.cumsum / repeat_interleave / bincount / arange + mask
// (core/multi_order_model.py:217-224,335,354-361,402-405, core/path_data.py:139-159), as one small kernel each.
//
//   counts -> offsets        single-pass exclusive scan (scan.cuh), int64 in, int64 out, {total, status} in the workspace
//   expand by offsets        out[j] = values[i] for offsets[i] <= j < offsets[i + 1]   (repeat_interleave)
//   walk chain               the edges p -> p + 1 of all walks laid end to end, links between walks dropped
//   bincount                 occurrence counts of int64 ids (integer atomics: exact, order-independent)
#include "scan.cuh"

namespace ppg {

namespace {

constexpr unsigned kStatusBadCount = 1u;   // a count was negative / an id was out of range (bit 0, as everywhere)

struct ResultWords {
  unsigned long long total;
  unsigned long long status;
};

struct CountProducer64 {
  const int64_t* counts;
  unsigned long long* status;
  int64_t min_count;
  __device__ unsigned long long operator()(int64_t i) const {
    const int64_t c = counts[i];
    if (c < min_count) {
      atomicOr(reinterpret_cast<unsigned*>(status), kStatusBadCount);
      return 0ull;
    }
    return static_cast<unsigned long long>(c);
  }
};
struct OffsetConsumer64 {
  int64_t* off;
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long v, unsigned long long prefix) const {
    off[i] = static_cast<int64_t>(prefix);
    if (i == n - 1) off[n] = static_cast<int64_t>(prefix + v);
  }
};

// index of the segment that holds position j: the last i with offsets[i] <= j (empty segments are skipped)
__device__ __forceinline__ int64_t segment_of(const int64_t* __restrict__ offsets, int64_t n, int64_t j) {
  int64_t lo = 0, hi = n;   // invariant: offsets[lo] <= j < offsets[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= j) lo = mid; else hi = mid;
  }
  return lo;
}

template <typename T>
__global__ void __launch_bounds__(256)
expand_offsets_kernel(const int64_t* __restrict__ offsets, int64_t n, int64_t total, const T* __restrict__ values,
                      T* __restrict__ out, int64_t* __restrict__ owner) {
  for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < total; j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = segment_of(offsets, n, j);
    if (out != nullptr) st_stream(out + j, values[i]);
    if (owner != nullptr) st_stream(owner + j, i);
  }
}

// position p of walk w is the source of an edge unless it is the walk's last position; the edges of walk w start at
// offsets[w] - w (every walk has at least one position)
__global__ void __launch_bounds__(256)
walk_chain_kernel(const int64_t* __restrict__ offsets, int64_t num_walks, int64_t total, int64_t base,
                  int64_t* __restrict__ edge_index, int64_t num_edges) {
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < total; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t w = segment_of(offsets, num_walks, p);
    if (p + 1 < offsets[w + 1]) {
      const int64_t k = p - w;
      st_stream(edge_index + k, base + p);
      st_stream(edge_index + num_edges + k, base + p + 1);
    }
  }
}

__global__ void __launch_bounds__(256)
bincount_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t num_bins, unsigned long long* __restrict__ counts,
                unsigned* __restrict__ status) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t v = ld_stream(ids + i);
    if (v < 0 || v >= num_bins) {
      atomicOr(status, kStatusBadCount);
      continue;
    }
    atomicAdd(counts + v, 1ull);
  }
}

}  // namespace
}  // namespace ppg

using namespace ppg;

extern "C" size_t ppg_counts_to_offsets_workspace_bytes(int64_t n) {
  Workspace ws(nullptr, 0);
  ws.take<ResultWords>(1);
  ws.take<unsigned long long>(scan_state_words(n));
  return align_up(ws.used, 256);
}

// offsets [n + 1] = exclusive prefix sums of counts [n] (offsets[n] = their sum, also left in the workspace's {total,
// status} words: ppg_result_read).  Counts below `min_count` set status bit 0 and count as 0.
extern "C" int ppg_counts_to_offsets(const int64_t* counts, int64_t n, int64_t min_count, void* workspace, size_t workspace_bytes,
                                     int64_t* offsets, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(n >= 0, PPG_ERR_INVALID, "counts_to_offsets: negative length");
  Workspace ws(workspace, workspace_bytes);
  ResultWords* result = ws.take<ResultWords>(1);
  unsigned long long* scan_ws = ws.take<unsigned long long>(scan_state_words(n));
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "counts_to_offsets: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, ws.used, stream));
  if (n == 0) {
    PPG_CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(int64_t), stream));
    return PPG_OK;
  }
  return launch_scan(CountProducer64{counts, &result->status, min_count}, OffsetConsumer64{offsets, n}, n, scan_ws, &result->total, stream);
}

// out_values[j] = values[i] (elements of 4 or 8 bytes; may be NULL) and owner[j] = i (may be NULL) for
// offsets[i] <= j < offsets[i + 1], j < total = offsets[n]
extern "C" int ppg_expand_offsets(const int64_t* offsets, int64_t n, int64_t total, const void* values, int value_bytes,
                                  void* out_values, int64_t* owner, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (total == 0 || n == 0) return PPG_OK;
  PPG_REQUIRE(out_values == nullptr || value_bytes == 4 || value_bytes == 8, PPG_ERR_INVALID, "expand_offsets: elements of %d bytes", value_bytes);
  const int grid = grid_for(total, 256);
  if (out_values != nullptr && value_bytes == 8)
    expand_offsets_kernel<unsigned long long><<<grid, 256, 0, stream>>>(offsets, n, total, static_cast<const unsigned long long*>(values),
                                                                        static_cast<unsigned long long*>(out_values), owner);
  else
    expand_offsets_kernel<uint32_t><<<grid, 256, 0, stream>>>(offsets, n, total, static_cast<const uint32_t*>(values),
                                                              static_cast<uint32_t*>(out_values), owner);
  PPG_LAUNCHED();
  return PPG_OK;
}

// edge_index [2, total - num_walks]: the links p -> p + 1 (plus `base`) inside every walk, walks laid end to end at
// offsets [num_walks + 1]; every walk must have at least one position (core/path_data.py:139-159)
extern "C" int ppg_walk_chain(const int64_t* offsets, int64_t num_walks, int64_t total, int64_t base, int64_t* edge_index,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(total >= num_walks, PPG_ERR_INVALID, "walk_chain: %lld positions for %lld walks", (long long)total, (long long)num_walks);
  if (total == num_walks) return PPG_OK;
  walk_chain_kernel<<<grid_for(total, 256), 256, 0, stream>>>(offsets, num_walks, total, base, edge_index, total - num_walks);
  PPG_LAUNCHED();
  return PPG_OK;
}

// counts [num_bins] (int64) of the ids; `status_word` (4 zeroed bytes... any zeroed 8-byte word) gets bit 0 if an id is outside [0, num_bins)
extern "C" int ppg_bincount(const int64_t* ids, int64_t n, int64_t num_bins, int64_t* counts, void* status_word, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(num_bins >= 0 && n >= 0, PPG_ERR_INVALID, "bincount: negative size");
  if (num_bins > 0) PPG_CUDA_TRY(cudaMemsetAsync(counts, 0, static_cast<size_t>(num_bins) * sizeof(int64_t), stream));
  if (n == 0) return PPG_OK;
  PPG_REQUIRE(num_bins > 0, PPG_ERR_INVALID, "bincount: ids present but no bins");
  bincount_kernel<<<grid_for(n, 256), 256, 0, stream>>>(ids, n, num_bins, reinterpret_cast<unsigned long long*>(counts),
                                                        static_cast<unsigned*>(status_word));
  PPG_LAUNCHED();
  return PPG_OK;
}
