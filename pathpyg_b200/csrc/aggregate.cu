// a4 aggregate_edge_index: distinct k-gram rows in lexicographic order (+ inverse), and
// (row, col) coalescing of the mapped edge index with a per-run weight reduction.
//
// Both halves are "pack -> onesweep radix sort -> run heads + look-back scan":
//   * a k-gram row [M,k] int64 is packed into ONE integer key holding only the significant bits of
//     every column (value - column minimum), so a lexicographic row sort is a single-key radix sort
//     over ceil(total_bits / 8) passes instead of a k-column comparison sort;
//   * the run-head flags of the sorted keys are scanned in a single pass whose consumer scatters the
//     dense rank (inverse index) / records the run starts, so no flag or rank array touches HBM;
//   * per-run weights are reduced left to right in stable sorted order (deterministic, and the same
//     order a stable CPU sort + sequential scatter_add produces).
//
// Algorithmic bytes at order k: 8(k+1)*E_{k-1} + 20*E_k + 8k*n_k + 20*E^_k.
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace ppg {

constexpr int kMaxPackWidth = 64;
constexpr unsigned kStatusIdOutOfRange = 1u;

struct ResultWords {
  unsigned long long total;
  unsigned long long status;
};

// ------------------------------------------------------------------ column statistics
constexpr int kStatCols = 8;

__global__ void __launch_bounds__(256)
rows_minmax_kernel(const int64_t* __restrict__ rows, int64_t M, int width, int col0, long long* __restrict__ mins,
                   long long* __restrict__ maxs, int* __restrict__ not_ascending) {
  long long lo[kStatCols], hi[kStatCols];
#pragma unroll
  for (int c = 0; c < kStatCols; ++c) {
    lo[c] = 0x7fffffffffffffffll;
    hi[c] = -0x7fffffffffffffffll - 1;
  }
  bool bad = false;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < M; i += stride) {
    const int64_t* r = rows + i * width;
#pragma unroll
    for (int c = 0; c < kStatCols; ++c) {
      if (col0 + c < width) {
        const long long v = r[col0 + c];
        lo[c] = v < lo[c] ? v : lo[c];
        hi[c] = v > hi[c] ? v : hi[c];
      }
    }
    if (col0 == 0 && i > 0) {  // lexicographic row(i) > row(i-1)?
      const int64_t* p = r - width;
      int cmp = 0;
      for (int c = 0; c < width && cmp == 0; ++c) cmp = r[c] > p[c] ? 1 : (r[c] < p[c] ? -1 : 0);
      bad |= cmp <= 0;
    }
  }
#pragma unroll
  for (int c = 0; c < kStatCols; ++c) {
    if (col0 + c < width) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const long long a = __shfl_xor_sync(kFullMask, lo[c], d);
        const long long b = __shfl_xor_sync(kFullMask, hi[c], d);
        lo[c] = a < lo[c] ? a : lo[c];
        hi[c] = b > hi[c] ? b : hi[c];
      }
      if (lane_id() == 0) {
        atomicMin(&mins[col0 + c], lo[c]);
        atomicMax(&maxs[col0 + c], hi[c]);
      }
    }
  }
  if (__any_sync(kFullMask, bad) && lane_id() == 0) atomicOr(not_ascending, 1);
}

__global__ void init_minmax_kernel(long long* mins, long long* maxs, int* flag, int width) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < width) {
    mins[c] = 0x7fffffffffffffffll;
    maxs[c] = -0x7fffffffffffffffll - 1;
  }
  if (c == 0) *flag = 0;
}

// ------------------------------------------------------------------ distinct rows
struct PackParams {
  long long col_min[kMaxPackWidth];
  int col_shift[kMaxPackWidth];
  int width;
};

__global__ void __launch_bounds__(256)
pack_rows_kernel(const int64_t* __restrict__ rows, int64_t M, PackParams pp, unsigned long long* __restrict__ keys,
                 unsigned long long* __restrict__ ghist0) {
  __shared__ unsigned s_hist[kRadix];
  Digit0Counter digit0;
  digit0.begin(s_hist);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x; base < M; base += stride) {
    const int64_t i = base + threadIdx.x;
    const bool valid = i < M;
    unsigned long long k = 0;
    if (valid) {
      const int64_t* r = rows + i * pp.width;
      for (int c = 0; c < pp.width; ++c)
        k |= static_cast<unsigned long long>(ld_stream(r + c) - pp.col_min[c]) << pp.col_shift[c];
      keys[i] = k;
    }
    digit0.count(static_cast<unsigned>(k) & (kRadix - 1), valid);
  }
  digit0.end(ghist0);
}

struct UniqueLayout {
  ResultWords* result;
  unsigned long long* scan_ws;
  unsigned long long* sort_ws;
  size_t zero_bytes;
  unsigned long long *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  uint32_t* rep;  // [M] original row index of the first occurrence of every distinct row, by rank
  int passes;

  UniqueLayout(Workspace& ws, int64_t M, int total_bits) {
    passes = sort_num_passes(total_bits);
    result = ws.take<ResultWords>(1);
    scan_ws = ws.take<unsigned long long>(scan_state_words(M));
    sort_ws = ws.take<unsigned long long>(sort_state_words(M, total_bits));
    zero_bytes = ws.used;
    keys_a = ws.take<unsigned long long>(static_cast<size_t>(M));
    keys_b = ws.take<unsigned long long>(static_cast<size_t>(M));
    vals_a = ws.take<uint32_t>(static_cast<size_t>(M));
    vals_b = ws.take<uint32_t>(static_cast<size_t>(M));
    rep = ws.take<uint32_t>(static_cast<size_t>(M));
  }
  const unsigned long long* sorted_keys() const { return (passes & 1) ? keys_b : keys_a; }
  const uint32_t* sorted_perm() const { return (passes & 1) ? vals_b : vals_a; }
};

struct RunHeadProducer {
  const unsigned long long* keys;
  __device__ unsigned long long operator()(int64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1ull : 0ull; }
};
struct RankScatterConsumer {  // inverse index + first occurrence per distinct row
  const uint32_t* perm;
  int64_t* inverse;
  uint32_t* rep;
  __device__ void operator()(int64_t i, unsigned long long head, unsigned long long prefix) const {
    const unsigned long long rank = prefix + head - 1;
    const uint32_t src = perm[i];
    inverse[src] = static_cast<int64_t>(rank);
    if (head) rep[rank] = src;
  }
};

__global__ void __launch_bounds__(256)
gather_rows_kernel(const int64_t* __restrict__ rows, const uint32_t* __restrict__ rep, int64_t n, int width,
                   int64_t* __restrict__ out) {
  const int64_t total = n * width;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / width;
    const int c = static_cast<int>(t - r * width);
    st_stream(out + t, rows[static_cast<int64_t>(rep[r]) * width + c]);
  }
}

// ------------------------------------------------------------------ coalesce
struct CoalesceLayout {
  ResultWords* result;
  unsigned long long* scan_ws;
  unsigned long long* sort_ws;
  size_t zero_bytes;
  unsigned long long *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  uint32_t* run_start;  // [E + 1] first sorted slot of every run, by output edge id
  int node_bits;
  int passes;

  CoalesceLayout(Workspace& ws, int64_t E, int64_t N) {
    node_bits = bits_for(N > 0 ? static_cast<uint64_t>(N - 1) : 0);
    passes = sort_num_passes(2 * node_bits);
    result = ws.take<ResultWords>(1);
    scan_ws = ws.take<unsigned long long>(scan_state_words(E));
    sort_ws = ws.take<unsigned long long>(sort_state_words(E, 2 * node_bits));
    zero_bytes = ws.used;
    keys_a = ws.take<unsigned long long>(static_cast<size_t>(E));
    keys_b = ws.take<unsigned long long>(static_cast<size_t>(E));
    vals_a = ws.take<uint32_t>(static_cast<size_t>(E));
    vals_b = ws.take<uint32_t>(static_cast<size_t>(E));
    run_start = ws.take<uint32_t>(static_cast<size_t>(E) + 1);
  }
  const unsigned long long* sorted_keys() const { return (passes & 1) ? keys_b : keys_a; }
  const uint32_t* sorted_perm() const { return (passes & 1) ? vals_b : vals_a; }
};

__global__ void __launch_bounds__(256)
edge_keys_kernel(const int64_t* __restrict__ ei, int64_t E, const int64_t* __restrict__ remap, int64_t remap_len,
                 int64_t num_nodes, int node_bits, unsigned long long* __restrict__ keys,
                 unsigned long long* __restrict__ status, unsigned long long* __restrict__ ghist0) {
  __shared__ unsigned s_hist[kRadix];
  Digit0Counter digit0;
  digit0.begin(s_hist);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x; base < E; base += stride) {
    const int64_t j = base + threadIdx.x;
    const bool valid = j < E;
    unsigned long long key = 0;
    if (valid) {
      int64_t r = ld_stream(ei + j);
      int64_t c = ld_stream(ei + E + j);
      bool ok = true;
      if (remap != nullptr) {
        ok = r >= 0 && r < remap_len && c >= 0 && c < remap_len;
        r = ok ? remap[r] : 0;
        c = ok ? remap[c] : 0;
      }
      ok = ok && r >= 0 && r < num_nodes && c >= 0 && c < num_nodes;
      if (!ok) {
        atomicOr(reinterpret_cast<unsigned*>(status), kStatusIdOutOfRange);
        r = c = 0;
      }
      key = (static_cast<unsigned long long>(r) << node_bits) | static_cast<unsigned long long>(c);
      keys[j] = key;
    }
    digit0.count(static_cast<unsigned>(key) & (kRadix - 1), valid);
  }
  digit0.end(ghist0);
}

struct RunStartConsumer {
  uint32_t* run_start;
  int64_t n;
  const uint32_t* perm;  // with `inverse`: the run (= output edge) every input edge fell into
  int64_t* inverse;      // nullable
  __device__ void operator()(int64_t i, unsigned long long head, unsigned long long prefix) const {
    if (head) run_start[prefix] = static_cast<uint32_t>(i);
    if (i == n - 1) run_start[prefix + head] = static_cast<uint32_t>(n);
    if (inverse != nullptr) inverse[perm[i]] = static_cast<int64_t>(prefix + head - 1);
  }
};

// k-gram of every edge of a layer: the (k-1)-gram of its source node + the last node of its target node
// (multi_order_model.py:114 applied to the DISTINCT edges: these rows are the next layer's node sequences)
__global__ void __launch_bounds__(256)
extend_rows_kernel(const int64_t* __restrict__ prev, int width, const int64_t* __restrict__ ei, int64_t n,
                   int64_t* __restrict__ out) {
  const int ow = width + 1;
  const int64_t total = n * ow;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t j = t / ow;
    const int c = static_cast<int>(t - j * ow);
    const int64_t v = c < width ? prev[ei[j] * width + c] : prev[ei[n + j] * width + (width - 1)];
    st_stream(out + t, v);
  }
}

__device__ __forceinline__ float mean_of(float s, uint32_t n) { return s / static_cast<float>(n); }
__device__ __forceinline__ double mean_of(double s, uint32_t n) { return s / static_cast<double>(n); }
// integers: floor division, as torch's div(rounding_mode="floor") inside PyG scatter(reduce="mean")
__device__ __forceinline__ long long mean_of(long long s, uint32_t n) {
  const long long c = n;
  long long q = s / c;
  if (s % c != 0 && s < 0) q -= 1;
  return q;
}
__device__ __forceinline__ int mean_of(int s, uint32_t n) { return static_cast<int>(mean_of(static_cast<long long>(s), n)); }

// one thread per output edge: decode (row, col) from the run's key, reduce the run's weights left to right
template <typename T, int REDUCE, bool UNIT>
__global__ void __launch_bounds__(256)
coalesce_fill_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ perm,
                     const uint32_t* __restrict__ run_start, int64_t num_out, int node_bits, const T* __restrict__ w,
                     int64_t* __restrict__ out_ei, T* __restrict__ out_w) {
  const unsigned long long mask = (1ull << node_bits) - 1;  // node_bits <= 32
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < num_out; r += stride) {
    const uint32_t a = run_start[r];
    const uint32_t b = run_start[r + 1];
    const unsigned long long k = keys[a];
    st_stream(out_ei + r, static_cast<int64_t>(k >> node_bits));
    st_stream(out_ei + num_out + r, static_cast<int64_t>(k & mask));
    T acc;
    if (UNIT) {
      acc = REDUCE == PPG_REDUCE_SUM ? static_cast<T>(b - a) : static_cast<T>(1);
    } else {
      acc = w[perm[a]];
      for (uint32_t i = a + 1; i < b; ++i) {
        const T x = w[perm[i]];
        if (REDUCE == PPG_REDUCE_MIN) acc = x < acc ? x : acc;
        else if (REDUCE == PPG_REDUCE_MAX) acc = x > acc ? x : acc;
        else acc += x;
      }
      if (REDUCE == PPG_REDUCE_MEAN) acc = mean_of(acc, b - a);
    }
    st_stream(out_w + r, acc);
  }
}

template <typename T, bool UNIT>
static int coalesce_fill_dispatch(const CoalesceLayout& L, int64_t num_out, const void* w, int reduce, int64_t* out_ei,
                                  void* out_w, cudaStream_t stream) {
  const int grid = grid_for(num_out, 256);
  const T* wt = static_cast<const T*>(w);
  T* ow = static_cast<T*>(out_w);
#define PPG_FILL(R)                                                                                        \
  coalesce_fill_kernel<T, R, UNIT><<<grid, 256, 0, stream>>>(L.sorted_keys(), L.sorted_perm(), L.run_start, \
                                                             num_out, L.node_bits, wt, out_ei, ow)
  switch (reduce) {
    case PPG_REDUCE_SUM: PPG_FILL(PPG_REDUCE_SUM); break;
    case PPG_REDUCE_MEAN: PPG_FILL(PPG_REDUCE_MEAN); break;
    case PPG_REDUCE_MIN: PPG_FILL(PPG_REDUCE_MIN); break;
    case PPG_REDUCE_MAX: PPG_FILL(PPG_REDUCE_MAX); break;
    default: PPG_REQUIRE(false, PPG_ERR_INVALID, "coalesce: unknown reduce code %d", reduce);
  }
#undef PPG_FILL
  PPG_LAUNCHED();
  return PPG_OK;
}

}  // namespace ppg

using namespace ppg;

// =================================================================== column statistics
extern "C" size_t ppg_rows_minmax_workspace_bytes(int64_t width) {
  return static_cast<size_t>(width < 1 ? 1 : width) * 2 * sizeof(long long) + 256;
}

extern "C" int ppg_rows_minmax(const int64_t* rows, int64_t M, int64_t width, void* workspace, size_t workspace_bytes,
                               int64_t* h_col_min, int64_t* h_col_max, int* h_strictly_ascending, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(width >= 1 && width < (1 << 20) && M >= 0, PPG_ERR_INVALID, "rows_minmax: bad shape [%lld, %lld]",
              (long long)M, (long long)width);
  PPG_REQUIRE(workspace_bytes >= ppg_rows_minmax_workspace_bytes(width), PPG_ERR_WORKSPACE,
              "rows_minmax: workspace too small");
  long long* mins = static_cast<long long*>(workspace);
  long long* maxs = mins + width;
  int* flag = reinterpret_cast<int*>(maxs + width);
  const int w = static_cast<int>(width);
  init_minmax_kernel<<<static_cast<unsigned>(ceil_div(w, 256)), 256, 0, stream>>>(mins, maxs, flag, w);
  PPG_LAUNCHED();
  if (M > 0) {
    for (int col0 = 0; col0 < w; col0 += kStatCols) {
      rows_minmax_kernel<<<grid_for(M, 256 * 4, kNumSMsB200 * 8), 256, 0, stream>>>(rows, M, w, col0, mins, maxs, flag);
      PPG_LAUNCHED();
    }
  }
  PPG_CUDA_TRY(cudaMemcpyAsync(h_col_min, mins, sizeof(long long) * width, cudaMemcpyDeviceToHost, stream));
  PPG_CUDA_TRY(cudaMemcpyAsync(h_col_max, maxs, sizeof(long long) * width, cudaMemcpyDeviceToHost, stream));
  int not_ascending = 0;
  PPG_CUDA_TRY(cudaMemcpyAsync(&not_ascending, flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PPG_CUDA_TRY(cudaStreamSynchronize(stream));
  *h_strictly_ascending = not_ascending ? 0 : 1;
  return PPG_OK;
}

// =================================================================== distinct rows
extern "C" size_t ppg_unique_rows_workspace_bytes(int64_t num_rows, int total_bits) {
  Workspace ws(nullptr, 0);
  UniqueLayout L(ws, num_rows < 0 ? 0 : num_rows, total_bits);
  return ws.used + 256;
}

extern "C" int ppg_unique_rows_sort(const int64_t* rows, int64_t M, int64_t width, const int64_t* h_col_min,
                                    const int* h_col_shift, int total_bits, void* workspace, size_t workspace_bytes,
                                    int64_t* out_inverse, int64_t* h_num_unique, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(M >= 0 && M < (1ll << 31), PPG_ERR_INVALID, "unique_rows: %lld rows outside [0, 2^31)", (long long)M);
  PPG_REQUIRE(width >= 1 && width <= kMaxPackWidth, PPG_ERR_INVALID, "unique_rows: width %lld outside [1, %d]",
              (long long)width, kMaxPackWidth);
  PPG_REQUIRE(total_bits >= 0 && total_bits <= 64, PPG_ERR_INVALID, "unique_rows: %d key bits exceed 64", total_bits);
  *h_num_unique = 0;
  if (M == 0) return PPG_OK;
  Workspace ws(workspace, workspace_bytes);
  UniqueLayout L(ws, M, total_bits);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "unique_rows: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));

  PackParams pp;
  pp.width = static_cast<int>(width);
  for (int c = 0; c < pp.width; ++c) {
    pp.col_min[c] = h_col_min[c];
    pp.col_shift[c] = h_col_shift[c];
    PPG_REQUIRE(h_col_shift[c] >= 0 && h_col_shift[c] < 64, PPG_ERR_INVALID, "unique_rows: bad shift for column %d", c);
  }
  pack_rows_kernel<<<grid_for(M, 256 * 4), 256, 0, stream>>>(rows, M, pp, L.keys_a, L.sort_ws);
  PPG_LAUNCHED();
  int in_b = 0;
  PPG_TRY(radix_sort_pairs<unsigned long long>(L.keys_a, L.keys_b, L.vals_a, L.vals_b, true, true, M, total_bits,
                                               L.sort_ws, &in_b, stream, nullptr, true));
  PPG_REQUIRE((in_b != 0) == ((L.passes & 1) != 0), PPG_ERR_CUDA, "unique_rows: internal buffer parity mismatch");
  PPG_TRY(launch_scan(RunHeadProducer{L.sorted_keys()}, RankScatterConsumer{L.sorted_perm(), out_inverse, L.rep}, M,
                      L.scan_ws, &L.result->total, stream));
  ResultWords h;
  PPG_TRY(read_back(&h, L.result, stream));
  *h_num_unique = static_cast<int64_t>(h.total);
  return PPG_OK;
}

extern "C" int ppg_unique_rows_gather(const int64_t* rows, int64_t M, int64_t width, const void* workspace,
                                      int total_bits, int64_t num_unique, int64_t* out_rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_unique == 0 || M == 0) return PPG_OK;
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  UniqueLayout L(ws, M, total_bits);
  gather_rows_kernel<<<grid_for(num_unique * width, 256 * 4), 256, 0, stream>>>(rows, L.rep, num_unique,
                                                                                static_cast<int>(width), out_rows);
  PPG_LAUNCHED();
  return PPG_OK;
}

// =================================================================== coalesce
extern "C" size_t ppg_coalesce_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
  Workspace ws(nullptr, 0);
  CoalesceLayout L(ws, num_edges < 0 ? 0 : num_edges, num_nodes < 0 ? 0 : num_nodes);
  return ws.used + 256;
}

extern "C" int ppg_coalesce_sort(const int64_t* edge_index, int64_t E, const int64_t* remap, int64_t remap_len,
                                 int64_t num_nodes, void* workspace, size_t workspace_bytes, int64_t* out_inverse,
                                 int64_t* h_num_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(E >= 0 && E < (1ll << 31), PPG_ERR_INVALID, "coalesce: %lld edges outside [0, 2^31)", (long long)E);
  PPG_REQUIRE(num_nodes >= 0 && num_nodes <= (1ll << 32), PPG_ERR_INVALID, "coalesce: num_nodes %lld outside [0, 2^32]",
              (long long)num_nodes);
  if (h_num_out != nullptr) *h_num_out = 0;
  if (E == 0) return PPG_OK;
  PPG_REQUIRE(num_nodes > 0, PPG_ERR_INVALID, "coalesce: edges present but num_nodes == 0");
  Workspace ws(workspace, workspace_bytes);
  CoalesceLayout L(ws, E, num_nodes);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "coalesce: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));

  edge_keys_kernel<<<grid_for(E, 256 * 4), 256, 0, stream>>>(edge_index, E, remap, remap_len, num_nodes, L.node_bits,
                                                             L.keys_a, &L.result->status, L.sort_ws);
  PPG_LAUNCHED();
  int in_b = 0;
  PPG_TRY(radix_sort_pairs<unsigned long long>(L.keys_a, L.keys_b, L.vals_a, L.vals_b, true, true, E, 2 * L.node_bits,
                                               L.sort_ws, &in_b, stream, nullptr, true));
  PPG_REQUIRE((in_b != 0) == ((L.passes & 1) != 0), PPG_ERR_CUDA, "coalesce: internal buffer parity mismatch");
  PPG_TRY(launch_scan(RunHeadProducer{L.sorted_keys()}, RunStartConsumer{L.run_start, E, L.sorted_perm(), out_inverse}, E,
                      L.scan_ws, &L.result->total, stream));
  if (h_num_out == nullptr) return PPG_OK;  // deferred: ppg_result_read
  ResultWords h;
  PPG_TRY(read_back(&h, L.result, stream));
  PPG_REQUIRE((h.status & kStatusIdOutOfRange) == 0, PPG_ERR_INVALID,
              "coalesce: mapped node id outside [0, num_nodes=%lld)", (long long)num_nodes);
  *h_num_out = static_cast<int64_t>(h.total);
  return PPG_OK;
}

extern "C" int ppg_coalesce_fill(const void* workspace, int64_t E, int64_t num_nodes, int64_t num_out,
                                 const void* weights, int dtype, int reduce, int64_t* out_edge_index, void* out_weights,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_out == 0 || E == 0) return PPG_OK;
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  CoalesceLayout L(ws, E, num_nodes);
  if (weights == nullptr) {
    PPG_REQUIRE(dtype == PPG_F32, PPG_ERR_INVALID, "coalesce: unit weights are float32");
    return coalesce_fill_dispatch<float, true>(L, num_out, nullptr, reduce, out_edge_index, out_weights, stream);
  }
  switch (dtype) {
    case PPG_F32: return coalesce_fill_dispatch<float, false>(L, num_out, weights, reduce, out_edge_index, out_weights, stream);
    case PPG_F64: return coalesce_fill_dispatch<double, false>(L, num_out, weights, reduce, out_edge_index, out_weights, stream);
    case PPG_I64: return coalesce_fill_dispatch<long long, false>(L, num_out, weights, reduce, out_edge_index, out_weights, stream);
    case PPG_I32: return coalesce_fill_dispatch<int, false>(L, num_out, weights, reduce, out_edge_index, out_weights, stream);
    default: PPG_REQUIRE(false, PPG_ERR_INVALID, "coalesce: unsupported dtype code %d", dtype);
  }
  return PPG_OK;
}

extern "C" int ppg_extend_rows(const int64_t* prev_rows, int64_t num_prev, int64_t width, const int64_t* edge_index,
                               int64_t num_edges, int64_t* out_rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)num_prev;
  PPG_REQUIRE(width >= 1 && width < (1 << 20) && num_edges >= 0, PPG_ERR_INVALID, "extend_rows: bad shape");
  if (num_edges == 0) return PPG_OK;
  extend_rows_kernel<<<grid_for(num_edges * (width + 1), 256 * 4), 256, 0, stream>>>(prev_rows, static_cast<int>(width),
                                                                                     edge_index, num_edges, out_rows);
  PPG_LAUNCHED();
  return PPG_OK;
}

// =================================================================== stand-alone pair sort (utility + bench probe)
namespace ppg {
struct SortLayout {
  unsigned long long* sort_ws;
  size_t zero_bytes;
  unsigned long long* keys_b;
  uint32_t *vals_a, *vals_b;
  SortLayout(Workspace& ws, int64_t n, int end_bit) {
    sort_ws = ws.take<unsigned long long>(sort_state_words(n, end_bit));
    zero_bytes = ws.used;
    keys_b = ws.take<unsigned long long>(static_cast<size_t>(n));
    vals_a = ws.take<uint32_t>(static_cast<size_t>(n));
    vals_b = ws.take<uint32_t>(static_cast<size_t>(n));
  }
};
}  // namespace ppg

extern "C" size_t ppg_sort_pairs_workspace_bytes(int64_t n, int end_bit) {
  Workspace ws(nullptr, 0);
  SortLayout L(ws, n < 0 ? 0 : n, end_bit);
  return ws.used + 256;
}

extern "C" int ppg_sort_pairs_u64(uint64_t* keys, uint32_t* out_perm, int64_t n, int end_bit, void* workspace,
                                  size_t workspace_bytes, float* h_pass_ms, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(n >= 0 && n < (1ll << 31) && end_bit >= 0 && end_bit <= 64, PPG_ERR_INVALID, "sort_pairs: bad arguments");
  if (n == 0) return PPG_OK;
  Workspace ws(workspace, workspace_bytes);
  SortLayout L(ws, n, end_bit);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "sort_pairs: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));
  int in_b = 0;
  PPG_TRY(radix_sort_pairs<unsigned long long>(reinterpret_cast<unsigned long long*>(keys), L.keys_b, L.vals_a, L.vals_b,
                                               true, true, n, end_bit, L.sort_ws, &in_b, stream, h_pass_ms));
  if (in_b) PPG_CUDA_TRY(cudaMemcpyAsync(keys, L.keys_b, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, stream));
  PPG_CUDA_TRY(cudaMemcpyAsync(out_perm, in_b ? L.vals_b : L.vals_a, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
  return PPG_OK;
}
