// a1 lift_order_temporal, a2 lift_order_edge_index, a3 aggregate_node_attributes.
//
// Both lifts are "count -> scan -> load-balanced expand":
//   count : per input edge e the number of continuations and the slot of the first one
//           (a2: out-degree / CSR pointer of dst(e); a1: two binary searches over the time-ordered
//           out-edges of dst(e)), fused into the single-pass look-back scan that turns the counts
//           into int64 output offsets;
//   expand: one CTA per 2048 OUTPUT columns. The sources overlapping the tile are found by a
//           32-ary cooperative search, marked in shared memory at their first output slot and
//           spread by a block-wide max-scan, so every output column is written exactly once with
//           coalesced 8-byte streaming stores regardless of how skewed the fan-out is.
//
// Algorithmic bytes: a2 16*E + 16*E', a1 24*m + 16*E2.
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace ppg {

constexpr int kExpandBlock = 256;
constexpr int kExpandItems = 8;
constexpr int kExpandTile = kExpandBlock * kExpandItems;

constexpr unsigned kStatusIdOutOfRange = 1u;

// result words shared by count kernels: [0] total, [1] status bits
struct ResultWords {
  unsigned long long total;
  unsigned long long status;
};

// ------------------------------------------------------------------ degree histogram
// deg[ids[i]] += 1 with one atomic per distinct id per warp (row-sorted input => ~1 atomic / warp).
template <bool EMIT_KEYS>
__global__ void __launch_bounds__(256)
degree_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t num_nodes, uint32_t* __restrict__ deg,
              uint32_t* __restrict__ keys_out, unsigned long long* __restrict__ status,
              unsigned long long* __restrict__ ghist0) {
  __shared__ unsigned s_hist[kRadix];
  Digit0Counter digit0;
  if (EMIT_KEYS) digit0.begin(s_hist);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x; base < n; base += stride) {
    const int64_t i = base + threadIdx.x;
    const bool valid = i < n;
    uint32_t key = 0;
    if (valid) {
      const int64_t v = ld_stream(ids + i);
      const bool ok = v >= 0 && v < num_nodes;
      key = ok ? static_cast<uint32_t>(v) : 0u;
      if (EMIT_KEYS) keys_out[i] = key;
      const unsigned active = __activemask();
      const unsigned tag = ok ? static_cast<unsigned>(v) : 0xffffffffu;
      const unsigned peers = __match_any_sync(active, tag);
      if (!ok) {
        atomicOr(reinterpret_cast<unsigned*>(status), kStatusIdOutOfRange);
      } else if (lane_id() == static_cast<unsigned>(__ffs(peers) - 1)) {
        atomicAdd(&deg[v], static_cast<uint32_t>(__popc(peers)));
      }
    }
    if (EMIT_KEYS) digit0.count(key & (kRadix - 1), valid);  // the keys feed a radix sort: count its first digit here
  }
  if (EMIT_KEYS) digit0.end(ghist0);
}

struct DegreeProducer {
  const uint32_t* deg;
  __device__ unsigned long long operator()(int64_t i) const { return deg[i]; }
};
struct PointerConsumer {  // CSR pointer array with n + 1 entries
  uint32_t* ptr;
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long v, unsigned long long prefix) const {
    ptr[i] = static_cast<uint32_t>(prefix);
    if (i == n - 1) ptr[n] = static_cast<uint32_t>(prefix + v);
  }
};
struct OffsetConsumer {  // int64 output offsets with n + 1 entries
  unsigned long long* off;
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long v, unsigned long long prefix) const {
    off[i] = prefix;
    if (i == n - 1) off[n] = prefix + v;
  }
};

// ------------------------------------------------------------------ load-balanced expand
// largest e in [0, E) with off[e] <= target; requires off[0] <= target < off[E]
__device__ __forceinline__ int64_t warp_search_last_le(const unsigned long long* __restrict__ off, int64_t E,
                                                       unsigned long long target) {
  int64_t lo = 0, hi = E;
  const unsigned lane = lane_id();
  while (hi - lo > 1) {
    const int64_t step = ceil_div(hi - lo, 32);
    const int64_t p = lo + static_cast<int64_t>(lane) * step;
    const bool le = p < hi && off[p] <= target;
    const int cnt = __popc(__ballot_sync(kFullMask, le));  // >= 1: lane 0 probes `lo`
    const int64_t nlo = lo + static_cast<int64_t>(cnt - 1) * step;
    const int64_t nhi = lo + static_cast<int64_t>(cnt) * step;
    lo = nlo;
    hi = nhi < hi ? nhi : hi;
  }
  return lo;
}

template <class TailMap>
__global__ void __launch_bounds__(kExpandBlock)
expand_kernel(const unsigned long long* __restrict__ off, int64_t E, int64_t total, int64_t* __restrict__ out0,
              int64_t* __restrict__ out1, TailMap tail) {
  __shared__ __align__(16) uint32_t s_mark[kExpandTile];
  __shared__ int64_t s_bound[2];
  __shared__ uint32_t s_warp_max[kExpandBlock / 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const unsigned lane = lane_id();
  const int64_t o0 = static_cast<int64_t>(blockIdx.x) * kExpandTile;
  const int64_t o1 = o0 + kExpandTile < total ? o0 + kExpandTile : total;

  for (int i = tid; i < kExpandTile; i += kExpandBlock) s_mark[i] = 0;
  if (warp < 2) {
    const int64_t r = warp_search_last_le(off, E, static_cast<unsigned long long>(warp == 0 ? o0 : o1 - 1));
    if (lane == 0) s_bound[warp] = r;
  }
  __syncthreads();
  const int64_t e_lo = s_bound[0];
  const int64_t e_hi = s_bound[1];

  // every source with at least one output in this tile marks its first slot inside the tile
  for (int64_t e = e_lo + tid; e <= e_hi; e += kExpandBlock) {
    const unsigned long long a = off[e];
    const unsigned long long b = off[e + 1];
    if (b > a) {
      const int64_t pos = static_cast<int64_t>(a) > o0 ? static_cast<int64_t>(a) - o0 : 0;
      s_mark[pos] = static_cast<uint32_t>(e - e_lo + 1);
    }
  }
  __syncthreads();

  // block-wide inclusive max-scan (thread t owns slots 8t .. 8t+7)
  uint32_t m[kExpandItems];
  {
    const uint4 q0 = reinterpret_cast<const uint4*>(s_mark)[tid * 2];
    const uint4 q1 = reinterpret_cast<const uint4*>(s_mark)[tid * 2 + 1];
    m[0] = q0.x; m[1] = q0.y; m[2] = q0.z; m[3] = q0.w;
    m[4] = q1.x; m[5] = q1.y; m[6] = q1.z; m[7] = q1.w;
  }
#pragma unroll
  for (int i = 1; i < kExpandItems; ++i) m[i] = max(m[i], m[i - 1]);
  uint32_t inc = m[kExpandItems - 1];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(kFullMask, inc, d);
    if (lane >= static_cast<unsigned>(d)) inc = max(inc, y);
  }
  uint32_t before = __shfl_up_sync(kFullMask, inc, 1);
  if (lane == 0) before = 0;
  if (lane == 31) s_warp_max[warp] = inc;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kExpandBlock / 32; ++w)
    if (w < warp) before = max(before, s_warp_max[w]);
#pragma unroll
  for (int i = 0; i < kExpandItems; ++i) m[i] = max(m[i], before);
  reinterpret_cast<uint4*>(s_mark)[tid * 2] = make_uint4(m[0], m[1], m[2], m[3]);
  reinterpret_cast<uint4*>(s_mark)[tid * 2 + 1] = make_uint4(m[4], m[5], m[6], m[7]);
  __syncthreads();

  // striped, coalesced write-out
#pragma unroll
  for (int i = 0; i < kExpandItems; ++i) {
    const int p = i * kExpandBlock + tid;
    const int64_t o = o0 + p;
    if (o < o1) {
      const int64_t e = e_lo + s_mark[p] - 1;
      const int64_t j = o - static_cast<int64_t>(off[e]);
      st_stream(out0 + o, e);
      st_stream(out1 + o, tail(e, j));
    }
  }
}

template <class TailMap>
inline int launch_expand(const unsigned long long* off, int64_t E, int64_t total, int64_t* out_index, TailMap tail,
                         cudaStream_t stream) {
  if (total == 0) return PPG_OK;
  const int64_t tiles = ceil_div(total, kExpandTile);
  PPG_REQUIRE(tiles < (1ll << 31), PPG_ERR_INVALID, "expand: %lld output columns are too many", (long long)total);
  expand_kernel<<<static_cast<unsigned>(tiles), kExpandBlock, 0, stream>>>(off, E, total, out_index, out_index + total,
                                                                            tail);
  PPG_LAUNCHED();
  return PPG_OK;
}

// ------------------------------------------------------------------ a2: line graph
struct LiftLayout {
  // zeroed region
  ResultWords* result;
  uint32_t* deg;
  unsigned long long* scan_ptr_ws;
  unsigned long long* scan_off_ws;
  size_t zero_bytes;
  // plain
  uint32_t* ptr;    // [N + 1]
  uint32_t* first;  // [E]     ptr[dst(e)]
  unsigned long long* off;  // [E + 1]

  LiftLayout(Workspace& ws, int64_t E, int64_t N) {
    result = ws.take<ResultWords>(1);
    deg = ws.take<uint32_t>(static_cast<size_t>(N));
    scan_ptr_ws = ws.take<unsigned long long>(scan_state_words(N));
    scan_off_ws = ws.take<unsigned long long>(scan_state_words(E));
    zero_bytes = ws.used;
    ptr = ws.take<uint32_t>(static_cast<size_t>(N) + 1);
    first = ws.take<uint32_t>(static_cast<size_t>(E));
    off = ws.take<unsigned long long>(static_cast<size_t>(E) + 1);
  }
};

struct LiftCountProducer {
  const int64_t* col;
  const uint32_t* ptr;
  uint32_t* first;
  int64_t num_nodes;
  unsigned long long* status;
  __device__ unsigned long long operator()(int64_t e) const {
    const int64_t c = ld_stream(col + e);
    if (c < 0 || c >= num_nodes) {
      atomicOr(reinterpret_cast<unsigned*>(status), kStatusIdOutOfRange);
      first[e] = 0;
      return 0;
    }
    const uint32_t a = ptr[c];
    first[e] = a;
    return ptr[c + 1] - a;
  }
};

struct LiftTail {
  const uint32_t* first;
  __device__ int64_t operator()(int64_t e, int64_t j) const { return static_cast<int64_t>(first[e]) + j; }
};

// ------------------------------------------------------------------ a1: temporal event graph
template <int MODE>
struct TimeWindow {
  int64_t delta_i;
  double delta_d;
  float delta_f;
  // raw 64-bit words of the time array
  __device__ bool after(unsigned long long tf, unsigned long long te) const {
    if (MODE == PPG_TIME_F64) return __longlong_as_double(tf) > __longlong_as_double(te);
    return static_cast<int64_t>(tf) > static_cast<int64_t>(te);
  }
  __device__ bool within(unsigned long long tf, unsigned long long te) const {
    if (MODE == PPG_TIME_F64) return __longlong_as_double(tf) <= __longlong_as_double(te) + delta_d;
    if (MODE == PPG_TIME_I64_F32DELTA)
      return __ll2float_rn(static_cast<int64_t>(tf)) <= __fadd_rn(__ll2float_rn(static_cast<int64_t>(te)), delta_f);
    return static_cast<int64_t>(tf) <= static_cast<int64_t>(te) + delta_i;
  }
};

struct TemporalLayout {
  ResultWords* result;
  uint32_t* deg;
  unsigned long long* scan_ptr_ws;
  unsigned long long* scan_off_ws;
  unsigned long long* sort_ws;
  size_t zero_bytes;
  uint32_t *keys_a, *keys_b, *vals_a, *vals_b;
  uint32_t* ptr;                  // [N + 1] CSR over source node
  unsigned long long* ts_sorted;  // [m] time of the edges grouped by source (time order inside a group)
  uint32_t* first;                // [m] slot (in grouped order) of the first continuation
  unsigned long long* off;        // [m + 1]
  int sort_bits;

  TemporalLayout(Workspace& ws, int64_t m, int64_t N) {
    sort_bits = bits_for(N > 0 ? static_cast<uint64_t>(N - 1) : 0);
    result = ws.take<ResultWords>(1);
    deg = ws.take<uint32_t>(static_cast<size_t>(N));
    scan_ptr_ws = ws.take<unsigned long long>(scan_state_words(N));
    scan_off_ws = ws.take<unsigned long long>(scan_state_words(m));
    sort_ws = ws.take<unsigned long long>(sort_state_words(m, sort_bits));
    zero_bytes = ws.used;
    keys_a = ws.take<uint32_t>(static_cast<size_t>(m));
    keys_b = ws.take<uint32_t>(static_cast<size_t>(m));
    vals_a = ws.take<uint32_t>(static_cast<size_t>(m));
    vals_b = ws.take<uint32_t>(static_cast<size_t>(m));
    ptr = ws.take<uint32_t>(static_cast<size_t>(N) + 1);
    ts_sorted = ws.take<unsigned long long>(static_cast<size_t>(m));
    first = ws.take<uint32_t>(static_cast<size_t>(m));
    off = ws.take<unsigned long long>(static_cast<size_t>(m) + 1);
  }
  // the permutation (edges grouped by source) ends in vals_b after an odd number of passes
  const uint32_t* grouped() const { return (sort_num_passes(sort_bits) & 1) ? vals_b : vals_a; }
};

__global__ void __launch_bounds__(256)
gather64_kernel(const unsigned long long* __restrict__ src, const uint32_t* __restrict__ idx, int64_t n,
                unsigned long long* __restrict__ dst) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = src[idx[i]];
}

template <int MODE>
struct TemporalCountProducer {
  const int64_t* dst;
  const unsigned long long* time;
  const uint32_t* ptr;
  const unsigned long long* ts_sorted;
  uint32_t* first;
  int64_t num_nodes;
  unsigned long long* status;
  TimeWindow<MODE> win;
  __device__ unsigned long long operator()(int64_t e) const {
    const int64_t v = ld_stream(dst + e);
    if (v < 0 || v >= num_nodes) {
      atomicOr(reinterpret_cast<unsigned*>(status), kStatusIdOutOfRange);
      first[e] = 0;
      return 0;
    }
    const unsigned long long te = ld_stream(time + e);
    uint32_t a = ptr[v];
    const uint32_t b = ptr[v + 1];
    // first out-edge of v strictly later than e
    uint32_t lo = a, hi = b;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (win.after(ts_sorted[mid], te)) hi = mid; else lo = mid + 1;
    }
    a = lo;
    // first out-edge of v beyond the window
    hi = b;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (win.within(ts_sorted[mid], te)) lo = mid + 1; else hi = mid;
    }
    first[e] = a;
    return lo - a;
  }
};

struct TemporalTail {
  const uint32_t* first;
  const uint32_t* grouped;
  __device__ int64_t operator()(int64_t e, int64_t j) const { return grouped[static_cast<int64_t>(first[e]) + j]; }
};

template <int MODE>
static int temporal_count_scan(const TemporalLayout& L, const int64_t* edge_index, const void* time, int64_t m,
                               int64_t N, int64_t delta_i, double delta_f, cudaStream_t stream) {
  TemporalCountProducer<MODE> prod{edge_index + m,
                                   static_cast<const unsigned long long*>(time),
                                   L.ptr,
                                   L.ts_sorted,
                                   L.first,
                                   N,
                                   &L.result->status,
                                   TimeWindow<MODE>{delta_i, delta_f, static_cast<float>(delta_f)}};
  return launch_scan(prod, OffsetConsumer{L.off, m}, m, L.scan_off_ws, &L.result->total, stream);
}

// ------------------------------------------------------------------ a3: per-edge attribute from its end points
// ids outside [0, num_attr) read slot 0 and raise the status bit (the reference fails with an IndexError there)
template <typename T, int RULE>
__global__ void __launch_bounds__(256)
pair_attributes_kernel(const int64_t* __restrict__ ei, int64_t E, const T* __restrict__ attr, int64_t num_attr,
                       T* __restrict__ out, unsigned* __restrict__ status) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  bool bad = false;
  auto fetch = [&](int64_t id) {
    const bool ok = id >= 0 && id < num_attr;
    bad |= !ok;
    return attr[ok ? id : 0];
  };
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < E; j += stride) {
    T r;
    if (RULE == PPG_PAIR_SRC) {
      r = fetch(ld_stream(ei + j));
    } else if (RULE == PPG_PAIR_DST) {
      r = fetch(ld_stream(ei + E + j));
    } else {
      const T a = fetch(ld_stream(ei + j));
      const T b = fetch(ld_stream(ei + E + j));
      if (RULE == PPG_PAIR_MAX) r = a > b ? a : (b > a ? b : (a == a ? a : b)); // torch.maximum propagates NaN
      else if (RULE == PPG_PAIR_MUL) r = a * b;
      else r = a + b;
    }
    st_stream(out + j, r);
  }
  if (bad && status != nullptr) atomicOr(status, kStatusIdOutOfRange);
}

template <typename T>
static int pair_attributes_dispatch(const int64_t* ei, int64_t E, const void* attr, int64_t num_attr, int rule, void* out,
                                    unsigned* status, cudaStream_t stream) {
  const int grid = grid_for(E, 256 * 4);
  const T* a = static_cast<const T*>(attr);
  T* o = static_cast<T*>(out);
  switch (rule) {
    case PPG_PAIR_SRC: pair_attributes_kernel<T, PPG_PAIR_SRC><<<grid, 256, 0, stream>>>(ei, E, a, num_attr, o, status); break;
    case PPG_PAIR_DST: pair_attributes_kernel<T, PPG_PAIR_DST><<<grid, 256, 0, stream>>>(ei, E, a, num_attr, o, status); break;
    case PPG_PAIR_MAX: pair_attributes_kernel<T, PPG_PAIR_MAX><<<grid, 256, 0, stream>>>(ei, E, a, num_attr, o, status); break;
    case PPG_PAIR_MUL: pair_attributes_kernel<T, PPG_PAIR_MUL><<<grid, 256, 0, stream>>>(ei, E, a, num_attr, o, status); break;
    case PPG_PAIR_ADD: pair_attributes_kernel<T, PPG_PAIR_ADD><<<grid, 256, 0, stream>>>(ei, E, a, num_attr, o, status); break;
    default: PPG_REQUIRE(false, PPG_ERR_INVALID, "Unknown aggregation method %d", rule);
  }
  PPG_LAUNCHED();
  return PPG_OK;
}

static int check_result(const ResultWords& h, const char* what) {
  PPG_REQUIRE((h.status & kStatusIdOutOfRange) == 0, PPG_ERR_INVALID, "%s: node id outside [0, num_nodes)", what);
  PPG_REQUIRE(h.total < (1ull << 62), PPG_ERR_INVALID, "%s: output size overflow", what);
  return PPG_OK;
}

}  // namespace ppg

using namespace ppg;

// =================================================================== a2
extern "C" size_t ppg_lift_order_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
  Workspace ws(nullptr, 0);
  LiftLayout L(ws, num_edges < 0 ? 0 : num_edges, num_nodes < 0 ? 0 : num_nodes);
  return ws.used + 256;
}

extern "C" int ppg_lift_order_count(const int64_t* edge_index, int64_t E, int64_t N, void* workspace,
                                    size_t workspace_bytes, int64_t* h_num_lifted, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(E >= 0 && N >= 0 && E < (1ll << 31) && N < (1ll << 31), PPG_ERR_INVALID,
              "lift_order: sizes E=%lld N=%lld outside [0, 2^31)", (long long)E, (long long)N);
  if (h_num_lifted != nullptr) *h_num_lifted = 0;
  if (E == 0) return PPG_OK;
  PPG_REQUIRE(N > 0, PPG_ERR_INVALID, "lift_order: num_nodes must be positive for a non-empty edge index");
  Workspace ws(workspace, workspace_bytes);
  LiftLayout L(ws, E, N);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "lift_order: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));

  degree_kernel<false><<<grid_for(E, 256 * 4), 256, 0, stream>>>(edge_index, E, N, L.deg, nullptr, &L.result->status, nullptr);
  PPG_LAUNCHED();
  PPG_TRY(launch_scan(DegreeProducer{L.deg}, PointerConsumer{L.ptr, N}, N, L.scan_ptr_ws, nullptr, stream));
  PPG_TRY(launch_scan(LiftCountProducer{edge_index + E, L.ptr, L.first, N, &L.result->status}, OffsetConsumer{L.off, E},
                      E, L.scan_off_ws, &L.result->total, stream));
  if (h_num_lifted == nullptr) return PPG_OK;  // deferred: the caller collects the count with ppg_result_read
  ResultWords h;
  PPG_TRY(read_back(&h, L.result, stream));
  PPG_TRY(check_result(h, "lift_order_edge_index"));
  *h_num_lifted = static_cast<int64_t>(h.total);
  return PPG_OK;
}

extern "C" int ppg_lift_order_fill(const void* workspace, int64_t E, int64_t N, int64_t num_lifted, int64_t* out_index,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_lifted == 0 || E == 0) return PPG_OK;
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  LiftLayout L(ws, E, N);
  return launch_expand(L.off, E, num_lifted, out_index, LiftTail{L.first}, stream);
}

// =================================================================== a3
extern "C" int ppg_pair_attributes(const int64_t* edge_index, int64_t E, const void* attr, int64_t num_attr, int dtype,
                                   int rule, void* out, void* status_word, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  unsigned* status = static_cast<unsigned*>(status_word);
  PPG_REQUIRE(rule >= PPG_PAIR_SRC && rule <= PPG_PAIR_ADD, PPG_ERR_INVALID, "Unknown aggregation method %d", rule);
  if (E == 0) return PPG_OK;
  PPG_REQUIRE(num_attr > 0, PPG_ERR_INVALID, "pair_attributes: edges present but the attribute tensor is empty");
  switch (dtype) {
    case PPG_F32: return pair_attributes_dispatch<float>(edge_index, E, attr, num_attr, rule, out, status, stream);
    case PPG_F64: return pair_attributes_dispatch<double>(edge_index, E, attr, num_attr, rule, out, status, stream);
    case PPG_I64: return pair_attributes_dispatch<long long>(edge_index, E, attr, num_attr, rule, out, status, stream);
    case PPG_I32: return pair_attributes_dispatch<int>(edge_index, E, attr, num_attr, rule, out, status, stream);
    default: PPG_REQUIRE(false, PPG_ERR_INVALID, "pair_attributes: unsupported dtype code %d", dtype);
  }
  return PPG_OK;
}

// =================================================================== a1
extern "C" size_t ppg_lift_temporal_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
  Workspace ws(nullptr, 0);
  TemporalLayout L(ws, num_edges < 0 ? 0 : num_edges, num_nodes < 0 ? 0 : num_nodes);
  return ws.used + 256;
}

// CSR over the source node, edges inside a group in time order (stable sort of a time-sorted stream).  Reads the
// SOURCE row only, so it can run while the target row and the time stamps are still on their way to the device.
static int temporal_group(TemporalLayout& L, const int64_t* src_row, int64_t m, int64_t N, void* workspace, cudaStream_t stream) {
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));
  degree_kernel<true><<<grid_for(m, 256 * 4), 256, 0, stream>>>(src_row, m, N, L.deg, L.keys_a, &L.result->status, L.sort_ws);
  PPG_LAUNCHED();
  PPG_TRY(launch_scan(DegreeProducer{L.deg}, PointerConsumer{L.ptr, N}, N, L.scan_ptr_ws, nullptr, stream));
  int in_b = 0;
  PPG_TRY(radix_sort_pairs<uint32_t>(L.keys_a, L.keys_b, L.vals_a, L.vals_b, true, true, m, L.sort_bits, L.sort_ws,
                                     &in_b, stream, nullptr, true));
  const uint32_t* grouped = in_b ? L.vals_b : L.vals_a;
  PPG_REQUIRE(grouped == L.grouped(), PPG_ERR_CUDA, "lift_order_temporal: internal buffer parity mismatch");
  return PPG_OK;
}

extern "C" int ppg_lift_temporal_group(const int64_t* src_row, int64_t m, int64_t N, void* workspace, size_t workspace_bytes,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(m > 0 && N > 0 && m < (1ll << 31) && N < (1ll << 31), PPG_ERR_INVALID,
              "lift_order_temporal: sizes m=%lld N=%lld outside (0, 2^31)", (long long)m, (long long)N);
  Workspace ws(workspace, workspace_bytes);
  TemporalLayout L(ws, m, N);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "lift_order_temporal: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  return temporal_group(L, src_row, m, N, workspace, stream);
}

extern "C" int ppg_lift_temporal_count(const int64_t* edge_index, const void* time, int64_t m, int64_t N, int time_mode,
                                       int64_t delta_i, double delta_f, void* workspace, size_t workspace_bytes,
                                       int64_t* h_num_pairs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool grouped_already = (time_mode & PPG_TIME_GROUPED) != 0;
  time_mode &= ~PPG_TIME_GROUPED;
  PPG_REQUIRE(m >= 0 && N >= 0 && m < (1ll << 31) && N < (1ll << 31), PPG_ERR_INVALID,
              "lift_order_temporal: sizes m=%lld N=%lld outside [0, 2^31)", (long long)m, (long long)N);
  PPG_REQUIRE(time_mode >= PPG_TIME_I64 && time_mode <= PPG_TIME_I64_F32DELTA, PPG_ERR_INVALID,
              "lift_order_temporal: unknown time mode %d", time_mode);
  if (h_num_pairs != nullptr) *h_num_pairs = 0;
  PPG_REQUIRE(m > 0 && N > 0, PPG_ERR_EMPTY, "lift_order_temporal: no time-respecting pair (empty input)");
  Workspace ws(workspace, workspace_bytes);
  TemporalLayout L(ws, m, N);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "lift_order_temporal: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  if (!grouped_already) PPG_TRY(temporal_group(L, edge_index, m, N, workspace, stream));
  const uint32_t* grouped = L.grouped();
  gather64_kernel<<<grid_for(m, 256 * 4), 256, 0, stream>>>(static_cast<const unsigned long long*>(time), grouped, m,
                                                             L.ts_sorted);
  PPG_LAUNCHED();

  switch (time_mode) {
    case PPG_TIME_I64: PPG_TRY(temporal_count_scan<PPG_TIME_I64>(L, edge_index, time, m, N, delta_i, delta_f, stream)); break;
    case PPG_TIME_F64: PPG_TRY(temporal_count_scan<PPG_TIME_F64>(L, edge_index, time, m, N, delta_i, delta_f, stream)); break;
    default: PPG_TRY(temporal_count_scan<PPG_TIME_I64_F32DELTA>(L, edge_index, time, m, N, delta_i, delta_f, stream)); break;
  }
  if (h_num_pairs == nullptr) return PPG_OK;  // deferred: ppg_result_read
  ResultWords h;
  PPG_TRY(read_back(&h, L.result, stream));
  PPG_TRY(check_result(h, "lift_order_temporal"));
  *h_num_pairs = static_cast<int64_t>(h.total);
  PPG_REQUIRE(h.total > 0, PPG_ERR_EMPTY, "lift_order_temporal: no time-respecting pair for this delta");
  return PPG_OK;
}

extern "C" int ppg_lift_temporal_fill(const void* workspace, int64_t m, int64_t N, int64_t num_pairs, int64_t* out_index,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_pairs == 0 || m == 0) return PPG_OK;
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  TemporalLayout L(ws, m, N);
  return launch_expand(L.off, m, num_pairs, out_index, TemporalTail{L.first, L.grouped()}, stream);
}

// Device arrays a grouped / counted temporal workspace holds, for the layer chain (chain.cu):
// out[0] ptr u32 [N + 1] (CSR over the source node), out[1] grouped u32 [m] (event at every grouped position),
// out[2] source node of every grouped position u32 [m], out[3] first u32 [m] (grouped position of an event's first
// continuation), out[4] off u64 [m + 1] (row pointer of the event graph), out[5] {total, status}
extern "C" int ppg_lift_temporal_views(const void* workspace, int64_t m, int64_t N, const void** out) {
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  TemporalLayout L(ws, m, N);
  const bool odd = (sort_num_passes(L.sort_bits) & 1) != 0;
  out[0] = L.ptr;
  out[1] = L.grouped();
  out[2] = odd ? L.keys_b : L.keys_a;
  out[3] = L.first;
  out[4] = L.off;
  out[5] = L.result;
  return PPG_OK;
}

// =================================================================== prefix-limited lifts
// The columns of a lift are ascending in the source, so "the lift of the first `limit` sources only" is a prefix of the
// full output: off[limit] columns.  The distributed lift (exchange.cu) needs only the paths that start before a cut.
__global__ void lift_limit_kernel(const unsigned long long* __restrict__ off, int64_t limit, ppg::ResultWords* result) {
  result->total = off[limit];
}

extern "C" int ppg_lift_limit(void* workspace, int temporal, int64_t num_sources, int64_t num_nodes, int64_t limit_sources,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(limit_sources >= 0 && limit_sources <= num_sources, PPG_ERR_INVALID, "lift_limit: %lld outside [0, %lld]",
              (long long)limit_sources, (long long)num_sources);
  if (num_sources == 0) return PPG_OK;
  Workspace ws(workspace, ~static_cast<size_t>(0));
  ResultWords* result;
  const unsigned long long* off;
  if (temporal) {
    TemporalLayout L(ws, num_sources, num_nodes);
    result = L.result;
    off = L.off;
  } else {
    LiftLayout L(ws, num_sources, num_nodes);
    result = L.result;
    off = L.off;
  }
  lift_limit_kernel<<<1, 1, 0, stream>>>(off, limit_sources, result);
  PPG_LAUNCHED();
  return PPG_OK;
}

// =================================================================== deferred count read-back
// Every workspace of a count -> fill pair starts with {total, status}.  A *_count / *_sort call with a NULL
// host pointer only enqueues its kernels; several such calls can be in flight on the stream before the
// host collects their results here with one synchronisation.
extern "C" int ppg_result_read(const void* workspace, int64_t* h_total, int* h_status_bits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ResultWords h;
  PPG_TRY(read_back(&h, static_cast<const ResultWords*>(workspace), stream));
  *h_total = static_cast<int64_t>(h.total);
  *h_status_bits = static_cast<int>(h.status);
  PPG_REQUIRE(h.total < (1ull << 62), PPG_ERR_INVALID, "result_read: output size overflow");
  return PPG_OK;
}
