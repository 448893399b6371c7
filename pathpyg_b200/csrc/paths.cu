// Shortest time-respecting paths between all first-order nodes: a consumer of the event
// graph a1 produces.  Reference: src/pathpyG/algorithms/temporal.py:57-107 (scipy dijkstra, unweighted, from
// every first-order node over the event DAG augmented with source and destination nodes).
//
// Here: bit-parallel multi-source breadth-first search over the event graph.  Sources are packed 32 per word;
// every event carries `words` frontier / visited / next words.  One level =
//   push   : one thread per event-graph edge (e -> f) ORs e's frontier words into f's `next` words
//            (integer atomics: order-independent), skipping events that are not in the frontier;
//   settle : one thread per (event, word): new bits = next & ~visited become the next frontier; every new bit
//            (source s reached event f for the first time, after `level` events) offers
//            (level << 32 | ~f) to best[s][dst(f)] with a 64-bit atomicMin.
// So best[s][v] ends up holding the smallest number of events on a time-respecting path s -> v and, among the
// events that end such a path, the one with the LARGEST index -- the predecessor scipy's dijkstra reports for the
// reference's augmented graph (checked against scipy on random instances in tests/).
// Traffic per level: E2 * (8 + words * 4) B read by push, m * words * 12 B by settle; levels = longest shortest path.
#include "common.cuh"

namespace ppg {

constexpr int kPathBlock = 256;
constexpr unsigned long long kUnreached = ~0ull;

__global__ void __launch_bounds__(kPathBlock)
msbfs_init_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t m, int64_t n, int64_t s0,
                  int64_t s1, int words, uint32_t* __restrict__ frontier, uint32_t* __restrict__ visited,
                  unsigned char* __restrict__ active, unsigned long long* __restrict__ best,
                  unsigned long long* __restrict__ flags) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < m; e += stride) {
    const int64_t s = src[e], v = dst[e];
    if (s < 0 || s >= n || v < 0 || v >= n) {
      atomicOr(flags + 1, 1ull);
      continue;
    }
    if (s >= s0 && s < s1) {
      const int64_t local = s - s0;
      const uint32_t bit = 1u << (local & 31);
      frontier[e * words + (local >> 5)] = bit;   // the words of an event are owned by this thread; they were zeroed
      visited[e * words + (local >> 5)] = bit;
      active[e] = 1;
      atomicMin(best + local * n + v, (1ull << 32) | (0xffffffffull - static_cast<unsigned long long>(e)));
      flags[0] = 1ull;  // the frontier is not empty
    }
  }
}

__global__ void __launch_bounds__(kPathBlock)
msbfs_push_kernel(const int64_t* __restrict__ from, const int64_t* __restrict__ to, int64_t num_pairs, int words,
                  const uint32_t* __restrict__ frontier, const unsigned char* __restrict__ active,
                  uint32_t* __restrict__ next) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < num_pairs; j += stride) {
    const int64_t e = ld_stream(from + j);
    if (!active[e]) continue;
    const int64_t f = ld_stream(to + j);
    const uint32_t* fe = frontier + e * words;
    uint32_t* nf = next + f * words;
    for (int w = 0; w < words; ++w) {
      const uint32_t bits = fe[w];
      if (bits) atomicOr(nf + w, bits);
    }
  }
}

__global__ void __launch_bounds__(kPathBlock)
msbfs_settle_kernel(const int64_t* __restrict__ dst, int64_t m, int64_t n, int words, unsigned level,
                    uint32_t* __restrict__ frontier, uint32_t* __restrict__ visited, uint32_t* __restrict__ next,
                    unsigned char* __restrict__ active, unsigned long long* __restrict__ best,
                    unsigned long long* __restrict__ any) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  bool found = false;
  for (int64_t f = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; f < m; f += stride) {
    bool live = false;
    const int64_t v = dst[f];
    for (int w = 0; w < words; ++w) {
      const int64_t i = f * words + w;
      const uint32_t offered = next[i];
      uint32_t fresh = 0;
      if (offered) {
        next[i] = 0;
        const uint32_t seen = visited[i];
        fresh = offered & ~seen;
        if (fresh) visited[i] = seen | fresh;
      }
      frontier[i] = fresh;
      if (fresh) {
        live = true;
        const unsigned long long offer = (static_cast<unsigned long long>(level) << 32) |
                                         (0xffffffffull - static_cast<unsigned long long>(f));
        uint32_t bits = fresh;
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          atomicMin(best + (static_cast<int64_t>(w) * 32 + b) * n + v, offer);
        }
      }
    }
    active[f] = live ? 1 : 0;
    found |= live;
  }
  if (__syncthreads_or(found) && threadIdx.x == 0) *any = 1ull;
}

// dist [rows, n] float64 (inf where unreachable, 0 on the diagonal), pred [rows, n] int64 (-1 / the node itself)
__global__ void __launch_bounds__(kPathBlock)
msbfs_finalize_kernel(const unsigned long long* __restrict__ best, const int64_t* __restrict__ src, int64_t rows, int64_t n,
                      int64_t s0, double* __restrict__ dist, int64_t* __restrict__ pred) {
  const int64_t total = rows * n;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t s = s0 + i / n, v = i % n;
    const unsigned long long b = best[i];
    if (s == v) {
      dist[i] = 0.0;
      pred[i] = s;
    } else if (b == kUnreached) {
      dist[i] = __longlong_as_double(0x7ff0000000000000ll);
      pred[i] = -1;
    } else {
      dist[i] = static_cast<double>(b >> 32);
      pred[i] = src[0xffffffffull - (b & 0xffffffffull)];
    }
  }
}

// closeness[v] = sum_{x != v} (n - 1) / dist[x, v], added in ascending x like Python's sum() over the column
// (reference centrality.py:322); one thread per column, consecutive threads read consecutive addresses
__global__ void __launch_bounds__(kPathBlock)
closeness_kernel(const double* __restrict__ dist, int64_t n, double* __restrict__ out) {
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const double scale = static_cast<double>(n - 1);
  double acc = 0.0;
  for (int64_t x = 0; x < n; ++x)
    if (x != v) acc += scale / dist[x * n + v];
  out[v] = acc;
}

__device__ __forceinline__ double block_sum_f64_paths(double x, double* s_part) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(kFullMask, x, d);
  if (lane_id() == 0) s_part[threadIdx.x >> 5] = x;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += s_part[w];  // every thread: same order, same value
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------ temporal betweenness (Brandes on the event DAG)
// Reference: src/pathpyG/algorithms/centrality.py:164-300.  One CTA per source node.  The event graph is a DAG whose
// edges run forward in time, and events sharing a time stamp never continue each other, so the time groups of the
// time-sorted event list are the levels of a topological order:
//   forward  (groups ascending) : dist[f] = fewest events from the source, sigma[f] = number of such shortest paths,
//                                 pulled from the predecessors of f (CSC of the event graph);
//   per node                    : dist_fo[x] = min dist over the events ending in x, sigma_fo[x] = their path count;
//   backward (groups descending): delta[v] = own(v) + sum over successors w one level deeper of
//                                 sigma[v] / sigma[w] * delta[w]   (CSR of the event graph),
//                                 own(v) = sigma[v] / sigma_fo[dst v] if v ends a shortest path to its node;
//   contrib[x] = sum of the pulled parts of the events ending in x (+ the source's own share and 1 - #reachable).
// Every sum runs over a fixed list in a fixed order: no floating-point atomics.
constexpr int kUnreachedDist = 0x3fffffff;
constexpr int kBrandesBlock = 256;

__global__ void __launch_bounds__(kBrandesBlock)
temporal_brandes_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t m, int64_t n,
                        const int32_t* __restrict__ group_off, int num_groups,
                        const int32_t* __restrict__ succ_ptr, const int64_t* __restrict__ succ,      // CSR of the event graph
                        const int32_t* __restrict__ pred_ptr, const int32_t* __restrict__ pred,      // CSC of the event graph
                        const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_event,    // events grouped by dst
                        const int32_t* __restrict__ sources, int batch,
                        int32_t* __restrict__ dist_all, double* __restrict__ sigma_all, double* __restrict__ delta_all,
                        double* __restrict__ pull_all, int32_t* __restrict__ dist_fo_all, double* __restrict__ sigma_fo_all,
                        double* __restrict__ contrib_all) {
  __shared__ double s_part[kBrandesBlock / 32];
  const int b = blockIdx.x;
  if (b >= batch) return;
  const int64_t s = sources[b];
  int32_t* dist = dist_all + static_cast<int64_t>(b) * m;
  double* sigma = sigma_all + static_cast<int64_t>(b) * m;
  double* delta = delta_all + static_cast<int64_t>(b) * m;
  double* pull = pull_all + static_cast<int64_t>(b) * m;
  int32_t* dist_fo = dist_fo_all + static_cast<int64_t>(b) * n;
  double* sigma_fo = sigma_fo_all + static_cast<int64_t>(b) * n;
  double* contrib = contrib_all + static_cast<int64_t>(b) * n;
  const int tid = threadIdx.x;

  // ---- forward: shortest distances and path counts, one time group after the other
  for (int g = 0; g < num_groups; ++g) {
    for (int32_t f = group_off[g] + tid; f < group_off[g + 1]; f += kBrandesBlock) {
      const bool first = src[f] == s;
      int32_t d = first ? 1 : kUnreachedDist;
      const int32_t p0 = pred_ptr[f], p1 = pred_ptr[f + 1];
      for (int32_t i = p0; i < p1; ++i) d = min(d, dist[pred[i]] + 1);
      double sg = 0.0;
      if (d < kUnreachedDist) {
        if (first && d == 1) sg = 1.0;
        for (int32_t i = p0; i < p1; ++i) {
          const int32_t e = pred[i];
          if (dist[e] + 1 == d) sg += sigma[e];
        }
      } else {
        d = kUnreachedDist;
      }
      dist[f] = d;
      sigma[f] = sg;
    }
    __syncthreads();
  }

  // ---- first-order nodes: distance and number of shortest paths, number of reachable nodes
  int reach = 0;
  for (int64_t x = tid; x < n; x += kBrandesBlock) {
    int32_t d = kUnreachedDist;
    double sg = 0.0;
    if (x == s) {
      d = 0;
      sg = 1.0;
    } else {
      const int32_t i0 = in_ptr[x], i1 = in_ptr[x + 1];
      for (int32_t i = i0; i < i1; ++i) d = min(d, dist[in_event[i]]);
      if (d < kUnreachedDist)
        for (int32_t i = i0; i < i1; ++i) {
          const int32_t w = in_event[i];
          if (dist[w] == d) sg += sigma[w];
        }
    }
    dist_fo[x] = d;
    sigma_fo[x] = sg;
    reach += d < kUnreachedDist;
  }
  __syncthreads();

  // ---- backward: dependencies, one time group after the other
  for (int g = num_groups - 1; g >= 0; --g) {
    for (int32_t v = group_off[g] + tid; v < group_off[g + 1]; v += kBrandesBlock) {
      const int32_t dv = dist[v];
      double p = 0.0, own = 0.0;
      if (dv < kUnreachedDist) {
        const double sv = sigma[v];
        for (int32_t i = succ_ptr[v]; i < succ_ptr[v + 1]; ++i) {
          const int64_t w = succ[i];
          if (dist[w] == dv + 1) p += sv / sigma[w] * delta[w];
        }
        const int64_t x = dst[v];
        if (x != s && dv == dist_fo[x]) own = sv / sigma_fo[x];
      }
      pull[v] = p;
      delta[v] = own + p;
    }
    __syncthreads();
  }

  // ---- contributions of this source to every node
  for (int64_t x = tid; x < n; x += kBrandesBlock) {
    double c = 0.0;
    for (int32_t i = in_ptr[x]; i < in_ptr[x + 1]; ++i) {
      const int32_t v = in_event[i];
      if (dist[v] < kUnreachedDist) c += pull[v];
    }
    contrib[x] = c;
  }
  // the source itself: its first events (dist 1, sigma 1) hand their dependency back, minus the reachable others
  double own_share = 0.0;
  for (int64_t e = tid; e < m; e += kBrandesBlock)
    if (src[e] == s) own_share += delta[e] / sigma[e];
  const double total = block_sum_f64_paths(own_share, s_part);
  __shared__ int s_reach[kBrandesBlock / 32];
  int r = reach;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(kFullMask, r, d);
  if (lane_id() == 0) s_reach[tid >> 5] = r;
  __syncthreads();
  if (tid == 0) {
    int rt = 0;
    for (int w = 0; w < kBrandesBlock / 32; ++w) rt += s_reach[w];
    contrib[s] += total + 1.0 - static_cast<double>(rt);
  }
}

// bw[x] += contrib[b][x] for b = 0 .. batch-1 in that order (sources ascending: a fixed order of additions)
__global__ void __launch_bounds__(kPathBlock)
brandes_accumulate_kernel(const double* __restrict__ contrib, int batch, int64_t n, double* __restrict__ bw) {
  const int64_t x = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (x >= n) return;
  double acc = bw[x];
  for (int b = 0; b < batch; ++b) acc += contrib[static_cast<int64_t>(b) * n + x];
  bw[x] = acc;
}

struct PathLayout {
  unsigned long long* flags;  // [0] frontier not empty, [1] status
  unsigned long long* best;   // [chunk, n]
  uint32_t* frontier;         // [m, words]
  uint32_t* visited;
  uint32_t* next;
  unsigned char* active;      // [m]
  PathLayout(Workspace& ws, int64_t m, int64_t n, int64_t chunk) {
    const size_t words = static_cast<size_t>(ceil_div(chunk, 32));
    flags = ws.take<unsigned long long>(2);
    best = ws.take<unsigned long long>(static_cast<size_t>(chunk) * n);
    frontier = ws.take<uint32_t>(static_cast<size_t>(m) * words);
    visited = ws.take<uint32_t>(static_cast<size_t>(m) * words);
    next = ws.take<uint32_t>(static_cast<size_t>(m) * words);
    active = ws.take<unsigned char>(static_cast<size_t>(m));
  }
};

}  // namespace ppg

using namespace ppg;

extern "C" size_t ppg_temporal_paths_workspace_bytes(int64_t num_events, int64_t num_nodes, int64_t chunk_sources) {
  Workspace ws(nullptr, 0);
  PathLayout layout(ws, num_events, num_nodes, chunk_sources);
  (void)layout;
  return ws.used + 256;
}

extern "C" int ppg_temporal_paths(const int64_t* edge_index, int64_t num_events, int64_t num_nodes,
                                  const int64_t* event_graph, int64_t num_pairs, int64_t source_begin, int64_t source_end,
                                  void* workspace, size_t workspace_bytes, double* out_dist, int64_t* out_pred,
                                  int* h_levels, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t m = num_events, n = num_nodes, chunk = source_end - source_begin;
  PPG_REQUIRE(m >= 0 && m < (1ll << 31) && n > 0 && chunk > 0 && source_begin >= 0 && source_end <= n, PPG_ERR_INVALID,
              "temporal_paths: bad sizes (m=%lld, n=%lld, sources [%lld, %lld))", (long long)m, (long long)n,
              (long long)source_begin, (long long)source_end);
  PPG_REQUIRE(source_begin % 32 == 0, PPG_ERR_INVALID, "temporal_paths: source_begin must be a multiple of 32");
  Workspace ws(workspace, workspace_bytes);
  PathLayout L(ws, m, n, chunk);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "temporal_paths: workspace of %zu bytes is too small (%zu needed)", workspace_bytes, ws.used);
  const int words = static_cast<int>(ceil_div(chunk, 32));
  const size_t state_bytes = static_cast<size_t>(m) * words * sizeof(uint32_t);
  PPG_CUDA_TRY(cudaMemsetAsync(L.flags, 0, 2 * sizeof(unsigned long long), stream));
  PPG_CUDA_TRY(cudaMemsetAsync(L.best, 0xff, static_cast<size_t>(chunk) * n * sizeof(unsigned long long), stream));
  if (m > 0) {
    PPG_CUDA_TRY(cudaMemsetAsync(L.frontier, 0, state_bytes, stream));
    PPG_CUDA_TRY(cudaMemsetAsync(L.visited, 0, state_bytes, stream));
    PPG_CUDA_TRY(cudaMemsetAsync(L.next, 0, state_bytes, stream));
    PPG_CUDA_TRY(cudaMemsetAsync(L.active, 0, static_cast<size_t>(m), stream));
    msbfs_init_kernel<<<grid_for(m, kPathBlock), kPathBlock, 0, stream>>>(edge_index, edge_index + m, m, n, source_begin,
                                                                          source_end, words, L.frontier, L.visited, L.active,
                                                                          L.best, L.flags);
    PPG_LAUNCHED();
  }
  unsigned long long h_flags[2] = {0, 0};
  PPG_CUDA_TRY(cudaMemcpyAsync(h_flags, L.flags, sizeof(h_flags), cudaMemcpyDeviceToHost, stream));
  PPG_CUDA_TRY(cudaStreamSynchronize(stream));
  PPG_REQUIRE(h_flags[1] == 0, PPG_ERR_INVALID, "temporal_paths: node id outside [0, %lld)", (long long)n);
  unsigned level = 1;  // events on the path so far
  while (h_flags[0] != 0 && num_pairs > 0) {
    PPG_REQUIRE(level < 0xfffffffeu, PPG_ERR_INVALID, "temporal_paths: path length overflow");
    PPG_CUDA_TRY(cudaMemsetAsync(L.flags, 0, sizeof(unsigned long long), stream));
    msbfs_push_kernel<<<grid_for(num_pairs, kPathBlock), kPathBlock, 0, stream>>>(event_graph, event_graph + num_pairs,
                                                                                  num_pairs, words, L.frontier, L.active, L.next);
    PPG_LAUNCHED();
    ++level;
    msbfs_settle_kernel<<<grid_for(m, kPathBlock), kPathBlock, 0, stream>>>(edge_index + m, m, n, words, level, L.frontier,
                                                                            L.visited, L.next, L.active, L.best, L.flags);
    PPG_LAUNCHED();
    PPG_CUDA_TRY(cudaMemcpyAsync(h_flags, L.flags, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    PPG_CUDA_TRY(cudaStreamSynchronize(stream));
  }
  msbfs_finalize_kernel<<<grid_for(chunk * n, kPathBlock * 4), kPathBlock, 0, stream>>>(L.best, edge_index, chunk, n,
                                                                                         source_begin, out_dist, out_pred);
  PPG_LAUNCHED();
  if (h_levels != nullptr) *h_levels = static_cast<int>(level);
  return PPG_OK;
}

extern "C" int ppg_temporal_closeness(const double* dist, int64_t num_nodes, double* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_nodes == 0) return PPG_OK;
  closeness_kernel<<<static_cast<unsigned>(ceil_div(num_nodes, kPathBlock)), kPathBlock, 0, stream>>>(dist, num_nodes, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" size_t ppg_temporal_betweenness_workspace_bytes(int64_t num_events, int64_t num_nodes, int64_t batch_sources) {
  Workspace ws(nullptr, 0);
  const size_t B = static_cast<size_t>(batch_sources), m = static_cast<size_t>(num_events), n = static_cast<size_t>(num_nodes);
  ws.take<int32_t>(B * m);
  ws.take<double>(B * m);
  ws.take<double>(B * m);
  ws.take<double>(B * m);
  ws.take<int32_t>(B * n);
  ws.take<double>(B * n);
  ws.take<double>(B * n);
  return ws.used + 256;
}

extern "C" int ppg_temporal_betweenness(const int64_t* edge_index, int64_t num_events, int64_t num_nodes,
                                        const int32_t* group_off, int64_t num_groups, const int32_t* succ_ptr,
                                        const int64_t* succ, const int32_t* pred_ptr, const int32_t* pred,
                                        const int32_t* in_ptr, const int32_t* in_event, const int32_t* sources,
                                        int64_t batch_sources, void* workspace, size_t workspace_bytes, double* inout_bw,
                                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t m = num_events, n = num_nodes, B = batch_sources;
  PPG_REQUIRE(m > 0 && m < (1ll << 31) && n > 0 && n < (1ll << 31) && B > 0 && B < (1 << 20) && num_groups > 0 &&
                  num_groups < (1ll << 31),
              PPG_ERR_INVALID, "temporal_betweenness: bad sizes (m=%lld, n=%lld, batch=%lld)", (long long)m, (long long)n, (long long)B);
  Workspace ws(workspace, workspace_bytes);
  int32_t* dist = ws.take<int32_t>(static_cast<size_t>(B * m));
  double* sigma = ws.take<double>(static_cast<size_t>(B * m));
  double* delta = ws.take<double>(static_cast<size_t>(B * m));
  double* pull = ws.take<double>(static_cast<size_t>(B * m));
  int32_t* dist_fo = ws.take<int32_t>(static_cast<size_t>(B * n));
  double* sigma_fo = ws.take<double>(static_cast<size_t>(B * n));
  double* contrib = ws.take<double>(static_cast<size_t>(B * n));
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "temporal_betweenness: workspace of %zu bytes is too small (%zu needed)", workspace_bytes, ws.used);
  temporal_brandes_kernel<<<static_cast<unsigned>(B), kBrandesBlock, 0, stream>>>(
      edge_index, edge_index + m, m, n, group_off, static_cast<int>(num_groups), succ_ptr, succ, pred_ptr, pred, in_ptr,
      in_event, sources, static_cast<int>(B), dist, sigma, delta, pull, dist_fo, sigma_fo, contrib);
  PPG_LAUNCHED();
  brandes_accumulate_kernel<<<static_cast<unsigned>(ceil_div(n, kPathBlock)), kPathBlock, 0, stream>>>(contrib, static_cast<int>(B), n, inout_bw);
  PPG_LAUNCHED();
  return PPG_OK;
}
