// a10 DBGNN.forward / a11 BipartiteGraphOperator building blocks.
//
// GCNConv and the bipartite operator are "gather rows by source, sum at the target".  The
// reference scatters (index_add_ over an edge list); here the edge list is regrouped ONCE per graph
// by target (stable radix sort => CSC view), after which every layer is a segment reduction:
// one lane group per target node walks its incoming edges, gathers the source rows with 16-byte
// loads and writes the output row exactly once -- no atomics, deterministic summation order
// (original edge order inside a target, self-loop last, like the reference's edge list).
//
//   gcn_norm   : deg[v] = sum_{u->v, u!=v} w + loop(v);  val = deg^-1/2[u] * w * deg^-1/2[v]
//                (PyG add_remaining_self_loops: an existing self-loop keeps its weight, others get 1)
//   spmm_csc   : out[v,:] = act( sum_i val_i * X[src_i,:] + self_v * X[v,:] + bias )
//   linear     : out = act( A1 W1^T + rowscale * (A2 W2^T + bias) )     (fp32 FMA, tiled)
//
// Algorithmic bytes per GCN layer: 20*e_sl + 4*H*e_sl + 8*H*n + 4*H*n.
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace ppg {

constexpr unsigned kStatusIdOutOfRange = 1u;

struct ResultWords {
  unsigned long long total;
  unsigned long long status;
};

// ------------------------------------------------------------------ CSC view of an edge list
struct CscLayout {
  ResultWords* result;
  uint32_t* deg;
  unsigned long long* scan_ws;
  unsigned long long* sort_ws;
  size_t zero_bytes;
  uint32_t *keys_a, *keys_b, *vals_a, *vals_b;
  int sort_bits;
  CscLayout(Workspace& ws, int64_t E, int64_t num_targets) {
    sort_bits = bits_for(num_targets > 0 ? static_cast<uint64_t>(num_targets - 1) : 0);
    result = ws.take<ResultWords>(1);
    deg = ws.take<uint32_t>(static_cast<size_t>(num_targets));
    scan_ws = ws.take<unsigned long long>(scan_state_words(num_targets));
    sort_ws = ws.take<unsigned long long>(sort_state_words(E, sort_bits));
    zero_bytes = ws.used;
    keys_a = ws.take<uint32_t>(static_cast<size_t>(E));
    keys_b = ws.take<uint32_t>(static_cast<size_t>(E));
    vals_a = ws.take<uint32_t>(static_cast<size_t>(E));
    vals_b = ws.take<uint32_t>(static_cast<size_t>(E));
  }
};

__global__ void __launch_bounds__(256)
target_keys_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t num_sources, int64_t num_targets,
                   uint32_t* __restrict__ keys, uint32_t* __restrict__ deg, unsigned long long* __restrict__ status,
                   unsigned long long* __restrict__ ghist0) {
  __shared__ unsigned s_hist[kRadix];
  Digit0Counter digit0;
  digit0.begin(s_hist);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x; base < E; base += stride) {
    const int64_t j = base + threadIdx.x;
    const bool valid = j < E;
    uint32_t key = 0;
    if (valid) {
      const int64_t r = ld_stream(ei + j);
      const int64_t c = ld_stream(ei + E + j);
      const bool ok = r >= 0 && r < num_sources && c >= 0 && c < num_targets;
      if (!ok) atomicOr(reinterpret_cast<unsigned*>(status), kStatusIdOutOfRange);
      key = ok ? static_cast<uint32_t>(c) : 0u;
      keys[j] = key;
      if (ok) atomicAdd(&deg[c], 1u);
    }
    digit0.count(key & (kRadix - 1), valid);
  }
  digit0.end(ghist0);
}

struct DegreeProducer32 {
  const uint32_t* deg;
  __device__ unsigned long long operator()(int64_t i) const { return deg[i]; }
};
struct PointerConsumer32 {
  int32_t* ptr;
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long v, unsigned long long prefix) const {
    ptr[i] = static_cast<int32_t>(prefix);
    if (i == n - 1) ptr[n] = static_cast<int32_t>(prefix + v);
  }
};

__global__ void __launch_bounds__(256)
csc_fill_kernel(const int64_t* __restrict__ ei, const uint32_t* __restrict__ perm, int64_t E, int64_t num_sources,
                int32_t* __restrict__ src, int32_t* __restrict__ eid) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < E; i += stride) {
    const uint32_t e = perm[i];
    eid[i] = static_cast<int32_t>(e);
    // an out-of-range source id was flagged by target_keys_kernel; it is clamped here so that consumers running
    // before a deferred status check stay inside their arrays
    const int64_t r = ei[e];
    src[i] = r >= 0 && r < num_sources ? static_cast<int32_t>(r) : 0;
  }
}

// ------------------------------------------------------------------ GCN normalisation on the CSC view
__global__ void __launch_bounds__(256)
gcn_degree_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ src, const int32_t* __restrict__ eid,
                  const float* __restrict__ w, int64_t n, float* __restrict__ dis, float* __restrict__ loop_w) {
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= n) return;
  float deg = 0.f, lw = 1.f;
  for (int32_t i = colptr[v]; i < colptr[v + 1]; ++i) {
    const float x = w ? w[eid[i]] : 1.f;
    if (src[i] == v) lw = x;  // existing self-loop keeps its weight (the last one wins)
    else deg += x;
  }
  deg += lw;  // the loop is appended after the other edges in the reference's edge list
  float d = 1.0f / sqrtf(deg);  // deg.pow(-0.5) as torch evaluates it (IEEE sqrt and divide)
  if (isinf(d)) d = 0.f;
  dis[v] = d;
  loop_w[v] = lw;
}

__global__ void __launch_bounds__(256)
gcn_values_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ src, const int32_t* __restrict__ eid,
                  const float* __restrict__ w, const float* __restrict__ dis, int64_t n, float* __restrict__ val,
                  float* __restrict__ self_val /* in: loop weight, out: normalised loop weight */,
                  float* __restrict__ val_edge /* nullable: the same coefficient by ORIGINAL edge id */) {
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const float dv = dis[v];
  for (int32_t i = colptr[v]; i < colptr[v + 1]; ++i) {
    const int32_t u = src[i];
    const float x = w ? w[eid[i]] : 1.f;
    const float c = (u == v) ? 0.f : dis[u] * x * dv;
    val[i] = c;
    if (val_edge != nullptr) val_edge[eid[i]] = c;
  }
  self_val[v] = dv * self_val[v] * dv;
}

__global__ void __launch_bounds__(256)
colptr_counts_kernel(const int32_t* __restrict__ colptr, int64_t n, float* __restrict__ out) {
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v < n) out[v] = static_cast<float>(colptr[v + 1] - colptr[v]);
}

// ------------------------------------------------------------------ activation
__device__ __forceinline__ float activate(float x, int act) {
  return (act == PPG_ACT_ELU && x <= 0.f) ? expm1f(x) : x;
}

// ------------------------------------------------------------------ segment-reduce SpMM
// LPN lanes cooperate on one target node; every lane owns VEC consecutive floats of a
// (LPN * VEC)-wide feature block and loops over the blocks if F is wider.
template <int LPN, int VEC>
__global__ void __launch_bounds__(256)
spmm_csc_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                const float* __restrict__ self_val, const float* __restrict__ X, int64_t n, int F,
                const float* __restrict__ bias, int act, float* __restrict__ out) {
  const int64_t group = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPN;
  const int g = threadIdx.x % LPN;
  if (group >= n) return;
  const int64_t v = group;
  const int32_t a = colptr[v];
  const int32_t b = colptr[v + 1];
  const float sv = self_val ? self_val[v] : 0.f;
  for (int f0 = g * VEC; f0 < F; f0 += LPN * VEC) {
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    int32_t i = a;
    for (; i + 4 <= b; i += 4) {  // four independent row gathers in flight
      int32_t s[4];
      float c[4];
      float x[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[u] = src[i + u];
        c[u] = val ? val[i + u] : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* row = X + static_cast<int64_t>(s[u]) * F + f0;
        if (VEC == 4) {
          const float4 q = *reinterpret_cast<const float4*>(row);
          x[u][0] = q.x; x[u][1 % VEC] = q.y; x[u][2 % VEC] = q.z; x[u][3 % VEC] = q.w;
        } else if (VEC == 2) {
          const float2 q = *reinterpret_cast<const float2*>(row);
          x[u][0] = q.x; x[u][1 % VEC] = q.y;
        } else {
          x[u][0] = row[0];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = fmaf(c[u], x[u][j], acc[j]);
    }
    for (; i < b; ++i) {
      const float c = val ? val[i] : 1.f;
      const float* row = X + static_cast<int64_t>(src[i]) * F + f0;
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = fmaf(c, row[j], acc[j]);
    }
    if (self_val) {
      const float* row = X + v * F + f0;
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = fmaf(sv, row[j], acc[j]);
    }
    float* o = out + v * F + f0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float r = acc[j];
      if (bias) r += bias[f0 + j];
      o[j] = activate(r, act);
    }
  }
}

// ------------------------------------------------------------------ tiled fp32 linear layer
// out[M,N] = act( A1[M,K1] W1[N,K1]^T + rowscale[M] * (A2[M,K2] W2[N,K2]^T + bias[N]) )
constexpr int kLinBM = 64, kLinBN = 64, kLinBK = 16;

__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ A1, const float* __restrict__ W1, int K1, const float* __restrict__ A2,
              const float* __restrict__ W2, int K2, const float* __restrict__ bias, const float* __restrict__ rowscale,
              int64_t M, int N, int act, float* __restrict__ out) {
  __shared__ float sA[kLinBK][kLinBM + 4];
  __shared__ float sW[kLinBK][kLinBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each a 4 x 4 micro tile
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kLinBM;
  const int n0 = blockIdx.y * kLinBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int part = 0; part < 2; ++part) {
    const float* A = part == 0 ? A1 : A2;
    const float* W = part == 0 ? W1 : W2;
    const int K = part == 0 ? K1 : K2;
    if (A == nullptr || K == 0) continue;
    for (int k0 = 0; k0 < K; k0 += kLinBK) {
      // 64 x 16 tile of A and of W: thread loads 4 elements of each (consecutive k inside a row)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int lin = tid + t * 256;
        const int r = lin / kLinBK, k = lin % kLinBK;
        const int64_t gm = m0 + r;
        const int gk = k0 + k;
        float a = 0.f, w = 0.f;
        if (gm < M && gk < K) {
          a = A[gm * K + gk];
          if (part == 1 && rowscale) a *= rowscale[gm];
        }
        if (n0 + r < N && gk < K) w = W[static_cast<int64_t>(n0 + r) * K + gk];
        sA[k][r] = a;
        sW[k][r] = w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kLinBK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
        const float4 w4 = *reinterpret_cast<const float4*>(&sW[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
    const float rs = rowscale ? rowscale[gm] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float r = acc[i][j];
      if (bias) r += rs * bias[gn];
      out[gm * N + gn] = activate(r, act);
    }
  }
}

}  // namespace ppg

using namespace ppg;

extern "C" size_t ppg_csc_workspace_bytes(int64_t num_edges, int64_t num_targets) {
  Workspace ws(nullptr, 0);
  CscLayout L(ws, num_edges < 0 ? 0 : num_edges, num_targets < 0 ? 0 : num_targets);
  return ws.used + 256;
}

extern "C" int ppg_csc_build_async(const int64_t* edge_index, int64_t E, int64_t num_sources, int64_t num_targets,
                             void* workspace, size_t workspace_bytes, int32_t* out_colptr, int32_t* out_src,
                             int32_t* out_eid, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(E >= 0 && E < (1ll << 31) && num_targets >= 0 && num_targets < (1ll << 31) && num_sources >= 0 &&
                  num_sources < (1ll << 31),
              PPG_ERR_INVALID, "csc_build: sizes outside [0, 2^31)");
  Workspace ws(workspace, workspace_bytes);
  CscLayout L(ws, E, num_targets);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "csc_build: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));
  if (num_targets == 0) {
    PPG_CUDA_TRY(cudaMemsetAsync(out_colptr, 0, sizeof(int32_t), stream));
    return PPG_OK;
  }
  if (E > 0) {
    target_keys_kernel<<<grid_for(E, 256 * 4), 256, 0, stream>>>(edge_index, E, num_sources, num_targets, L.keys_a, L.deg,
                                                                 &L.result->status, L.sort_ws);
    PPG_LAUNCHED();
  }
  PPG_TRY(launch_scan(DegreeProducer32{L.deg}, PointerConsumer32{out_colptr, num_targets}, num_targets, L.scan_ws,
                      nullptr, stream));
  if (E > 0) {
    int in_b = 0;
    PPG_TRY(radix_sort_pairs<uint32_t>(L.keys_a, L.keys_b, L.vals_a, L.vals_b, true, true, E, L.sort_bits, L.sort_ws,
                                       &in_b, stream, nullptr, true));
    csc_fill_kernel<<<grid_for(E, 256 * 4), 256, 0, stream>>>(edge_index, in_b ? L.vals_b : L.vals_a, E, num_sources, out_src,
                                                              out_eid);
    PPG_LAUNCHED();
  }
  return PPG_OK;  // status bits stay in the workspace head: ppg_result_read
}

extern "C" int ppg_csc_build(const int64_t* edge_index, int64_t E, int64_t num_sources, int64_t num_targets, void* workspace,
                             size_t workspace_bytes, int32_t* out_colptr, int32_t* out_src, int32_t* out_eid, void* stream_) {
  PPG_TRY(ppg_csc_build_async(edge_index, E, num_sources, num_targets, workspace, workspace_bytes, out_colptr, out_src,
                              out_eid, stream_));
  ResultWords h;
  PPG_TRY(read_back(&h, static_cast<const ResultWords*>(workspace), static_cast<cudaStream_t>(stream_)));
  PPG_REQUIRE((h.status & kStatusIdOutOfRange) == 0, PPG_ERR_INVALID, "csc_build: node id out of range");
  return PPG_OK;
}

extern "C" int ppg_gcn_norm(const int32_t* colptr, const int32_t* src, const int32_t* eid, const float* edge_weight,
                            int64_t n, int64_t E, float* scratch_dis, float* out_val, float* out_self,
                            float* out_val_edge, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)E;
  if (n == 0) return PPG_OK;
  const unsigned grid = static_cast<unsigned>(ceil_div(n, 256));
  gcn_degree_kernel<<<grid, 256, 0, stream>>>(colptr, src, eid, edge_weight, n, scratch_dis, out_self);
  PPG_LAUNCHED();
  gcn_values_kernel<<<grid, 256, 0, stream>>>(colptr, src, eid, edge_weight, scratch_dis, n, out_val, out_self, out_val_edge);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_colptr_counts(const int32_t* colptr, int64_t n, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) return PPG_OK;
  colptr_counts_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, stream>>>(colptr, n, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

template <int LPN, int VEC>
static int launch_spmm(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val,
                       const float* X, int64_t n, int F, const float* bias, int act, float* out, cudaStream_t stream) {
  const int64_t threads = n * LPN;
  spmm_csc_kernel<LPN, VEC><<<static_cast<unsigned>(ceil_div(threads, 256)), 256, 0, stream>>>(
      colptr, src, val, self_val, X, n, F, bias, act, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_spmm_csc(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val,
                            const float* X, int64_t num_targets, int64_t F, const float* bias, int act, float* out,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(F >= 1 && F < (1 << 30), PPG_ERR_INVALID, "spmm: feature width %lld out of range", (long long)F);
  PPG_REQUIRE(num_targets * 32 < (1ll << 40), PPG_ERR_INVALID, "spmm: too many target nodes");
  if (num_targets == 0) return PPG_OK;
  const int f = static_cast<int>(F);
  const bool al16 = (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if (f % 4 == 0 && al16) {
    if (f <= 32) return launch_spmm<8, 4>(colptr, src, val, self_val, X, num_targets, f, bias, act, out, stream);
    if (f <= 64) return launch_spmm<16, 4>(colptr, src, val, self_val, X, num_targets, f, bias, act, out, stream);
    return launch_spmm<32, 4>(colptr, src, val, self_val, X, num_targets, f, bias, act, out, stream);
  }
  if (f <= 8) return launch_spmm<8, 1>(colptr, src, val, self_val, X, num_targets, f, bias, act, out, stream);
  if (f <= 16) return launch_spmm<16, 1>(colptr, src, val, self_val, X, num_targets, f, bias, act, out, stream);
  return launch_spmm<32, 1>(colptr, src, val, self_val, X, num_targets, f, bias, act, out, stream);
}

extern "C" int ppg_linear(const float* A1, const float* W1, int64_t M, int64_t K1, const float* A2, const float* W2,
                          int64_t K2, const float* bias, const float* rowscale, int64_t N, int act, float* out,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(K1 >= 0 && K2 >= 0 && K1 < (1 << 30) && K2 < (1 << 30) && N >= 1 && N < (1 << 30), PPG_ERR_INVALID,
              "linear: bad shape");
  if (M == 0) return PPG_OK;
  dim3 grid(static_cast<unsigned>(ceil_div(M, kLinBM)), static_cast<unsigned>(ceil_div(N, kLinBN)));
  PPG_REQUIRE(grid.y < 65536, PPG_ERR_INVALID, "linear: output width %lld too large", (long long)N);
  linear_kernel<<<grid, 256, 0, stream>>>(A1, W1, static_cast<int>(K1), A2, W2, static_cast<int>(K2), bias, rowscale, M,
                                          static_cast<int>(N), act, out);
  PPG_LAUNCHED();
  return PPG_OK;
}
