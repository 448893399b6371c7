// Fused GCN layer / bipartite layer for the common DBGNN widths (F, H in {16, 32, 64}).
//
//   GCN       : out[v,:] = act( (sum_i val_i X[src_i,:] + self_v X[v,:]) W^T + b )
//   bipartite : out[v,:] = act( (sum_i X_h[src_i,:]) W1^T + indeg(v) (X[v,:] W2^T + b1 + b2) )
//
// (A X) W^T == A (X W^T): aggregating first keeps the transform on the 128-row tile that is already in
// shared memory, so the intermediate [n, F] matrix of the unfused path (one HBM write + one HBM read per
// layer) disappears and the output row is written exactly once, bias and ELU applied in registers.
//
// One persistent CTA (256 threads) per SM slot loops over 128-node tiles:
//   phase 1  lane groups of F/4 lanes own a few consecutive nodes each and walk the contiguous CSC slot range
//            of those nodes as one flat list, 8 source-row gathers (16-byte loads) in flight per group
//            independent of the node degrees, parking every finished node's row in shared memory;
//   phase 2  128 x H x K tile product on the FMA pipe from shared memory (W^T staged once per CTA),
//            (H/8) x 4 register micro-tile per thread, float4 epilogue stores.
// Per-layer HBM traffic: 8 e + (4F) e_sl gathered + 4F n (self rows) + 4H n written.
#include "common.cuh"

namespace ppg {

constexpr int kFusedTile = 128;
constexpr int kFusedThreads = 256;
constexpr int kGatherBatch = 8;  // row gathers a lane group keeps in flight

__device__ __forceinline__ float fused_activate(float x, int act) {
  return (act == PPG_ACT_ELU && x <= 0.f) ? expm1f(x) : x;
}

__device__ __forceinline__ void fma4(float4& acc, float c, const float4& x) {
  acc.x = fmaf(c, x.x, acc.x);
  acc.y = fmaf(c, x.y, acc.y);
  acc.z = fmaf(c, x.z, acc.z);
  acc.w = fmaf(c, x.w, acc.w);
}

template <int F, int H, bool BIP>
__global__ void __launch_bounds__(kFusedThreads, 2)
gcn_fused_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                 const float* __restrict__ self_val, const float* __restrict__ X, const float* __restrict__ X2,
                 const float* __restrict__ W, const float* __restrict__ W2, const float* __restrict__ bias, int64_t n,
                 int act, float* __restrict__ out) {
  constexpr int K = BIP ? 2 * F : F;
  constexpr int LDA = K + 4;            // row stride of the aggregated tile (floats): 16-byte aligned rows
  constexpr int LPN = F / 4;            // lanes per node
  constexpr int GPW = 32 / LPN;         // lane groups per warp
  constexpr int NPG = kFusedTile / ((kFusedThreads / 32) * GPW);  // consecutive nodes owned by a lane group
  constexpr int CG = H / 4;             // column groups (4 output columns per thread)
  constexpr int RG = kFusedThreads / CG;
  constexpr int RPT = kFusedTile / RG;  // output rows per thread
  static_assert(F % 4 == 0 && 32 % LPN == 0 && kFusedTile % RG == 0 && RPT >= 1, "unsupported width");

  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                    // [128][LDA]
  float* sW = sA + kFusedTile * LDA;   // [K][H] = W^T (bipartite: [W1 | W2]^T)
  __shared__ int32_t s_ptr[kFusedTile + 1];  // CSC pointers of the tile's nodes

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  for (int idx = tid; idx < F * H; idx += kFusedThreads) {
    const int k = idx / H, c = idx % H;
    sW[k * H + c] = W[c * F + k];
    if (BIP) sW[(F + k) * H + c] = W2[c * F + k];
  }

  const int g = lane % LPN;
  const int grp = lane / LPN;
  const int cg = tid % CG;
  const int rg = tid / CG;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias != nullptr) b4 = *reinterpret_cast<const float4*>(bias + cg * 4);

  const int64_t num_tiles = ceil_div(n, kFusedTile);
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kFusedTile;

    // ---------------- phase 1: segment-reduce the incoming rows of 128 target nodes
    // Each lane group owns NPG consecutive nodes and walks THEIR contiguous CSC slot range as one flat
    // list, kGatherBatch row gathers in flight at a time, whatever the individual in-degrees are (mean
    // in-degree of a De Bruijn layer is 2-3: a per-node loop would expose one DRAM latency per edge).
    // A node's edges are summed in slot order by one lane group (deterministic), its own row first.
    if (tid <= kFusedTile) {
      const int64_t v = row0 + tid;
      s_ptr[tid] = colptr[v < n ? v : n];
    }
    __syncthreads();
    {
      const int r_lo = (warp * GPW + grp) * NPG;
      const int32_t e_lo = s_ptr[r_lo];
      const int32_t e_hi = s_ptr[r_lo + NPG];
      // own rows: sA[r] = self_v X[v]   (bipartite: sA[r] = [0 | indeg(v) X2[v]])
      {
        float4 own[NPG];
        float coef[NPG];
#pragma unroll
        for (int q = 0; q < NPG; ++q) {
          const int64_t v = row0 + r_lo + q;
          own[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          coef[q] = 0.f;
          if (v < n) {
            if (BIP) {
              own[q] = *reinterpret_cast<const float4*>(X2 + v * F + g * 4);
              coef[q] = static_cast<float>(s_ptr[r_lo + q + 1] - s_ptr[r_lo + q]);
            } else if (self_val != nullptr) {
              own[q] = *reinterpret_cast<const float4*>(X + v * F + g * 4);
              coef[q] = self_val[v];
            }
          }
        }
#pragma unroll
        for (int q = 0; q < NPG; ++q) {
          const float4 o4 = make_float4(coef[q] * own[q].x, coef[q] * own[q].y, coef[q] * own[q].z, coef[q] * own[q].w);
          float* dst = sA + (r_lo + q) * LDA + g * 4;
          if (BIP) {
            *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(dst + F) = o4;
          } else {
            *reinterpret_cast<float4*>(dst) = o4;
          }
        }
      }
      int r = r_lo;
      int32_t nb = s_ptr[r + 1];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // software pipeline: the (src, val) words of the NEXT batch are requested before the rows of the current
      // batch are consumed, so every batch after the first costs one memory latency (the row gather), not two
      int32_t sidx_n[kGatherBatch];
      float c_n[kGatherBatch];
#pragma unroll
      for (int u = 0; u < kGatherBatch; ++u) {
        const bool in = e_lo + u < e_hi;
        sidx_n[u] = in ? src[e_lo + u] : 0;
        c_n[u] = in ? (val != nullptr ? val[e_lo + u] : 1.f) : 0.f;
      }
      for (int32_t i = e_lo; i < e_hi; i += kGatherBatch) {
        int32_t sidx[kGatherBatch];
        float c[kGatherBatch];
        float4 x[kGatherBatch];
#pragma unroll
        for (int u = 0; u < kGatherBatch; ++u) {
          sidx[u] = sidx_n[u];
          c[u] = c_n[u];
        }
#pragma unroll
        for (int u = 0; u < kGatherBatch; ++u)
          x[u] = (i + u < e_hi) ? *reinterpret_cast<const float4*>(X + static_cast<int64_t>(sidx[u]) * F + g * 4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kGatherBatch; ++u) {
          const int32_t nx = i + kGatherBatch + u;
          const bool in = nx < e_hi;
          sidx_n[u] = in ? src[nx] : 0;
          c_n[u] = in ? (val != nullptr ? val[nx] : 1.f) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kGatherBatch; ++u) {
          if (i + u < e_hi) {
            while (i + u >= nb) {  // the slot belongs to a later node: park the finished one
              float4* dst = reinterpret_cast<float4*>(sA + r * LDA + g * 4);
              float4 t = *dst;
              t.x += acc.x; t.y += acc.y; t.z += acc.z; t.w += acc.w;
              *dst = t;
              acc = make_float4(0.f, 0.f, 0.f, 0.f);
              ++r;
              nb = s_ptr[r + 1];
            }
            fma4(acc, c[u], x[u]);
          }
        }
      }
      if (e_hi > e_lo) {  // the node the walk ended in (nodes after it have no edges)
        float4* dst = reinterpret_cast<float4*>(sA + r * LDA + g * 4);
        float4 t = *dst;
        t.x += acc.x; t.y += acc.y; t.z += acc.z; t.w += acc.w;
        *dst = t;
      }
    }
    __syncthreads();

    // ---------------- phase 2: [128 x K] . [K x H] on the FMA pipe
    float4 o[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < K; k += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(sW + (k + 0) * H + cg * 4);
      const float4 w1 = *reinterpret_cast<const float4*>(sW + (k + 1) * H + cg * 4);
      const float4 w2 = *reinterpret_cast<const float4*>(sW + (k + 2) * H + cg * 4);
      const float4 w3 = *reinterpret_cast<const float4*>(sW + (k + 3) * H + cg * 4);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(sA + (rg * RPT + i) * LDA + k);
        fma4(o[i], a.x, w0);
        fma4(o[i], a.y, w1);
        fma4(o[i], a.z, w2);
        fma4(o[i], a.w, w3);
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int64_t v = row0 + rg * RPT + i;
      if (v < n) {
        float scale = 1.f;
        if (BIP) scale = static_cast<float>(colptr[v + 1] - colptr[v]);
        float4 r;
        r.x = fused_activate(o[i].x + scale * b4.x, act);
        r.y = fused_activate(o[i].y + scale * b4.y, act);
        r.z = fused_activate(o[i].z + scale * b4.z, act);
        r.w = fused_activate(o[i].w + scale * b4.w, act);
        *reinterpret_cast<float4*>(out + v * H + cg * 4) = r;
      }
    }
    __syncthreads();  // the next tile overwrites sA
  }
}

template <int F, int H, bool BIP>
static int launch_fused(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val,
                        const float* X, const float* X2, const float* W, const float* W2, const float* bias, int64_t n,
                        int act, float* out, cudaStream_t stream) {
  constexpr int K = BIP ? 2 * F : F;
  constexpr size_t smem = (static_cast<size_t>(kFusedTile) * (K + 4) + static_cast<size_t>(K) * H) * sizeof(float);
  auto kern = gcn_fused_kernel<F, H, BIP>;
  static bool configured = false;
  if (!configured) {
    PPG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  const int64_t tiles = ceil_div(n, kFusedTile);
  const int64_t slots = static_cast<int64_t>(kNumSMsB200) * 2;
  const unsigned grid = static_cast<unsigned>(tiles < slots ? tiles : slots);
  kern<<<grid, kFusedThreads, smem, stream>>>(colptr, src, val, self_val, X, X2, W, W2, bias, n, act, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

static bool fused_width_ok(int64_t w) { return w == 16 || w == 32 || w == 64; }

template <bool BIP>
static int dispatch_fused(int64_t F, int64_t H, const int32_t* colptr, const int32_t* src, const float* val,
                          const float* self_val, const float* X, const float* X2, const float* W, const float* W2,
                          const float* bias, int64_t n, int act, float* out, cudaStream_t stream) {
#define PPG_FUSED_CASE(FF, HH) \
  if (F == FF && H == HH) return launch_fused<FF, HH, BIP>(colptr, src, val, self_val, X, X2, W, W2, bias, n, act, out, stream)
  PPG_FUSED_CASE(16, 16); PPG_FUSED_CASE(16, 32); PPG_FUSED_CASE(16, 64);
  PPG_FUSED_CASE(32, 16); PPG_FUSED_CASE(32, 32); PPG_FUSED_CASE(32, 64);
  PPG_FUSED_CASE(64, 16); PPG_FUSED_CASE(64, 32); PPG_FUSED_CASE(64, 64);
#undef PPG_FUSED_CASE
  PPG_REQUIRE(false, PPG_ERR_INVALID, "fused layer: widths F=%lld H=%lld not in {16,32,64}", (long long)F, (long long)H);
}

}  // namespace ppg

using namespace ppg;

extern "C" int ppg_gcn_fused_supported(int64_t F, int64_t H) { return fused_width_ok(F) && fused_width_ok(H) ? 1 : 0; }

extern "C" int ppg_gcn_layer_fused(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val,
                                   const float* X, const float* W, const float* bias, int64_t n, int64_t F, int64_t H,
                                   int act, float* out, void* stream_) {
  if (n == 0) return PPG_OK;
  return dispatch_fused<false>(F, H, colptr, src, val, self_val, X, nullptr, W, nullptr, bias, n, act, out,
                               static_cast<cudaStream_t>(stream_));
}

extern "C" int ppg_bipartite_fused(const int32_t* colptr, const int32_t* src, const float* X_h, const float* X,
                                   const float* W1, const float* W2, const float* bias12, int64_t n, int64_t F, int64_t H,
                                   int act, float* out, void* stream_) {
  if (n == 0) return PPG_OK;
  return dispatch_fused<true>(F, H, colptr, src, nullptr, nullptr, X_h, X, W1, W2, bias12, n, act, out,
                              static_cast<cudaStream_t>(stream_));
}
