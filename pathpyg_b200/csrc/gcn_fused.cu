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
//   phase 1  F/4 lanes per node walk the node's CSC segment, gather source rows with 16-byte loads
//            (4 edges in flight per lane group) and park the aggregated row in shared memory;
//   phase 2  128 x H x K tile product on the FMA pipe from shared memory (W^T staged once per CTA),
//            (H/8) x 4 register micro-tile per thread, float4 epilogue stores.
// Per-layer HBM traffic: 8 e + (4F) e_sl gathered + 4F n (self rows) + 4H n written.
#include "common.cuh"

namespace ppg {

constexpr int kFusedTile = 128;
constexpr int kFusedThreads = 256;

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
  constexpr int GPW = 32 / LPN;         // nodes a warp aggregates concurrently
  constexpr int CG = H / 4;             // column groups (4 output columns per thread)
  constexpr int RG = kFusedThreads / CG;
  constexpr int RPT = kFusedTile / RG;  // output rows per thread
  static_assert(F % 4 == 0 && 32 % LPN == 0 && kFusedTile % RG == 0 && RPT >= 1, "unsupported width");

  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                    // [128][LDA]
  float* sW = sA + kFusedTile * LDA;   // [K][H] = W^T (bipartite: [W1 | W2]^T)

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
    for (int r = warp * GPW + grp; r < kFusedTile; r += (kFusedThreads / 32) * GPW) {
      const int64_t v = row0 + r;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < n) {
        const int32_t a = colptr[v];
        const int32_t b = colptr[v + 1];
        int32_t i = a;
        for (; i + 4 <= b; i += 4) {
          int32_t s[4];
          float c[4];
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            s[u] = src[i + u];
            c[u] = val != nullptr ? val[i + u] : 1.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const float4*>(X + static_cast<int64_t>(s[u]) * F + g * 4);
#pragma unroll
          for (int u = 0; u < 4; ++u) fma4(acc, c[u], x[u]);
        }
        for (; i < b; ++i) {
          const float c = val != nullptr ? val[i] : 1.f;
          const float4 x = *reinterpret_cast<const float4*>(X + static_cast<int64_t>(src[i]) * F + g * 4);
          fma4(acc, c, x);
        }
        if (BIP) {
          const float deg = static_cast<float>(b - a);
          const float4 x = *reinterpret_cast<const float4*>(X2 + v * F + g * 4);
          own = make_float4(deg * x.x, deg * x.y, deg * x.z, deg * x.w);
        } else if (self_val != nullptr) {
          const float4 x = *reinterpret_cast<const float4*>(X + v * F + g * 4);
          fma4(acc, self_val[v], x);
        }
      }
      *reinterpret_cast<float4*>(sA + r * LDA + g * 4) = acc;
      if (BIP) *reinterpret_cast<float4*>(sA + r * LDA + F + g * 4) = own;
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
