// Backward building blocks of DBGNN (autograd of a10 / a11; the reference trains with plain torch autograd,
// docs/tutorial/dbgnn.ipynb cell 42, through PyG's GCNConv and MessagePassing.propagate).
//
// For Y = act(A X W^T + b) with A the (normalised) target-grouped adjacency:
//     dPre = dY * act'(pre)                   act_backward_kernel   (ELU: act' = 1 if Y > 0 else Y + 1)
//     db   = column sums of dPre              same kernel, per-CTA partials + fixed-order reduction
//     G    = A^T dPre                         the forward segment-reduce kernel on the source-grouped view
//     dW   = G^T X                            atb_kernel            (tall-skinny, reduction over the node dimension)
//     dX   = G W                              the forward linear kernel
// No atomics on floating point anywhere: every sum has a fixed order, so gradients are reproducible run to run.
#include "common.cuh"

namespace ppg {

// ------------------------------------------------------------------ activation backward + bias gradient
constexpr int kActRows = 32;  // rows a CTA handles per loop trip; blockDim = (32 column lanes) x 8 row groups

__global__ void __launch_bounds__(256)
act_backward_kernel(const float* __restrict__ dY, const float* __restrict__ Y, const float* __restrict__ rowscale,
                    int64_t M, int H, int act, float* __restrict__ dPre, float* __restrict__ dPreScaled,
                    float* __restrict__ partials /* [gridDim.x][H] */) {
  extern __shared__ float s_col[];  // [8][H]
  const int lane = threadIdx.x & 31;
  const int rgrp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < 8 * H; c += 256) s_col[c] = 0.f;
  __syncthreads();
  for (int c0 = 0; c0 < H; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    if (c < H) {
      for (int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + rgrp; r < M; r += static_cast<int64_t>(gridDim.x) * 8) {
        const float g = dY[r * H + c];
        float d = g;
        if (act == PPG_ACT_ELU) {
          const float y = Y[r * H + c];
          d = y > 0.f ? g : g * (y + 1.f);
        }
        if (dPre != nullptr) dPre[r * H + c] = d;
        if (rowscale != nullptr) {
          d *= rowscale[r];
          if (dPreScaled != nullptr) dPreScaled[r * H + c] = d;
        }
        acc += d;
      }
      s_col[rgrp * H + c] = acc;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += s_col[g * H + c];
    partials[static_cast<size_t>(blockIdx.x) * H + c] = t;
  }
}

// out[c] = sum over p (ascending) of partials[p][c]
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, int num_partials, int64_t width, float* __restrict__ out) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= width) return;
  float t = 0.f;
  for (int p = 0; p < num_partials; ++p) t += partials[static_cast<size_t>(p) * width + c];
  out[c] = t;
}

// ------------------------------------------------------------------ out[H,F] = A[M,H]^T B[M,F]
// CTA (bx, by, bz): output tile 64 x 64 at (by*64, bz*64), row chunk bx.  256 threads, 4 x 4 micro tile each.
constexpr int kAtbRows = 32;

__global__ void __launch_bounds__(256)
atb_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t M, int H, int F, int64_t rows_per_chunk,
           float* __restrict__ partials /* [gridDim.x][H*F] */) {
  __shared__ __align__(16) float sA[kAtbRows][64 + 4];
  __shared__ __align__(16) float sB[kAtbRows][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int h0 = blockIdx.y * 64, f0 = blockIdx.z * 64;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * rows_per_chunk;
  const int64_t r_end = r_begin + rows_per_chunk < M ? r_begin + rows_per_chunk : M;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += kAtbRows) {
    for (int idx = tid; idx < kAtbRows * 64; idx += 256) {
      const int rr = idx / 64, cc = idx % 64;
      const int64_t r = r0 + rr;
      sA[rr][cc] = (r < r_end && h0 + cc < H) ? A[r * H + h0 + cc] : 0.f;
      sB[rr][cc] = (r < r_end && f0 + cc < F) ? B[r * F + f0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < kAtbRows; ++rr) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[rr][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[rr][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* p = partials + static_cast<size_t>(blockIdx.x) * H * F;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int h = h0 + ty * 4 + i;
    if (h >= H) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = f0 + tx * 4 + j;
      if (f < F) p[static_cast<size_t>(h) * F + f] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256)
gather_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, float* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[idx[i]];
}

inline int reduce_chunks(int64_t M) {
  int64_t c = ceil_div(M, 1024);  // at least 1024 rows per partial
  if (c < 1) c = 1;
  if (c > kNumSMsB200 * 2) c = kNumSMsB200 * 2;
  return static_cast<int>(c);
}

}  // namespace ppg

using namespace ppg;

extern "C" size_t ppg_act_backward_workspace_bytes(int64_t M, int64_t H) {
  return static_cast<size_t>(reduce_chunks(M)) * static_cast<size_t>(H < 1 ? 1 : H) * sizeof(float) + 256;
}

extern "C" int ppg_act_backward(const float* dY, const float* Y, const float* rowscale, int64_t M, int64_t H, int act,
                                float* dPre, float* dPreScaled, float* out_colsum, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(M >= 0 && H >= 1 && H <= 1024, PPG_ERR_INVALID, "act_backward: bad shape [%lld, %lld]", (long long)M, (long long)H);
  PPG_REQUIRE(act == PPG_ACT_NONE || (act == PPG_ACT_ELU && Y != nullptr), PPG_ERR_INVALID, "act_backward: ELU needs the forward output");
  PPG_REQUIRE(workspace_bytes >= ppg_act_backward_workspace_bytes(M, H), PPG_ERR_WORKSPACE, "act_backward: workspace too small");
  const int chunks = reduce_chunks(M);
  float* partials = static_cast<float*>(workspace);
  const int h = static_cast<int>(H);
  act_backward_kernel<<<chunks, 256, 8 * h * sizeof(float), stream>>>(dY, Y, rowscale, M, h, act, dPre, dPreScaled, partials);
  PPG_LAUNCHED();
  reduce_partials_kernel<<<static_cast<unsigned>(ceil_div(H, 256)), 256, 0, stream>>>(partials, chunks, H, out_colsum);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" size_t ppg_atb_workspace_bytes(int64_t M, int64_t H, int64_t F) {
  return static_cast<size_t>(reduce_chunks(M)) * static_cast<size_t>(H < 1 ? 1 : H) * static_cast<size_t>(F < 1 ? 1 : F) *
             sizeof(float) + 256;
}

extern "C" int ppg_atb(const float* A, const float* B, int64_t M, int64_t H, int64_t F, float* out, void* workspace,
                       size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(M >= 0 && H >= 1 && F >= 1 && H <= 4096 && F <= 4096, PPG_ERR_INVALID, "atb: bad shape M=%lld H=%lld F=%lld",
              (long long)M, (long long)H, (long long)F);
  PPG_REQUIRE(workspace_bytes >= ppg_atb_workspace_bytes(M, H, F), PPG_ERR_WORKSPACE, "atb: workspace too small");
  const int chunks = reduce_chunks(M);
  const int64_t rows_per_chunk = ceil_div(ceil_div(M > 0 ? M : 1, chunks), kAtbRows) * kAtbRows;
  float* partials = static_cast<float*>(workspace);
  dim3 grid(chunks, static_cast<unsigned>(ceil_div(H, 64)), static_cast<unsigned>(ceil_div(F, 64)));
  atb_kernel<<<grid, 256, 0, stream>>>(A, B, M, static_cast<int>(H), static_cast<int>(F), rows_per_chunk, partials);
  PPG_LAUNCHED();
  reduce_partials_kernel<<<static_cast<unsigned>(ceil_div(H * F, 256)), 256, 0, stream>>>(partials, chunks, H * F, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_gather_f32(const float* src, const int32_t* idx, int64_t n, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) return PPG_OK;
  gather_f32_kernel<<<grid_for(n, 256 * 4), 256, 0, stream>>>(src, idx, n, out);
  PPG_LAUNCHED();
  return PPG_OK;
}
