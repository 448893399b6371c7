// Fused GCN layer with the dense transform on the 5th-generation tensor cores (tcgen05, accumulator in TMEM).
//
//   out[v,:] = act( (sum_i val_i X[src_i,:] + self_v X[v,:]) W^T + b )        F in {32, 64}, H in {16, 32, 64}
//
// Same tiling as gcn_fused.cu (persistent CTAs over 128-node tiles; phase 1 = flat segmented walk over the
// tile's CSC slots with 8 row gathers in flight per lane group), but the [128 x F] . [F x H] product of
// phase 2 no longer runs on the FMA pipe, where it costs as much time as the gathers (2*128*64*64 flop per
// tile = 4096 FMA-pipe cycles per SM): one thread issues tcgen05.mma (kind::tf32, M = 128, N = H, K = 8
// per instruction) on operands staged in shared memory in the canonical K-major SWIZZLE_128B layout, the
// accumulator lives in TMEM, and the epilogue reads it back with tcgen05.ld (bias + ELU in registers).
//
// fp32 accuracy (north_star: 1e-5 relative): TF32 keeps 10 mantissa bits, so every operand is split
// x = hi + lo with hi = x truncated to TF32 (exactly representable) and lo = x - hi (exact in fp32), and
// three MMAs accumulate A_hi W_hi + A_lo W_hi + A_hi W_lo into the same TMEM tile; the dropped term
// A_lo W_lo is below 2^-21 relative, products of 11-bit significands are exact in the fp32 accumulator.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ppg {

constexpr int kTcTile = 128;
constexpr int kTcThreads = 256;
constexpr int kTcGatherBatch = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// byte offset of element (row, k) of a [rows x K] fp32 operand in the K-major SWIZZLE_128B canonical layout:
// 128-byte K blocks (32 floats) are the outer dimension; inside a block every row is 128 contiguous bytes
// and the 16-byte chunk index is XORed with (row % 8)  (Swizzle<3,4,3> on byte addresses).
__device__ __forceinline__ uint32_t swz_offset(int rows, int row, int k) {
  const int kb = k >> 5, kk = k & 31;
  return static_cast<uint32_t>(kb * rows * 128 + row * 128 + ((((kk >> 2) ^ (row & 7)) << 4) | ((kk & 3) << 2)));
}

// shared-memory matrix descriptor (sm_100): start address >> 4 in [0,14), leading byte offset [16,30) (unused for
// swizzled K-major), stride byte offset [32,46) = 8 rows * 128 B, version 1 at [46,48), SWIZZLE_128B = 2 at [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// instruction descriptor: D = F32 (1 at [4,6)), A = B = TF32 (2 at [7,10) and [10,13)), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1 << 26)) __trap();  // a lost commit must fail loudly, not hang the device
  }
}

__host__ __device__ constexpr int tc_a_region_bytes(int F, int H) {
  const int a = 2 * kTcTile * F * 4, o = kTcTile * (H + 4) * 4;
  return ((a > o ? a : o) + 1023) / 1024 * 1024;
}

template <int N>
struct TmemLoad;
template <>
struct TmemLoad<8> {
  __device__ static void run(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  }
};
template <>
struct TmemLoad<16> {
  __device__ static void run(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  }
};
template <>
struct TmemLoad<32> {
  __device__ static void run(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
        "[%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
  }
};

// ELU: expm1 for x <= 0 without the slow library path.  |x| < 0.125: degree-5 Taylor polynomial (truncation
// < 5e-8 relative); below: exp(x) - 1 with ex2.approx (result magnitude >= 0.117, absolute error ~2e-7: < 2e-6 relative).
__device__ __forceinline__ float tc_activate(float x, int act) {
  // branch-free: both sides are evaluated for every element anyway (a warp holds positive and negative values), and a
  // branch costs a reconvergence barrier per element
  float p = fmaf(x, 1.f / 120.f, 1.f / 24.f);
  p = fmaf(p, x, 1.f / 6.f);
  p = fmaf(p, x, 0.5f);
  p = fmaf(p, x, 1.f);
  p *= x;
  const float e = __expf(x) - 1.f;
  const float neg = x > -0.125f ? p : e;
  return (act != PPG_ACT_ELU || x > 0.f) ? x : neg;
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}

template <int F, int H>
__global__ void __launch_bounds__(kTcThreads, 2)
gcn_tc_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ src, const float* __restrict__ val,
              const float* __restrict__ self_val, const float* __restrict__ X, const float* __restrict__ W,
              const float* __restrict__ bias, int64_t n, int act, float* __restrict__ out) {
  static_assert(F % 32 == 0 && (H == 16 || H == 32 || H == 64), "unsupported width");
  constexpr int LPN = F / 4;                                   // lanes per lane group (one float4 each)
  constexpr int GPW = 32 / LPN;                                // lane groups per warp
  constexpr int NPG = kTcTile / ((kTcThreads / 32) * GPW);     // consecutive nodes owned by a lane group
  constexpr int A_BYTES = kTcTile * F * 4;
  constexpr int W_BYTES = H * F * 4;
  constexpr int TMEM_COLS = H < 32 ? 32 : H;
  constexpr int CPT = H / 2;                                   // accumulator columns per thread in the epilogue
  constexpr uint32_t IDESC = umma_idesc_tf32(kTcTile, H);
  constexpr int A_REGION = tc_a_region_bytes(F, H);  // both A parts, or the staged output tile if that is larger

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sAhi = base;
  unsigned char* sAlo = base + A_BYTES;
  unsigned char* sWhi = base + A_REGION;
  unsigned char* sWlo = sWhi + W_BYTES;
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ uint32_t s_tmem_base;
  __shared__ int32_t s_ptr[kTcTile + 1];
  __shared__ float s_bias[H];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // ---- one-time setup: W^T split into TF32 hi / lo parts, bias, TMEM allocation, completion barrier
  for (int idx = tid; idx < H * F; idx += kTcThreads) {
    const int h = idx / F, k = idx % F;
    float hi, lo;
    split_tf32(W[idx], hi, lo);
    const uint32_t off = swz_offset(H, h, k);
    *reinterpret_cast<float*>(sWhi + off) = hi;
    *reinterpret_cast<float*>(sWlo + off) = lo;
  }
  if (tid < H) s_bias[tid] = bias != nullptr ? bias[tid] : 0.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t mbar = smem_u32(&s_mbar);
  uint32_t parity = 0;

  const int g = lane % LPN;
  const int grp = lane / LPN;
  const int64_t num_tiles = ceil_div(n, kTcTile);
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kTcTile;
    PPG_TRACE(static_cast<unsigned>(tile), 0);

    // ---------------- phase 1: segment-reduce the incoming rows of 128 target nodes (see gcn_fused.cu)
    if (tid <= kTcTile) {
      const int64_t v = row0 + tid;
      s_ptr[tid] = colptr[v < n ? v : n];
    }
    __syncthreads();
    {
      const int r_lo = (warp * GPW + grp) * NPG;
      const int32_t e_lo = s_ptr[r_lo];
      const int32_t e_hi = s_ptr[r_lo + NPG];
      // own rows first: self_v X[v], parked (full fp32) in the hi buffer until the node is finished
      {
        float4 own[NPG];
        float coef[NPG];
#pragma unroll
        for (int q = 0; q < NPG; ++q) {
          const int64_t v = row0 + r_lo + q;
          own[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          coef[q] = 0.f;
          if (v < n && self_val != nullptr) {
            own[q] = *reinterpret_cast<const float4*>(X + v * F + g * 4);
            coef[q] = self_val[v];
          }
        }
#pragma unroll
        for (int q = 0; q < NPG; ++q)
          *reinterpret_cast<float4*>(sAhi + swz_offset(kTcTile, r_lo + q, g * 4)) =
              make_float4(coef[q] * own[q].x, coef[q] * own[q].y, coef[q] * own[q].z, coef[q] * own[q].w);
      }
      int r = r_lo;
      int32_t nb = s_ptr[r + 1];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      auto finish_node = [&](int row, const float4& a) {
        const uint32_t off = swz_offset(kTcTile, row, g * 4);
        const float4 o = *reinterpret_cast<const float4*>(sAhi + off);
        float4 hi, lo;
        split_tf32(o.x + a.x, hi.x, lo.x);
        split_tf32(o.y + a.y, hi.y, lo.y);
        split_tf32(o.z + a.z, hi.z, lo.z);
        split_tf32(o.w + a.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(sAhi + off) = hi;
        *reinterpret_cast<float4*>(sAlo + off) = lo;
      };
      // software pipeline: the (src, val) words of the NEXT batch are requested before the rows of the current
      // batch are consumed, so every batch after the first costs one memory latency (the row gather), not two
      int32_t sidx_n[kTcGatherBatch];
      float c_n[kTcGatherBatch];
#pragma unroll
      for (int u = 0; u < kTcGatherBatch; ++u) {
        const bool in = e_lo + u < e_hi;
        sidx_n[u] = in ? src[e_lo + u] : 0;
        c_n[u] = in ? (val != nullptr ? val[e_lo + u] : 1.f) : 0.f;
      }
      for (int32_t i = e_lo; i < e_hi; i += kTcGatherBatch) {
        int32_t sidx[kTcGatherBatch];
        float c[kTcGatherBatch];
        float4 x[kTcGatherBatch];
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u) {
          sidx[u] = sidx_n[u];
          c[u] = c_n[u];
        }
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u)
          x[u] = (i + u < e_hi) ? *reinterpret_cast<const float4*>(X + static_cast<int64_t>(sidx[u]) * F + g * 4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u) {
          const int32_t nx = i + kTcGatherBatch + u;
          const bool in = nx < e_hi;
          sidx_n[u] = in ? src[nx] : 0;
          c_n[u] = in ? (val != nullptr ? val[nx] : 1.f) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u) {
          if (i + u < e_hi) {
            while (i + u >= nb) {  // the slot belongs to a later node: finish the current one
              finish_node(r, acc);
              acc = make_float4(0.f, 0.f, 0.f, 0.f);
              ++r;
              nb = s_ptr[r + 1];
            }
            acc.x = fmaf(c[u], x[u].x, acc.x);
            acc.y = fmaf(c[u], x[u].y, acc.y);
            acc.z = fmaf(c[u], x[u].z, acc.z);
            acc.w = fmaf(c[u], x[u].w, acc.w);
          }
        }
      }
      for (; r < r_lo + NPG; ++r) {  // the node the walk ended in, and the edge-less nodes after it
        finish_node(r, acc);
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // generic-proxy writes of the operands -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    PPG_TRACE(static_cast<unsigned>(tile), 1);  // rows gathered

    // ---------------- phase 2: D[128 x H] (TMEM) = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T, one issuing thread
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo), w_hi = smem_u32(sWhi), w_lo = smem_u32(sWlo);
      uint32_t accumulate = 0;
#pragma unroll
      for (int j = 0; j < F / 8; ++j) {  // K = 8 per tf32 instruction: 32 bytes along the 128-byte swizzled row
        const uint32_t a_off = (j >> 2) * (kTcTile * 128) + (j & 3) * 32;
        const uint32_t w_off = (j >> 2) * (H * 128) + (j & 3) * 32;
        umma_tf32(tmem_base, umma_desc(a_hi + a_off), umma_desc(w_hi + w_off), IDESC, accumulate);
        accumulate = 1;
        umma_tf32(tmem_base, umma_desc(a_lo + a_off), umma_desc(w_hi + w_off), IDESC, 1);
        umma_tf32(tmem_base, umma_desc(a_hi + a_off), umma_desc(w_lo + w_off), IDESC, 1);
      }
      // arrives on the barrier when all MMAs above have completed (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
    }
    mbar_wait(mbar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    PPG_TRACE(static_cast<unsigned>(tile), 2);  // product in TMEM

    // ---------------- epilogue: TMEM -> registers (thread = one row, CPT consecutive columns) -> bias + act ->
    // shared memory (the operand buffers are idle once the MMAs have completed) -> coalesced 16-byte stores
    {
      constexpr int LDO = H + 4;  // padded row stride (floats) of the staged output tile
      float* sOut = reinterpret_cast<float*>(base);
      uint32_t d[CPT];
      const int col0 = (warp >> 2) * CPT;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(col0);
      TmemLoad<CPT>::run(taddr, d);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* so = sOut + ((warp & 3) * 32 + lane) * LDO + col0;
#pragma unroll
      for (int c = 0; c < CPT; c += 4) {
        float4 r4;
        r4.x = tc_activate(__uint_as_float(d[c + 0]) + s_bias[col0 + c + 0], act);
        r4.y = tc_activate(__uint_as_float(d[c + 1]) + s_bias[col0 + c + 1], act);
        r4.z = tc_activate(__uint_as_float(d[c + 2]) + s_bias[col0 + c + 2], act);
        r4.w = tc_activate(__uint_as_float(d[c + 3]) + s_bias[col0 + c + 3], act);
        *reinterpret_cast<float4*>(so + c) = r4;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      const int64_t rows_here = n - row0 < kTcTile ? n - row0 : kTcTile;
      float* o = out + row0 * H;
      for (int idx = tid; idx < kTcTile * (H / 4); idx += kTcThreads) {
        const int r = idx / (H / 4), c4 = idx % (H / 4);
        if (r < rows_here) *reinterpret_cast<float4*>(o + static_cast<int64_t>(r) * H + c4 * 4) = *reinterpret_cast<const float4*>(sOut + r * LDO + c4 * 4);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // TMEM tile and operand buffers are free for the next tile
    PPG_TRACE(static_cast<unsigned>(tile), 3);  // tile written
  }

  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised variant (the default): one persistent CTA per SM, 21 warps.
//   warps 0-3   epilogue : tcgen05.ld of accumulator stage a (warp w reads TMEM lanes 32w..32w+31) -> bias + ELU ->
//                          per-warp padded staging in shared memory -> coalesced 16-byte stores of 32 finished rows
//   warp  4     MMA      : one thread; waits for an operand stage, issues the 3 x F/8 tcgen05.mma into accumulator
//                          stage a, commits to the stage's `done` barrier
//   warps 5-12  gather team 0 (tiles 0, 2, 4, ... of the CTA -> operand stage 0 / accumulator 0)
//   warps 13-20 gather team 1 (tiles 1, 3, 5, ...            -> operand stage 1 / accumulator 1)
// The three phases of a tile (segment-reduce gather, tensor-core product, epilogue) belong to different warps and
// hand over through mbarriers, so the gathers of tile i + 1 and i + 2 are in flight while tile i is multiplied and
// stored: the load pipe never waits for the tensor core or the store, which the single-role kernel below did for
// more than half of every tile.  Barriers per stage s: full[s] (8 gather warps arrive), done[s] (tcgen05.commit:
// the MMAs have read the operands AND the accumulator is complete -- the gather team and the epilogue both wait on
// it), tempty[s] (4 epilogue warps arrive once their TMEM loads have completed).
constexpr int kWsTeamWarps = 8;
constexpr int kWsThreads = (4 + 1 + 2 * kWsTeamWarps) * 32;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kWsTeamWarps * 32) : "memory");
}

template <int F, int H>
__global__ void __launch_bounds__(kWsThreads, 1)
gcn_tc_ws_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                 const float* __restrict__ self_val, const float* __restrict__ X, const float* __restrict__ W,
                 const float* __restrict__ bias, int64_t n, int act, float* __restrict__ out) {
  static_assert(F % 32 == 0 && (H == 16 || H == 32 || H == 64), "unsupported width");
  constexpr int LPN = F / 4;                               // lanes per lane group (one float4 each)
  constexpr int GPW = 32 / LPN;                            // lane groups per warp
  constexpr int NPG = kTcTile / (kWsTeamWarps * GPW);      // consecutive nodes owned by a lane group
  constexpr int A_BYTES = kTcTile * F * 4;
  constexpr int W_BYTES = H * F * 4;
  constexpr int TMEM_COLS = 2 * H < 32 ? 32 : 2 * H;
  constexpr int LDO = H + 4;                               // padded row stride (floats) of the epilogue staging
  constexpr uint32_t IDESC = umma_idesc_tf32(kTcTile, H);
  constexpr int CH = H < 32 ? H : 32;                      // accumulator columns per tcgen05.ld

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sA = base;                                // [stage][hi | lo][A_BYTES]
  unsigned char* sWhi = base + 4 * A_BYTES;
  unsigned char* sWlo = sWhi + W_BYTES;
  float* sOut = reinterpret_cast<float*>(sWlo + W_BYTES);  // [4 warps][32][LDO]
  __shared__ __align__(8) unsigned long long s_bar[6];     // full[2], done[2], tempty[2]
  __shared__ uint32_t s_tmem_base;
  __shared__ int32_t s_ptr[2][2][kTcTile + 1];
  __shared__ float s_bias[H];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // ---- one-time setup: W^T split into TF32 hi / lo parts, bias, TMEM allocation, barriers
  for (int idx = tid; idx < H * F; idx += kWsThreads) {
    const int h = idx / F, k = idx % F;
    float hi, lo;
    split_tf32(W[idx], hi, lo);
    const uint32_t off = swz_offset(H, h, k);
    *reinterpret_cast<float*>(sWhi + off) = hi;
    *reinterpret_cast<float*>(sWlo + off) = lo;
  }
  if (tid < H) s_bias[tid] = bias != nullptr ? bias[tid] : 0.f;
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&s_bar[s]), kWsTeamWarps);
      mbar_init(smem_u32(&s_bar[2 + s]), 1);
      mbar_init(smem_u32(&s_bar[4 + s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t bar_full = smem_u32(&s_bar[0]), bar_done = smem_u32(&s_bar[2]), bar_tempty = smem_u32(&s_bar[4]);
  const int64_t num_tiles = ceil_div(n, kTcTile);

  if (warp < 4) {
    // ================================================================== epilogue
    float* so = sOut + warp * (32 * LDO);
    for (int64_t j = 0;; ++j) {
      const int64_t tile = blockIdx.x + j * gridDim.x;
      if (tile >= num_tiles) break;
      const int s = static_cast<int>(j & 1);
      const uint32_t k = static_cast<uint32_t>(j >> 1);
      mbar_wait(bar_done + 8 * s, k & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(s * H);
#pragma unroll
      for (int c0 = 0; c0 < H; c0 += CH) {
        uint32_t d[CH];
        TmemLoad<CH>::run(taddr + c0, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < CH; c += 4) {
          float4 r4;
          r4.x = tc_activate(__uint_as_float(d[c + 0]) + s_bias[c0 + c + 0], act);
          r4.y = tc_activate(__uint_as_float(d[c + 1]) + s_bias[c0 + c + 1], act);
          r4.z = tc_activate(__uint_as_float(d[c + 2]) + s_bias[c0 + c + 2], act);
          r4.w = tc_activate(__uint_as_float(d[c + 3]) + s_bias[c0 + c + 3], act);
          *reinterpret_cast<float4*>(so + lane * LDO + c0 + c) = r4;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * s);   // the accumulator stage may be overwritten
      const int64_t row0 = tile * kTcTile + warp * 32;
      const int64_t rows_here = n - row0 < 32 ? n - row0 : 32;
      float* o = out + row0 * H;
#pragma unroll 4
      for (int idx = lane; idx < 32 * (H / 4); idx += 32) {
        const int r = idx / (H / 4), c4 = idx % (H / 4);
        if (r < rows_here) __stcs(reinterpret_cast<float4*>(o + static_cast<int64_t>(r) * H + c4 * 4),
                                  *reinterpret_cast<const float4*>(so + r * LDO + c4 * 4));
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t w_hi = smem_u32(sWhi), w_lo = smem_u32(sWlo);
      for (int64_t j = 0;; ++j) {
        const int64_t tile = blockIdx.x + j * gridDim.x;
        if (tile >= num_tiles) break;
        const int s = static_cast<int>(j & 1);
        const uint32_t k = static_cast<uint32_t>(j >> 1);
        mbar_wait(bar_tempty + 8 * s, (k & 1) ^ 1);
        mbar_wait(bar_full + 8 * s, k & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(sA + s * 2 * A_BYTES), a_lo = a_hi + A_BYTES;
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(s * H);
        uint32_t accumulate = 0;
#pragma unroll
        for (int q = 0; q < F / 8; ++q) {  // K = 8 per tf32 instruction: 32 bytes along the 128-byte swizzled row
          const uint32_t a_off = (q >> 2) * (kTcTile * 128) + (q & 3) * 32;
          const uint32_t w_off = (q >> 2) * (H * 128) + (q & 3) * 32;
          umma_tf32(tmem_d, umma_desc(a_hi + a_off), umma_desc(w_hi + w_off), IDESC, accumulate);
          accumulate = 1;
          umma_tf32(tmem_d, umma_desc(a_lo + a_off), umma_desc(w_hi + w_off), IDESC, 1);
          umma_tf32(tmem_d, umma_desc(a_hi + a_off), umma_desc(w_lo + w_off), IDESC, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done + 8 * s) : "memory");
      }
    }
  } else {
    // ================================================================== gather teams
    const int team = (warp - 5) / kWsTeamWarps;
    const int tw = (warp - 5) % kWsTeamWarps;
    const int tt = tw * 32 + lane;                        // thread index inside the team
    const int g = lane % LPN;
    const int r_lo = (tw * GPW + lane / LPN) * NPG;
    unsigned char* sAhi = sA + team * 2 * A_BYTES;
    unsigned char* sAlo = sAhi + A_BYTES;
    bool first = true;
    for (int64_t j = team;; j += 2) {
      const int64_t tile = blockIdx.x + j * gridDim.x;
      if (tile >= num_tiles) break;
      const uint32_t k = static_cast<uint32_t>(j >> 1);
      const int64_t row0 = tile * kTcTile;
      int32_t* ptr = s_ptr[team][k & 1];
      if (first) {
        if (tt <= kTcTile) {
          const int64_t v = row0 + tt;
          ptr[tt] = colptr[v < n ? v : n];
        }
        team_sync(team);
        first = false;
      }
      // the CSC pointers of the team's NEXT tile travel while this one is gathered
      const int64_t next_row0 = (tile + 2 * static_cast<int64_t>(gridDim.x)) * kTcTile;
      int32_t ptr_next = 0;
      if (tt <= kTcTile && next_row0 < n) {
        const int64_t v = next_row0 + tt;
        ptr_next = colptr[v < n ? v : n];
      }
      const int32_t e_lo = ptr[r_lo];
      const int32_t e_hi = ptr[r_lo + NPG];
      // own rows: self_v X[v] (requested before the operand stage is known to be free)
      float4 own[NPG];
#pragma unroll
      for (int q = 0; q < NPG; ++q) {
        const int64_t v = row0 + r_lo + q;
        own[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < n && self_val != nullptr) {
          const float4 xv = *reinterpret_cast<const float4*>(X + v * F + g * 4);
          const float cf = self_val[v];
          own[q] = make_float4(cf * xv.x, cf * xv.y, cf * xv.z, cf * xv.w);
        }
      }
      int32_t sidx_n[kTcGatherBatch];
      float c_n[kTcGatherBatch];
#pragma unroll
      for (int u = 0; u < kTcGatherBatch; ++u) {
        const bool in = e_lo + u < e_hi;
        sidx_n[u] = in ? src[e_lo + u] : 0;
        c_n[u] = in ? (val != nullptr ? val[e_lo + u] : 1.f) : 0.f;
      }
      // the MMAs of the tile that used this operand stage last have completed
      mbar_wait(bar_done + 8 * team, (k & 1) ^ 1);
#pragma unroll
      for (int q = 0; q < NPG; ++q) *reinterpret_cast<float4*>(sAhi + swz_offset(kTcTile, r_lo + q, g * 4)) = own[q];
      int r = r_lo;
      int32_t nb = ptr[r + 1];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      auto finish_node = [&](int row, const float4& a) {
        const uint32_t off = swz_offset(kTcTile, row, g * 4);
        const float4 o = *reinterpret_cast<const float4*>(sAhi + off);
        float4 hi, lo;
        split_tf32(o.x + a.x, hi.x, lo.x);
        split_tf32(o.y + a.y, hi.y, lo.y);
        split_tf32(o.z + a.z, hi.z, lo.z);
        split_tf32(o.w + a.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(sAhi + off) = hi;
        *reinterpret_cast<float4*>(sAlo + off) = lo;
      };
      for (int32_t i = e_lo; i < e_hi; i += kTcGatherBatch) {
        int32_t sidx[kTcGatherBatch];
        float c[kTcGatherBatch];
        float4 x[kTcGatherBatch];
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u) {
          sidx[u] = sidx_n[u];
          c[u] = c_n[u];
        }
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u)
          x[u] = (i + u < e_hi) ? *reinterpret_cast<const float4*>(X + static_cast<int64_t>(sidx[u]) * F + g * 4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u) {
          const int32_t nx = i + kTcGatherBatch + u;
          const bool in = nx < e_hi;
          sidx_n[u] = in ? src[nx] : 0;
          c_n[u] = in ? (val != nullptr ? val[nx] : 1.f) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kTcGatherBatch; ++u) {
          if (i + u < e_hi) {
            while (i + u >= nb) {  // the slot belongs to a later node: finish the current one
              finish_node(r, acc);
              acc = make_float4(0.f, 0.f, 0.f, 0.f);
              ++r;
              nb = ptr[r + 1];
            }
            acc.x = fmaf(c[u], x[u].x, acc.x);
            acc.y = fmaf(c[u], x[u].y, acc.y);
            acc.z = fmaf(c[u], x[u].z, acc.z);
            acc.w = fmaf(c[u], x[u].w, acc.w);
          }
        }
      }
      for (; r < r_lo + NPG; ++r) {  // the node the walk ended in, and the edge-less nodes after it
        finish_node(r, acc);
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // generic-proxy writes of the operands -> visible to the tensor core (async proxy), then hand the stage over
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * team);
      if (tt <= kTcTile) s_ptr[team][(k + 1) & 1][tt] = ptr_next;
      team_sync(team);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Staged variant (the default): one persistent CTA per SM, 3 producer warps + 16 consumer warps + 1 MMA warp; the row gather is
// asynchronous, holds no registers and runs ahead of the arithmetic across tile boundaries.
//   * The rows of a tile's CSC slots, in slot order, are cut into chunks of 128; a chunk fills one of kSgStages row
//     buffers in shared memory (128 x F floats, plain row-major).
//   * Producer warps: one per stage.  Its lanes read the chunk's src words (coalesced) while the stage is awaited, the
//     warp requests the 128 rows with cp.async.cg (16 bytes per lane, L1 bypassed) and the slot values with 4-byte
//     cp.async, waits for them (cp.async.wait_all) and signals the stage's `full` barrier with one arrival.  Producers
//     only wait for their stage, so the rows of tiles t + 1, t + 2 travel while tile t is multiplied and written.
//   * Consumer warps: a lane group (F/4 lanes, 16 bytes of the row each) owns 128/groups consecutive nodes and keeps
//     their sums in registers; the nodes' own rows (contiguous in X) are requested with plain loads at the top of the
//     tile and used at its end; per stage a thread adds the slots of its nodes that fall into the chunk (ld.shared.v4
//     + 4 FMA per row and lane), then the warp releases the stage (`empty` barrier).  Same summation order as the
//     other two kernels: bit-identical.  A warp starts the next tile without waiting for the others (the CSC
//     pointers are stored two tiles ahead in one of three buffers).
//   * One more warp issues the MMAs (the issue of 24 tcgen05.mma blocks its thread for 1.7 us per tile) into one of
//     two TMEM accumulators as soon as every consumer warp has arrived on the `operands ready` barrier; while they run
//     the consumers convert, activate and store tile t - 1 straight from TMEM to global memory (thread = one row, 16
//     consecutive columns).  The operand buffers are rewritten only after the MMAs of the previous tile have
//     completed.  No CTA-wide barrier inside the tile loop: every hand-over is an mbarrier.
// Why: the single-role kernel spends 7.8 of 12.2 us per tile in a chain of four dependent memory latencies (pointers ->
// (src, val) -> first row batch -> second row batch) at 25 % occupancy and executes about 20 000 warp instructions per
// tile (profiles/r02aa_gcn_trace.log, r02aa ncu capture); here a row costs one LDGSTS per 16 lanes on the producer
// side and one LDS.128 + 4 FFMA per lane on the consumer side.
constexpr int kSgConsumers = 512;
#ifndef PPG_SG_CHUNK
#define PPG_SG_CHUNK 128
#endif
#ifndef PPG_SG_STAGES
#define PPG_SG_STAGES 3
#endif
constexpr int kSgChunk = PPG_SG_CHUNK;     // rows per stage (multiple of 32, at most 128; experiment builds: make variant)
constexpr int kSgStages = PPG_SG_STAGES;
constexpr int kSgProducers = 32 * kSgStages;                   // one producer warp per stage
constexpr int kSgThreads = kSgConsumers + kSgProducers + 32;   // + the warp whose lane 0 issues the MMAs

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// mbarrier wait of a role that has nothing else to do: try_wait (which may suspend the thread in hardware), then sleep
// between polls -- the waiting role must not take issue slots from the working one
#ifndef PPG_SG_SLEEP
#define PPG_SG_SLEEP 128
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // a parked warp polls every PPG_SG_SLEEP ns: re-polling at once (every arrival on any barrier of the CTA wakes it) took
  // a third of all issued instructions away from the warps that had work
  int spins = 0;
  do {
    __nanosleep(PPG_SG_SLEEP);
    if (++spins > (1 << 25)) __trap();  // a lost arrival must fail loudly (after ~4 s), not hang the device
  } while (!mbar_try_wait(bar, parity));
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSgConsumers) : "memory"); }

__host__ __device__ constexpr size_t sg_smem_bytes(int F, int H) {
  return 2 * static_cast<size_t>(kTcTile) * F * 4 + 2 * static_cast<size_t>(H) * F * 4 +
         static_cast<size_t>(kSgStages) * kSgChunk * F * 4 + static_cast<size_t>(kSgStages) * kSgChunk * 4 + 1024;
}

template <int F, int H>
__global__ void __launch_bounds__(kSgThreads, 1)
gcn_tc_staged_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                     const float* __restrict__ self_val, const float* __restrict__ X, const float* __restrict__ W,
                     const float* __restrict__ bias, int64_t n, int act, float* __restrict__ out) {
  static_assert(F % 32 == 0 && (H == 16 || H == 32 || H == 64), "unsupported width");
  constexpr int LPN = F / 4;                                   // lanes per row (one float4 each)
  constexpr int GPW = 32 / LPN;                                // rows per warp instruction
  constexpr int GROUPS = kSgConsumers / LPN;                   // consumer lane groups
  constexpr int NPG = kTcTile / GROUPS;                        // consecutive nodes per lane group
  constexpr int ROW_BYTES = F * 4;
  constexpr int STAGE_BYTES = kSgChunk * ROW_BYTES;
  constexpr int A_BYTES = kTcTile * F * 4;
  constexpr int W_BYTES = H * F * 4;
  constexpr int TMEM_COLS = 2 * H < 32 ? 32 : 2 * H;           // two accumulators
  constexpr int CPT = H >= 32 ? H / 4 : 8;                     // accumulator columns per thread in the epilogue
  constexpr int CBLOCKS = H / CPT;                             // column blocks (warp / 4 below CBLOCKS takes part)
  constexpr uint32_t IDESC = umma_idesc_tf32(kTcTile, H);
  constexpr int C = kSgChunk;
  static_assert(NPG >= 1 && C % 16 == 0 && C <= 128, "geometry");

  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sAhi = base;
  unsigned char* sAlo = base + A_BYTES;
  unsigned char* sWhi = base + 2 * A_BYTES;
  unsigned char* sWlo = sWhi + W_BYTES;
  unsigned char* sRows = sWlo + W_BYTES;                                          // [kSgStages][C][ROW_BYTES]
  float* sVal = reinterpret_cast<float*>(sRows + kSgStages * STAGE_BYTES);        // [kSgStages][C]
  __shared__ __align__(8) unsigned long long s_mbar;                // MMAs of a tile complete (phase = tile)
  __shared__ __align__(8) unsigned long long s_aready;              // operands of a tile complete (phase = tile)
  __shared__ __align__(8) unsigned long long s_full[kSgStages];     // rows of the stage have landed
  __shared__ __align__(8) unsigned long long s_empty[kSgStages];    // every consumer warp has left the stage
  __shared__ uint32_t s_tmem_base;
  __shared__ int32_t s_ptr[3][kTcTile + 1];                         // CSC pointers of tiles t, t + 1, t + 2
  __shared__ float s_bias[H];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int64_t num_tiles = ceil_div(n, kTcTile);
  const int64_t my_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto row0_of = [&](int64_t t) { return (static_cast<int64_t>(blockIdx.x) + t * gridDim.x) * kTcTile; };

  // ---- one-time setup: W^T split into TF32 hi / lo parts, bias, TMEM allocation, barriers, pointers of tile 0
  for (int idx = tid; idx < H * F; idx += kSgThreads) {
    const int h = idx / F, k = idx % F;
    float hi, lo;
    split_tf32(W[idx], hi, lo);
    const uint32_t off = swz_offset(H, h, k);
    *reinterpret_cast<float*>(sWhi + off) = hi;
    *reinterpret_cast<float*>(sWlo + off) = lo;
  }
  if (tid < H) s_bias[tid] = bias != nullptr ? bias[tid] : 0.f;
  if (tid <= kTcTile) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int64_t v = row0_of(b) + tid;
      s_ptr[b][tid] = colptr[(b < my_tiles && v < n) ? v : n];
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(&s_mbar), 1);
    mbar_init(smem_u32(&s_aready), kSgConsumers / 32);
    for (int s = 0; s < kSgStages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), kSgConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t mbar = smem_u32(&s_mbar), aready = smem_u32(&s_aready);
  const uint32_t bar_full = smem_u32(&s_full[0]), bar_empty = smem_u32(&s_empty[0]);
  const uint32_t rows_base = smem_u32(sRows), val_base = smem_u32(sVal);

  if (warp == (kSgConsumers + kSgProducers) / 32) {
    // ================================================================== MMA issuer
    // D[128 x H] (TMEM, accumulator t & 1) = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T once the operands of tile t are ready
    if (lane == 0) {
      const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo), w_hi = smem_u32(sWhi), w_lo = smem_u32(sWlo);
      for (int64_t t = 0; t < my_tiles; ++t) {
        mbar_wait_parked(aready, static_cast<uint32_t>(t & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>((t & 1) * H);
        uint32_t accumulate = 0;
#pragma unroll
        for (int j = 0; j < F / 8; ++j) {  // K = 8 per tf32 instruction: 32 bytes along the 128-byte swizzled row
          const uint32_t a_off = (j >> 2) * (kTcTile * 128) + (j & 3) * 32;
          const uint32_t w_off = (j >> 2) * (H * 128) + (j & 3) * 32;
          umma_tf32(tmem_d, umma_desc(a_hi + a_off), umma_desc(w_hi + w_off), IDESC, accumulate);
          accumulate = 1;
          umma_tf32(tmem_d, umma_desc(a_lo + a_off), umma_desc(w_hi + w_off), IDESC, 1);
          umma_tf32(tmem_d, umma_desc(a_hi + a_off), umma_desc(w_lo + w_off), IDESC, 1);
        }
        // arrives on the barrier when all MMAs above have completed (implies tcgen05.fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
      }
    }
  } else if (warp >= kSgConsumers / 32) {
    // ================================================================== producers
    // Producer warp pw owns stage pw, i.e. the chunks with (running chunk number) % stages == pw: it requests all rows of
    // the chunk, waits for them and signals the stage with ONE arrival (every arrival on any mbarrier of the CTA wakes
    // the warps parked on the others; 128 per-thread arrivals per chunk cost a third of all issued instructions in
    // re-polls -- DESIGN.md section 4.2).  A warp that owns its stage sees every phase of the stage's barriers, which a
    // parity wait needs.
    constexpr int WPL = C / 32;                       // src words (= rows) of a chunk per lane
    static_assert(C % 32 == 0, "chunk geometry");
    const int pw = warp - kSgConsumers / 32;
    const int h = lane / LPN, c16 = lane % LPN;
    uint32_t gchunk = 0;
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int64_t r0 = row0_of(t), r1 = r0 + kTcTile;
      const int32_t E0 = colptr[r0 < n ? r0 : n], E1 = colptr[r1 < n ? r1 : n];
      const int nchunks = (E1 - E0 + C - 1) / C;
      for (int k = 0; k < nchunks; ++k, ++gchunk) {
        const uint32_t st = gchunk % kSgStages, use = gchunk / kSgStages;
        if (st != static_cast<uint32_t>(pw)) continue;
        const int32_t c0 = E0 + k * C;
        int32_t idx[WPL];
#pragma unroll
        for (int i = 0; i < WPL; ++i) idx[i] = c0 + i * 32 + lane < E1 ? src[c0 + i * 32 + lane] : 0;   // lands while the stage is awaited
        mbar_wait_parked(bar_empty + 8 * st, (use & 1) ^ 1);
        PPG_TRACE_IF(lane == 0 && k == 0, static_cast<unsigned>(r0 / kTcTile), 4);   // first stage of the tile acquired
        const uint32_t stage = rows_base + st * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < WPL; ++i) {
          const int32_t e = c0 + i * 32 + lane;
          if (val != nullptr && e < E1) cp_async4(val_base + (st * C + i * 32 + lane) * 4, val + e);
#pragma unroll
          for (int j = 0; j < 32 / GPW; ++j) {
            const int item = j * GPW + h;
            const int32_t row = __shfl_sync(kFullMask, idx[i], item);
            if (c0 + i * 32 + item < E1)
              cp_async16(stage + static_cast<uint32_t>((i * 32 + item) * ROW_BYTES + c16 * 16), X + static_cast<int64_t>(row) * F + c16 * 4);
          }
        }
        PPG_TRACE_IF(lane == 0 && k + 1 == nchunks, static_cast<unsigned>(r0 / kTcTile), 5);   // last chunk of the tile requested
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * st);
      }
    }
  } else {
    // ================================================================== consumers
    const int c16 = tid % LPN;
    const int G = tid / LPN;
    uint32_t gchunk = 0;
    float cf[NPG], cf_next[NPG];   // coefficients of the own rows, fetched one tile ahead
    auto load_self = [&](int64_t tt, float* coef) {
      const int64_t r0 = row0_of(tt);
#pragma unroll
      for (int j = 0; j < NPG; ++j) {
        const int64_t v = r0 + G * NPG + j;
        coef[j] = (self_val != nullptr && tt < my_tiles && v < n) ? self_val[v] : 0.f;
      }
    };
    // epilogue of tile tt: TMEM accumulator tt & 1 -> registers (thread = one row, CPT consecutive columns) -> bias +
    // act -> global memory
    auto epilogue = [&](int64_t tt) {
      if ((warp >> 2) < CBLOCKS) {
        uint32_t d[CPT];
        const int col0 = (warp >> 2) * CPT;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((tt & 1) * H + col0);
        TmemLoad<CPT>::run(taddr, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float4 q[CPT / 4];
#pragma unroll
        for (int c = 0; c < CPT; c += 4) {
          q[c / 4].x = tc_activate(__uint_as_float(d[c + 0]) + s_bias[col0 + c + 0], act);
          q[c / 4].y = tc_activate(__uint_as_float(d[c + 1]) + s_bias[col0 + c + 1], act);
          q[c / 4].z = tc_activate(__uint_as_float(d[c + 2]) + s_bias[col0 + c + 2], act);
          q[c / 4].w = tc_activate(__uint_as_float(d[c + 3]) + s_bias[col0 + c + 3], act);
        }
        // lanes 2i and 2i + 1 trade halves so that every store instruction writes whole 32-byte sectors: the even lane
        // keeps float4 0 (and 2) of both rows, the odd lane float4 1 (and 3)
        const bool odd = lane & 1;
        const int64_t row_e = row0_of(tt) + (warp & 3) * 32 + (lane & ~1);
#pragma unroll
        for (int c = 0; c < CPT / 4; c += 2) {
          const float4 give = odd ? q[c] : q[c + 1];
          float4 got;
          got.x = __shfl_xor_sync(kFullMask, give.x, 1);
          got.y = __shfl_xor_sync(kFullMask, give.y, 1);
          got.z = __shfl_xor_sync(kFullMask, give.z, 1);
          got.w = __shfl_xor_sync(kFullMask, give.w, 1);
          const float4 for_even_row = odd ? got : q[c];        // even lane: own float4 c; odd lane: the even row's float4 c + 1
          const float4 for_odd_row = odd ? q[c + 1] : got;     // even lane: the odd row's float4 c; odd lane: own float4 c + 1
          float* o = out + row_e * H + col0 + (c + (odd ? 1 : 0)) * 4;
          if (row_e < n) __stcs(reinterpret_cast<float4*>(o), for_even_row);
          if (row_e + 1 < n) __stcs(reinterpret_cast<float4*>(o + H), for_odd_row);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };
    load_self(0, cf_next);
    const int v0 = G * NPG;                       // this lane group's nodes: v0 .. v0 + NPG - 1
    const int w0 = warp * (GPW * NPG);            // this warp's nodes: w0 .. w0 + GPW * NPG - 1
    for (int64_t t = 0; t < my_tiles; ++t) {
      const int b = static_cast<int>(t & 1);
      const int pb = static_cast<int>(t % 3);
      const int64_t row0 = row0_of(t);
      // No wait here: a warp starts to reduce tile t while slower warps are still on tile t - 1 (the rows of a warp's
      // nodes vary from tile to tile; the common point of a tile is the MMA, one phase later).  The pointers of this
      // tile were stored two tiles ago and acquired in the previous iteration (see below).
      PPG_TRACE(static_cast<unsigned>(row0 / kTcTile), 0);
      // pointers of tile t + 2 (registers until this tile's rows are reduced), own rows of this tile (used at its end)
      int32_t p_next = 0;
      if (tid <= kTcTile) {
        const int64_t v = row0_of(t + 2) + tid;
        p_next = colptr[(t + 2 < my_tiles && v < n) ? v : n];
      }
      const int32_t* ptr = s_ptr[pb];
      const int32_t E0 = ptr[0], E1 = ptr[kTcTile];
      const int32_t wlo = ptr[w0], whi = ptr[w0 + GPW * NPG];   // slots of the whole warp
      const int nchunks = (E1 - E0 + C - 1) / C;
      int32_t pn[NPG + 1];
      float4 own[NPG], acc[NPG];
#pragma unroll
      for (int j = 0; j < NPG; ++j) cf[j] = cf_next[j];
      load_self(t + 1, cf_next);
#pragma unroll
      for (int j = 0; j <= NPG; ++j) pn[j] = ptr[v0 + j];
#pragma unroll
      for (int j = 0; j < NPG; ++j) {
        own[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (self_val != nullptr && row0 + v0 + j < n) own[j] = __ldg(reinterpret_cast<const float4*>(X + (row0 + v0 + j) * F + c16 * 4));
        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }

      // ---------------- phase 1: segment-reduce the incoming rows of 128 target nodes out of the stages
      for (int k = 0; k < nchunks; ++k, ++gchunk) {
        const uint32_t st = gchunk % kSgStages, use = gchunk / kSgStages;
        const int32_t c0 = E0 + k * C;
        const int32_t c1 = c0 + C < E1 ? c0 + C : E1;
        // every warp observes every phase of the `full` barrier (a parity wait is only meaningful to a thread that has
        // seen the previous phase complete), but only touches the chunks that hold slots of its own nodes
        mbar_wait_parked(bar_full + 8 * st, use & 1);
        if (wlo < c1 && whi > c0) {
          const uint32_t stage = rows_base + st * STAGE_BYTES + c16 * 16;
          const uint32_t sv = val_base + st * C * 4;
#pragma unroll
          for (int j = 0; j < NPG; ++j) {
            const int32_t a = pn[j] > c0 ? pn[j] : c0;
            const int32_t z = pn[j + 1] < c1 ? pn[j + 1] : c1;
            for (int32_t e = a; e < z; ++e) {
              const float4 x = lds128(stage + static_cast<uint32_t>((e - c0) * ROW_BYTES));
              const float c = val != nullptr ? lds32(sv + static_cast<uint32_t>((e - c0) * 4)) : 1.f;
              acc[j].x = fmaf(c, x.x, acc[j].x);
              acc[j].y = fmaf(c, x.y, acc[j].y);
              acc[j].z = fmaf(c, x.z, acc[j].z);
              acc[j].w = fmaf(c, x.w, acc[j].w);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * st);
      }
      PPG_TRACE(static_cast<unsigned>(row0 / kTcTile), 1);  // rows reduced
      // the MMAs of tile t - 1 have read the operand buffers (and their accumulator is complete)
      if (t > 0) {
        mbar_wait_parked(mbar, static_cast<uint32_t>((t - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ... which implies that every warp has handed tile t - 1 over: this wait returns at once and acquires what the
        // other warps stored before their arrival (the pointers of tile t + 1)
        mbar_wait_parked(aready, static_cast<uint32_t>((t - 1) & 1));
      }
#pragma unroll
      for (int j = 0; j < NPG; ++j) {   // operand rows: TF32 hi / lo parts of self_v X[v] + sum (no FMA contraction: same bits as the other kernels)
        const uint32_t off = swz_offset(kTcTile, v0 + j, c16 * 4);
        float4 h4, l4;
        split_tf32(__fadd_rn(__fmul_rn(cf[j], own[j].x), acc[j].x), h4.x, l4.x);
        split_tf32(__fadd_rn(__fmul_rn(cf[j], own[j].y), acc[j].y), h4.y, l4.y);
        split_tf32(__fadd_rn(__fmul_rn(cf[j], own[j].z), acc[j].z), h4.z, l4.z);
        split_tf32(__fadd_rn(__fmul_rn(cf[j], own[j].w), acc[j].w), h4.w, l4.w);
        sts128(smem_u32(sAhi) + off, h4);
        sts128(smem_u32(sAlo) + off, l4);
      }
      if (tid <= kTcTile) s_ptr[(pb + 2) % 3][tid] = p_next;   // buffer of tile t - 1, which every warp read before its arrival for tile t - 1
      // generic-proxy writes of the operands -> visible to the tensor core (async proxy); hand the tile to the MMA warp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(aready);
      PPG_TRACE(static_cast<unsigned>(row0 / kTcTile), 2);  // operands handed over
      // ---------------- epilogue of the PREVIOUS tile while the tensor core works on this one
      if (t > 0) epilogue(t - 1);
      PPG_TRACE(static_cast<unsigned>(row0 / kTcTile), 3);  // previous tile written
    }
    mbar_wait_parked(mbar, static_cast<uint32_t>((my_tiles - 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(my_tiles - 1);
    consumer_sync();
    if (warp == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
  }
}

template <int F, int H>
static int launch_tc_staged(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val, const float* X,
                            const float* W, const float* bias, int64_t n, int act, float* out, cudaStream_t stream) {
  constexpr size_t smem = sg_smem_bytes(F, H);
  auto kern = gcn_tc_staged_kernel<F, H>;
  static bool configured = false;
  if (!configured) {
    PPG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  const int64_t tiles = ceil_div(n, kTcTile);
  unsigned grid = static_cast<unsigned>(tiles < kNumSMsB200 ? tiles : kNumSMsB200);
  if (const char* g = getenv("PPG_GCN_TC_GRID")) {   // test hook: few CTAs with many tiles each (sanitizer runs on small graphs)
    const long v = atol(g);
    if (v >= 1 && v < static_cast<long>(grid)) grid = static_cast<unsigned>(v);
  }
  kern<<<grid, kSgThreads, smem, stream>>>(colptr, src, val, self_val, X, W, bias, n, act, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

template <int F, int H>
static int launch_tc_ws(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val, const float* X,
                        const float* W, const float* bias, int64_t n, int act, float* out, cudaStream_t stream) {
  constexpr size_t smem = 4 * static_cast<size_t>(kTcTile) * F * 4 + 2 * static_cast<size_t>(H) * F * 4 +
                          4 * 32 * static_cast<size_t>(H + 4) * 4 + 1024;
  auto kern = gcn_tc_ws_kernel<F, H>;
  static bool configured = false;
  if (!configured) {
    PPG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  const int64_t tiles = ceil_div(n, kTcTile);
  const unsigned grid = static_cast<unsigned>(tiles < kNumSMsB200 ? tiles : kNumSMsB200);
  kern<<<grid, kWsThreads, smem, stream>>>(colptr, src, val, self_val, X, W, bias, n, act, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

template <int F, int H>
static int launch_tc(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val, const float* X,
                     const float* W, const float* bias, int64_t n, int act, float* out, cudaStream_t stream) {
  constexpr size_t smem = static_cast<size_t>(tc_a_region_bytes(F, H)) + 2 * static_cast<size_t>(H) * F * 4 + 1024;
  auto kern = gcn_tc_kernel<F, H>;
  static bool configured = false;
  if (!configured) {
    PPG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  const int64_t tiles = ceil_div(n, kTcTile);
  const int64_t slots = static_cast<int64_t>(kNumSMsB200) * 2;
  const unsigned grid = static_cast<unsigned>(tiles < slots ? tiles : slots);
  kern<<<grid, kTcThreads, smem, stream>>>(colptr, src, val, self_val, X, W, bias, n, act, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

}  // namespace ppg

using namespace ppg;

extern "C" int ppg_gcn_tc_supported(int64_t F, int64_t H) {
  return ((F == 32 || F == 64) && (H == 16 || H == 32 || H == 64)) ? 1 : 0;
}

extern "C" int ppg_gcn_layer_tc(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val,
                                const float* X, const float* W, const float* bias, int64_t n, int64_t e, int64_t F,
                                int64_t H, int act, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) return PPG_OK;
  // Three kernels with identical results (bit for bit): the staged kernel (producer warps stream the rows into shared
  // memory stages with cp.async and run ahead across tiles, consumer warps reduce, a dedicated warp issues the MMAs),
  // the single-role one (every warp gathers a batch of rows into registers, then one thread issues the MMAs, then
  // every warp runs the epilogue) and the warp-specialised one (gather warps fill operand buffer i + 1 while the MMA /
  // epilogue warps work on tile i, two TMEM accumulators).  Sparse graphs (at most 4 slots per node: the higher-order
  // De Bruijn layers, 217 against 309 us at cfg2) take the staged kernel, dense ones (the first-order layer of cfg2, 10
  // slots per node: 76 against 92 us) the single-role kernel, where one lane group streams a long slot range with 8
  // loads in flight while a stage of the staged kernel would serve only one or two consumer warps.
  // PPG_GCN_TC=staged|single|ws forces one (PPG_GCN_TC_WS=1 is the older spelling of ws); a feature matrix that is not
  // 16-byte aligned takes the single-role kernel.  Measurements: DESIGN.md section 4.2.
  const int variant = [] {   // read per call: the tests and the A/B scripts switch kernels inside one process
    const char* v = getenv("PPG_GCN_TC");
    const char* w = getenv("PPG_GCN_TC_WS");
    if (v != nullptr && strcmp(v, "staged") == 0) return 0;
    if (v != nullptr && strcmp(v, "single") == 0) return 1;
    if ((v != nullptr && strcmp(v, "ws") == 0) || (w != nullptr && w[0] == '1')) return 2;
    return -1;
  }();
  const bool aligned = (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  int which = variant >= 0 ? variant : (e > 4 * n ? 1 : 0);
  if (which == 0 && !aligned) which = 1;
  int rc = -1;
  profile_pass_begin(stream);
#define PPG_TC_CASE(FF, HH)                                                                                        \
  if (rc < 0 && F == FF && H == HH)                                                                                \
    rc = which == 0   ? launch_tc_staged<FF, HH>(colptr, src, val, self_val, X, W, bias, n, act, out, stream)        \
         : which == 2 ? launch_tc_ws<FF, HH>(colptr, src, val, self_val, X, W, bias, n, act, out, stream)          \
                      : launch_tc<FF, HH>(colptr, src, val, self_val, X, W, bias, n, act, out, stream)
  PPG_TC_CASE(32, 16); PPG_TC_CASE(32, 32); PPG_TC_CASE(32, 64);
  PPG_TC_CASE(64, 16); PPG_TC_CASE(64, 32); PPG_TC_CASE(64, 64);
#undef PPG_TC_CASE
  profile_pass_end(stream, n, 0, PPG_PROFILE_GCN_LAYER);
  PPG_REQUIRE(rc >= 0, PPG_ERR_INVALID, "tensor-core layer: widths F=%lld H=%lld not supported", (long long)F, (long long)H);
  return rc;
}
