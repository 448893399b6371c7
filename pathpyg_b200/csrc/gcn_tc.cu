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

// ELU: expm1 for x <= 0 without the slow library path.  |x| < 0.5: degree-8 Taylor polynomial (truncation
// < 3e-8 relative); below: exp(x) - 1 with ex2.approx (result magnitude >= 0.39, so no cancellation).
__device__ __forceinline__ float tc_activate(float x, int act) {
  if (act != PPG_ACT_ELU || x > 0.f) return x;
  float p = fmaf(x, 1.f / 40320.f, 1.f / 5040.f);
  p = fmaf(p, x, 1.f / 720.f);
  p = fmaf(p, x, 1.f / 120.f);
  p = fmaf(p, x, 1.f / 24.f);
  p = fmaf(p, x, 1.f / 6.f);
  p = fmaf(p, x, 0.5f);
  p = fmaf(p, x, 1.f);
  p *= x;
  const float e = __expf(x) - 1.f;
  return x > -0.5f ? p : e;
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
                                const float* X, const float* W, const float* bias, int64_t n, int64_t F, int64_t H,
                                int act, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) return PPG_OK;
  // Two kernels with identical results (bit for bit): the single-role one (every warp gathers, then one thread issues
  // the MMAs, then every warp runs the epilogue) and the warp-specialised one (gather warps fill operand buffer i + 1
  // while the MMA / epilogue warps work on tile i, two TMEM accumulators).  Measured on B200 at the order-2 layer of
  // cfg2 (n = 1M, e = 1.8M, 64 -> 64; profiles/r02b_gcn_ws_ab.log): 329 us against 340 us -- the gather itself, not the
  // serialisation of the phases, is what bounds the layer, so the single-role kernel is the default; PPG_GCN_TC_WS=1
  // selects the other.
  static const bool specialised = [] {
    const char* e = getenv("PPG_GCN_TC_WS");
    return e != nullptr && e[0] == '1';
  }();
  int rc = -1;
  profile_pass_begin(stream);
#define PPG_TC_CASE(FF, HH)                                                                                        \
  if (rc < 0 && F == FF && H == HH)                                                                                \
    rc = specialised ? launch_tc_ws<FF, HH>(colptr, src, val, self_val, X, W, bias, n, act, out, stream)           \
                     : launch_tc<FF, HH>(colptr, src, val, self_val, X, W, bias, n, act, out, stream)
  PPG_TC_CASE(32, 16); PPG_TC_CASE(32, 32); PPG_TC_CASE(32, 64);
  PPG_TC_CASE(64, 16); PPG_TC_CASE(64, 32); PPG_TC_CASE(64, 64);
#undef PPG_TC_CASE
  profile_pass_end(stream, n, 0, PPG_PROFILE_GCN_LAYER);
  PPG_REQUIRE(rc >= 0, PPG_ERR_INVALID, "tensor-core layer: widths F=%lld H=%lld not supported", (long long)F, (long long)H);
  return rc;
}
