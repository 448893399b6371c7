// Single-pass exclusive prefix sum with decoupled look-back.
//
// One kernel reads its input once and writes its output once: every tile publishes its
// aggregate, then a warp walks back over the preceding tiles' published words (32 at a time)
// until it meets an inclusive prefix. Tiles take their id from an atomic counter, so a tile only
// ever waits on tiles that are already resident -> no deadlock regardless of block scheduling.
//
// The input is not an array but a Producer functor (idx -> count), and the result is handed to a
// Consumer functor (idx, count, exclusive prefix): the per-item work of the calling stage (degree
// look-ups, binary searches, run-head detection, rank scatter) is fused into the scan pass.
#pragma once
#include "common.cuh"

namespace ppg {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

__host__ __device__ inline int64_t scan_num_tiles(int64_t n) { return n > 0 ? ceil_div(n, kScanTile) : 1; }
// workspace words: [0] tile counter (as u64), [1..] one state word per tile
inline size_t scan_state_words(int64_t n) { return 1 + static_cast<size_t>(scan_num_tiles(n)); }

template <class Producer, class Consumer>
__global__ void __launch_bounds__(kScanBlock)
scan_lookback_kernel(Producer produce, Consumer consume, int64_t n, unsigned long long* __restrict__ ws,
                     unsigned code_partial, unsigned code_inclusive, unsigned long long* __restrict__ total_out) {
  constexpr int NW = kScanBlock / 32;
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_warp_total[NW];
  __shared__ unsigned long long s_tile_prefix;

  unsigned* tile_counter = reinterpret_cast<unsigned*>(ws);
  unsigned long long* state = ws + 1;

  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const int64_t warp_base = static_cast<int64_t>(tile) * kScanTile + static_cast<int64_t>(warp) * (32 * kScanItems);

  unsigned long long v[kScanItems];
  unsigned long long excl[kScanItems];
  unsigned long long carry = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t idx = warp_base + i * 32 + lane;
    v[i] = idx < n ? static_cast<unsigned long long>(produce(idx)) : 0ull;
  }
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const unsigned long long inc = warp_inclusive_sum(v[i]);
    excl[i] = carry + inc - v[i];
    carry += __shfl_sync(kFullMask, inc, 31);
  }
  if (lane == 0) s_warp_total[warp] = carry;
  __syncthreads();

  unsigned long long warp_offset = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const unsigned long long t = s_warp_total[w];
    if (w < warp) warp_offset += t;
    tile_total += t;
  }

  if (warp == 0) {
    if (tile == 0) {
      if (lane == 0) {
        state_store(&state[0], code_inclusive, tile_total);
        s_tile_prefix = 0;
      }
    } else {
      if (lane == 0) state_store(&state[tile], code_partial, tile_total);
      unsigned long long acc = 0;
      int64_t newest = static_cast<int64_t>(tile) - 1;  // lane 0 looks at `newest`, lane l at newest - l
      while (true) {
        const int64_t q = newest - lane;
        unsigned code = code_inclusive;
        unsigned long long val = 0;
        if (q >= 0) {
          unsigned long long w;
          do {
            w = state_load(&state[q]);
            code = static_cast<unsigned>(w >> 56);
          } while (code != code_partial && code != code_inclusive);
          val = w & kStateValueMask;
        }
        const unsigned incl = __ballot_sync(kFullMask, code == code_inclusive);
        if (incl) {
          const unsigned first = __ffs(incl) - 1;  // nearest tile that already holds an inclusive prefix
          acc += warp_sum(lane <= first ? val : 0ull);
          break;
        }
        acc += warp_sum(val);
        newest -= 32;
      }
      if (lane == 0) {
        state_store(&state[tile], code_inclusive, acc + tile_total);
        s_tile_prefix = acc;
      }
    }
  }
  __syncthreads();

  const unsigned long long prefix = s_tile_prefix + warp_offset;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t idx = warp_base + i * 32 + lane;
    if (idx < n) consume(idx, v[i], prefix + excl[i]);
  }
  if (total_out != nullptr && threadIdx.x == 0 &&
      static_cast<int64_t>(tile) == scan_num_tiles(n) - 1) {
    *total_out = s_tile_prefix + tile_total;
  }
}

// The state words handed to a scan must have been zeroed (one cudaMemsetAsync per C-ABI call covers
// all scans of that call). `code` must be unique among the scans sharing a state region; distinct
// scans normally get distinct regions and may all use code 1.
template <class Producer, class Consumer>
inline int launch_scan(Producer produce, Consumer consume, int64_t n, unsigned long long* zeroed_ws,
                       unsigned long long* total_out, cudaStream_t stream, unsigned code = 1) {
  const int64_t tiles = scan_num_tiles(n);
  scan_lookback_kernel<<<static_cast<unsigned>(tiles), kScanBlock, 0, stream>>>(
      produce, consume, n, zeroed_ws, 2 * code - 1, 2 * code, total_out);
  PPG_LAUNCHED();
  return PPG_OK;
}

}  // namespace ppg
