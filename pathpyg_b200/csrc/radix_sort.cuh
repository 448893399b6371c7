// Stable least-significant-digit radix sort, one sweep per 8-bit digit ("onesweep").
//
// HBM traffic for n (key, payload) pairs over P digit passes:
//   histogram pass : read keys once, count digit 0       (sizeof(Key) * n)
//   each digit pass: read pairs once, write pairs once   (2 * (sizeof(Key) + 4) * n)
// The per-pass global scatter offsets come from (a) the digit histogram of the pass -- digit 0 is counted up front,
// every pass counts the digits of the NEXT pass while it holds the keys in registers -- and (b) a decoupled
// look-back chain over the per-tile digit counts, so no pass reads its keys twice.
// Only the significant key bits [0, end_bit) are sorted: P = ceil(end_bit / 8).
//
// Within a tile the ranking is stable: a warp ranks its 32*ITEMS keys round by round; lanes holding the
// same digit find each other through a per-warp mask word in shared memory (atomic OR of the lane bit) and
// the highest of them bumps the warp's digit counter; the tile is reordered through shared memory, and
// every digit's run is written out as one contiguous (coalesced) segment.
//
// Look-back: thread d owns digit d.  Two levels (tiles inside a group of 16, then groups), every walk with
// kLookBatch independent loads in flight consumed in order -- see the comment in the kernel.  A tile counts its
// digits with shared-memory atomics and publishes them BEFORE it ranks its keys: when it looks back after the
// ranking, its predecessors' counts have long been published and the walk does not poll (B200, 64M (u64, u32)
// pairs: 559 us per pass against 650 us when the counts were published after the ranking).
//
// Tile size: 256 threads x 16 keys (the per-tile fixed costs -- look-back, digit scans, counter reset -- are
// amortised over 4096 keys); 256 x 8 up to 512K keys, where more tiles fill more SMs.
//
// The payload is an optional u32 per key; the first pass can synthesise it as the element index
// (IOTA) so that an index permutation costs no read.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace ppg {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortBlock = 256;  // == kRadix: thread d owns digit d in the per-digit phases
constexpr int kMaxPasses = 8;
constexpr int kLookBatch = 8;   // look-back words in flight per thread
constexpr int kLookGroup = 16;  // tiles per look-back group
constexpr int64_t kSmallSortLimit = 512ll << 10;

// development override: PPG_SORT_ITEMS=8|16 forces the tile size
inline int sort_items_override() {
  static const int v = [] {
    const char* e = getenv("PPG_SORT_ITEMS");
    return e ? atoi(e) : 0;
  }();
  return v;
}
inline int sort_items_for(int64_t n) {
  const int o = sort_items_override();
  if (o == 8 || o == 16) return o;
  return n <= kSmallSortLimit ? 8 : 16;
}
inline int64_t sort_num_tiles(int64_t n) {
  return n > 0 ? ceil_div(n, static_cast<int64_t>(kSortBlock) * sort_items_for(n)) : 1;
}
inline int sort_num_passes(int end_bit) { return end_bit <= 0 ? 1 : static_cast<int>(ceil_div(end_bit, kRadixBits)); }

inline size_t sort_num_groups(int64_t n) { return static_cast<size_t>(ceil_div(sort_num_tiles(n), kLookGroup)); }
// zero-initialised words a sort needs:
// [P*256 histogram][P tile counters (u64 each)][tiles*256 tile words][groups*256 group words]
inline size_t sort_state_words(int64_t n, int end_bit) {
  const size_t P = static_cast<size_t>(sort_num_passes(end_bit));
  return P * kRadix + P + (static_cast<size_t>(sort_num_tiles(n)) + sort_num_groups(n)) * kRadix;
}

// Digit-0 counts taken by the kernel that PRODUCES the keys (one shared-memory atomic per key, one global atomic
// per non-empty bin and CTA), so that a sort never reads its keys just to count them.  count() must be reached
// by all 32 lanes of a warp (pass valid = false for lanes past the end).
struct Digit0Counter {
  unsigned* s;
  __device__ __forceinline__ void begin(unsigned* smem256) {
    s = smem256;
    for (int i = threadIdx.x; i < kRadix; i += blockDim.x) s[i] = 0;
    __syncthreads();
  }
  __device__ __forceinline__ void count(unsigned digit, bool valid) {
    const unsigned live = __ballot_sync(kFullMask, valid);
    if (live == 0) return;
    const int first = __ffs(live) - 1;
    const unsigned d0 = __shfl_sync(kFullMask, digit, first);
    if (__all_sync(kFullMask, !valid || digit == d0)) {
      if (static_cast<int>(lane_id()) == first) atomicAdd(&s[digit], static_cast<unsigned>(__popc(live)));
    } else if (valid) {
      atomicAdd(&s[digit], 1u);
    }
  }
  __device__ __forceinline__ void end(unsigned long long* ghist) {
    __syncthreads();
    for (int i = threadIdx.x; i < kRadix; i += blockDim.x) {
      const unsigned c = s[i];
      if (c) atomicAdd(&ghist[i], static_cast<unsigned long long>(c));
    }
  }
};

template <typename KeyT>
__global__ void __launch_bounds__(256)
radix_histogram_kernel(const KeyT* __restrict__ keys, int64_t n, int num_passes, unsigned long long* __restrict__ ghist) {
  __shared__ unsigned s_hist[kMaxPasses * kRadix];
  for (int i = threadIdx.x; i < num_passes * kRadix; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const KeyT k = ld_stream(keys + i);
    for (int p = 0; p < num_passes; ++p) {
      const unsigned d = static_cast<unsigned>(k >> (p * kRadixBits)) & (kRadix - 1);
      // high digits of small-range keys are often identical across a warp: one atomic instead of 32
      const unsigned active = __activemask();
      const unsigned d0 = __shfl_sync(active, d, __ffs(active) - 1);
      if (__all_sync(active, d == d0)) {
        if (lane_id() == static_cast<unsigned>(__ffs(active) - 1)) atomicAdd(&s_hist[p * kRadix + d], __popc(active));
      } else {
        atomicAdd(&s_hist[p * kRadix + d], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < num_passes * kRadix; i += blockDim.x) {
    const unsigned c = s_hist[i];
    if (c) atomicAdd(&ghist[i], static_cast<unsigned long long>(c));
  }
}

// Resident CTAs per SM the register allocation is capped for: the phases of a tile (load, rank, look-back, scan,
// reorder, write) are a dependent chain, so throughput comes from tiles in different phases sharing an SM.
// 16 keys per thread: 3 CTAs (<= 85 registers, 59 KB shared memory each; measured on B200 at 64M pairs: 608 us per
// pass against 704 us at 2 CTAs and 641 us at 4 CTAs with spills); 8 keys: 4 CTAs (64 registers).
#ifndef PPG_SORT_CTAS16
#define PPG_SORT_CTAS16 3
#endif
template <int ITEMS>
struct SortMinCtas {
  static constexpr int value = ITEMS >= 16 ? PPG_SORT_CTAS16 : 4;
};

// bytes of the buffer the tile is reordered through; the peer-mask words of the ranking ([NW][2][256] u32) live in
// it before that, so it is at least as large as they are (only a keys-only sort of small 32-bit tiles is smaller)
template <typename KeyT, bool HAS_VALUES, int ITEMS>
constexpr size_t onesweep_reorder_bytes() {
  constexpr size_t tile = static_cast<size_t>(kSortBlock * ITEMS) * (sizeof(KeyT) + (HAS_VALUES ? sizeof(uint32_t) : 0));
  constexpr size_t masks = static_cast<size_t>(kSortBlock / 32) * 2 * kRadix * sizeof(uint32_t);
  return tile > masks ? tile : masks;
}

template <typename KeyT, bool HAS_VALUES, bool IOTA, int ITEMS, int MIN_CTAS>
__global__ void __launch_bounds__(kSortBlock, MIN_CTAS)
onesweep_pass_kernel(const KeyT* __restrict__ keys_in, KeyT* __restrict__ keys_out,
                     const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                     const unsigned long long* __restrict__ ghist,  // this pass: 256 digit counts
                     unsigned long long* __restrict__ ghist_next,   // next pass: accumulated here (nullptr on the last pass)
                     unsigned* __restrict__ tile_counter, unsigned long long* __restrict__ state,
                     unsigned long long* __restrict__ gstate, unsigned code_partial, unsigned code_inclusive) {
  constexpr int NW = kSortBlock / 32;
  constexpr int TILE = kSortBlock * ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyT* s_keys = reinterpret_cast<KeyT*>(smem_raw);
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);
  // [NW][256] per-warp digit counts: a warp holds at most 32 * ITEMS <= 512 keys of one digit, a tile 4096: 16 bits
  uint16_t* s_whist = reinterpret_cast<uint16_t*>(smem_raw + onesweep_reorder_bytes<KeyT, HAS_VALUES, ITEMS>());
  long long* s_gbase = reinterpret_cast<long long*>(s_whist + NW * kRadix);  // [256] global slot of tile slot 0 of a digit
  // [NW][2][256] per-warp peer masks of the ranking rounds: they are dead before the tile is reordered, so they
  // live in the buffer the reorder fills (onesweep_reorder_bytes is never smaller than they are)
  uint32_t* s_wmask = reinterpret_cast<uint32_t*>(smem_raw);
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_scan[2 * NW];
  __shared__ unsigned s_next[kRadix];  // digit counts of this tile for the NEXT pass
  __shared__ unsigned s_cnt[kRadix];   // digit counts of this tile for THIS pass, taken before the ranking

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const unsigned lane = lane_id();

  for (int i = tid; i < NW * kRadix; i += kSortBlock) {
    s_wmask[i] = 0;
    s_wmask[NW * kRadix + i] = 0;
  }
  for (int i = tid; i < NW * kRadix / 2; i += kSortBlock) reinterpret_cast<uint32_t*>(s_whist)[i] = 0;
  s_next[tid] = 0;
  s_cnt[tid] = 0;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const int64_t tile_base = static_cast<int64_t>(tile) * TILE;
  const int64_t warp_base = tile_base + static_cast<int64_t>(warp) * (32 * ITEMS);
  const int64_t remaining = n - tile_base;
  const int valid_in_tile = remaining >= TILE ? TILE : static_cast<int>(remaining);

  PPG_TRACE(tile, 0);
  // ---- load (warp-striped => coalesced; element order inside a warp is (item, lane))
  KeyT key[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int64_t idx = warp_base + i * 32 + lane;
    key[i] = idx < n ? ld_stream(keys_in + idx) : static_cast<KeyT>(~static_cast<KeyT>(0));
  }

  // ---- digit counts, taken while the keys are in registers (one shared-memory atomic per key and digit; lanes
  // that share a digit hit one address, which the shared-memory atomic unit combines -- measured on B200, a
  // warp-uniformity vote in front of the atomic costs more on random digits than it saves on uniform ones):
  //  * the NEXT pass's digits -> ghist_next: the sort reads its keys once per pass and never for a histogram of
  //    its own (only digit 0 is counted up front);
  //  * THIS pass's digits of the tile -> published BEFORE the ranking.  The ranking yields the same counts, but
  //    2-3 us later; published early, the counts of a tile's predecessors are long out when it looks back after its
  //    own ranking, so the look-back no longer polls (the polls were a quarter of the kernel's instructions).
  if (ghist_next != nullptr) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const unsigned d1 = static_cast<unsigned>(key[i] >> (shift + kRadixBits)) & (kRadix - 1);
      if (warp_base + i * 32 + lane < n) atomicAdd(&s_next[d1], 1u);
    }
  }
  {
    const bool full_tile = valid_in_tile == TILE;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const unsigned d = static_cast<unsigned>(key[i] >> shift) & (kRadix - 1);
      if (full_tile || warp_base + i * 32 + lane < n) atomicAdd(&s_cnt[d], 1u);
    }
  }
  __syncthreads();

  // ---- two-level decoupled look-back, part 1: publish (thread d owns digit d)
  // Tiles are grouped by kLookGroup.  A tile sums the partial counts of the earlier tiles of its own group
  // (all loads independent), the last tile of a group publishes the group aggregate, and the prefix over
  // earlier groups comes from a look-back over the group words (aggregate -> inclusive, as in single-level
  // decoupled look-back).  When all tiles of a sort start together (n up to a few million keys: tiles ~ SM
  // slots) a single-level walk is a dependent chain of ~tiles/16 L2 round trips and reads ~tiles^2/4 words;
  // here it is <= 2 + groups/8 round trips and <= (kLookGroup + groups) words per tile and digit.
  const unsigned grp_id = tile / kLookGroup;
  const bool closes_group = tile % kLookGroup == kLookGroup - 1;
  unsigned long long* my_group = gstate + static_cast<size_t>(grp_id) * kRadix + tid;
  const unsigned long long count = s_cnt[tid];  // valid keys of the tile with digit `tid`
  state_store(state + static_cast<size_t>(tile) * kRadix + tid, code_partial, count);
  auto sum_group_predecessors = [&]() {  // keys with this digit in the earlier tiles of this group
    unsigned long long sum = 0;
    int64_t q = static_cast<int64_t>(tile) - 1;
    const int64_t stop = static_cast<int64_t>(grp_id) * kLookGroup;
    while (q >= stop) {
      unsigned long long w[kLookBatch];
#pragma unroll
      for (int j = 0; j < kLookBatch; ++j) {
        const int64_t qq = q - j;
        w[j] = qq >= stop ? state_load(state + static_cast<size_t>(qq) * kRadix + tid) : 0ull;
      }
      int consumed = 0;
#pragma unroll
      for (int j = 0; j < kLookBatch; ++j) {
        if (consumed == j && q - j >= stop && static_cast<unsigned>(w[j] >> 56) == code_partial) {
          sum += w[j] & kStateValueMask;
          consumed = j + 1;
        }
      }
      q -= consumed;  // unpublished words are polled again
    }
    return sum;
  };
  unsigned long long within = 0;
  if (closes_group) {
    // the tile that closes a group sums the group right away, so that the group's aggregate is out early too
    within = sum_group_predecessors();
    state_store(my_group, grp_id == 0 ? code_inclusive : code_partial, within + count);
  }

  PPG_TRACE(tile, 1);
  // ---- stable rank inside the warp, digit counts per warp
  // Lanes holding the same digit find each other through shared memory: every lane ORs its lane bit into the
  // warp's mask word of its digit, the warp synchronises, and the word read back IS the peer set (one RED.OR +
  // one LDS per key instead of 8 ballots + 8 logic ops).  The highest peer bumps the warp's digit counter and
  // clears the mask word; the mask arrays alternate between rounds, so the clear of round i is ordered before
  // the ORs of round i + 2 by the warp barriers of round i + 1.  Rounds run in item order => stable.
  // ranks (< 512 inside the warp, < 4096 inside the tile) are kept two per register
  uint32_t rank2[ITEMS / 2];
  uint16_t* my_hist = s_whist + warp * kRadix;
  uint32_t* my_mask = s_wmask + warp * (2 * kRadix);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const unsigned d = static_cast<unsigned>(key[i] >> shift) & (kRadix - 1);
    uint32_t* m = my_mask + (i & 1) * kRadix + d;
    atomicOr(m, 1u << lane);
    __syncwarp();
    const unsigned peers = *m;
    const uint32_t before = my_hist[d];
    __syncwarp();
    if (lane == static_cast<unsigned>(31 - __clz(peers))) {
      my_hist[d] = static_cast<uint16_t>(before + static_cast<uint32_t>(__popc(peers)));
      *m = 0;
    }
    const uint32_t r = before + __popc(peers & lanemask_lt());
    rank2[i >> 1] = (i & 1) ? (rank2[i >> 1] | (r << 16)) : r;
  }
  __syncthreads();

  PPG_TRACE(tile, 2);
  if (ghist_next != nullptr) {
    const unsigned c = s_next[tid];
    if (c) atomicAdd(&ghist_next[tid], static_cast<unsigned long long>(c));
  }
  // ---- per digit (thread d): exclusive over warps (the out-of-range slots of the last tile carry all-ones keys:
  // they sit at the very end of the top digit's run, after every valid slot, and are never written out)
  {
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint32_t t = s_whist[w * kRadix + tid];
      s_whist[w * kRadix + tid] = static_cast<uint16_t>(sum);
      sum += t;
    }
  }
  // ---- look-back, part 2: the predecessors' counts have been out since before their ranking
  if (!closes_group) within = sum_group_predecessors();
  unsigned long long before_group = 0;  // keys with this digit in earlier groups
  if (grp_id != 0) {
    int64_t q = static_cast<int64_t>(grp_id) - 1;
    bool done = false;
    while (!done) {
      unsigned long long w[kLookBatch];
#pragma unroll
      for (int j = 0; j < kLookBatch; ++j) {
        const int64_t qq = q - j;
        w[j] = qq >= 0 ? state_load(gstate + static_cast<size_t>(qq) * kRadix + tid) : 0ull;
      }
      int consumed = 0;
#pragma unroll
      for (int j = 0; j < kLookBatch; ++j) {
        if (!done && consumed == j) {
          const unsigned code = static_cast<unsigned>(w[j] >> 56);
          if (code == code_inclusive) {
            before_group += w[j] & kStateValueMask;
            done = true;
          } else if (code == code_partial) {
            before_group += w[j] & kStateValueMask;
            consumed = j + 1;
          }  // else: not published yet -> poll again from this group
        }
      }
      q -= consumed;  // group 0 always ends up inclusive, so the walk ends at q >= 0
    }
    if (closes_group) state_store(my_group, code_inclusive, before_group + within + count);
  }
  const unsigned long long prev = before_group + within;  // keys with this digit in earlier tiles

  PPG_TRACE(tile, 3);
  // ---- block exclusive scans over the 256 digits: slot of the digit in the tile, and in the output
  unsigned long long g = ghist[tid];
  {
    unsigned long long a = warp_inclusive_sum(count);
    unsigned long long b = warp_inclusive_sum(g);
    if (lane == 31) {
      s_scan[warp] = a;
      s_scan[NW + warp] = b;
    }
    __syncthreads();
    unsigned long long oa = 0, ob = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      if (w < warp) {
        oa += s_scan[w];
        ob += s_scan[NW + w];
      }
    }
    const unsigned long long bin_start = oa + a - count;
    const unsigned long long out_start = ob + b - g;
    // the digit's first slot goes into the per-warp offsets: one look-up per key in the reorder instead of two
#pragma unroll
    for (int w = 0; w < NW; ++w) s_whist[w * kRadix + tid] = static_cast<uint16_t>(s_whist[w * kRadix + tid] + bin_start);
    s_gbase[tid] = static_cast<long long>(out_start + prev) - static_cast<long long>(bin_start);
  }
  __syncthreads();

  PPG_TRACE(tile, 4);
  // ---- reorder the tile through shared memory
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const unsigned d = static_cast<unsigned>(key[i] >> shift) & (kRadix - 1);
    const uint32_t r = (i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu);
    const uint32_t pos = my_hist[d] + r;
    rank2[i >> 1] = (i & 1) ? ((rank2[i >> 1] & 0xffffu) | (pos << 16)) : ((rank2[i >> 1] & 0xffff0000u) | pos);
    s_keys[pos] = key[i];
  }
  if (HAS_VALUES) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int64_t idx = warp_base + i * 32 + lane;
      uint32_t v = 0;
      if (idx < n) v = IOTA ? static_cast<uint32_t>(idx) : ld_stream(vals_in + idx);
      s_vals[(i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xffffu)] = v;
    }
  }
  __syncthreads();

  PPG_TRACE(tile, 5);
  // ---- write every digit's run to its global segment
  for (int j = tid; j < valid_in_tile; j += kSortBlock) {
    const KeyT k = s_keys[j];
    const unsigned d = static_cast<unsigned>(k >> shift) & (kRadix - 1);
    const long long dst = s_gbase[d] + j;
    keys_out[dst] = k;
    if (HAS_VALUES) vals_out[dst] = s_vals[j];
  }
  PPG_TRACE(tile, 6);
}

template <typename KeyT, bool HAS_VALUES, int ITEMS>
constexpr size_t onesweep_smem_bytes() {
  return onesweep_reorder_bytes<KeyT, HAS_VALUES, ITEMS>() + (kSortBlock / 32) * kRadix * sizeof(uint16_t) +
         kRadix * sizeof(long long);
}

template <typename KeyT, bool HAS_VALUES, bool IOTA, int ITEMS>
inline int launch_onesweep_pass_items(const KeyT* kin, KeyT* kout, const uint32_t* vin, uint32_t* vout, int64_t n,
                                      int shift, const unsigned long long* ghist, unsigned long long* ghist_next,
                                      unsigned* counter, unsigned long long* state, unsigned long long* gstate,
                                      unsigned pass, cudaStream_t stream) {
  auto kern = onesweep_pass_kernel<KeyT, HAS_VALUES, IOTA, ITEMS, SortMinCtas<ITEMS>::value>;
  constexpr size_t smem = onesweep_smem_bytes<KeyT, HAS_VALUES, ITEMS>();
  static bool configured = false;  // per instantiation
  if (!configured) {
    PPG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // all resident-CTA slots the register file allows must also fit in shared memory (one wave at cfg2 sizes)
    PPG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  kern<<<static_cast<unsigned>(sort_num_tiles(n)), kSortBlock, smem, stream>>>(
      kin, kout, vin, vout, n, shift, ghist, ghist_next, counter, state, gstate, 2 * pass + 1, 2 * pass + 2);
  PPG_LAUNCHED();
  return PPG_OK;
}

template <typename KeyT, bool HAS_VALUES, bool IOTA>
inline int launch_onesweep_pass(const KeyT* kin, KeyT* kout, const uint32_t* vin, uint32_t* vout, int64_t n, int shift,
                                const unsigned long long* ghist, unsigned long long* ghist_next, unsigned* counter,
                                unsigned long long* state, unsigned long long* gstate, unsigned pass, cudaStream_t stream) {
  if (sort_items_for(n) == 8)
    return launch_onesweep_pass_items<KeyT, HAS_VALUES, IOTA, 8>(kin, kout, vin, vout, n, shift, ghist, ghist_next, counter, state, gstate, pass, stream);
  return launch_onesweep_pass_items<KeyT, HAS_VALUES, IOTA, 16>(kin, kout, vin, vout, n, shift, ghist, ghist_next, counter, state, gstate, pass, stream);
}

// Sorts the significant bits [0, end_bit) of keys_a (n elements) with an optional u32 payload.
//   keys_a/vals_a : input, clobbered (used as the ping buffer)     keys_b/vals_b : pong buffer
//   vals_a == nullptr with has_values => payload synthesised as the element index on the first pass
//   zeroed_ws     : sort_state_words(n, end_bit) words, zero on entry
//   *in_b         : 1 if the sorted result ended in the *_b buffers
//   h_pass_ms     : optional host array [P]: CUDA-event time of every digit pass (synchronises; bench probe)
//   digit0_counted: the producer of the keys already added the digit-0 counts to zeroed_ws[0..255] (Digit0Counter)
template <typename KeyT>
inline int radix_sort_pairs(KeyT* keys_a, KeyT* keys_b, uint32_t* vals_a, uint32_t* vals_b, bool has_values,
                            bool iota_payload, int64_t n, int end_bit, unsigned long long* zeroed_ws, int* in_b,
                            cudaStream_t stream, float* h_pass_ms = nullptr, bool digit0_counted = false) {
  const int P = sort_num_passes(end_bit);
  PPG_REQUIRE(P <= kMaxPasses, PPG_ERR_INVALID, "radix sort: %d key bits need more than %d passes", end_bit, kMaxPasses);
  PPG_REQUIRE(n < (1ll << 31), PPG_ERR_INVALID, "radix sort: %lld elements exceed the 2^31 limit", (long long)n);
  cudaEvent_t ev[kMaxPasses + 1];
  if (h_pass_ms != nullptr)
    for (int p = 0; p <= P; ++p) PPG_CUDA_TRY(cudaEventCreate(&ev[p]));
  unsigned long long* ghist = zeroed_ws;
  unsigned long long* counters = zeroed_ws + static_cast<size_t>(P) * kRadix;
  unsigned long long* state = counters + P;
  unsigned long long* gstate = state + static_cast<size_t>(sort_num_tiles(n)) * kRadix;
  *in_b = 0;
  if (n == 0) return PPG_OK;

  // digit 0 only: every pass counts the digits of the next one while it holds the keys
  if (!digit0_counted) {
    radix_histogram_kernel<KeyT><<<grid_for(n, 256 * 8, kNumSMsB200 * 8), 256, 0, stream>>>(keys_a, n, 1, ghist);
    PPG_LAUNCHED();
  }

  KeyT* kin = keys_a;
  KeyT* kout = keys_b;
  uint32_t* vin = vals_a;
  uint32_t* vout = vals_b;
  if (h_pass_ms != nullptr) PPG_CUDA_TRY(cudaEventRecord(ev[0], stream));
  for (int p = 0; p < P; ++p) {
    unsigned* counter = reinterpret_cast<unsigned*>(counters + p);
    const unsigned long long* h = ghist + static_cast<size_t>(p) * kRadix;
    unsigned long long* hn = p + 1 < P ? ghist + static_cast<size_t>(p + 1) * kRadix : nullptr;
    const int shift = p * kRadixBits;
    profile_pass_begin(stream);
    if (!has_values) {
      PPG_TRY((launch_onesweep_pass<KeyT, false, false>(kin, kout, nullptr, nullptr, n, shift, h, hn, counter, state, gstate, p, stream)));
    } else if (p == 0 && iota_payload) {
      PPG_TRY((launch_onesweep_pass<KeyT, true, true>(kin, kout, nullptr, vout, n, shift, h, hn, counter, state, gstate, p, stream)));
    } else {
      PPG_TRY((launch_onesweep_pass<KeyT, true, false>(kin, kout, vin, vout, n, shift, h, hn, counter, state, gstate, p, stream)));
    }
    KeyT* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = (p == 0 && iota_payload) ? vals_a : vin;
    vin = vout; vout = tv;
    *in_b ^= 1;
    profile_pass_end(stream, n, 2 * static_cast<int>(sizeof(KeyT) + (has_values ? sizeof(uint32_t) : 0)));
    if (h_pass_ms != nullptr) PPG_CUDA_TRY(cudaEventRecord(ev[p + 1], stream));
  }
  if (h_pass_ms != nullptr) {
    PPG_CUDA_TRY(cudaStreamSynchronize(stream));
    for (int p = 0; p < P; ++p) PPG_CUDA_TRY(cudaEventElapsedTime(&h_pass_ms[p], ev[p], ev[p + 1]));
    for (int p = 0; p <= P; ++p) cudaEventDestroy(ev[p]);
  }
  return PPG_OK;
}

}  // namespace ppg
