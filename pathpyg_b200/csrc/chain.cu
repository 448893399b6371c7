// a4-a6 without a global sort per order: the De Bruijn layers of consecutive orders built by EXPANDING the line graph
// in the merged order of its source nodes.
//
// The reference (lift_order.py:109-152 called from multi_order_model.py:124-192) lifts the line graph L_k of order k
// (one column per pair (e, f) of consecutive level-(k-1) items, ascending in (e, f)), maps both ends through the
// inverse index of layer k-1 and coalesces: a sort of E_k keys (row, col) = (id(e), id(f)).  aggregate.cu does that
// with an LSD radix sort: ceil(2 bits(n_k) / 8) passes of 24 B per pair and direction.
//
// Here the sort is replaced by the ORDER OF GENERATION.  Layer k-1 was produced in (row, col) order, so the stable order
// P of the level-(k-1) items by their merged id id(e) is known (it is the order in which they were written).  Expanding
// the items in THAT order -- for s = 0, 1, ...: e = P[s], emit (e, f) for the continuations f of e -- yields the pairs
// grouped by row id(e), rows ascending, and inside a row in ascending (e, f): exactly the state of a stable sort by
// row.  What is left is a stable sort by col INSIDE every row, and rows are short (out-degree x multiplicity): a CTA
// stages a tile of whole rows in shared memory and ranks every pair among the pairs of its row by counting
// (rank = #{j in row: (col_j, j) < (col_i, i)}), which needs no key bits, no passes and no global traffic.  Rows longer
// than `heavy` pairs (hubs) are left in generation order, listed, and put in order afterwards by ONE radix sort over
// only those pairs (ppg_chain_heavy_fix).
//
//   level k-1 -> k:  count (item order)  : deg[e] = #continuations, ptr_k = exclusive scan            (lift.cu for k = 2)
//                    count (P order)     : offP = exclusive scan of deg[P[s]]; tile boundaries
//                    tiles               : expand + rank inside rows -> (row, col, label) in final order; tail of every pair
//                    heads               : run heads -> merged id of every pair (by slot and by label), run starts
//                    fill                : merged edges + weights (sum of a run in slot order = stream order)
//
// Labels: a pair keeps the index it has in the reference's line graph (label = ptr_k[e] + j), so `inverse_idx` of the
// next layer is the id array itself, and ties inside a run are in label order = the summation order of the reference's
// coalesce (stable sort + scatter_add).
//
// Algorithmic bytes per pair and level (c = continuation factor E_k / E_{k-1}): reads 4/c (P) + 16/c (first, ptr pair) +
// 8/c + 4 (id of the continuation) , writes 12 (row, col, label) + 4 (tail) + 8 (ids) -- about 40-50 B against
// 48 * passes of the sort it replaces.
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace ppg {

// Tile geometry, measured on B200 (device time of from_temporal_graph at cfg3 / cfg5; `make variant`):
//   256 threads x 1024 + 256 slots: 6.13 / 107.9 ms      512 x 1280 + 256: 5.72 / 104.1 ms  (the default)
//   256 x  768 + 256: 6.17 / 109.0 ms                    512 x  768 + 256: 5.93 / 112.7 ms
//   512 x 2304 + 256: 6.15 / 108.4 ms                   1024 x 1792 + 256: 5.90 / 109.8 ms
#ifndef PPG_CHAIN_BLOCK
#define PPG_CHAIN_BLOCK 512
#endif
#ifndef PPG_CHAIN_TILE
#define PPG_CHAIN_TILE 1280
#endif
constexpr int kChainBlock = PPG_CHAIN_BLOCK;
constexpr int kChainTile = PPG_CHAIN_TILE;               // nominal slots (pairs) per tile
constexpr int kChainHeavyMax = 256;                      // rows above this never take the in-tile ranking
constexpr int kChainCap = kChainTile + kChainHeavyMax;   // slots staged per chunk: every non-heavy row of a tile fits
constexpr int kChainPerThread = kChainCap / kChainBlock; // 5
static_assert(kChainCap % kChainBlock == 0 && kChainCap < 65535, "chunk geometry");

// result words of one level (device int64[8], zeroed by the caller)
enum { kResHeads = 0, kResStatus = 1, kResHeavySlots = 2, kResHeavyRows = 3 };
constexpr unsigned long long kChainStatusIdOutOfRange = 1ull;

// ------------------------------------------------------------------ offsets of a scan
struct ChainOffsetConsumer {
  unsigned long long* off;
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long v, unsigned long long prefix) const {
    off[i] = prefix;
    if (i == n - 1) off[n] = prefix + v;
  }
};

// ------------------------------------------------------------------ count in merged (P) order
struct ChainSortedProducer {
  const uint32_t* P;                     // [n] item at every position of the merged order
  const uint32_t* first;                 // [n] by item
  const unsigned long long* ptr_next;    // [n + 1] by item
  const float* w_item;                   // [n] by item or nullptr
  int64_t limit;                         // items >= limit are not expanded
  uint32_t* firstP;
  uint32_t* lblP;
  float* wP;
  __device__ unsigned long long operator()(int64_t s) const {
    const uint32_t q = ld_stream(P + s);
    const unsigned long long a = ptr_next[q];
    const unsigned long long b = ptr_next[q + 1];
    firstP[s] = first[q];
    lblP[s] = static_cast<uint32_t>(a);
    if (wP != nullptr) wP[s] = w_item[q];
    return static_cast<int64_t>(q) < limit ? b - a : 0ull;
  }
};
struct ChainSortedConsumer {
  unsigned long long* offP;  // [n + 1]
  uint2* srcbound;           // [tiles]: first source of the row that holds slot t * kChainTile, and of the next row
  const uint32_t* rowid;     // [n] row of every source
  const uint32_t* run_start; // [rows + 1]
  int64_t n;
  __device__ void operator()(int64_t s, unsigned long long v, unsigned long long prefix) const {
    offP[s] = prefix;
    if (s == n - 1) offP[n] = prefix + v;
    if (v) {
      const unsigned long long t0 = (prefix + kChainTile - 1) / kChainTile;
      if (t0 * kChainTile < prefix + v) {
        const uint32_t r0 = rowid[s];
        const uint2 b = make_uint2(run_start[r0], run_start[r0 + 1]);
        for (unsigned long long t = t0; t * kChainTile < prefix + v; ++t) srcbound[t] = b;
      }
    }
  }
};

// ------------------------------------------------------------------ tiles
struct ChainTileArgs {
  // sources (items of the previous level) in merged order
  const unsigned long long* offP;  // [n_sources + 1] first slot                       (FIRST: identity)
  const uint32_t* firstP;          // [n_sources] first continuation                   (FIRST: identity)
  const uint32_t* lblP;            // [n_sources] label of the first slot              (FIRST: event at the position)
  const float* wP;                 // [n_sources] weight of the source or nullptr      (FIRST: by event)
  const uint32_t* run_start;       // [n_rows + 1] first source of every row
  const uint32_t* rowid;           // [n_sources] row of every source
  const unsigned long long* node_prev;  // by item: merged id | row pointer of THIS level << 32 ([items + 1]; FIRST, DIST: unused)
  const uint32_t* via;             // continuation position -> item (temporal level) or nullptr
  const int64_t* dst;              // FIRST: target node of every event
  const uint32_t* srcbound;        // [tiles][2] (see ChainSortedConsumer)              (FIRST: unused)
  int64_t n_sources, n_rows, n_slots, num_cols;
  int heavy;
  uint32_t *rowS, *colS, *labS;
  float* wS;
  // what the next level needs (NEXT): by label the merged id | number of continuations << 32 (FIRST: the id word only),
  // and in slot order -- the merged order P of the next level -- the first continuation and their number
  unsigned long long* node_out;
  uint32_t *firstS, *degS;
  uint2* heavy_list;
  unsigned long long* result;
  // run heads, fused: valid when no row of the level is heavy (otherwise ppg_chain_heavy_fix + ppg_chain_heads redo them)
  unsigned long long* tile_state;  // [tiles] zeroed
  unsigned code_partial, code_inclusive;
  uint32_t *idS, *run_start_out;
  // distributed build (DIST): ids are global, rows are local, the slots leave as 16-byte records
  const uint4* info;               // by item: {global id, last first-order node, row pointer of THIS level, -} ([items + 1])
  const uint32_t* row_value;       // by local row: its global id (the record's row)
  uint4* rec;                      // out [n_slots]: {col id, row id, last node of the pair, weight}
  uint4* info_out;                 // by label [n_slots + 1]: word 2 receives the pair's number of continuations (or nullptr)
  int w_stride;                    // wP[s * w_stride] (DIST: the weight word of the previous level's records)
};

// largest s in [lo, hi) with off[s] <= target; requires off[lo] <= target
__device__ __forceinline__ int64_t warp_search_range(const unsigned long long* __restrict__ off, int64_t lo, int64_t hi,
                                                     unsigned long long target) {
  const unsigned lane = lane_id();
  while (hi - lo > 1) {
    const int64_t step = ceil_div(hi - lo, 32);
    const int64_t p = lo + static_cast<int64_t>(lane) * step;
    const bool le = p < hi && off[p] <= target;
    const int cnt = __popc(__ballot_sync(kFullMask, le));
    const int64_t nlo = lo + static_cast<int64_t>(cnt - 1) * step;
    const int64_t nhi = lo + static_cast<int64_t>(cnt) * step;
    lo = nlo;
    hi = nhi < hi ? nhi : hi;
  }
  return lo;
}

// Look-back over the tiles of a launch: the number of run heads (merged edges) in the tiles before this one.  Tiles are
// taken in blockIdx order (CTAs are dispatched in that order, so a tile only waits for tiles that are running or done);
// the words carry a code per call like the scan's (common.cuh).  A tile publishes its count as soon as it has it and
// resolves its prefix after it has written everything that does not need it.
__device__ __forceinline__ void chain_tile_publish(unsigned long long* __restrict__ state, unsigned tile, unsigned long long total,
                                                   unsigned code_partial, unsigned code_inclusive) {
  state_store(&state[tile], tile == 0 ? code_inclusive : code_partial, total);
}
// called by warp 0, some time after chain_tile_publish: by then the predecessors' words are usually out
__device__ __forceinline__ unsigned long long chain_tile_prefix(unsigned long long* __restrict__ state, unsigned tile,
                                                                unsigned long long total, unsigned code_partial,
                                                                unsigned code_inclusive) {
  const unsigned lane = lane_id();
  if (tile == 0) return 0;
  unsigned long long acc = 0;
  int64_t newest = static_cast<int64_t>(tile) - 1;
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
      const unsigned first = __ffs(incl) - 1;
      acc += warp_sum(lane <= first ? val : 0ull);
      break;
    }
    acc += warp_sum(val);
    newest -= 32;
  }
  if (lane == 0) state_store(&state[tile], code_inclusive, acc + total);
  return acc;
}

template <bool FIRST, bool DIST, bool NEXT>
__global__ void __launch_bounds__(kChainBlock)
chain_tile_kernel(ChainTileArgs a) {
  extern __shared__ __align__(16) uint32_t chain_smem[];
  uint32_t* s_mark = chain_smem;                                   // (row's first slot + 1) << 16 | (source's first slot + 1)
  uint32_t* s_col = s_mark + kChainCap;
  int32_t* s_soff = reinterpret_cast<int32_t*>(s_col + kChainCap);  // at a source's first slot: its first slot - chunk base (may be < 0)
  uint32_t* s_first = reinterpret_cast<uint32_t*>(s_soff + kChainCap);
  uint32_t* s_lbl = s_first + kChainCap;
  float* s_w = reinterpret_cast<float*>(s_lbl + kChainCap);
  uint32_t* s_row = reinterpret_cast<uint32_t*>(s_w + kChainCap);   // at a row's first slot: the row
  uint32_t* s_len = s_row + kChainCap;                              // at a row's first slot: its length; later: run index of every slot
  uint16_t* s_perm = reinterpret_cast<uint16_t*>(s_len + kChainCap);
  uint32_t* s_nfirst = reinterpret_cast<uint32_t*>(s_perm + kChainCap);  // NEXT only: first continuation at the next level ...
  uint32_t* s_ndeg = s_nfirst + kChainCap;                               // ... and their number
  __shared__ int64_t s_geo[4];
  __shared__ int64_t s_bound[2];
  __shared__ uint32_t s_warp_max[kChainBlock / 32];
  __shared__ uint32_t s_warp_cnt[kChainBlock / 32];
  __shared__ unsigned long long s_base_id;
  __shared__ unsigned long long s_total;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const unsigned lane = lane_id();
  auto off = [&](int64_t s) -> int64_t { return FIRST ? s : static_cast<int64_t>(a.offP[s]); };

  // ---- tile geometry: the rows whose first slot lies in [t T, (t + 1) T).  The source that holds slot t T belongs to row
  // r0; the tile starts with r0 if that row starts exactly at t T, else with the next row (the count pass left the
  // first sources of both rows in srcbound: two dependent loads here).
  if (tid < 2) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) + tid;
    const int64_t target = t * kChainTile;
    int64_t src = a.n_sources, slot = a.n_slots;
    if (target < a.n_slots) {
      int64_t rs0, rs1;
      if (FIRST) {
        const int64_t r0 = a.rowid[target];
        rs0 = a.run_start[r0];
        rs1 = a.run_start[r0 + 1];
      } else {
        const uint2 b = reinterpret_cast<const uint2*>(a.srcbound)[t];
        rs0 = b.x;
        rs1 = b.y;
      }
      const int64_t o0 = off(rs0), o1 = off(rs1);
      src = o0 == target ? rs0 : rs1;
      slot = o0 == target ? o0 : o1;
    }
    s_geo[tid] = src;
    s_geo[2 + tid] = slot;
  }
  __syncthreads();
  const int64_t sa = s_geo[0], sb = s_geo[1];
  const int64_t base = s_geo[2], end = s_geo[3];
  const int64_t n = end - base;
  const bool last_tile = blockIdx.x == gridDim.x - 1;

  if (n == 0) {  // a row of an earlier tile covers this window: nothing to do but to keep the look-back chain whole
    if (warp == 0) {
      if (lane == 0) chain_tile_publish(a.tile_state, blockIdx.x, 0ull, a.code_partial, a.code_inclusive);
      const unsigned long long before = chain_tile_prefix(a.tile_state, blockIdx.x, 0ull, a.code_partial, a.code_inclusive);
      if (last_tile && lane == 0) {
        a.result[kResHeads] = before;
        if (a.run_start_out != nullptr) a.run_start_out[before] = static_cast<uint32_t>(a.n_slots);
      }
    }
    return;
  }

  for (int64_t c0 = 0; c0 < n; c0 += kChainCap) {
    const int64_t cb = base + c0;
    const int cn = static_cast<int>(n - c0 < kChainCap ? n - c0 : kChainCap);
    for (int i = tid; i < kChainCap; i += kChainBlock) {
      s_mark[i] = 0;
      s_row[i] = 0;
    }
    // sources of this chunk: all of the tile's, unless the tile is longer than a chunk (then its last row is heavy)
    if (n <= kChainCap) {
      if (tid == 0) {
        s_bound[0] = sa;
        s_bound[1] = sb - 1;
      }
    } else if (FIRST) {
      if (tid == 0) {
        s_bound[0] = cb;
        s_bound[1] = cb + cn - 1;
      }
    } else if (warp < 2) {
      const unsigned long long target = static_cast<unsigned long long>(warp == 0 ? cb : cb + cn - 1);
      const int64_t r = (warp == 0 && c0 == 0) ? sa : warp_search_range(a.offP, sa, sb, target);
      if (lane == 0) s_bound[warp] = r;
    }
    __syncthreads();
    const int64_t s_lo = s_bound[0], s_hi = s_bound[1];
    for (int64_t s = s_lo + tid; s <= s_hi; s += kChainBlock) {
      const int64_t o0 = off(s), o1 = off(s + 1);
      const uint32_t rid = a.rowid[s];
      const int64_t x = o0 - cb;
      if (c0 == 0 && x < cn && (s == 0 || a.rowid[s - 1] != rid)) {
        // first source of a row: the row's slots (if it has any) start here; of several empty rows and one that owns
        // the slot, the last (largest) is the owner
        atomicOr(&s_mark[x], static_cast<uint32_t>(x + 1) << 16);
        atomicMax(&s_row[x], rid);
      }
      if (o1 > o0 && o1 > cb && x < cn) {
        const int f0 = x > 0 ? static_cast<int>(x) : 0;
        s_soff[f0] = static_cast<int32_t>(x);
        s_first[f0] = FIRST ? static_cast<uint32_t>(s) : a.firstP[s];
        const uint32_t lbl = a.lblP[s];
        s_lbl[f0] = lbl;
        if (DIST || a.wS != nullptr) s_w[f0] = FIRST ? a.wP[lbl] : a.wP[s * a.w_stride];
        uint32_t mark = static_cast<uint32_t>(f0 + 1);
        if (c0 > 0 && f0 == 0) {  // chunks after the first hold only the tile's last (heavy) row
          mark |= 1u << 16;
          s_row[0] = rid;
        }
        atomicOr(&s_mark[f0], mark);
      }
    }
    __syncthreads();

    // ---- both marks spread to the right: inclusive max-scan on the two 16-bit halves
    {
      uint32_t m[kChainPerThread];
#pragma unroll
      for (int i = 0; i < kChainPerThread; ++i) m[i] = s_mark[tid * kChainPerThread + i];
#pragma unroll
      for (int i = 1; i < kChainPerThread; ++i) m[i] = __vmaxu2(m[i], m[i - 1]);
      uint32_t inc = m[kChainPerThread - 1];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(kFullMask, inc, d);
        if (lane >= static_cast<unsigned>(d)) inc = __vmaxu2(inc, y);
      }
      uint32_t before = __shfl_up_sync(kFullMask, inc, 1);
      if (lane == 0) before = 0;
      if (lane == 31) s_warp_max[warp] = inc;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < kChainBlock / 32; ++w)
        if (w < warp) before = __vmaxu2(before, s_warp_max[w]);
#pragma unroll
      for (int i = 0; i < kChainPerThread; ++i) s_mark[tid * kChainPerThread + i] = __vmaxu2(m[i], before);
    }
    __syncthreads();

    // ---- every slot: its continuation, the continuation's merged id (the column), its tail; the last slot of a row
    // records the row's length
#pragma unroll
    for (int k = 0; k < kChainPerThread; ++k) {
      const int i = k * kChainBlock + tid;
      if (i < cn) {
        const uint32_t mk = s_mark[i];
        const int f0 = static_cast<int>(mk & 0xffffu) - 1;
        const int g0 = static_cast<int>(mk >> 16) - 1;
        const int64_t jj = static_cast<int64_t>(i) - s_soff[f0];
        const uint32_t label = s_lbl[f0] + static_cast<uint32_t>(jj);
        uint32_t col;
        if (FIRST) {
          const int64_t v = a.dst[label];
          col = static_cast<uint32_t>(v);
          if (v < 0 || v >= a.num_cols) {
            atomicOr(a.result + kResStatus, kChainStatusIdOutOfRange);
            col = 0;
          }
        } else {
          const uint32_t g = s_first[f0] + static_cast<uint32_t>(jj);
          const uint32_t item = a.via != nullptr ? a.via[g] : g;
          if (DIST) {  // the other words of the continuation's info are fetched again when the record is written
            col = a.info[item].x;
          } else {     // merged id | first level-k item whose source is this item << 32
            const unsigned long long nd = a.node_prev[item];
            col = static_cast<uint32_t>(nd);
#ifndef PPG_NEXT_REREAD
            if (NEXT) {
              const uint32_t p0 = static_cast<uint32_t>(nd >> 32);
              s_nfirst[i] = p0;
              s_ndeg[i] = static_cast<uint32_t>(a.node_prev[item + 1] >> 32) - p0;
            }
#endif
          }
        }
        s_col[i] = col;
        if (i == cn - 1 || static_cast<int>(s_mark[i + 1] >> 16) - 1 != g0) {
          int64_t len = static_cast<int64_t>(i) + 1 - g0;
          if (i == cn - 1 && c0 + cn < n) len = (n - c0) - g0;   // the tile's last row runs on into the next chunk
          if (c0 > 0 && g0 == 0) {
            len = 0xffffffffll;                                  // ... and this is one of its later chunks
          } else if (len > a.heavy) {
            const unsigned long long at = atomicAdd(a.result + kResHeavyRows, 1ull);
            atomicAdd(a.result + kResHeavySlots, static_cast<unsigned long long>(len));
            a.heavy_list[at] = make_uint2(static_cast<uint32_t>(cb + g0), static_cast<uint32_t>(len));
          }
          s_len[g0] = len > 0xffffffffll ? 0xffffffffu : static_cast<uint32_t>(len);
        }
      }
    }
    __syncthreads();

    // ---- stable rank of every slot among the slots of its row (rows above `heavy` keep their order)
#pragma unroll
    for (int k = 0; k < kChainPerThread; ++k) {
      const int i = k * kChainBlock + tid;
      if (i < cn) {
        const int g0 = static_cast<int>(s_mark[i] >> 16) - 1;
        const uint32_t len = s_len[g0];
        int at = i;
        if (len > 1 && len <= static_cast<uint32_t>(a.heavy)) {
          const uint32_t key = s_col[i];
          int rank = 0;
          const int g1 = g0 + static_cast<int>(len);
          for (int j = g0; j < g1; ++j) {
            const uint32_t c = s_col[j];
            rank += (c < key || (c == key && j < i)) ? 1 : 0;
          }
          at = g0 + rank;
        }
        s_perm[at] = static_cast<uint16_t>(i);
      }
    }
    __syncthreads();

    // ---- run heads of the chunk in its final order (a tile starts with a row, so its first slot is a head): every
    // slot gets the index of its run inside the chunk
    {
      const int p0 = tid * kChainPerThread;
      uint32_t prev_row = 0xffffffffu, prev_col = 0;
      if (p0 > 0 && p0 <= cn) {
        prev_row = s_mark[p0 - 1] >> 16;
        prev_col = s_col[s_perm[p0 - 1]];
      }
      uint32_t run[kChainPerThread];
      uint32_t cnt = 0;
#pragma unroll
      for (int k = 0; k < kChainPerThread; ++k) {
        const int p = p0 + k;
        if (p < cn) {
          const uint32_t row = s_mark[p] >> 16;
          const uint32_t col = s_col[s_perm[p]];
          cnt += (row != prev_row || col != prev_col) ? 1u : 0u;
          prev_row = row;
          prev_col = col;
        }
        run[k] = cnt;
      }
      const uint32_t inc = warp_inclusive_sum(cnt);
      if (lane == 31) s_warp_cnt[warp] = inc;
      __syncthreads();   // also: every thread has read s_len for the ranking
      uint32_t before = inc - cnt;
      uint32_t total = 0;
#pragma unroll
      for (int w = 0; w < kChainBlock / 32; ++w) {
        if (w < warp) before += s_warp_cnt[w];
        total += s_warp_cnt[w];
      }
#pragma unroll
      for (int k = 0; k < kChainPerThread; ++k)
        if (p0 + k < cn) s_len[p0 + k] = before + run[k] - 1;
      if (c0 == 0 && tid == 0) {
        s_total = total;
        chain_tile_publish(a.tile_state, blockIdx.x, total, a.code_partial, a.code_inclusive);
      }
    }
    __syncthreads();

    // ---- write the chunk in its final order: first what does not depend on the tiles before this one ...
    uint32_t label[kChainPerThread];
    uint32_t ndeg[kChainPerThread];
#pragma unroll
    for (int k = 0; k < kChainPerThread; ++k) {
      const int p = k * kChainBlock + tid;
      ndeg[k] = 0;
      if (p < cn) {
        const int i = s_perm[p];
        const int g0 = static_cast<int>(s_mark[p] >> 16) - 1;
        const int f0 = static_cast<int>(s_mark[i] & 0xffffu) - 1;
        const int64_t jj = static_cast<int64_t>(i) - s_soff[f0];
        label[k] = s_lbl[f0] + static_cast<uint32_t>(jj);
        if (NEXT) {
#ifndef PPG_NEXT_REREAD
          ndeg[k] = s_ndeg[i];
          st_stream(a.firstS + cb + p, s_nfirst[i]);
#else
          // experiment (-DPPG_NEXT_REREAD): the continuation's node word again (L1 / L2) rather than two more shared-memory
          // words per slot, which cost the fifth resident CTA.  Measured on B200: 6.66 vs 6.13 ms at cfg3, 108.9 vs 107.9 ms at
          // cfg5 -- the extra dependent load in the store phase costs more than the fifth CTA gains; not the default.
          const uint32_t g = s_first[f0] + static_cast<uint32_t>(jj);
          const uint32_t item = a.via != nullptr ? a.via[g] : g;
          const uint32_t p0 = static_cast<uint32_t>(a.node_prev[item] >> 32);
          ndeg[k] = static_cast<uint32_t>(a.node_prev[item + 1] >> 32) - p0;
          st_stream(a.firstS + cb + p, p0);
#endif
          st_stream(a.degS + cb + p, ndeg[k]);
        }
        st_stream(a.labS + cb + p, label[k]);
        if (DIST) {
          // the continuation's info words again (this CTA read them a moment ago: L1 / L2), so that neither the last
          // node nor the next level's counts have to be parked in shared memory
          const uint32_t g = s_first[f0] + static_cast<uint32_t>(jj);
          const uint32_t item = a.via != nullptr ? a.via[g] : g;
          const uint4 nf = a.info[item];
          if (a.firstS != nullptr) {
            const uint32_t deg = a.info[item + 1].z - nf.z;
            st_stream(a.firstS + cb + p, nf.z);
            st_stream(a.degS + cb + p, deg);
            reinterpret_cast<uint32_t*>(a.info_out + label[k])[2] = deg;
          }
          __stcs(a.rec + cb + p, make_uint4(s_col[i], a.row_value[s_row[g0]], nf.y, __float_as_uint(s_w[f0])));
        } else {
          st_stream(a.rowS + cb + p, s_row[g0]);
          st_stream(a.colS + cb + p, s_col[i]);
          if (a.wS != nullptr) st_stream(a.wS + cb + p, s_w[f0]);
        }
      }
    }
    // ... then the merged ids, once the number of run heads before this tile is known
    if (c0 == 0 && warp == 0) {
      const unsigned long long total = s_total;
      const unsigned long long prefix = chain_tile_prefix(a.tile_state, blockIdx.x, total, a.code_partial, a.code_inclusive);
      if (lane == 0) {
        s_base_id = prefix;
        if (last_tile) {
          a.result[kResHeads] = prefix + total;
          if (a.run_start_out != nullptr) a.run_start_out[prefix + total] = static_cast<uint32_t>(a.n_slots);
        }
      }
    }
    __syncthreads();
    const unsigned long long base_id = s_base_id;
#pragma unroll
    for (int k = 0; k < kChainPerThread; ++k) {
      const int p = k * kChainBlock + tid;
      if (p < cn) {
        const uint32_t run = s_len[p];
        const uint32_t id = static_cast<uint32_t>(base_id + run);
        if (a.idS != nullptr) st_stream(a.idS + cb + p, id);
        if (FIRST) {
          if (a.node_out != nullptr) reinterpret_cast<uint32_t*>(a.node_out)[2 * static_cast<size_t>(label[k])] = id;
        } else if (NEXT) {
          a.node_out[label[k]] = static_cast<unsigned long long>(id) | (static_cast<unsigned long long>(ndeg[k]) << 32);
        }
        if (a.run_start_out != nullptr && (p == 0 || s_len[p - 1] != run)) a.run_start_out[id] = static_cast<uint32_t>(cb + p);
      }
    }
    __syncthreads();
  }
}

constexpr size_t kChainTileSmem = static_cast<size_t>(kChainCap) * (8 * 4 + 2);
constexpr size_t kChainTileSmemNext = kChainTileSmem + static_cast<size_t>(kChainCap) * 8;

// ------------------------------------------------------------------ heads
struct ChainHeadProducer {
  const uint32_t* rowS;
  const uint32_t* colS;
  __device__ unsigned long long operator()(int64_t i) const {
    if (i == 0) return 1ull;
    return (rowS[i] != rowS[i - 1] || colS[i] != colS[i - 1]) ? 1ull : 0ull;
  }
};
struct ChainHeadConsumer {
  const uint32_t* labS;
  uint32_t* idS;        // merged id of every slot (nullable)
  uint32_t* id_item;    // merged id of every item, by label, `id_stride` words apart (nullable)
  int id_stride;
  uint32_t* run_start;  // [heads + 1]
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long head, unsigned long long prefix) const {
    const uint32_t id = static_cast<uint32_t>(prefix + head - 1);
    if (head) run_start[prefix] = static_cast<uint32_t>(i);
    if (i == n - 1) run_start[prefix + head] = static_cast<uint32_t>(n);
    if (idS != nullptr) idS[i] = id;
    if (id_item != nullptr) id_item[static_cast<size_t>(ld_stream(labS + i)) * id_stride] = id;
  }
};

// ------------------------------------------------------------------ fill
__global__ void __launch_bounds__(256)
chain_fill_kernel(const uint32_t* __restrict__ rowS, const uint32_t* __restrict__ colS, const float* __restrict__ wS,
                  const uint32_t* __restrict__ run_start, int64_t num_out, int64_t* __restrict__ out_ei,
                  float* __restrict__ out_w) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < num_out; r += stride) {
    const uint32_t a = run_start[r];
    const uint32_t b = run_start[r + 1];
    st_stream(out_ei + r, static_cast<int64_t>(rowS[a]));
    st_stream(out_ei + num_out + r, static_cast<int64_t>(colS[a]));
    float acc;
    if (wS == nullptr) {
      acc = static_cast<float>(b - a);  // unit weights: the sum of b - a ones is exact in fp32 up to 2^24 and rounds like it above
      if (b - a > (1u << 24)) {
        acc = 0.f;
        for (uint32_t i = a; i < b; ++i) acc += 1.f;
      }
    } else {
      acc = wS[a];
      for (uint32_t i = a + 1; i < b; ++i) acc += wS[i];  // slot order = label order = stream order
    }
    st_stream(out_w + r, acc);
  }
}

__global__ void __launch_bounds__(256)
chain_widen_kernel(const uint32_t* __restrict__ in, int in_stride, int64_t n, int64_t* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    st_stream(out + i, static_cast<int64_t>(ld_stream(in + i * in_stride)));
}

// level 1: the row pointer of the event graph (u64, lift.cu) into the ptr word of the events' node words
__global__ void __launch_bounds__(256)
chain_node_ptr_kernel(const unsigned long long* __restrict__ off, int64_t n_plus_1, uint32_t* __restrict__ words, int word_stride) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_plus_1; i += stride)
    words[i * word_stride] = static_cast<uint32_t>(ld_stream(off + i));
}

// label-order scan of the continuation counts the tiles left in the per-item words: count -> row pointer of the next
// level.  The word of item i is words[i * stride] (node words: the high word, stride 2; info words of the distributed
// build: the third word, stride 4).
struct NodeDegProducer {
  const uint32_t* words;
  int stride;
  __device__ unsigned long long operator()(int64_t i) const { return words[i * stride]; }
};
struct NodePtrConsumer {
  uint32_t* words;
  int stride;
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long v, unsigned long long prefix) const {
    words[i * stride] = static_cast<uint32_t>(prefix);
    if (i == n - 1) words[n * stride] = static_cast<uint32_t>(prefix + v);
  }
};

// count pass of levels >= 3 in merged order: the counts are already in that order (degS of the previous level's tiles);
// one gather per source fetches the label of its first pair
struct ChainSortedProducer2 {
  const uint32_t* P;
  const uint32_t* degS;
  const uint32_t* ptr_words;   // by label, `stride` words apart: row pointer of the next level (after the label-order scan)
  int stride;
  int64_t limit;
  uint32_t* lblP;
  __device__ unsigned long long operator()(int64_t s) const {
    const uint32_t q = ld_stream(P + s);
    lblP[s] = ptr_words[static_cast<int64_t>(q) * stride];
    return static_cast<int64_t>(q) < limit ? ld_stream(degS + s) : 0u;
  }
};

// ------------------------------------------------------------------ heavy rows
struct HeavyLayout {
  unsigned long long* scan_ws;
  unsigned long long* sort_ws;
  size_t zero_bytes;
  unsigned long long* hoff;  // [rows + 1]
  unsigned long long *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  uint32_t* t_lab;
  uint32_t* t_extra[3];
  float* t_w;
  uint32_t* t_dest;
  int bits;
  HeavyLayout(Workspace& ws, int64_t slots, int64_t rows, int64_t n_slots) {
    bits = 32 + bits_for(n_slots > 0 ? static_cast<uint64_t>(n_slots - 1) : 0);
    scan_ws = ws.take<unsigned long long>(scan_state_words(rows));
    sort_ws = ws.take<unsigned long long>(sort_state_words(slots, bits));
    zero_bytes = ws.used;
    hoff = ws.take<unsigned long long>(static_cast<size_t>(rows) + 1);
    keys_a = ws.take<unsigned long long>(static_cast<size_t>(slots));
    keys_b = ws.take<unsigned long long>(static_cast<size_t>(slots));
    vals_a = ws.take<uint32_t>(static_cast<size_t>(slots));
    vals_b = ws.take<uint32_t>(static_cast<size_t>(slots));
    t_lab = ws.take<uint32_t>(static_cast<size_t>(slots));
    for (int j = 0; j < 3; ++j) t_extra[j] = ws.take<uint32_t>(static_cast<size_t>(slots));
    t_w = ws.take<float>(static_cast<size_t>(slots));
    t_dest = ws.take<uint32_t>(static_cast<size_t>(slots));
  }
};
struct HeavyExtras {  // further slot arrays that travel with the pairs of a heavy row (firstS, degS, lastS)
  uint32_t* p[3];
};
struct HeavyLenProducer {
  const uint2* list;
  __device__ unsigned long long operator()(int64_t k) const { return list[k].y; }
};

// compact index c -> (heavy row k, offset): key = first slot of the row << 32 | col, payload = the slot
__global__ void __launch_bounds__(256)
heavy_compact_kernel(const uint2* __restrict__ list, const unsigned long long* __restrict__ hoff, int64_t rows, int64_t slots,
                     const uint32_t* __restrict__ colS, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < slots; c += stride) {
    int64_t lo = 0, hi = rows;  // last k with hoff[k] <= c
    while (hi - lo > 1) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (hoff[mid] <= static_cast<unsigned long long>(c)) lo = mid; else hi = mid;
    }
    const uint2 row = list[lo];
    const uint32_t slot = row.x + static_cast<uint32_t>(c - static_cast<int64_t>(hoff[lo]));
    keys[c] = (static_cast<unsigned long long>(row.x) << 32) | colS[slot];
    vals[c] = slot;
  }
}

__global__ void __launch_bounds__(256)
heavy_gather_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t slots,
                    const uint32_t* __restrict__ labS, const float* __restrict__ wS, HeavyExtras extraS,
                    uint32_t* __restrict__ t_lab, float* __restrict__ t_w, HeavyExtras t_extra,
                    uint32_t* __restrict__ t_dest) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < slots; i += stride) {
    const unsigned long long k = keys[i];
    const unsigned long long row_key = k & 0xffffffff00000000ull;
    int64_t lo = -1, hi = i;  // first index with key >= row_key
    while (hi - lo > 1) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (keys[mid] >= row_key) hi = mid; else lo = mid;
    }
    const uint32_t src = vals[i];
    t_dest[i] = static_cast<uint32_t>(k >> 32) + static_cast<uint32_t>(i - hi);
    t_lab[i] = labS[src];
    if (wS != nullptr) t_w[i] = wS[src];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (extraS.p[j] != nullptr) t_extra.p[j][i] = extraS.p[j][src];
  }
}

__global__ void __launch_bounds__(256)
heavy_write_kernel(const unsigned long long* __restrict__ keys, int64_t slots, const uint32_t* __restrict__ t_lab,
                   const float* __restrict__ t_w, HeavyExtras t_extra, const uint32_t* __restrict__ t_dest,
                   uint32_t* __restrict__ colS, uint32_t* __restrict__ labS, float* __restrict__ wS, HeavyExtras extraS) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < slots; i += stride) {
    const uint32_t d = t_dest[i];
    colS[d] = static_cast<uint32_t>(keys[i]);
    labS[d] = t_lab[i];
    if (wS != nullptr) wS[d] = t_w[i];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (extraS.p[j] != nullptr) extraS.p[j][d] = t_extra.p[j][i];
  }
}

// ================================================================== distributed build
// Every rank expands ITS items in the order of their GLOBAL merged ids, so its pairs come out as 16-byte records sorted by
// (row, col) in global ids, ties in stream order; the owners hold ascending row ranges, hence the records of one owner are
// one contiguous range of the sender's buffer: no partition pass, no pack pass.  An owner READS one sorted run per sender
// straight from the senders' buffers (NVLink peer memory), merges the runs in shared-memory tiles of whole row ranges
// (rank of a record = its index in its own run + binary searches in the other runs: senders in rank order = stream
// order, so equal keys stay in the single-device summation order) and WRITES the merged-edge index of every record
// straight into the sender's answer buffer: the transfer in both directions is the merge kernel's own load and store
// stream.
#ifndef PPG_MERGE_BLOCK
#define PPG_MERGE_BLOCK 512
#endif
#ifndef PPG_MERGE_TILE
#define PPG_MERGE_TILE 1024
#endif
constexpr int kMergeBlock = PPG_MERGE_BLOCK;
constexpr int kMergeTile = PPG_MERGE_TILE;      // records a tile is sized for
constexpr int kMergeCap = 2 * PPG_MERGE_TILE;   // records a tile can hold (uniform row ranges: load varies)
constexpr int kMergeRows = 2048;                // rows per tile up to which the merge counts the records of every row
constexpr size_t kMergeSmem = static_cast<size_t>(kMergeCap) * (8 + 4 + 4 + 4 + 2) + (kMergeRows + 1) * 4;
constexpr int kMergePerThread = kMergeCap / kMergeBlock;
constexpr int kMaxRanks = PPG_ROUTE_MAX_RANKS;
constexpr unsigned long long kMergeStatusOverflow = 2ull;

// first slot of every destination: rows are ascending (rows[i * stride]), owners hold ascending row ranges
__global__ void chain_dest_bounds_kernel(const uint32_t* __restrict__ rows, int stride, int64_t n, const int64_t* __restrict__ offsets,
                                         int world, int64_t* __restrict__ dstart, int64_t* __restrict__ counts) {
  __shared__ int64_t s_start[kMaxRanks + 1];
  const int d = threadIdx.x;
  if (d <= world) {
    int64_t lo = 0, hi = n;  // first slot whose row is >= offsets[d]
    if (d == world) lo = n;
    const int64_t target = offsets[d];
    while (lo < hi) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (static_cast<int64_t>(rows[mid * stride]) >= target) hi = mid; else lo = mid + 1;
    }
    if (d == 0) lo = 0;
    s_start[d] = lo;
    dstart[d] = lo;
  }
  __syncthreads();
  if (d < world) counts[d] = s_start[d + 1] - s_start[d];
}

// level 1: slot arrays -> records
__global__ void __launch_bounds__(256)
chain_pack_kernel(const uint32_t* __restrict__ rowS, const uint32_t* __restrict__ colS, const uint32_t* __restrict__ lastS,
                  const float* __restrict__ wS, int64_t n, uint4* __restrict__ rec) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += stride)
    __stcs(rec + p, make_uint4(ld_stream(colS + p), ld_stream(rowS + p), ld_stream(lastS + p),
                               wS != nullptr ? __float_as_uint(ld_stream(wS + p)) : __float_as_uint(1.f)));
}

// heavy rows of a level whose slots are records: the same fix as ppg_chain_heavy_fix, the record travels as a whole
__global__ void __launch_bounds__(256)
heavy_compact_rec_kernel(const uint2* __restrict__ list, const unsigned long long* __restrict__ hoff, int64_t rows, int64_t slots,
                         const uint4* __restrict__ rec, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < slots; c += stride) {
    int64_t lo = 0, hi = rows;
    while (hi - lo > 1) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (hoff[mid] <= static_cast<unsigned long long>(c)) lo = mid; else hi = mid;
    }
    const uint2 row = list[lo];
    const uint32_t slot = row.x + static_cast<uint32_t>(c - static_cast<int64_t>(hoff[lo]));
    keys[c] = (static_cast<unsigned long long>(row.x) << 32) | rec[slot].x;
    vals[c] = slot;
  }
}
__global__ void __launch_bounds__(256)
heavy_gather_rec_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t slots,
                        const uint4* __restrict__ rec, const uint32_t* __restrict__ labS, HeavyExtras extraS,
                        uint4* __restrict__ t_rec, uint32_t* __restrict__ t_lab, HeavyExtras t_extra, uint32_t* __restrict__ t_dest) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < slots; i += stride) {
    const unsigned long long k = keys[i];
    const unsigned long long row_key = k & 0xffffffff00000000ull;
    int64_t lo = -1, hi = i;
    while (hi - lo > 1) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (keys[mid] >= row_key) hi = mid; else lo = mid;
    }
    const uint32_t src = vals[i];
    t_dest[i] = static_cast<uint32_t>(k >> 32) + static_cast<uint32_t>(i - hi);
    t_rec[i] = rec[src];
    t_lab[i] = labS[src];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (extraS.p[j] != nullptr) t_extra.p[j][i] = extraS.p[j][src];
  }
}
__global__ void __launch_bounds__(256)
heavy_write_rec_kernel(int64_t slots, const uint4* __restrict__ t_rec, const uint32_t* __restrict__ t_lab, HeavyExtras t_extra,
                       const uint32_t* __restrict__ t_dest, uint4* __restrict__ rec, uint32_t* __restrict__ labS, HeavyExtras extraS) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < slots; i += stride) {
    const uint32_t d = t_dest[i];
    rec[d] = t_rec[i];
    labS[d] = t_lab[i];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (extraS.p[j] != nullptr) extraS.p[j][d] = t_extra.p[j][i];
  }
}

// The owners' answers (merged-edge index of every record, slot order) -> global ids.  One scan turns them into the row
// structure of the next level (local rows = distinct global ids among the local items) and the info word of every item
// {global id, last node, number of continuations (row pointer after ppg_chain_scan_nodes), -}.
struct UnpackProducer {
  const uint32_t* back;
  const int64_t* dstart;
  const int64_t* edge_offsets;
  int world;
  __device__ unsigned long long gid(int64_t i) const {
    int d = 0;
    for (int j = 1; j < world; ++j) d += i >= dstart[j] ? 1 : 0;
    return static_cast<unsigned long long>(edge_offsets[d]) + back[i];
  }
  __device__ unsigned long long operator()(int64_t i) const { return (i == 0 || gid(i) != gid(i - 1)) ? 1ull : 0ull; }
};
struct UnpackConsumer {
  UnpackProducer ids;
  const uint32_t* labS;
  const uint32_t* lastS;             // last node of every slot, `last_stride` words apart (records: the third word)
  int last_stride;
  const uint32_t* degS;              // continuations of every slot at the next level or nullptr
  uint32_t* rowid;                   // local row of every slot
  uint32_t* run_start;               // [local rows + 1]
  uint32_t* row_value;               // global id of every local row
  uint4* info_item;                  // by label
  int64_t n;
  __device__ void operator()(int64_t i, unsigned long long head, unsigned long long prefix) const {
    const unsigned long long g = ids.gid(i);
    const unsigned long long r = prefix + head - 1;
    rowid[i] = static_cast<uint32_t>(r);
    if (head) {
      run_start[r] = static_cast<uint32_t>(i);
      row_value[r] = static_cast<uint32_t>(g);
    }
    if (i == n - 1) run_start[r + 1] = static_cast<uint32_t>(n);
    uint32_t* out = reinterpret_cast<uint32_t*>(info_item + labS[i]);
    *reinterpret_cast<uint2*>(out) = make_uint2(static_cast<uint32_t>(g), lastS[i * last_stride]);
    if (degS != nullptr) out[2] = degS[i];
  }
};

// ---- owner side
struct MergeRuns {
  const uint4* run[kMaxRanks];   // first record of every sender's run for this owner (peer memory or a local copy)
  uint32_t len[kMaxRanks];
  uint32_t* back[kMaxRanks];     // where the merged-edge index of every record of the run goes
  int world;
};

__global__ void __launch_bounds__(256)
merge_bounds_kernel(MergeRuns runs, long long row_lo, long long rows_per_tile, int64_t n_tiles, uint32_t* __restrict__ bnd) {
  const int64_t total = (n_tiles + 1) * runs.world;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < total; k += stride) {
    const int64_t t = k / runs.world;
    const int s = static_cast<int>(k - t * runs.world);
    int64_t lo = 0, hi = runs.len[s];
    if (t == n_tiles) {
      lo = hi;
    } else {
      const uint4* __restrict__ rec = runs.run[s];
      const long long target = row_lo + t * rows_per_tile;
      while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (static_cast<long long>(rec[mid].y) >= target) hi = mid; else lo = mid + 1;
      }
    }
    bnd[k] = static_cast<uint32_t>(lo);
  }
}

struct MergeArgs {
  MergeRuns runs;
  const uint32_t* bnd;  // [(tiles + 1) * world]: first record of every run in every tile
  long long row_lo, rows_owned, total_nodes, rows_per_tile;
  uint32_t *row_m, *col_m, *last_m;
  float* w_m;
  unsigned long long* tile_state;
  unsigned long long* result;  // [0] merged edges, [1] status
};

__global__ void __launch_bounds__(kMergeBlock)
merge_tile_kernel(MergeArgs a) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(merge_smem);   // (row - first row of the tile) << 32 | col
  float* s_w = reinterpret_cast<float*>(s_key + kMergeCap);
  uint32_t* s_last = reinterpret_cast<uint32_t*>(s_w + kMergeCap);
  uint32_t* s_run = s_last + kMergeCap;
  uint32_t* s_hist = s_run + kMergeCap;                                             // [kMergeRows + 1] records per row of the tile
  uint16_t* s_perm = reinterpret_cast<uint16_t*>(s_hist + kMergeRows + 1);
  uint16_t* s_grp = reinterpret_cast<uint16_t*>(s_run);                             // records grouped by row (dead before s_run is written)
  __shared__ uint32_t s_lo[kMaxRanks], s_off[kMaxRanks + 1];
  __shared__ uint32_t s_warp_cnt[kMergeBlock / 32];
  __shared__ unsigned long long s_base_id;
  __shared__ unsigned long long s_total;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const unsigned lane = lane_id();
  const unsigned tile = blockIdx.x;
  const bool last_tile = tile == gridDim.x - 1;
  const int world = a.runs.world;
  if (tid == 0) {
    uint32_t run = 0;
    for (int s = 0; s < world; ++s) {
      const uint32_t lo = a.bnd[static_cast<size_t>(tile) * world + s];
      const uint32_t hi = a.bnd[static_cast<size_t>(tile + 1) * world + s];
      s_lo[s] = lo;
      s_off[s] = run;
      run += hi - lo;
    }
    s_off[world] = run;
  }
  __syncthreads();
  const int n = static_cast<int>(s_off[world]);
  if (n == 0 || n > kMergeCap) {
    if (warp == 0) {
      if (n > kMergeCap && lane == 0) atomicOr(a.result + 1, kMergeStatusOverflow);
      if (lane == 0) chain_tile_publish(a.tile_state, tile, 0ull, 1, 2);
      const unsigned long long before = chain_tile_prefix(a.tile_state, tile, 0ull, 1, 2);
      if (last_tile && lane == 0) a.result[0] = before;
    }
    return;
  }
  const long long row0 = a.row_lo + static_cast<long long>(tile) * a.rows_per_tile;
  const bool by_rows = a.rows_per_tile <= kMergeRows;   // few enough rows per tile to count the records of every row
  if (by_rows) {
    for (int r = tid; r <= kMergeRows; r += kMergeBlock) s_hist[r] = 0;
    __syncthreads();
  }
  // ---- load the slices, sender after sender (remote senders: NVLink peer loads)
  for (int i = tid; i < n; i += kMergeBlock) {
    int s = 0;
    for (int j = 1; j < world; ++j) s += static_cast<uint32_t>(i) >= s_off[j] ? 1 : 0;
    const uint4 r = a.runs.run[s][s_lo[s] + (i - s_off[s])];
    const long long row = static_cast<long long>(r.y);
    if (row < a.row_lo || row >= a.row_lo + a.rows_owned || static_cast<long long>(r.x) >= a.total_nodes)
      atomicOr(a.result + 1, kChainStatusIdOutOfRange);
    long long local = row - row0;
    if (local < 0 || local >= a.rows_per_tile) local = 0;   // a foreign row (flagged above) must not leave the tile's tables
    s_key[i] = (static_cast<unsigned long long>(local) << 32) | r.x;
    s_w[i] = __uint_as_float(r.w);
    s_last[i] = r.z;
    if (by_rows) atomicAdd(&s_hist[local], 1u);
  }
  __syncthreads();
  if (by_rows) {
    // ---- merge by counting: records per row -> row segments -> every record ranked inside its row by (col, arrival
    // index).  The arrival index orders senders by rank and a sender's records by position, i.e. stream order, so the
    // result is the stable merge of the runs -- at a cost of two shared-memory atomics and one pass over a (short) row
    // per record, where the search-based merge below pays (world - 1) binary searches.
    {
      constexpr int PER = (kMergeRows + kMergeBlock - 1) / kMergeBlock;
      uint32_t v[PER];
      uint32_t sum = 0;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int r = tid * PER + k;
        v[k] = r < kMergeRows ? s_hist[r] : 0u;
        sum += v[k];
      }
      const uint32_t inc = warp_inclusive_sum(sum);
      if (lane == 31) s_warp_cnt[warp] = inc;
      __syncthreads();
      uint32_t before = inc - sum;
#pragma unroll
      for (int w = 0; w < kMergeBlock / 32; ++w)
        if (w < warp) before += s_warp_cnt[w];
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int r = tid * PER + k;
        if (r < kMergeRows) s_hist[r] = before;   // first slot of the row; the scatter below advances it to the row's end
        before += v[k];
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += kMergeBlock) s_grp[atomicAdd(&s_hist[s_key[i] >> 32], 1u)] = static_cast<uint16_t>(i);
    __syncthreads();
    for (int i = tid; i < n; i += kMergeBlock) {
      const unsigned long long key = s_key[i];
      const uint32_t row = static_cast<uint32_t>(key >> 32);
      const int lo = row ? static_cast<int>(s_hist[row - 1]) : 0, hi = static_cast<int>(s_hist[row]);
      int rank = 0;
      for (int j = lo; j < hi; ++j) {
        const int other = s_grp[j];
        const unsigned long long k2 = s_key[other];
        rank += (k2 < key || (k2 == key && other < i)) ? 1 : 0;
      }
      s_perm[lo + rank] = static_cast<uint16_t>(i);
    }
    __syncthreads();   // s_grp (aliased with s_run) is dead from here on
  } else
  // ---- rank of every record in the merge of the runs (stable: earlier senders first among equal keys)
  for (int i = tid; i < n; i += kMergeBlock) {
    int s = 0;
    for (int j = 1; j < world; ++j) s += static_cast<uint32_t>(i) >= s_off[j] ? 1 : 0;
    const unsigned long long key = s_key[i];
    int rank = i - static_cast<int>(s_off[s]);
    for (int q = 0; q < world; ++q) {
      if (q == s) continue;
      int lo = static_cast<int>(s_off[q]), hi = static_cast<int>(s_off[q + 1]);
      const int begin = lo;
      if (q < s) {  // records of earlier senders with key <= this key come first
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_key[mid] <= key) lo = mid + 1; else hi = mid;
        }
      } else {      // of later senders: key < this key
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_key[mid] < key) lo = mid + 1; else hi = mid;
        }
      }
      rank += lo - begin;
    }
    s_perm[rank] = static_cast<uint16_t>(i);
  }
  __syncthreads();
  // ---- run heads in merged order (a tile starts with a new row, so its first record is a head)
  {
    const int p0 = tid * kMergePerThread;
    unsigned long long prev = ~0ull;
    if (p0 > 0 && p0 <= n) prev = s_key[s_perm[p0 - 1]];
    uint32_t run[kMergePerThread];
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < kMergePerThread; ++k) {
      const int p = p0 + k;
      if (p < n) {
        const unsigned long long key = s_key[s_perm[p]];
        cnt += (p == 0 || key != prev) ? 1u : 0u;
        prev = key;
      }
      run[k] = cnt;
    }
    const uint32_t inc = warp_inclusive_sum(cnt);
    if (lane == 31) s_warp_cnt[warp] = inc;
    __syncthreads();
    uint32_t before = inc - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < kMergeBlock / 32; ++w) {
      if (w < warp) before += s_warp_cnt[w];
      total += s_warp_cnt[w];
    }
#pragma unroll
    for (int k = 0; k < kMergePerThread; ++k)
      if (p0 + k < n) s_run[p0 + k] = before + run[k] - 1;
    if (tid == 0) {
      s_total = total;
      chain_tile_publish(a.tile_state, tile, total, 1, 2);
    }
  }
  __syncthreads();
  if (warp == 0) {
    const unsigned long long total = s_total;
    const unsigned long long prefix = chain_tile_prefix(a.tile_state, tile, total, 1, 2);
    if (lane == 0) {
      s_base_id = prefix;
      if (last_tile) a.result[0] = prefix + total;
    }
  }
  __syncthreads();
  const unsigned long long base_id = s_base_id;
  for (int p = tid; p < n; p += kMergeBlock) {
    const int i = s_perm[p];
    const uint32_t run = s_run[p];
    const uint32_t id = static_cast<uint32_t>(base_id + run);
    int s = 0;
    for (int j = 1; j < world; ++j) s += static_cast<uint32_t>(i) >= s_off[j] ? 1 : 0;
    a.runs.back[s][s_lo[s] + (i - s_off[s])] = id;   // remote senders: NVLink peer store
    if (p == 0 || s_run[p - 1] != run) {
      const unsigned long long key = s_key[i];
      float acc = s_w[i];
      for (int q = p + 1; q < n && s_run[q] == run; ++q) acc += s_w[s_perm[q]];   // merged order = stream order
      a.row_m[id] = static_cast<uint32_t>(static_cast<long long>(key >> 32) + row0);
      a.col_m[id] = static_cast<uint32_t>(key);
      a.last_m[id] = s_last[i];
      a.w_m[id] = acc;
    }
  }
}

__global__ void __launch_bounds__(256)
merge_sorted_fill_kernel(const uint32_t* __restrict__ row_m, const uint32_t* __restrict__ col_m, const float* __restrict__ w_m,
                         const uint32_t* __restrict__ last_m, int64_t num_out, int64_t* __restrict__ out_ei,
                         float* __restrict__ out_w, int64_t* __restrict__ out_last) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < num_out; r += stride) {
    st_stream(out_ei + r, static_cast<int64_t>(row_m[r]));
    st_stream(out_ei + num_out + r, static_cast<int64_t>(col_m[r]));
    st_stream(out_w + r, w_m[r]);
    st_stream(out_last + r, static_cast<int64_t>(last_m[r]));
  }
}

static int launch_tiles(ChainTileArgs& a, bool first, cudaStream_t stream) {
  if (a.n_slots == 0) return PPG_OK;
  a.code_partial = 1;
  a.code_inclusive = 2;
  PPG_REQUIRE(a.tile_state != nullptr, PPG_ERR_INVALID, "chain: tile state is required");
  const int64_t tiles = ceil_div(a.n_slots, kChainTile);
  PPG_REQUIRE(tiles < (1ll << 31), PPG_ERR_INVALID, "chain: %lld pairs are too many", (long long)a.n_slots);
  PPG_CUDA_TRY(cudaMemsetAsync(a.tile_state, 0, static_cast<size_t>(tiles) * sizeof(unsigned long long), stream));
  // algorithmic bytes of the launch (what the kernel must read and write once), for the live roofline of bench.py
  const bool dist = !first && a.info != nullptr;
  if (a.w_stride < 1) a.w_stride = 1;
  const long long per_source = first ? 4 + 4 + 8 + (a.wS != nullptr ? 4 : 0) : 8 + 4 + 4 + 4 + (a.wS != nullptr ? 4 : 0);
  const bool next = !first && !dist && a.node_out != nullptr;
  const long long per_slot = dist ? 16 + (a.via != nullptr ? 4 : 0) + 16 + 4 + (a.firstS != nullptr ? 12 : 0)
                                  : (first ? 0 : 8 + (a.via != nullptr ? 4 : 0)) + 12 + (a.wS != nullptr ? 4 : 0) +
                                        (a.idS != nullptr ? 4 : 0) + (a.node_out != nullptr ? (first ? 4 : 8) : 0) + (next ? 8 : 0);
  const long long launch_bytes = per_source * a.n_sources + per_slot * a.n_slots;
  static bool configured_all = false;   // shared memory beyond the 48 KB default (larger tile geometries, NEXT)
  if (!configured_all) {
    PPG_CUDA_TRY(cudaFuncSetAttribute(chain_tile_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kChainTileSmem)));
    PPG_CUDA_TRY(cudaFuncSetAttribute(chain_tile_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kChainTileSmem)));
    PPG_CUDA_TRY(cudaFuncSetAttribute(chain_tile_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kChainTileSmem)));
    configured_all = true;
  }
  profile_pass_begin(stream);
  if (first) {
    chain_tile_kernel<true, false, false><<<static_cast<unsigned>(tiles), kChainBlock, kChainTileSmem, stream>>>(a);
  } else if (dist) {
    chain_tile_kernel<false, true, false><<<static_cast<unsigned>(tiles), kChainBlock, kChainTileSmem, stream>>>(a);
  } else if (next) {
    static bool configured = false;
    if (!configured) {
      PPG_CUDA_TRY(cudaFuncSetAttribute(chain_tile_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kChainTileSmemNext)));
      configured = true;
    }
#ifndef PPG_NEXT_REREAD
    chain_tile_kernel<false, false, true><<<static_cast<unsigned>(tiles), kChainBlock, kChainTileSmemNext, stream>>>(a);
#else
    chain_tile_kernel<false, false, true><<<static_cast<unsigned>(tiles), kChainBlock, kChainTileSmem, stream>>>(a);
#endif
  } else {
    chain_tile_kernel<false, false, false><<<static_cast<unsigned>(tiles), kChainBlock, kChainTileSmem, stream>>>(a);
  }
  profile_pass_end(stream, a.n_slots, static_cast<int>((launch_bytes + a.n_slots / 2) / a.n_slots), PPG_PROFILE_CHAIN_TILES);
  PPG_LAUNCHED();
  return PPG_OK;
}

}  // namespace ppg

using namespace ppg;

extern "C" int ppg_chain_heavy_default(void) { return kChainHeavyMax; }
extern "C" int ppg_chain_tile_slots(void) { return kChainTile; }

// Level 1: rows = first-order nodes, items = events grouped by source node (arrays of ppg_lift_temporal_group).
extern "C" int ppg_chain_first_tiles(const int64_t* edge_index, int64_t m, int64_t N, const uint32_t* ptr1,
                                     const uint32_t* grouped, const uint32_t* sorted_src, const float* weights, int heavy,
                                     uint32_t* rowS, uint32_t* colS, uint32_t* labS, float* wS, uint32_t* idS,
                                     void* node_out, uint32_t* run_start, void* tile_state, void* heavy_list,
                                     int64_t* result, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(m > 0 && N > 0 && m < (1ll << 31) && N < (1ll << 31), PPG_ERR_INVALID, "chain: sizes m=%lld N=%lld outside (0, 2^31)",
              (long long)m, (long long)N);
  PPG_REQUIRE(heavy >= 1 && heavy <= kChainHeavyMax, PPG_ERR_INVALID, "chain: heavy threshold %d outside [1, %d]", heavy, kChainHeavyMax);
  PPG_REQUIRE((weights == nullptr) == (wS == nullptr), PPG_ERR_INVALID, "chain: weights in and out must come together");
  ChainTileArgs a = {};
  a.lblP = grouped;
  a.wP = weights;
  a.run_start = ptr1;
  a.rowid = sorted_src;
  a.dst = edge_index + m;
  a.n_sources = m;
  a.n_rows = N;
  a.n_slots = m;
  a.num_cols = N;
  a.heavy = heavy;
  a.rowS = rowS;
  a.colS = colS;
  a.labS = labS;
  a.wS = wS;
  a.heavy_list = static_cast<uint2*>(heavy_list);
  a.result = reinterpret_cast<unsigned long long*>(result);
  a.idS = idS;
  a.node_out = static_cast<unsigned long long*>(node_out);
  a.run_start_out = run_start;
  a.tile_state = static_cast<unsigned long long*>(tile_state);
  return launch_tiles(a, true, stream);
}

extern "C" size_t ppg_chain_scan_workspace_bytes(int64_t n) { return (scan_state_words(n < 0 ? 0 : n) + 4) * sizeof(unsigned long long); }

// Run heads of the slots in final order -> merged ids; result[kResHeads] = number of merged edges.
extern "C" int ppg_chain_heads(const uint32_t* rowS, const uint32_t* colS, const uint32_t* labS, int64_t n, void* workspace,
                               size_t workspace_bytes, uint32_t* idS, uint32_t* id_item, int id_stride, uint32_t* run_start,
                               int64_t* result, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(n >= 0 && n < (1ll << 31), PPG_ERR_INVALID, "chain: %lld slots outside [0, 2^31)", (long long)n);
  const size_t need = ppg_chain_scan_workspace_bytes(n);
  PPG_REQUIRE(workspace_bytes >= need, PPG_ERR_WORKSPACE, "chain_heads: workspace %zu < %zu bytes", workspace_bytes, need);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, stream));
  if (n == 0) {
    PPG_CUDA_TRY(cudaMemsetAsync(result + kResHeads, 0, sizeof(int64_t), stream));
    PPG_CUDA_TRY(cudaMemsetAsync(run_start, 0, sizeof(uint32_t), stream));
    return PPG_OK;
  }
  return launch_scan(ChainHeadProducer{rowS, colS}, ChainHeadConsumer{labS, idS, id_item, id_stride < 1 ? 1 : id_stride, run_start, n}, n,
                     static_cast<unsigned long long*>(workspace), reinterpret_cast<unsigned long long*>(result) + kResHeads, stream);
}

extern "C" int ppg_chain_fill(const uint32_t* rowS, const uint32_t* colS, const float* wS, const uint32_t* run_start,
                              int64_t num_out, int64_t* out_edge_index, float* out_weights, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_out == 0) return PPG_OK;
  chain_fill_kernel<<<grid_for(num_out, 256), 256, 0, stream>>>(rowS, colS, wS, run_start, num_out, out_edge_index, out_weights);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_chain_widen(const uint32_t* in, int in_stride, int64_t n, int64_t* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) return PPG_OK;
  chain_widen_kernel<<<grid_for(n, 256 * 4), 256, 0, stream>>>(in, in_stride < 1 ? 1 : in_stride, n, out);
  PPG_LAUNCHED();
  return PPG_OK;
}

// Count pass in merged order: slot offsets of the expansion, per-source data in that order, tile boundaries.
extern "C" int ppg_chain_count_sorted(const uint32_t* P, int64_t n, const uint32_t* first, const void* ptr_next,
                                      const float* w_item, int64_t limit, const uint32_t* rowid, const uint32_t* run_start,
                                      void* workspace, size_t workspace_bytes, void* offP, uint32_t* firstP, uint32_t* lblP,
                                      float* wP, uint32_t* srcbound, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(n > 0 && n < (1ll << 31), PPG_ERR_INVALID, "chain: %lld sources outside (0, 2^31)", (long long)n);
  const size_t need = ppg_chain_scan_workspace_bytes(n);
  PPG_REQUIRE(workspace_bytes >= need, PPG_ERR_WORKSPACE, "chain_count_sorted: workspace %zu < %zu bytes", workspace_bytes, need);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, stream));
  return launch_scan(ChainSortedProducer{P, first, static_cast<const unsigned long long*>(ptr_next), w_item, limit, firstP, lblP, wP},
                     ChainSortedConsumer{static_cast<unsigned long long*>(offP), reinterpret_cast<uint2*>(srcbound), rowid, run_start, n}, n,
                     static_cast<unsigned long long*>(workspace), nullptr, stream);
}

// Tiles of level k >= 2: `via` maps a continuation position to its item (temporal level: the events grouped by source).
extern "C" int ppg_chain_tiles(int64_t n_sources, int64_t n_rows, int64_t n_slots, const void* offP, const uint32_t* firstP,
                               const uint32_t* lblP, const float* wP, const uint32_t* run_start, const uint32_t* rowid,
                               const void* node_prev, const uint32_t* via, const uint32_t* srcbound, int heavy, uint32_t* rowS,
                               uint32_t* colS, uint32_t* labS, float* wS, uint32_t* idS, void* node_out, uint32_t* firstS,
                               uint32_t* degS, uint32_t* run_start_out, void* tile_state, void* heavy_list, int64_t* result,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(n_sources > 0 && n_sources < (1ll << 31) && n_slots >= 0 && n_slots < (1ll << 31) && n_rows > 0, PPG_ERR_INVALID,
              "chain: sizes outside [0, 2^31)");
  PPG_REQUIRE(heavy >= 1 && heavy <= kChainHeavyMax, PPG_ERR_INVALID, "chain: heavy threshold %d outside [1, %d]", heavy, kChainHeavyMax);
  PPG_REQUIRE((wP == nullptr) == (wS == nullptr), PPG_ERR_INVALID, "chain: weights in and out must come together");
  ChainTileArgs a = {};
  a.offP = static_cast<const unsigned long long*>(offP);
  a.firstP = firstP;
  a.lblP = lblP;
  a.wP = wP;
  a.run_start = run_start;
  a.rowid = rowid;
  PPG_REQUIRE(node_prev != nullptr && (node_out == nullptr) == (firstS == nullptr) && (firstS == nullptr) == (degS == nullptr),
              PPG_ERR_INVALID, "chain: node words of the previous level are required; node_out, firstS and degS come together");
  a.node_prev = static_cast<const unsigned long long*>(node_prev);
  a.via = via;
  a.srcbound = srcbound;
  a.n_sources = n_sources;
  a.n_rows = n_rows;
  a.n_slots = n_slots;
  a.heavy = heavy;
  a.rowS = rowS;
  a.colS = colS;
  a.labS = labS;
  a.wS = wS;
  a.node_out = static_cast<unsigned long long*>(node_out);
  a.firstS = firstS;
  a.degS = degS;
  a.heavy_list = static_cast<uint2*>(heavy_list);
  a.result = reinterpret_cast<unsigned long long*>(result);
  a.idS = idS;
  a.run_start_out = run_start_out;
  a.tile_state = static_cast<unsigned long long*>(tile_state);
  return launch_tiles(a, false, stream);
}

extern "C" size_t ppg_chain_heavy_workspace_bytes(int64_t heavy_slots, int64_t heavy_rows, int64_t n_slots) {
  Workspace ws(nullptr, 0);
  HeavyLayout L(ws, heavy_slots, heavy_rows, n_slots);
  return ws.used + 256;
}

// Rows the tiles left in generation order: one stable radix sort over their pairs only, written back in place.
extern "C" int ppg_chain_heavy_fix(const void* heavy_list, int64_t heavy_rows, int64_t heavy_slots, int64_t n_slots,
                                   uint32_t* colS, uint32_t* labS, float* wS, uint32_t* extra0, uint32_t* extra1,
                                   uint32_t* extra2, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (heavy_rows == 0 || heavy_slots == 0) return PPG_OK;
  Workspace ws(workspace, workspace_bytes);
  HeavyLayout L(ws, heavy_slots, heavy_rows, n_slots);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "chain_heavy_fix: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));
  const uint2* list = static_cast<const uint2*>(heavy_list);
  PPG_TRY(launch_scan(HeavyLenProducer{list}, ChainOffsetConsumer{L.hoff, heavy_rows}, heavy_rows, L.scan_ws, nullptr, stream));
  heavy_compact_kernel<<<grid_for(heavy_slots, 256 * 4), 256, 0, stream>>>(list, L.hoff, heavy_rows, heavy_slots, colS, L.keys_a, L.vals_a);
  PPG_LAUNCHED();
  int in_b = 0;
  PPG_TRY(radix_sort_pairs<unsigned long long>(L.keys_a, L.keys_b, L.vals_a, L.vals_b, true, false, heavy_slots, L.bits, L.sort_ws,
                                               &in_b, stream));
  const unsigned long long* keys = in_b ? L.keys_b : L.keys_a;
  const uint32_t* vals = in_b ? L.vals_b : L.vals_a;
  const HeavyExtras extras = {{extra0, extra1, extra2}};
  const HeavyExtras temps = {{L.t_extra[0], L.t_extra[1], L.t_extra[2]}};
  heavy_gather_kernel<<<grid_for(heavy_slots, 256 * 4), 256, 0, stream>>>(keys, vals, heavy_slots, labS, wS, extras, L.t_lab, L.t_w, temps, L.t_dest);
  PPG_LAUNCHED();
  heavy_write_kernel<<<grid_for(heavy_slots, 256 * 4), 256, 0, stream>>>(keys, heavy_slots, L.t_lab, L.t_w, temps, L.t_dest, colS, labS, wS, extras);
  PPG_LAUNCHED();
  return PPG_OK;
}

// =================================================================== distributed build (see the kernels above)
// Tiles of a level >= 2 on a rank of a distributed build: `info` by item = {global id, last node, row pointer of this
// level, -}; rowid / run_start / row_value describe the LOCAL rows (distinct global ids among the local items, from
// ppg_chain_unpack); the slots leave as records (rec), with labS and -- if a level follows -- firstS / degS.
// wP: weight of every source, w_stride words apart (the weight word of the previous level's records).
extern "C" int ppg_chain_tiles_dist(int64_t n_sources, int64_t n_slots, const void* offP, const uint32_t* firstP,
                                    const uint32_t* lblP, const float* wP, int w_stride, const uint32_t* run_start,
                                    const uint32_t* rowid, const uint32_t* row_value, const void* info, const uint32_t* via,
                                    const uint32_t* srcbound, int heavy, void* rec, uint32_t* labS, uint32_t* firstS, uint32_t* degS,
                                    void* info_out, void* tile_state, void* heavy_list, int64_t* result, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(n_sources > 0 && n_sources < (1ll << 31) && n_slots >= 0 && n_slots < (1ll << 31), PPG_ERR_INVALID,
              "chain: sizes outside [0, 2^31)");
  PPG_REQUIRE(heavy >= 1 && heavy <= kChainHeavyMax, PPG_ERR_INVALID, "chain: heavy threshold %d outside [1, %d]", heavy, kChainHeavyMax);
  PPG_REQUIRE(wP != nullptr && info != nullptr && row_value != nullptr && rec != nullptr && (firstS == nullptr) == (degS == nullptr) &&
                  (firstS == nullptr) == (info_out == nullptr),
              PPG_ERR_INVALID, "chain (distributed): weights, info words, row values and the record buffer are required");
  ChainTileArgs a = {};
  a.offP = static_cast<const unsigned long long*>(offP);
  a.firstP = firstP;
  a.lblP = lblP;
  a.wP = wP;
  a.w_stride = w_stride;
  a.run_start = run_start;
  a.rowid = rowid;
  a.row_value = row_value;
  a.info = static_cast<const uint4*>(info);
  a.via = via;
  a.srcbound = srcbound;
  a.n_sources = n_sources;
  a.n_rows = 1;
  a.n_slots = n_slots;
  a.heavy = heavy;
  a.rec = static_cast<uint4*>(rec);
  a.labS = labS;
  a.firstS = firstS;
  a.degS = degS;
  a.info_out = static_cast<uint4*>(info_out);
  a.heavy_list = static_cast<uint2*>(heavy_list);
  a.result = reinterpret_cast<unsigned long long*>(result);
  a.tile_state = static_cast<unsigned long long*>(tile_state);
  return launch_tiles(a, false, stream);
}

// First slot of every destination rank (dstart [world + 1], counts [world]; device int64) among slots whose rows
// (rows[i * stride]: stride 1 for a slot array, 4 from the row word of the records) ascend.
extern "C" int ppg_chain_dest_bounds(const uint32_t* rows, int stride, int64_t n_slots, const int64_t* offsets, int world,
                                     int64_t* dstart, int64_t* counts, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(world >= 1 && world <= kMaxRanks, PPG_ERR_INVALID, "chain: %d ranks outside [1, %d]", world, kMaxRanks);
  chain_dest_bounds_kernel<<<1, 32, 0, stream>>>(rows, stride < 1 ? 1 : stride, n_slots, offsets, world, dstart, counts);
  PPG_LAUNCHED();
  return PPG_OK;
}

// Level 1: slot arrays -> records {col, row, last node, weight} in slot order.
extern "C" int ppg_chain_pack(const uint32_t* rowS, const uint32_t* colS, const uint32_t* lastS, const float* wS, int64_t n_slots,
                              void* rec, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_slots == 0) return PPG_OK;
  chain_pack_kernel<<<grid_for(n_slots, 256 * 4), 256, 0, stream>>>(rowS, colS, lastS, wS, n_slots, static_cast<uint4*>(rec));
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" size_t ppg_chain_heavy_records_workspace_bytes(int64_t heavy_slots, int64_t heavy_rows, int64_t n_slots) {
  Workspace ws(nullptr, 0);
  HeavyLayout L(ws, heavy_slots, heavy_rows, n_slots);
  ws.take<uint4>(static_cast<size_t>(heavy_slots));
  return ws.used + 256;
}

// ppg_chain_heavy_fix for a level whose slots are records (distributed build, levels >= 2).
extern "C" int ppg_chain_heavy_fix_records(const void* heavy_list, int64_t heavy_rows, int64_t heavy_slots, int64_t n_slots,
                                           void* rec_, uint32_t* labS, uint32_t* firstS, uint32_t* degS, void* workspace,
                                           size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (heavy_rows == 0 || heavy_slots == 0) return PPG_OK;
  Workspace ws(workspace, workspace_bytes);
  HeavyLayout L(ws, heavy_slots, heavy_rows, n_slots);
  uint4* t_rec = ws.take<uint4>(static_cast<size_t>(heavy_slots));
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "chain_heavy_fix_records: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));
  uint4* rec = static_cast<uint4*>(rec_);
  const uint2* list = static_cast<const uint2*>(heavy_list);
  PPG_TRY(launch_scan(HeavyLenProducer{list}, ChainOffsetConsumer{L.hoff, heavy_rows}, heavy_rows, L.scan_ws, nullptr, stream));
  heavy_compact_rec_kernel<<<grid_for(heavy_slots, 256 * 4), 256, 0, stream>>>(list, L.hoff, heavy_rows, heavy_slots, rec, L.keys_a, L.vals_a);
  PPG_LAUNCHED();
  int in_b = 0;
  PPG_TRY(radix_sort_pairs<unsigned long long>(L.keys_a, L.keys_b, L.vals_a, L.vals_b, true, false, heavy_slots, L.bits, L.sort_ws,
                                               &in_b, stream));
  const unsigned long long* keys = in_b ? L.keys_b : L.keys_a;
  const uint32_t* vals = in_b ? L.vals_b : L.vals_a;
  const HeavyExtras extras = {{firstS, degS, nullptr}};
  const HeavyExtras temps = {{L.t_extra[0], L.t_extra[1], L.t_extra[2]}};
  heavy_gather_rec_kernel<<<grid_for(heavy_slots, 256 * 4), 256, 0, stream>>>(keys, vals, heavy_slots, rec, labS, extras, t_rec, L.t_lab,
                                                                             temps, L.t_dest);
  PPG_LAUNCHED();
  heavy_write_rec_kernel<<<grid_for(heavy_slots, 256 * 4), 256, 0, stream>>>(heavy_slots, t_rec, L.t_lab, temps, L.t_dest, rec, labS, extras);
  PPG_LAUNCHED();
  return PPG_OK;
}

// The owners' answers (back [n_slots], slot order) -> local row structure of the next level + info word of every item
// (words 0, 1: global id, last node; word 2: degS if given).  lastS: last node of every slot, last_stride words apart.
// result[0] = number of local rows.
extern "C" int ppg_chain_unpack(const uint32_t* back, int64_t n_slots, const int64_t* dstart, const int64_t* edge_offsets, int world,
                                const uint32_t* labS, const uint32_t* lastS, int last_stride, const uint32_t* degS, void* workspace,
                                size_t workspace_bytes, uint32_t* rowid, uint32_t* run_start, uint32_t* row_value, void* info_item,
                                int64_t* result, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(world >= 1 && world <= kMaxRanks && n_slots >= 0 && n_slots < (1ll << 31), PPG_ERR_INVALID, "chain_unpack: bad sizes");
  const size_t need = ppg_chain_scan_workspace_bytes(n_slots);
  PPG_REQUIRE(workspace_bytes >= need, PPG_ERR_WORKSPACE, "chain_unpack: workspace %zu < %zu bytes", workspace_bytes, need);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, stream));
  if (n_slots == 0) {
    PPG_CUDA_TRY(cudaMemsetAsync(result, 0, sizeof(int64_t), stream));
    PPG_CUDA_TRY(cudaMemsetAsync(run_start, 0, sizeof(uint32_t), stream));
    return PPG_OK;
  }
  UnpackProducer ids{back, dstart, edge_offsets, world};
  return launch_scan(ids, UnpackConsumer{ids, labS, lastS, last_stride < 1 ? 1 : last_stride, degS, rowid, run_start, row_value,
                                         static_cast<uint4*>(info_item), n_slots},
                     n_slots, static_cast<unsigned long long*>(workspace), reinterpret_cast<unsigned long long*>(result), stream);
}

extern "C" int64_t ppg_merge_sorted_tiles(int64_t num_records) { return num_records > 0 ? ceil_div(num_records, kMergeTile) : 0; }

// Owner side: `world` runs of records, run s = h_runs[s][0 .. h_run_len[s]) sorted by (row, col) -- device addresses, peer
// memory of the senders or a local copy -- -> merged edges in compact arrays (row_m, col_m, w_m, last_m: capacity = total
// records) and, for every record, the index of its merged edge stored at h_back[s][position in the run] (the sender's
// answer buffer, or a local one).  result[0] = merged edges, result[1] = status (1: id out of range, 2: a row range did
// not fit a tile -- merge these records with ppg_merge_records_sort instead).  bounds: u32 [(tiles + 1) * world],
// tile_state: u64 [tiles] with tiles = ppg_merge_sorted_tiles(total records).
extern "C" int ppg_merge_sorted(void* const* h_runs, const int64_t* h_run_len, void* const* h_back, int world, int64_t row_lo,
                                int64_t rows_owned, int64_t total_nodes, uint32_t* bounds, void* tile_state, uint32_t* row_m,
                                uint32_t* col_m, float* w_m, uint32_t* last_m, int64_t* result, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(world >= 1 && world <= kMaxRanks, PPG_ERR_INVALID, "merge_sorted: %d ranks outside [1, %d]", world, kMaxRanks);
  MergeArgs a = {};
  int64_t num_records = 0;
  for (int s = 0; s < world; ++s) {
    PPG_REQUIRE(h_run_len[s] >= 0 && h_run_len[s] < (1ll << 31), PPG_ERR_INVALID, "merge_sorted: bad run length");
    a.runs.run[s] = static_cast<const uint4*>(h_runs[s]);
    a.runs.len[s] = static_cast<uint32_t>(h_run_len[s]);
    a.runs.back[s] = static_cast<uint32_t*>(h_back[s]);
    num_records += h_run_len[s];
  }
  a.runs.world = world;
  PPG_REQUIRE(num_records < (1ll << 31) && total_nodes <= (1ll << 32), PPG_ERR_INVALID, "merge_sorted: bad sizes");
  PPG_CUDA_TRY(cudaMemsetAsync(result, 0, 2 * sizeof(int64_t), stream));
  if (num_records == 0) return PPG_OK;
  const int64_t tiles = ppg_merge_sorted_tiles(num_records);
  const int64_t rows_per_tile = rows_owned > 0 ? ceil_div(rows_owned, tiles) : 1;
  PPG_CUDA_TRY(cudaMemsetAsync(tile_state, 0, static_cast<size_t>(tiles) * sizeof(unsigned long long), stream));
  merge_bounds_kernel<<<grid_for((tiles + 1) * world, 256), 256, 0, stream>>>(a.runs, row_lo, rows_per_tile, tiles, bounds);
  PPG_LAUNCHED();
  a.bnd = bounds;
  a.row_lo = row_lo;
  a.rows_owned = rows_owned;
  a.total_nodes = total_nodes;
  a.rows_per_tile = rows_per_tile;
  a.row_m = row_m;
  a.col_m = col_m;
  a.w_m = w_m;
  a.last_m = last_m;
  a.tile_state = static_cast<unsigned long long*>(tile_state);
  a.result = reinterpret_cast<unsigned long long*>(result);
  profile_pass_begin(stream);
  static bool configured = false;
  if (!configured) {
    PPG_CUDA_TRY(cudaFuncSetAttribute(merge_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMergeSmem)));
    configured = true;
  }
  merge_tile_kernel<<<static_cast<unsigned>(tiles), kMergeBlock, kMergeSmem, stream>>>(a);
  profile_pass_end(stream, num_records, 16 + 4 + 16, PPG_PROFILE_MERGE_TILES);   // record in, merged index out, merged edge out (upper bound)
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_merge_sorted_fill(const uint32_t* row_m, const uint32_t* col_m, const float* w_m, const uint32_t* last_m,
                                     int64_t num_out, int64_t* out_edge_index, float* out_weights, int64_t* out_last, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_out == 0) return PPG_OK;
  merge_sorted_fill_kernel<<<grid_for(num_out, 256 * 4), 256, 0, stream>>>(row_m, col_m, w_m, last_m, num_out, out_edge_index, out_weights,
                                                                          out_last);
  PPG_LAUNCHED();
  return PPG_OK;
}

// Level 1: ptr word of the events' node words <- row pointer of the event graph (u64 [m + 1] of ppg_lift_temporal_views).
extern "C" int ppg_chain_node_ptr(const void* off, int64_t num_items, void* node, int stride, int word, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  chain_node_ptr_kernel<<<grid_for(num_items + 1, 256 * 4), 256, 0, stream>>>(static_cast<const unsigned long long*>(off), num_items + 1,
                                                                              static_cast<uint32_t*>(node) + word, stride);
  PPG_LAUNCHED();
  return PPG_OK;
}

// Label-order scan of the continuation counts in the node words (in place: count -> row pointer of the next level;
// node[num_items] receives the total, which is also written to *total).
extern "C" int ppg_chain_scan_nodes(void* node, int stride, int word, int64_t num_items, void* workspace, size_t workspace_bytes,
                                    int64_t* total, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(num_items >= 0 && num_items < (1ll << 31), PPG_ERR_INVALID, "chain: %lld items outside [0, 2^31)", (long long)num_items);
  const size_t need = ppg_chain_scan_workspace_bytes(num_items);
  PPG_REQUIRE(workspace_bytes >= need, PPG_ERR_WORKSPACE, "chain_scan_nodes: workspace %zu < %zu bytes", workspace_bytes, need);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, stream));
  if (num_items == 0) {
    PPG_CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(int64_t), stream));
    PPG_CUDA_TRY(cudaMemsetAsync(static_cast<uint32_t*>(node) + word, 0, sizeof(uint32_t), stream));
    return PPG_OK;
  }
  uint32_t* words = static_cast<uint32_t*>(node) + word;
  return launch_scan(NodeDegProducer{words, stride}, NodePtrConsumer{words, stride, num_items}, num_items,
                     static_cast<unsigned long long*>(workspace), reinterpret_cast<unsigned long long*>(total), stream);
}

// Count pass of levels >= 3 in merged order (see ChainSortedProducer2): degS of the previous level's tiles, limit, node
// words after ppg_chain_scan_nodes -> offP u64 [n + 1], lblP [n], srcbound.
extern "C" int ppg_chain_count_sorted_next(const uint32_t* P, int64_t n, const uint32_t* degS, const void* node, int stride,
                                           int word, int64_t limit,
                                           const uint32_t* rowid, const uint32_t* run_start, void* workspace, size_t workspace_bytes,
                                           void* offP, uint32_t* lblP, uint32_t* srcbound, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(n > 0 && n < (1ll << 31), PPG_ERR_INVALID, "chain: %lld sources outside (0, 2^31)", (long long)n);
  const size_t need = ppg_chain_scan_workspace_bytes(n);
  PPG_REQUIRE(workspace_bytes >= need, PPG_ERR_WORKSPACE, "chain_count_sorted_next: workspace %zu < %zu bytes", workspace_bytes, need);
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, stream));
  return launch_scan(ChainSortedProducer2{P, degS, static_cast<const uint32_t*>(node) + word, stride, limit, lblP},
                     ChainSortedConsumer{static_cast<unsigned long long*>(offP), reinterpret_cast<uint2*>(srcbound), rowid, run_start, n}, n,
                     static_cast<unsigned long long*>(workspace), nullptr, stream);
}
