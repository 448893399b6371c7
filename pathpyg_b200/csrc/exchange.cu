// e: cross-partition exchange of lifted edges.  The reference is single-device: nothing here has a
// counterpart in /root/reference; the contract is that the distributed layers, gathered, are bit-identical to the
// single-device ones.
//
// A rank holds the line graph L_k of its (own + ghost) slice of the stream.  Every edge (p -> q) of L_k becomes the
// 16-byte record {id(q), id(p), last node of q, weight} and travels to the rank that owns row id(p) of layer k:
//
//   route_count  : per chunk of edges, how many go to each rank                      (read 8 B + one 8 B gather / edge)
//   route_scan   : chunk x rank exclusive scan -> first slot of every (chunk, rank) in the send buffer
//   route_pack   : STABLE partition by owner: records of one destination keep their edge order, and chunks are in
//                  edge order, so a destination receives from rank r exactly r's edges in stream order.  The merge
//                  at the owner is a stable sort, hence weights are summed in global stream order -- the summation
//                  order of the single-device path -- whatever the weights are.
//   merge_*      : owner side: key = (row - first owned row, col) in as few bits as the owned range needs,
//                  onesweep radix sort, run heads -> merged edges + the merged-edge index of every record
//   route_unpack : sender side: the returned merged-edge index (+ the owner's global offset) is the id of the
//                  next level's line-graph node; stored with the node's last first-order node as one 8-byte word
//                  so that the next level's pack needs ONE random 8-byte gather per edge end point.
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace ppg {

constexpr int kRouteBlock = 256;
constexpr int kRouteItems = 4;
constexpr int kRouteTile = kRouteBlock * kRouteItems;
constexpr int kRouteMaxRanks = PPG_ROUTE_MAX_RANKS;
constexpr unsigned kStatusIdOutOfRange = 1u;

struct ResultWords {
  unsigned long long total;
  unsigned long long status;
};

struct RouteGeometry {
  int chunks;
  int64_t chunk_edges;  // multiple of kRouteTile
};
inline RouteGeometry route_geometry(int64_t E) {
  const int64_t tiles = E > 0 ? ceil_div(E, kRouteTile) : 1;
  int64_t chunks = tiles < kNumSMsB200 * 4 ? tiles : kNumSMsB200 * 4;
  const int64_t per_chunk = ceil_div(tiles, chunks);
  chunks = ceil_div(tiles, per_chunk);
  return {static_cast<int>(chunks), per_chunk * kRouteTile};
}

struct RouteLayout {
  ResultWords* result;
  uint32_t* counts;    // [chunks][kRouteMaxRanks]
  uint32_t* base;      // [chunks][kRouteMaxRanks] first slot of the chunk's records inside the destination's segment
  long long* segment;  // [kRouteMaxRanks + 1] first slot of every destination's segment in the send buffer
  size_t zero_bytes;
  RouteLayout(Workspace& ws, int64_t E) {
    const RouteGeometry g = route_geometry(E);
    result = ws.take<ResultWords>(1);
    zero_bytes = ws.used;
    counts = ws.take<uint32_t>(static_cast<size_t>(g.chunks) * kRouteMaxRanks);
    base = ws.take<uint32_t>(static_cast<size_t>(g.chunks) * kRouteMaxRanks);
    segment = ws.take<long long>(kRouteMaxRanks + 1);
  }
};

struct RouteInput {
  const int64_t* li0;                   // source line node of every edge (ascending)
  const int64_t* li1;                   // target line node
  int64_t E;
  const unsigned long long* node_info;  // [line nodes] id << 32 | last node; nullptr: id = the node itself, last = li1
  const float* w;                       // nullptr: unit weights
  int64_t own_prefix;                   // edges whose source line node (first level: whose position) is >= own_prefix travel with weight 0
  const int64_t* offsets;               // device [world + 1]: first row owned by every rank
  int world;
  int64_t chunk_edges;
};

__device__ __forceinline__ int route_owner(unsigned long long id, const long long* s_off, int world) {
  int d = 0;
  for (int j = 1; j < world; ++j) d += static_cast<long long>(id) >= s_off[j] ? 1 : 0;
  return d;
}

__global__ void __launch_bounds__(kRouteBlock)
route_count_kernel(RouteInput in, uint32_t* __restrict__ counts) {
  __shared__ long long s_off[kRouteMaxRanks + 1];
  __shared__ unsigned s_cnt[kRouteMaxRanks];
  const int tid = threadIdx.x;
  if (tid <= in.world) s_off[tid] = in.offsets[tid];
  if (tid < kRouteMaxRanks) s_cnt[tid] = 0;
  __syncthreads();
  const int64_t begin = static_cast<int64_t>(blockIdx.x) * in.chunk_edges;
  const int64_t end = begin + in.chunk_edges < in.E ? begin + in.chunk_edges : in.E;
  for (int64_t b = begin; b < end; b += kRouteBlock) {
    const int64_t e = b + tid;
    const bool valid = e < end;
    unsigned d = 0xffffffffu;
    if (valid) {
      const int64_t s = ld_stream(in.li0 + e);
      const unsigned long long id = in.node_info != nullptr ? in.node_info[s] >> 32 : static_cast<unsigned long long>(s);
      d = static_cast<unsigned>(route_owner(id, s_off, in.world));
    }
    const unsigned peers = __match_any_sync(kFullMask, d);
    if (valid && lane_id() == static_cast<unsigned>(__ffs(peers) - 1)) atomicAdd(&s_cnt[d], static_cast<unsigned>(__popc(peers)));
  }
  __syncthreads();
  if (tid < kRouteMaxRanks) counts[static_cast<size_t>(blockIdx.x) * kRouteMaxRanks + tid] = s_cnt[tid];
}

// warp d scans the chunk counts of destination d; thread 0 lays the destination segments out back to back
__global__ void __launch_bounds__(32 * kRouteMaxRanks)
route_scan_kernel(const uint32_t* __restrict__ counts, int chunks, int world, uint32_t* __restrict__ base,
                  long long* __restrict__ segment, int64_t* __restrict__ totals_out) {
  __shared__ long long s_total[kRouteMaxRanks];
  const int d = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  unsigned carry = 0;
  for (int b = 0; b < chunks; b += 32) {
    const int c = b + static_cast<int>(lane);
    const unsigned v = c < chunks ? counts[static_cast<size_t>(c) * kRouteMaxRanks + d] : 0u;
    const unsigned inc = warp_inclusive_sum(v);
    if (c < chunks) base[static_cast<size_t>(c) * kRouteMaxRanks + d] = carry + inc - v;
    carry += __shfl_sync(kFullMask, inc, 31);
  }
  if (lane == 0) s_total[d] = carry;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int j = 0; j < kRouteMaxRanks; ++j) {
      segment[j] = run;
      if (j < world) totals_out[j] = s_total[j];
      run += s_total[j];
    }
    segment[kRouteMaxRanks] = run;
  }
}

// Where the records of every destination go.  local != nullptr: one send buffer on this device, destination d at
// segment[d] (the collective moves it).  Otherwise peer[d] is rank d's receive buffer mapped into this process (NVLink
// peer memory), already advanced to the first slot reserved for this sender: the partition kernel stores straight
// into the owners' memory, so the transfer IS the kernel's write stream and needs no collective and no staging copy.
struct RouteTargets {
  uint4* local;
  uint4* peer[kRouteMaxRanks];
};

__global__ void __launch_bounds__(kRouteBlock)
route_pack_kernel(RouteInput in, const uint32_t* __restrict__ base, const long long* __restrict__ segment,
                  RouteTargets targets, uint32_t* __restrict__ slot, uint32_t* __restrict__ last_next) {
  constexpr int NW = kRouteBlock / 32;
  __shared__ long long s_off[kRouteMaxRanks + 1];
  __shared__ unsigned s_cursor[kRouteMaxRanks];
  __shared__ unsigned s_wcount[NW][kRouteMaxRanks];
  __shared__ unsigned s_wbase[NW][kRouteMaxRanks];
  __shared__ uint4* s_target[kRouteMaxRanks];   // s_target[d] + record index = where the record goes
  // the tile's records grouped by destination: they leave as contiguous runs (coalesced 512-byte warp stores; over
  // NVLink that is the difference between 64-byte fragments and full packets)
  __shared__ __align__(16) uint4 s_rec[kRouteTile];
  __shared__ unsigned s_tfirst[kRouteMaxRanks];      // record index (in the destination's stream) of the tile's first record
  __shared__ unsigned s_dstart[kRouteMaxRanks + 1];  // first slot of every destination in s_rec
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const unsigned lane = lane_id();
  if (tid <= in.world) s_off[tid] = in.offsets[tid];
  if (tid < kRouteMaxRanks) {
    s_cursor[tid] = base[static_cast<size_t>(blockIdx.x) * kRouteMaxRanks + tid] + static_cast<unsigned>(segment[tid]);
    s_target[tid] = targets.local != nullptr ? targets.local : targets.peer[tid] - segment[tid];
  }
  const int64_t begin = static_cast<int64_t>(blockIdx.x) * in.chunk_edges;
  const int64_t end = begin + in.chunk_edges < in.E ? begin + in.chunk_edges : in.E;
  for (int64_t tile = begin; tile < end; tile += kRouteTile) {
    if (tid < NW * kRouteMaxRanks) (&s_wcount[0][0])[tid] = 0;
    __syncthreads();  // also: s_off / s_cursor are set, and the previous tile has read its s_wbase
    uint4 rec[kRouteItems] = {};
    unsigned dest[kRouteItems], rank[kRouteItems];
    int64_t pos[kRouteItems];
    unsigned long long info_s[kRouteItems], info_t[kRouteItems];
    // all gathers of the tile's items are issued before any is consumed
#pragma unroll
    for (int i = 0; i < kRouteItems; ++i) {
      pos[i] = tile + warp * (32 * kRouteItems) + i * 32 + lane;
      info_s[i] = info_t[i] = 0;
      if (pos[i] < end) {
        const int64_t s = ld_stream(in.li0 + pos[i]);
        const int64_t t = ld_stream(in.li1 + pos[i]);
        if (in.node_info != nullptr) {
          info_s[i] = in.node_info[s];
          info_t[i] = in.node_info[t];
        } else {  // first level: the line nodes are the first-order nodes themselves
          info_s[i] = static_cast<unsigned long long>(s) << 32;
          info_t[i] = (static_cast<unsigned long long>(t) << 32) | static_cast<unsigned long long>(t);
        }
        float w = in.w != nullptr ? ld_stream(in.w + pos[i]) : 1.f;
        // paths that start with a ghost event only collect their ids (first level: the edges ARE the events)
        if ((in.node_info != nullptr ? s : pos[i]) >= in.own_prefix) w = 0.f;
        rec[i].w = __float_as_uint(w);
      }
    }
#pragma unroll
    for (int i = 0; i < kRouteItems; ++i) {
      const bool valid = pos[i] < end;
      rec[i].x = static_cast<uint32_t>(info_t[i] >> 32);
      rec[i].y = static_cast<uint32_t>(info_s[i] >> 32);
      rec[i].z = static_cast<uint32_t>(info_t[i]);
      dest[i] = valid ? static_cast<unsigned>(route_owner(info_s[i] >> 32, s_off, in.world)) : 0xffffffffu;
      const unsigned peers = __match_any_sync(kFullMask, dest[i]);
      const int leader = __ffs(peers) - 1;
      unsigned before = 0;
      if (valid && static_cast<int>(lane) == leader) {
        before = s_wcount[warp][dest[i]];
        s_wcount[warp][dest[i]] = before + static_cast<unsigned>(__popc(peers));
      }
      before = __shfl_sync(kFullMask, before, leader);
      rank[i] = before + static_cast<unsigned>(__popc(peers & lanemask_lt()));
      __syncwarp();
    }
    __syncthreads();
    if (tid < kRouteMaxRanks) {
      unsigned run = s_cursor[tid];
      s_tfirst[tid] = run;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        s_wbase[w][tid] = run;
        run += tid < in.world ? s_wcount[w][tid] : 0u;
      }
      s_cursor[tid] = run;
      // exclusive scan of the tile's per-destination counts over the (<= 16) destinations
      const unsigned cnt = run - s_tfirst[tid];
      unsigned inc = cnt;
#pragma unroll
      for (int d = 1; d < kRouteMaxRanks; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffu, inc, d, kRouteMaxRanks);
        if (tid >= d) inc += y;
      }
      s_dstart[tid] = inc - cnt;
      if (tid == kRouteMaxRanks - 1) s_dstart[kRouteMaxRanks] = inc;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRouteItems; ++i) {
      if (pos[i] < end) {
        const unsigned at = s_wbase[warp][dest[i]] + rank[i];
        s_rec[s_dstart[dest[i]] + (at - s_tfirst[dest[i]])] = rec[i];
        st_stream(slot + pos[i], at);
        st_stream(last_next + pos[i], rec[i].z);
      }
    }
    __syncthreads();
    const unsigned in_tile = s_dstart[kRouteMaxRanks];
    for (unsigned j = tid; j < in_tile; j += kRouteBlock) {
      int d = 0;
      for (int q = 1; q < in.world; ++q) d += j >= s_dstart[q] ? 1 : 0;
      s_target[d][s_tfirst[d] + (j - s_dstart[d])] = s_rec[j];
    }
  }
}

__global__ void __launch_bounds__(256)
route_unpack_kernel(const uint32_t* __restrict__ back, const uint32_t* __restrict__ slot,
                    const uint32_t* __restrict__ last_next, int64_t E, const long long* __restrict__ segment,
                    const int64_t* __restrict__ edge_offsets, int world, unsigned long long* __restrict__ node_info_out) {
  __shared__ long long s_seg[kRouteMaxRanks + 1];
  __shared__ long long s_off[kRouteMaxRanks + 1];
  if (threadIdx.x <= kRouteMaxRanks) s_seg[threadIdx.x] = segment[threadIdx.x];
  if (threadIdx.x <= world) s_off[threadIdx.x] = edge_offsets[threadIdx.x];
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < E; e += stride) {
    const uint32_t at = ld_stream(slot + e);
    int d = 0;
    for (int j = 1; j < world; ++j) d += static_cast<long long>(at) >= s_seg[j] ? 1 : 0;
    const unsigned long long id = static_cast<unsigned long long>(s_off[d]) + back[at];
    st_stream(node_info_out + e, (id << 32) | static_cast<unsigned long long>(ld_stream(last_next + e)));
  }
}

// ------------------------------------------------------------------ owner side
struct MergeLayout {
  ResultWords* result;
  unsigned long long* scan_ws;
  unsigned long long* sort_ws;
  size_t zero_bytes;
  unsigned long long *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  uint32_t* run_start;  // [R + 1]
  int row_bits, col_bits, passes;
  MergeLayout(Workspace& ws, int64_t R, int64_t rows_owned, int64_t total_nodes) {
    row_bits = bits_for(rows_owned > 0 ? static_cast<uint64_t>(rows_owned - 1) : 0);
    col_bits = bits_for(total_nodes > 0 ? static_cast<uint64_t>(total_nodes - 1) : 0);
    passes = sort_num_passes(row_bits + col_bits);
    result = ws.take<ResultWords>(1);
    scan_ws = ws.take<unsigned long long>(scan_state_words(R));
    sort_ws = ws.take<unsigned long long>(sort_state_words(R, row_bits + col_bits));
    zero_bytes = ws.used;
    keys_a = ws.take<unsigned long long>(static_cast<size_t>(R));
    keys_b = ws.take<unsigned long long>(static_cast<size_t>(R));
    vals_a = ws.take<uint32_t>(static_cast<size_t>(R));
    vals_b = ws.take<uint32_t>(static_cast<size_t>(R));
    run_start = ws.take<uint32_t>(static_cast<size_t>(R) + 1);
  }
  const unsigned long long* sorted_keys() const { return (passes & 1) ? keys_b : keys_a; }
  const uint32_t* sorted_perm() const { return (passes & 1) ? vals_b : vals_a; }
};

__global__ void __launch_bounds__(256)
merge_keys_kernel(const uint4* __restrict__ records, int64_t R, long long row_lo, long long rows_owned,
                  long long total_nodes, int col_bits, unsigned long long* __restrict__ keys,
                  unsigned long long* __restrict__ status, unsigned long long* __restrict__ ghist0) {
  __shared__ unsigned s_hist[kRadix];
  Digit0Counter digit0;
  digit0.begin(s_hist);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x; b < R; b += stride) {
    const int64_t i = b + threadIdx.x;
    const bool valid = i < R;
    unsigned long long key = 0;
    if (valid) {
      const uint4 r = __ldcs(records + i);
      long long row = static_cast<long long>(r.y) - row_lo;
      long long col = r.x;
      if (row < 0 || row >= rows_owned || col >= total_nodes) {
        atomicOr(reinterpret_cast<unsigned*>(status), kStatusIdOutOfRange);
        row = col = 0;
      }
      key = (static_cast<unsigned long long>(row) << col_bits) | static_cast<unsigned long long>(col);
      keys[i] = key;
    }
    digit0.count(static_cast<unsigned>(key) & (kRadix - 1), valid);
  }
  digit0.end(ghist0);
}

struct MergeHeadProducer {
  const unsigned long long* keys;
  __device__ unsigned long long operator()(int64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1ull : 0ull; }
};
struct MergeRunConsumer {
  uint32_t* run_start;
  int64_t n;
  const uint32_t* perm;
  uint32_t* inverse;  // merged-edge index (local to this owner) of every record, in arrival order
  __device__ void operator()(int64_t i, unsigned long long head, unsigned long long prefix) const {
    if (head) run_start[prefix] = static_cast<uint32_t>(i);
    if (i == n - 1) run_start[prefix + head] = static_cast<uint32_t>(n);
    inverse[perm[i]] = static_cast<uint32_t>(prefix + head - 1);
  }
};

__global__ void __launch_bounds__(256)
merge_fill_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ perm,
                  const uint32_t* __restrict__ run_start, const uint4* __restrict__ records, int64_t num_out,
                  long long row_lo, int col_bits, int64_t* __restrict__ out_ei, float* __restrict__ out_w,
                  int64_t* __restrict__ out_last) {
  const unsigned long long mask = col_bits >= 64 ? ~0ull : ((1ull << col_bits) - 1);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < num_out; r += stride) {
    const uint32_t a = run_start[r];
    const uint32_t b = run_start[r + 1];
    const unsigned long long k = keys[a];
    st_stream(out_ei + r, static_cast<int64_t>((k >> col_bits) + static_cast<unsigned long long>(row_lo)));
    st_stream(out_ei + num_out + r, static_cast<int64_t>(k & mask));
    const uint4 first = records[perm[a]];
    float acc = __uint_as_float(first.w);
    for (uint32_t i = a + 1; i < b; ++i) acc += __uint_as_float(records[perm[i]].w);  // arrival order = stream order
    st_stream(out_w + r, acc);
    st_stream(out_last + r, static_cast<int64_t>(first.z));
  }
}

// node sequences of the owned rows of the next layer: the row of the merged edge's source ++ its last node
__global__ void __launch_bounds__(256)
extend_owned_rows_kernel(const int64_t* __restrict__ prev, int width, long long prev_row_lo,
                         const int64_t* __restrict__ src_ids, const int64_t* __restrict__ last, int64_t n,
                         int64_t* __restrict__ out) {
  const int ow = width + 1;
  const int64_t total = n * ow;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t j = t / ow;
    const int c = static_cast<int>(t - j * ow);
    st_stream(out + t, c < width ? prev[(src_ids[j] - prev_row_lo) * width + c] : last[j]);
  }
}

}  // namespace ppg

using namespace ppg;

// =================================================================== sender side
extern "C" size_t ppg_route_workspace_bytes(int64_t num_edges) {
  Workspace ws(nullptr, 0);
  RouteLayout L(ws, num_edges < 0 ? 0 : num_edges);
  return ws.used + 256;
}

static int route_input(RouteInput& in, const int64_t* line_index, int64_t E, const void* node_info, const float* weights,
                       int64_t own_prefix, const int64_t* offsets, int world) {
  PPG_REQUIRE(E >= 0 && E < (1ll << 31), PPG_ERR_INVALID, "route: %lld edges outside [0, 2^31)", (long long)E);
  PPG_REQUIRE(world >= 1 && world <= kRouteMaxRanks, PPG_ERR_INVALID, "route: %d ranks outside [1, %d]", world, kRouteMaxRanks);
  in.li0 = line_index;
  in.li1 = line_index + E;
  in.E = E;
  in.node_info = static_cast<const unsigned long long*>(node_info);
  in.w = weights;
  in.own_prefix = own_prefix;
  in.offsets = offsets;
  in.world = world;
  in.chunk_edges = route_geometry(E).chunk_edges;
  return PPG_OK;
}

extern "C" int ppg_route_count(const int64_t* line_index, int64_t E, const void* node_info, const int64_t* offsets,
                               int world, void* workspace, size_t workspace_bytes, int64_t* out_counts, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RouteInput in;
  PPG_TRY(route_input(in, line_index, E, node_info, nullptr, 0, offsets, world));
  Workspace ws(workspace, workspace_bytes);
  RouteLayout L(ws, E);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "route: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  const RouteGeometry g = route_geometry(E);
  if (E == 0) {  // one empty chunk: the scan writes zero totals and an all-zero segment table
    PPG_CUDA_TRY(cudaMemsetAsync(L.counts, 0, sizeof(uint32_t) * kRouteMaxRanks, stream));
  } else {
    route_count_kernel<<<g.chunks, kRouteBlock, 0, stream>>>(in, L.counts);
    PPG_LAUNCHED();
  }
  route_scan_kernel<<<1, 32 * kRouteMaxRanks, 0, stream>>>(L.counts, g.chunks, world, L.base, L.segment, out_counts);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_route_pack(const int64_t* line_index, int64_t E, const void* node_info, const float* weights,
                              int64_t own_prefix, const int64_t* offsets, int world, const void* workspace,
                              void* out_records, void* const* h_peer_records, uint32_t* out_slot, uint32_t* out_last,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (E == 0) return PPG_OK;
  RouteInput in;
  PPG_TRY(route_input(in, line_index, E, node_info, weights, own_prefix, offsets, world));
  PPG_REQUIRE((out_records != nullptr) != (h_peer_records != nullptr), PPG_ERR_INVALID,
              "route_pack: give either a local record buffer or the peers' receive buffers");
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  RouteLayout L(ws, E);
  RouteTargets targets = {};
  targets.local = static_cast<uint4*>(out_records);
  if (h_peer_records != nullptr)
    for (int d = 0; d < world; ++d) targets.peer[d] = static_cast<uint4*>(h_peer_records[d]);
  route_pack_kernel<<<route_geometry(E).chunks, kRouteBlock, 0, stream>>>(in, L.base, L.segment, targets, out_slot, out_last);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_route_unpack(const void* workspace, int64_t E, const uint32_t* back, const uint32_t* slot,
                                const uint32_t* last, const int64_t* edge_offsets, int world, void* out_node_info,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (E == 0) return PPG_OK;
  PPG_REQUIRE(world >= 1 && world <= kRouteMaxRanks, PPG_ERR_INVALID, "route: %d ranks outside [1, %d]", world, kRouteMaxRanks);
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  RouteLayout L(ws, E);
  route_unpack_kernel<<<grid_for(E, 256 * 4), 256, 0, stream>>>(back, slot, last, E, L.segment, edge_offsets, world,
                                                               static_cast<unsigned long long*>(out_node_info));
  PPG_LAUNCHED();
  return PPG_OK;
}

// =================================================================== owner side
extern "C" size_t ppg_merge_records_workspace_bytes(int64_t num_records, int64_t rows_owned, int64_t total_nodes) {
  Workspace ws(nullptr, 0);
  MergeLayout L(ws, num_records < 0 ? 0 : num_records, rows_owned, total_nodes);
  return ws.used + 256;
}

extern "C" int ppg_merge_records_sort(const void* records, int64_t R, int64_t row_lo, int64_t rows_owned,
                                      int64_t total_nodes, void* workspace, size_t workspace_bytes, uint32_t* out_inverse,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(R >= 0 && R < (1ll << 31), PPG_ERR_INVALID, "merge_records: %lld records outside [0, 2^31)", (long long)R);
  PPG_REQUIRE(total_nodes >= 0 && total_nodes <= (1ll << 32) && rows_owned >= 0, PPG_ERR_INVALID,
              "merge_records: %lld nodes exceed the 32-bit record fields", (long long)total_nodes);
  Workspace ws(workspace, workspace_bytes);
  MergeLayout L(ws, R, rows_owned, total_nodes);
  PPG_REQUIRE(ws.fits(), PPG_ERR_WORKSPACE, "merge_records: workspace %zu < %zu bytes", workspace_bytes, ws.used);
  PPG_REQUIRE(L.row_bits + L.col_bits <= 64, PPG_ERR_INVALID, "merge_records: key needs more than 64 bits");
  PPG_CUDA_TRY(cudaMemsetAsync(workspace, 0, L.zero_bytes, stream));  // result.total = 0 when R == 0
  if (R == 0) return PPG_OK;
  merge_keys_kernel<<<grid_for(R, 256 * 4), 256, 0, stream>>>(static_cast<const uint4*>(records), R, row_lo, rows_owned,
                                                              total_nodes, L.col_bits, L.keys_a, &L.result->status, L.sort_ws);
  PPG_LAUNCHED();
  int in_b = 0;
  PPG_TRY(radix_sort_pairs<unsigned long long>(L.keys_a, L.keys_b, L.vals_a, L.vals_b, true, true, R, L.row_bits + L.col_bits,
                                               L.sort_ws, &in_b, stream, nullptr, true));
  PPG_REQUIRE((in_b != 0) == ((L.passes & 1) != 0), PPG_ERR_CUDA, "merge_records: internal buffer parity mismatch");
  PPG_TRY(launch_scan(MergeHeadProducer{L.sorted_keys()}, MergeRunConsumer{L.run_start, R, L.sorted_perm(), out_inverse}, R,
                      L.scan_ws, &L.result->total, stream));
  return PPG_OK;
}

extern "C" int ppg_merge_records_fill(const void* workspace, const void* records, int64_t R, int64_t row_lo,
                                      int64_t rows_owned, int64_t total_nodes, int64_t num_out, int64_t* out_edge_index,
                                      float* out_weights, int64_t* out_last, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_out == 0 || R == 0) return PPG_OK;
  Workspace ws(const_cast<void*>(workspace), ~static_cast<size_t>(0));
  MergeLayout L(ws, R, rows_owned, total_nodes);
  merge_fill_kernel<<<grid_for(num_out, 256), 256, 0, stream>>>(L.sorted_keys(), L.sorted_perm(), L.run_start,
                                                                static_cast<const uint4*>(records), num_out, row_lo, L.col_bits,
                                                                out_edge_index, out_weights, out_last);
  PPG_LAUNCHED();
  return PPG_OK;
}

extern "C" int ppg_extend_owned_rows(const int64_t* prev_rows, int64_t width, int64_t prev_row_lo, const int64_t* src_ids,
                                     const int64_t* last, int64_t n, int64_t* out_rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PPG_REQUIRE(width >= 1 && width < (1 << 20) && n >= 0, PPG_ERR_INVALID, "extend_owned_rows: bad shape");
  if (n == 0) return PPG_OK;
  extend_owned_rows_kernel<<<grid_for(n * (width + 1), 256 * 4), 256, 0, stream>>>(prev_rows, static_cast<int>(width), prev_row_lo,
                                                                                   src_ids, last, n, out_rows);
  PPG_LAUNCHED();
  return PPG_OK;
}
