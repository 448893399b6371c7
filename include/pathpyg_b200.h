/* pathpyg_b200 -- C ABI of the B200 (sm_100a) lift -> DBGNN hot path.
 *
 * The reference (pathpy/pathpyG, /root/reference) has no FFI: its hot path is Python calling
 * torch / torch_geometric ops.  This header is the boundary a maintainer would bind instead
 * (ctypes stub in INTEGRATION.md): plain device pointers, element counts, a cudaStream_t passed
 * as void*, int status codes.  No torch types, no hidden global state (last-error text is
 * thread-local), no host synchronisation except the single count read-back of each
 * count -> allocate -> fill pair.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (host);
 *   - index tensors are int64 and row-major exactly as the reference lays them out
 *     (edge_index [2,E]: row 0 then row 1; node_sequence [M,k]: row after row);
 *   - `workspace` comes from the caller (size from the matching *_workspace_bytes); the library
 *     never allocates device memory, so the caller's allocator (torch's caching allocator) owns it;
 *   - `stream` is the cudaStream_t the work is enqueued on;
 *   - return 0 on success; PPG_ERR_INVALID maps to ValueError, everything else to RuntimeError;
 *     ppg_last_error() holds the message.
 *
 * Limits (checked): element counts per call < 2^31; packed sort keys <= 64 bits.
 */
#ifndef PATHPYG_B200_H_
#define PATHPYG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPG_ABI_VERSION 1

#define PPG_OK 0
#define PPG_ERR_INVALID 1   /* bad argument / id out of range          -> ValueError   */
#define PPG_ERR_CUDA 2      /* CUDA runtime failure                    -> RuntimeError */
#define PPG_ERR_WORKSPACE 3 /* workspace too small                     -> RuntimeError */
#define PPG_ERR_EMPTY 4     /* temporal lift found no pair (reference: torch.cat([]), temporal.py:53) */

/* element types of attribute / weight arrays */
#define PPG_F32 0
#define PPG_F64 1
#define PPG_I64 2
#define PPG_I32 3

/* aggregate_node_attributes rules (lift_order.py:33-44) */
#define PPG_PAIR_SRC 0
#define PPG_PAIR_DST 1
#define PPG_PAIR_MAX 2
#define PPG_PAIR_MUL 3
#define PPG_PAIR_ADD 4

/* coalesce reductions (lift_order.py:139-144 -> torch_geometric.utils.coalesce(reduce=...)) */
#define PPG_REDUCE_SUM 0
#define PPG_REDUCE_MEAN 1
#define PPG_REDUCE_MIN 2
#define PPG_REDUCE_MAX 3

/* how `t_f <= t_e + delta` is evaluated (torch type promotion at temporal.py:30,43) */
#define PPG_TIME_I64 0          /* int64 time, integer delta: exact int64 arithmetic            */
#define PPG_TIME_F64 1          /* float64 time: t_e + (double)(float)delta                     */
#define PPG_TIME_I64_F32DELTA 2 /* int64 time, float delta: (float)t_f <= (float)t_e + (float)delta */
#define PPG_TIME_GROUPED 0x100  /* or'ed into time_mode: ppg_lift_temporal_group already ran on this workspace */

int ppg_abi_version(void);
/* Deferred count read-back.  ppg_lift_order_count, ppg_lift_temporal_count and ppg_coalesce_sort accept a NULL
 * host count pointer: they then only enqueue their kernels, and the host collects {count, status bits}
 * (bit 0: a node id was out of range) later with this call -- one synchronisation for several pending ops. */
int ppg_result_read(const void* workspace, int64_t* h_total, int* h_status_bits, void* stream);
const char* ppg_last_error(void);
/* number of kernels this library has launched in this process (monotonic; for bench accounting) */
unsigned long long ppg_launch_count(void);
/* Opt-in timing of the hot kernels with CUDA events on the launching stream: between _begin and _end every launch of a
 * radix digit pass, a chain tile kernel, an owner merge kernel and a fused tensor-core GCN layer is bracketed by an event
 * pair (at most 512 launches); _end synchronises the device and returns, per launch, its duration, its kind, the number
 * of items (pairs / slots / records / nodes) and the algorithmic bytes per item (0 for the GCN layer: the caller knows its
 * edge count).  Benchmark instrumentation; off by default. */
#define PPG_PROFILE_DIGIT_PASS 0
#define PPG_PROFILE_CHAIN_TILES 1
#define PPG_PROFILE_MERGE_TILES 2
#define PPG_PROFILE_GCN_LAYER 3
int ppg_profile_begin(void);
int ppg_profile_end(float* h_ms, int64_t* h_items, int* h_bytes_per_item, int* h_kind, int capacity, int* h_count);

/* ---------------------------------------------------------------------------------------------
 * a2  lift_order_edge_index            reference: src/pathpyG/algorithms/lift_order.py:48-79
 *   edge_index [2,E] sorted by row; output [2,E'] with E' = sum_e outdeg(dst e), columns ascending
 *   in (e, f):  (e, ptr[dst e] + j).
 * ------------------------------------------------------------------------------------------- */
size_t ppg_lift_order_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int ppg_lift_order_count(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, void* workspace,
                         size_t workspace_bytes, int64_t* h_num_lifted, void* stream);
int ppg_lift_order_fill(const void* workspace, int64_t num_edges, int64_t num_nodes, int64_t num_lifted,
                        int64_t* out_index, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3  aggregate_node_attributes        reference: src/pathpyG/algorithms/lift_order.py:10-45
 *   out[j] = rule(attr[edge_index[0][j]], attr[edge_index[1][j]])
 *   ids outside [0, num_attr) read attr[0] and set bit 0 of *status_word (device uint32, zeroed by the caller,
 *   may be NULL when the caller knows the ids are in range); the reference raises IndexError there.
 * ------------------------------------------------------------------------------------------- */
int ppg_pair_attributes(const int64_t* edge_index, int64_t num_edges, const void* attr, int64_t num_attr, int dtype,
                        int rule, void* out, void* status_word, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a1  lift_order_temporal              reference: src/pathpyG/algorithms/temporal.py:17-54
 *   edge_index [2,m] and time [m] sorted by time; (e -> f) iff dst(e) == src(f) and
 *   t_e < t_f <= t_e + delta; output [2,E2] ascending in (e, f).  PPG_ERR_EMPTY if E2 == 0.
 * ------------------------------------------------------------------------------------------- */
size_t ppg_lift_temporal_workspace_bytes(int64_t num_edges, int64_t num_nodes);
/* Optional first half of ppg_lift_temporal_count: groups the events by source node.  It reads the SOURCE row only
 * (edge_index[0], num_edges entries), so a caller that uploads its inputs can start it while the target row and the
 * time stamps are still being copied; pass time_mode | PPG_TIME_GROUPED to the count call on the same workspace. */
int ppg_lift_temporal_group(const int64_t* source_row, int64_t num_edges, int64_t num_nodes, void* workspace,
                            size_t workspace_bytes, void* stream);
int ppg_lift_temporal_count(const int64_t* edge_index, const void* time, int64_t num_edges, int64_t num_nodes,
                            int time_mode, int64_t delta_i, double delta_f, void* workspace, size_t workspace_bytes,
                            int64_t* h_num_pairs, void* stream);
int ppg_lift_temporal_fill(const void* workspace, int64_t num_edges, int64_t num_nodes, int64_t num_pairs,
                           int64_t* out_index, void* stream);

/* Device arrays of a grouped / counted temporal workspace (consumed by the layer chain below): out[0] CSR pointer over the
 * source node u32 [N+1], out[1] event at every grouped position u32 [m], out[2] source node of every grouped position
 * u32 [m], out[3] grouped position of every event's first continuation u32 [m], out[4] row pointer of the event graph
 * u64 [m+1], out[5] {total, status}. */
int ppg_lift_temporal_views(const void* workspace, int64_t num_edges, int64_t num_nodes, const void** out);

/* ---------------------------------------------------------------------------------------------
 * a4-a6  layer chain without a global sort per order      reference: MultiOrderModel.from_temporal_graph
 *        src/pathpyG/core/multi_order_model.py:124-192 = lift_order_temporal (temporal.py:17-54) / lift_order_edge_index
 *        (lift_order.py:48-79) + aggregate_edge_index (lift_order.py:109-152) per order
 *
 *   The line graph of every order is expanded in the merged (row, col) order of its source items, which leaves the pairs
 *   grouped by row; a tile of whole rows is ranked by column in shared memory (csrc/chain.cu).  Items keep the index
 *   ("label") they have in the reference's line graph.  Per level the caller provides
 *     slot arrays  [pairs]: rowS, colS (merged ids of both ends), labS (label of the pair at every slot; it is the merged
 *                            order P of the next level), wS (optional weights), idS (merged id of every slot), and for
 *                            the next level firstS / degS (first continuation of the slot's pair and their number)
 *     node words   u64 [pairs + 1] by label: merged id (low word; read with stride 2 it is inverse_idx of the next layer)
 *                            | number of continuations, then -- after ppg_chain_scan_nodes -- row pointer of the next
 *                            level (high word)
 *     run_start    [merged + 1]: first slot of every merged edge
 *     result       int64[8], zeroed: [0] merged edges, [1] status bits (1: node id out of range), [2] pairs in rows
 *                  longer than `heavy`, [3] such rows (listed in heavy_list: {first slot, length} u32 pairs, capacity
 *                  pairs / (heavy + 1) + 2)
 *   ppg_chain_first_tiles   level 1: rows = source nodes; ptr1 / grouped / sorted_src from ppg_lift_temporal_views; writes
 *                           the id word of the events' node words, ppg_chain_node_ptr the other word (from the event
 *                           graph's row pointer)
 *   ppg_chain_count_sorted  level 2: continuation counts of the events in merged order P (events >= limit are not
 *                           expanded): offP u64 [n+1], firstP, lblP, wP [n], srcbound [ceil(pairs / tile)][2] (rowid /
 *                           run_start: the previous level's idS / run_start)
 *   ppg_chain_count_sorted_next  levels >= 3: the counts are already in merged order (degS); one gather per source
 *   ppg_chain_tiles         expand + rank: `via` maps a continuation position to its item (level 2: `grouped`);
 *                           node_prev: node words of the previous level; the run heads are found in the same kernel
 *                           (tile_state: 8 bytes per tile of ppg_chain_tile_slots() pairs) -> idS, node_out, run_start,
 *                           result[0]: final unless result[3] != 0
 *   ppg_chain_scan_nodes    label-order scan of the counts in the node words -> row pointer of the next level, total
 *   ppg_chain_heavy_fix     rows the tiles left in generation order: one radix sort over their pairs; then
 *   ppg_chain_heads         run heads of the slots again -> idS, id word of the node words, run_start, result[0]
 *   ppg_chain_fill          merged edges [2, merged] int64 + weights (unit weights when wS is NULL)
 * ------------------------------------------------------------------------------------------- */
int ppg_chain_heavy_default(void);
int ppg_chain_tile_slots(void);
size_t ppg_chain_scan_workspace_bytes(int64_t num_items);
int ppg_chain_first_tiles(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, const uint32_t* ptr1,
                          const uint32_t* grouped, const uint32_t* sorted_src, const float* weights, int heavy, uint32_t* rowS,
                          uint32_t* colS, uint32_t* labS, float* wS, uint32_t* idS, void* node_out, uint32_t* run_start,
                          void* tile_state, void* heavy_list, int64_t* result, void* stream);
int ppg_chain_count_sorted(const uint32_t* P, int64_t num_items, const uint32_t* first, const void* ptr_next,
                           const float* w_item, int64_t limit, const uint32_t* rowid, const uint32_t* run_start,
                           void* workspace, size_t workspace_bytes, void* offP, uint32_t* firstP, uint32_t* lblP, float* wP,
                           uint32_t* srcbound, void* stream);
int ppg_chain_tiles(int64_t num_sources, int64_t num_rows, int64_t num_slots, const void* offP, const uint32_t* firstP,
                    const uint32_t* lblP, const float* wP, const uint32_t* run_start, const uint32_t* rowid,
                    const void* node_prev, const uint32_t* via, const uint32_t* srcbound, int heavy, uint32_t* rowS,
                    uint32_t* colS, uint32_t* labS, float* wS, uint32_t* idS, void* node_out, uint32_t* firstS, uint32_t* degS,
                    uint32_t* run_start_out, void* tile_state, void* heavy_list, int64_t* result, void* stream);
/* per-item words: u32 words `stride` apart, the one in question at index `word` (node words: stride 2, pointer word 1;
 * info words of the distributed build: stride 4, pointer word 2) */
int ppg_chain_node_ptr(const void* off, int64_t num_items, void* node, int stride, int word, void* stream);
int ppg_chain_scan_nodes(void* node, int stride, int word, int64_t num_items, void* workspace, size_t workspace_bytes,
                         int64_t* total, void* stream);
int ppg_chain_count_sorted_next(const uint32_t* P, int64_t num_items, const uint32_t* degS, const void* node, int stride,
                                int word, int64_t limit, const uint32_t* rowid, const uint32_t* run_start, void* workspace,
                                size_t workspace_bytes, void* offP, uint32_t* lblP, uint32_t* srcbound, void* stream);
int ppg_chain_heads(const uint32_t* rowS, const uint32_t* colS, const uint32_t* labS, int64_t num_slots, void* workspace,
                    size_t workspace_bytes, uint32_t* idS, uint32_t* id_item, int id_stride, uint32_t* run_start,
                    int64_t* result, void* stream);
size_t ppg_chain_heavy_workspace_bytes(int64_t heavy_slots, int64_t heavy_rows, int64_t num_slots);
int ppg_chain_heavy_fix(const void* heavy_list, int64_t heavy_rows, int64_t heavy_slots, int64_t num_slots, uint32_t* colS,
                        uint32_t* labS, float* wS, uint32_t* extra0, uint32_t* extra1, uint32_t* extra2, void* workspace,
                        size_t workspace_bytes, void* stream);
int ppg_chain_fill(const uint32_t* rowS, const uint32_t* colS, const float* wS, const uint32_t* run_start, int64_t num_out,
                   int64_t* out_edge_index, float* out_weights, void* stream);
int ppg_chain_widen(const uint32_t* in, int in_stride, int64_t n, int64_t* out, void* stream);

/* The same chain on a rank of a distributed build (no counterpart in /root/reference).  A rank expands its
 * items in the order of their GLOBAL merged ids, so its pairs leave as 16-byte records {col, row, last node, weight} sorted
 * by (row, col) in global ids, and the records of one owner (owners hold ascending row ranges) are one contiguous range of
 * the sender's buffer; an owner reads one sorted run per sender (peer memory), merges them in shared-memory tiles and
 * stores the merged-edge index of every record into the sender's answer buffer.
 *   info words  uint4 [items + 1] by label: {global id, last node, number of continuations -> row pointer, -}
 *   ppg_chain_pack          level 1: the slot arrays of ppg_chain_first_tiles -> records
 *   ppg_chain_tiles_dist    levels >= 2: as ppg_chain_tiles with info words for node words and LOCAL rows (rowid /
 *                           run_start / row_value from ppg_chain_unpack); writes records, labS, firstS / degS and the
 *                           count word of the new items' info words (info_out)
 *   ppg_chain_heavy_fix_records  ppg_chain_heavy_fix for record slots
 *   ppg_chain_dest_bounds   first slot of every destination rank: dstart [world + 1], counts [world] (device int64)
 *   ppg_merge_sorted        owner: merge the runs -> compact merged arrays + the merged edge of every record (into
 *                           h_back[s]); result[1] & 2: a tile overflowed, use ppg_merge_records_* on a copy instead
 *   ppg_chain_unpack        sender: returned indices -> global ids, local rows of the next level, info words */
int ppg_chain_pack(const uint32_t* rowS, const uint32_t* colS, const uint32_t* lastS, const float* wS, int64_t num_slots,
                   void* records, void* stream);
int ppg_chain_tiles_dist(int64_t num_sources, int64_t num_slots, const void* offP, const uint32_t* firstP, const uint32_t* lblP,
                         const float* wP, int w_stride, const uint32_t* run_start, const uint32_t* rowid,
                         const uint32_t* row_value, const void* info, const uint32_t* via, const uint32_t* srcbound, int heavy,
                         void* records, uint32_t* labS, uint32_t* firstS, uint32_t* degS, void* info_out, void* tile_state,
                         void* heavy_list, int64_t* result, void* stream);
size_t ppg_chain_heavy_records_workspace_bytes(int64_t heavy_slots, int64_t heavy_rows, int64_t num_slots);
int ppg_chain_heavy_fix_records(const void* heavy_list, int64_t heavy_rows, int64_t heavy_slots, int64_t num_slots, void* records,
                                uint32_t* labS, uint32_t* firstS, uint32_t* degS, void* workspace, size_t workspace_bytes,
                                void* stream);
int ppg_chain_dest_bounds(const uint32_t* rows, int stride, int64_t num_slots, const int64_t* offsets, int world,
                          int64_t* dstart, int64_t* counts, void* stream);
int ppg_chain_unpack(const uint32_t* back, int64_t num_slots, const int64_t* dstart, const int64_t* edge_offsets, int world,
                     const uint32_t* labS, const uint32_t* lastS, int last_stride, const uint32_t* degS, void* workspace,
                     size_t workspace_bytes, uint32_t* rowid, uint32_t* run_start, uint32_t* row_value, void* info_item,
                     int64_t* result, void* stream);
int64_t ppg_merge_sorted_tiles(int64_t num_records);
int ppg_merge_sorted(void* const* h_runs, const int64_t* h_run_len, void* const* h_back, int world, int64_t row_lo,
                     int64_t rows_owned, int64_t total_nodes, uint32_t* bounds, void* tile_state, uint32_t* row_m, uint32_t* col_m,
                     float* w_m, uint32_t* last_m, int64_t* result, void* stream);
int ppg_merge_sorted_fill(const uint32_t* row_m, const uint32_t* col_m, const float* w_m, const uint32_t* last_m,
                          int64_t num_out, int64_t* out_edge_index, float* out_weights, int64_t* out_last, void* stream);

/* ---------------------------------------------------------------------------------------------
 * e   cross-partition exchange of lifted edges (the reference is single-device, there is no
 *     counterpart in /root/reference: the gathered result must equal MultiOrderModel.from_temporal_graph,
 *     src/pathpyG/core/multi_order_model.py:124-192, bit for bit)
 *
 *   ppg_lift_limit        after ppg_lift_order_count (temporal = 0) / ppg_lift_temporal_count (temporal = 1): make
 *                         the pending count the number of columns whose source is < limit_sources (a prefix of the
 *                         output, which is ascending in the source); fill with that count writes exactly that prefix.
 *   ppg_route_count       line_index [2,E] of one level; node_info [line nodes] u64 = id << 32 | last first-order node
 *                         (NULL on the first level: id = the value itself); offsets [world + 1] (device) = first row
 *                         owned by every rank.  out_counts [world] (device int64) = records per destination.
 *   ppg_route_pack        stable partition by destination: out_records [E] 16-byte records {id(target), id(source),
 *                         last node of target, float32 weight bits} grouped by destination rank in edge order;
 *                         instead of out_records, h_peer_records (HOST array of `world` DEVICE pointers) names, per
 *                         destination rank, the first slot reserved for this sender in that rank's receive buffer
 *                         (peer memory mapped over NVLink): the kernel then stores into the owners' memory directly;
 *                         sources (first level: positions) >= own_prefix carry weight 0; out_slot [E] = record index;
 *                         out_last [E] = last first-order node of the edge's path.
 *   ppg_route_unpack      back [E] = merged-edge index returned by the owner for every record (record order);
 *                         edge_offsets [world + 1] (device) = global index of every owner's first merged edge;
 *                         out_node_info [E] = the node_info of the NEXT level.
 *   ppg_merge_records_*   owner side: rows [row_lo, row_lo + rows_owned), columns < total_nodes (<= 2^32);
 *                         sort leaves {merged count, status} at the head of the workspace (ppg_result_read) and
 *                         out_inverse [R] = merged-edge index (local) of every record; fill writes the merged edges
 *                         (global ids, (row, col)-sorted), their weights summed in arrival order and their last node.
 *   ppg_extend_owned_rows out_rows [n, width + 1] = prev_rows[src_ids - prev_row_lo] ++ last
 * ------------------------------------------------------------------------------------------- */
#define PPG_ROUTE_MAX_RANKS 16
int ppg_lift_limit(void* workspace, int temporal, int64_t num_sources, int64_t num_nodes, int64_t limit_sources,
                   void* stream);
size_t ppg_route_workspace_bytes(int64_t num_edges);
int ppg_route_count(const int64_t* line_index, int64_t num_edges, const void* node_info, const int64_t* offsets,
                    int world, void* workspace, size_t workspace_bytes, int64_t* out_counts, void* stream);
int ppg_route_pack(const int64_t* line_index, int64_t num_edges, const void* node_info, const float* weights,
                   int64_t own_prefix, const int64_t* offsets, int world, const void* workspace, void* out_records,
                   void* const* h_peer_records, uint32_t* out_slot, uint32_t* out_last, void* stream);
int ppg_route_unpack(const void* workspace, int64_t num_edges, const uint32_t* back, const uint32_t* slot,
                     const uint32_t* last, const int64_t* edge_offsets, int world, void* out_node_info, void* stream);
size_t ppg_merge_records_workspace_bytes(int64_t num_records, int64_t rows_owned, int64_t total_nodes);
int ppg_merge_records_sort(const void* records, int64_t num_records, int64_t row_lo, int64_t rows_owned,
                           int64_t total_nodes, void* workspace, size_t workspace_bytes, uint32_t* out_inverse,
                           void* stream);
int ppg_merge_records_fill(const void* workspace, const void* records, int64_t num_records, int64_t row_lo,
                           int64_t rows_owned, int64_t total_nodes, int64_t num_out, int64_t* out_edge_index,
                           float* out_weights, int64_t* out_last, void* stream);
int ppg_extend_owned_rows(const int64_t* prev_rows, int64_t width, int64_t prev_row_lo, const int64_t* src_ids,
                          const int64_t* last, int64_t n, int64_t* out_rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4  aggregate_edge_index             reference: src/pathpyG/algorithms/lift_order.py:109-152
 *   step 1: distinct rows of node_sequence [M,k] in lexicographic order + inverse
 *           (torch.unique(dim=0, return_inverse=True), :133);
 *   step 2: map the edge index through a table and coalesce duplicates (PyG coalesce, :139-144),
 *           output (row, col)-sorted as Graph.__init__ leaves it (core/graph.py:103-105).
 * ------------------------------------------------------------------------------------------- */
size_t ppg_rows_minmax_workspace_bytes(int64_t width);
int ppg_rows_minmax(const int64_t* rows, int64_t num_rows, int64_t width, void* workspace, size_t workspace_bytes,
                    int64_t* h_col_min, int64_t* h_col_max, int* h_strictly_ascending, void* stream);

/* rows are packed into one key: sum_c (rows[i][c] - h_col_min[c]) << h_col_shift[c], total_bits <= 64 */
size_t ppg_unique_rows_workspace_bytes(int64_t num_rows, int total_bits);
int ppg_unique_rows_sort(const int64_t* rows, int64_t num_rows, int64_t width, const int64_t* h_col_min,
                         const int* h_col_shift, int total_bits, void* workspace, size_t workspace_bytes,
                         int64_t* out_inverse, int64_t* h_num_unique, void* stream);
int ppg_unique_rows_gather(const int64_t* rows, int64_t num_rows, int64_t width, const void* workspace, int total_bits,
                           int64_t num_unique, int64_t* out_rows, void* stream);

/* remap == NULL: edge ids are used as they are; else id -> remap[id] (remap_len entries).
 * Mapped ids must be < num_nodes (EdgeIndex.validate(), core/graph.py:107) else PPG_ERR_INVALID.
 * out_inverse (nullable) [num_edges]: the output edge every input edge was merged into.  Because the
 * distinct edges of layer k-1 in (row, col) order ARE the distinct k-grams in lexicographic order, this is
 * the `inverse_idx` of layer k (torch.unique(node_sequence, dim=0), lift_order.py:133) -- no second sort. */
size_t ppg_coalesce_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int ppg_coalesce_sort(const int64_t* edge_index, int64_t num_edges, const int64_t* remap, int64_t remap_len,
                      int64_t num_nodes, void* workspace, size_t workspace_bytes, int64_t* out_inverse,
                      int64_t* h_num_out, void* stream);
/* weights == NULL: unit weights (the reference's torch.ones default, lift_order.py:130-131), out dtype f32 */
int ppg_coalesce_fill(const void* workspace, int64_t num_edges, int64_t num_nodes, int64_t num_out,
                      const void* weights, int dtype, int reduce, int64_t* out_edge_index, void* out_weights,
                      void* stream);

/* out_rows[j] = prev_rows[edge_index[0][j]] ++ last(prev_rows[edge_index[1][j]])   (multi_order_model.py:114):
 * node sequences [num_edges, width+1] of the next layer from the distinct edges of this one. */
int ppg_extend_rows(const int64_t* prev_rows, int64_t num_prev, int64_t width, const int64_t* edge_index,
                    int64_t num_edges, int64_t* out_rows, void* stream);

/* Stable sort of the low `end_bit` bits of 64-bit keys (in place) + the permutation that sorts them:
 * the radix sort underneath a4 / the CSC build, exposed for the containers' sort_by and for bench.py.
 * h_pass_ms (host, nullable, ceil(end_bit/8) floats): CUDA-event time of every digit pass (synchronises). */
size_t ppg_sort_pairs_workspace_bytes(int64_t n, int end_bit);
int ppg_sort_pairs_u64(uint64_t* keys, uint32_t* out_perm, int64_t n, int end_bit, void* workspace,
                       size_t workspace_bytes, float* h_pass_ms, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a10/a11  DBGNN.forward building blocks      reference: src/pathpyG/nn/dbgnn.py:32-151 and
 *   torch_geometric 2.7.0 GCNConv / MessagePassing.propagate (not vendored, see oracle/pyg.py)
 * ------------------------------------------------------------------------------------------- */
#define PPG_ACT_NONE 0
#define PPG_ACT_ELU 1

/* Target-grouped (CSC) view of an edge list [2,E] (stable: original edge order inside a target):
 * out_colptr [num_targets+1], out_src [E] source node per slot, out_eid [E] original edge per slot. */
size_t ppg_csc_workspace_bytes(int64_t num_edges, int64_t num_targets);
int ppg_csc_build(const int64_t* edge_index, int64_t num_edges, int64_t num_sources, int64_t num_targets,
                  void* workspace, size_t workspace_bytes, int32_t* out_colptr, int32_t* out_src, int32_t* out_eid,
                  void* stream);
/* Same without the host synchronisation: the kernels are only enqueued (so the call can sit inside a CUDA graph),
 * out-of-range ids are clamped to 0 in the outputs and flagged in the workspace head; collect the flag later with
 * ppg_result_read(workspace, ...) (status bit 0). */
int ppg_csc_build_async(const int64_t* edge_index, int64_t num_edges, int64_t num_sources, int64_t num_targets,
                        void* workspace, size_t workspace_bytes, int32_t* out_colptr, int32_t* out_src, int32_t* out_eid,
                        void* stream);

/* gcn_norm with add_remaining_self_loops(fill_value=1): out_val [E] per CSC slot (0 on self-loop
 * slots), out_self [n] normalised self-loop weight; scratch_dis [n]; edge_weight NULL = ones;
 * out_val_edge (nullable) [E]: the same coefficient indexed by ORIGINAL edge id (for the backward view). */
int ppg_gcn_norm(const int32_t* colptr, const int32_t* src, const int32_t* eid, const float* edge_weight, int64_t n,
                 int64_t num_edges, float* scratch_dis, float* out_val, float* out_self, float* out_val_edge,
                 void* stream);

/* out[v] = float(colptr[v+1] - colptr[v]) */
int ppg_colptr_counts(const int32_t* colptr, int64_t n, float* out, void* stream);

/* out[v,:] = act( sum_{slots i of v} val[i] * X[src[i],:] + self_val[v] * X[v,:] + bias )
 * val NULL = 1, self_val NULL = no self term, bias NULL = 0;  X [*,F] row-major, out [num_targets,F] */
int ppg_spmm_csc(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val, const float* X,
                 int64_t num_targets, int64_t F, const float* bias, int act, float* out, void* stream);

/* out[M,N] = act( A1[M,K1] W1[N,K1]^T + rowscale[M] * (A2[M,K2] W2[N,K2]^T + bias[N]) )
 * A2 / W2 / bias / rowscale may be NULL (rowscale NULL = 1) */
int ppg_linear(const float* A1, const float* W1, int64_t M, int64_t K1, const float* A2, const float* W2, int64_t K2,
               const float* bias, const float* rowscale, int64_t N, int act, float* out, void* stream);

/* Fused layers for widths F, H in {16, 32, 64} (ppg_gcn_fused_supported): aggregate the 128-node tile in
 * shared memory, transform it there, apply bias + activation, write the output row once.
 *   gcn       : out = act( (sum_i val_i X[src_i] + self_v X[v]) W^T + bias )         W [H,F]
 *   bipartite : out = act( (sum_i X_h[src_i]) W1^T + indeg(v) (X[v] W2^T + bias12) ) W1, W2 [H,F] */
int ppg_gcn_fused_supported(int64_t F, int64_t H);
int ppg_gcn_layer_fused(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val,
                        const float* X, const float* W, const float* bias, int64_t n, int64_t F, int64_t H, int act,
                        float* out, void* stream);
int ppg_bipartite_fused(const int32_t* colptr, const int32_t* src, const float* X_h, const float* X, const float* W1,
                        const float* W2, const float* bias12, int64_t n, int64_t F, int64_t H, int act, float* out,
                        void* stream);

/* The same GCN layer with the dense transform on the tcgen05 tensor cores (3xTF32 split, fp32-accurate),
 * accumulator in TMEM; F in {32, 64}, H in {16, 32, 64} (ppg_gcn_tc_supported).  `e` = number of CSC slots
 * (colptr[n]) or 0 if the caller does not know it: sparse graphs (e <= 4 n) take the kernel that streams the rows
 * through shared-memory stages, dense ones the kernel that gathers them into registers; the results are the same
 * bit for bit. */
int ppg_gcn_tc_supported(int64_t F, int64_t H);
int ppg_gcn_layer_tc(const int32_t* colptr, const int32_t* src, const float* val, const float* self_val, const float* X,
                     const float* W, const float* bias, int64_t n, int64_t e, int64_t F, int64_t H, int act, float* out,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of a10/a11 (the reference trains through torch autograd of PyG's GCNConv / propagate;
 * docs/tutorial/dbgnn.ipynb cell 42).  For Y = act(A X W^T + b):
 *   dPre = dY * act'(pre), db = colsum(dPre)            ppg_act_backward
 *   G = A^T dPre                                        ppg_spmm_csc on the source-grouped view
 *   dW = G^T X                                          ppg_atb
 *   dX = G W                                            ppg_linear
 * All reductions have a fixed order (no floating-point atomics).
 * ------------------------------------------------------------------------------------------- */
/* dPre[M,H] = dY * act'(.) (ELU: 1 if Y > 0 else Y + 1; Y may be NULL for PPG_ACT_NONE); with rowscale:
 * dPreScaled = rowscale[:,None] * dPre; out_colsum[H] = column sums of dPreScaled if rowscale else of dPre.
 * dPre / dPreScaled may be NULL (not written). */
size_t ppg_act_backward_workspace_bytes(int64_t M, int64_t H);
int ppg_act_backward(const float* dY, const float* Y, const float* rowscale, int64_t M, int64_t H, int act, float* dPre,
                     float* dPreScaled, float* out_colsum, void* workspace, size_t workspace_bytes, void* stream);
/* out[H,F] = A[M,H]^T B[M,F] */
size_t ppg_atb_workspace_bytes(int64_t M, int64_t H, int64_t F);
int ppg_atb(const float* A, const float* B, int64_t M, int64_t H, int64_t F, float* out, void* workspace,
            size_t workspace_bytes, void* stream);
/* out[i] = src[idx[i]] */
int ppg_gather_f32(const float* src, const int32_t* idx, int64_t n, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Consumers of the built layers
 *   Graph.degrees / transition_probabilities      reference: src/pathpyG/core/graph.py:486-533
 *   MultiOrderModel.get_mon_dof                   reference: src/pathpyG/core/multi_order_model.py:243-312
 *   MultiOrderModel.get_*_log_likelihood          reference: src/pathpyG/core/multi_order_model.py:314-409
 * ------------------------------------------------------------------------------------------- */
/* out_ptr[v] = first slot with sorted_ids[slot] >= v, v in [0, num_nodes]: the CSR pointer of a SORTED id column */
int ppg_sorted_ids_ptr(const int64_t* sorted_ids, int64_t num_ids, int64_t num_nodes, int32_t* out_ptr, void* stream);
/* out[s] = sum_{i in [ptr[s], ptr[s+1])} weights[perm ? perm[i] : i], added in slot order (the order of a sequential
 * scatter_add); weights NULL: out[s] = ptr[s+1] - ptr[s].  Segments longer than 64 slots use a fixed fp64 tree. */
size_t ppg_segment_sum_workspace_bytes(int64_t num_slots);
int ppg_segment_sum(const int32_t* ptr, const int32_t* perm, const float* weights, int64_t num_segments, int64_t num_slots,
                    void* workspace, size_t workspace_bytes, float* out, void* stream);
/* out[e] = (weights ? weights[e] : 1) / denom[ids[e]]     (graph.py:531-533) */
int ppg_edge_ratio(const int64_t* ids, const float* weights, const float* denom, int64_t num_edges, float* out, void* stream);
/* h_num_walks[k-1]   = number of walks with k edges          (= columns of the k-th line-graph lift, :285-291),
 * h_num_sources[k-1] = number of nodes that start such a walk (= non-empty rows of A^k, :294-303), k = 1..max_len.
 * Counts are exact below 2^64.  Synchronises (host outputs). */
size_t ppg_walk_counts_workspace_bytes(int64_t num_nodes, int max_len);
int ppg_walk_counts(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int max_len, void* workspace,
                    size_t workspace_bytes, int64_t* h_num_walks, int64_t* h_num_sources, void* stream);
/* *h_out = sum_i freq[i] * logf(prob[j]),  j = i, idx[i] or idx2[idx[i]] (idx / idx2 nullable): every term formed in
 * fp32 like torch.mul(frequencies, torch.log(p[...])) (:339,367-368,395-396), accumulated in fp64 in a fixed order.
 * PPG_ERR_INVALID if an index is out of range.  Synchronises (host output). */
size_t ppg_weighted_log_sum_workspace_bytes(void);
int ppg_weighted_log_sum(const float* freq, const float* prob, const int64_t* idx, const int64_t* idx2, int64_t n,
                         int64_t prob_len, int64_t idx2_len, void* workspace, size_t workspace_bytes, double* h_out,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-walk bookkeeping of a7 / a9: what the reference writes as torch.cumsum, repeat_interleave, bincount and
 * arange + mask   reference: src/pathpyG/core/multi_order_model.py:217-224,335,354-361,402-405,
 *                            src/pathpyG/core/path_data.py:139-159
 * ------------------------------------------------------------------------------------------- */
/* offsets [n + 1] = exclusive prefix sums of counts [n] (offsets[n] = their sum; the workspace starts with {total,
 * status}: ppg_result_read).  A count below min_count sets status bit 0 and counts as 0. */
size_t ppg_counts_to_offsets_workspace_bytes(int64_t n);
int ppg_counts_to_offsets(const int64_t* counts, int64_t n, int64_t min_count, void* workspace, size_t workspace_bytes,
                          int64_t* offsets, void* stream);
/* repeat_interleave: out_values[j] = values[i] (elements of 4 or 8 bytes; nullable) and owner[j] = i (nullable) for
 * offsets[i] <= j < offsets[i + 1], j < total = offsets[n] */
int ppg_expand_offsets(const int64_t* offsets, int64_t n, int64_t total, const void* values, int value_bytes,
                       void* out_values, int64_t* owner, void* stream);
/* edge_index [2, total - num_walks]: the links p -> p + 1 (plus base) inside every walk, the walks laid end to end at
 * offsets [num_walks + 1]; every walk has at least one position */
int ppg_walk_chain(const int64_t* offsets, int64_t num_walks, int64_t total, int64_t base, int64_t* edge_index, void* stream);
/* counts [num_bins] (int64) = occurrences of every id; bit 0 of the zeroed status_word is set if an id is outside
 * [0, num_bins) */
int ppg_bincount(const int64_t* ids, int64_t n, int64_t num_bins, int64_t* counts, void* status_word, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Shortest time-respecting paths   reference: src/pathpyG/algorithms/temporal.py:57-107
 *   edge_index [2,m] time-sorted events, event_graph [2,num_pairs] = the a1 output for the same delta.
 *   For the sources [source_begin, source_end) (source_begin a multiple of 32):
 *     out_dist [source_end - source_begin, n] float64: fewest events on a time-respecting path s -> v
 *              (0 on the diagonal, +inf if there is none);
 *     out_pred [same] int64: source node of the event that ends such a path (the one with the largest event
 *              index, which is what scipy's dijkstra reports for the reference's augmented graph), s itself on
 *              the diagonal, -1 if unreachable.
 *   h_levels (nullable): number of breadth-first levels run.  Synchronises once per level.
 * ------------------------------------------------------------------------------------------- */
size_t ppg_temporal_paths_workspace_bytes(int64_t num_events, int64_t num_nodes, int64_t chunk_sources);
int ppg_temporal_paths(const int64_t* edge_index, int64_t num_events, int64_t num_nodes, const int64_t* event_graph,
                       int64_t num_pairs, int64_t source_begin, int64_t source_end, void* workspace,
                       size_t workspace_bytes, double* out_dist, int64_t* out_pred, int* h_levels, void* stream);
/* temporal_betweenness_centrality (centrality.py:164-300: Brandes on the event DAG), one batch of source nodes:
 *   group_off [num_groups+1]: boundaries of the runs of equal time stamps in the time-sorted event list;
 *   succ_ptr [m+1] / succ: CSR of the event graph (the a1 output is sorted by its first row: succ = its second row);
 *   pred_ptr [m+1] / pred: its CSC (ppg_csc_build);  in_ptr [n+1] / in_event: the events grouped by target node;
 *   sources [batch]: node ids in ascending order;  inout_bw [n] float64: the batch's contributions are ADDED.
 * All sums run in a fixed order (no floating-point atomics).  Does not synchronise. */
size_t ppg_temporal_betweenness_workspace_bytes(int64_t num_events, int64_t num_nodes, int64_t batch_sources);
int ppg_temporal_betweenness(const int64_t* edge_index, int64_t num_events, int64_t num_nodes, const int32_t* group_off,
                             int64_t num_groups, const int32_t* succ_ptr, const int64_t* succ, const int32_t* pred_ptr,
                             const int32_t* pred, const int32_t* in_ptr, const int32_t* in_event, const int32_t* sources,
                             int64_t batch_sources, void* workspace, size_t workspace_bytes, double* inout_bw, void* stream);
/* out[v] = sum_{x != v} (n - 1) / dist[x, v] in ascending x (temporal_closeness_centrality, centrality.py:320-322) */
int ppg_temporal_closeness(const double* dist, int64_t num_nodes, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PATHPYG_B200_H_ */
