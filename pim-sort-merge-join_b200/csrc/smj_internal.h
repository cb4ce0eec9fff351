// smj_internal.h -- host-side internals shared by the .cu translation units of libsmj.so.
// Nothing here crosses the C-ABI (include/smj.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/smj.h"

typedef unsigned long long u64;
typedef unsigned int u32;

#define SMJ_RADIX_BITS 8
#define SMJ_RADIX 256
#define SMJ_KEY_PASSES 4            // 32-bit keys, 8-bit digits
#define SMJ_MAX_SORT_ROWS ((1u << 30) - 1) // look-back status words keep 30-bit counts

// One per GPU driven by this process.
struct SmjCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaMemPool_t pool = nullptr;    // private stream-ordered pool of the output buffers
    size_t l2_fetch_saved = 0;       // the device's cudaLimitMaxL2FetchGranularity before this library changed it
    bool l2_fetch_set = false;
    u32 *d_err = nullptr;            // device-side consistency flag (bounded spins, bad partitions)
    void *h_pinned = nullptr;        // small pinned mailbox for counts / histograms
    size_t h_pinned_bytes = 0;
    int64_t launches = 0;
    // grow-only workspace slots (no cudaMalloc in steady state)
    static const int kSlots = 32;
    void *slot[kSlots] = {};
    size_t slot_bytes[kSlots] = {};
    cudaEvent_t ev[16] = {};
    // per-run accounting of the dominant kernel (radix scatter pass): events bracket every pass launch
    static const int kMaxTimedPasses = 16;
    cudaEvent_t pass_ev[2 * kMaxTimedPasses] = {};
    u32 pass_items[kMaxTimedPasses] = {};
    int pass_count = 0;
    bool radix_attr_set = false;
    bool select_attr_set = false;
    bool run_planned = false;        // the last smj_run_single pipeline ran with device sort plans
    bool stage_from_pass = false;    // ... and its select|sort|join boundaries are the events around the radix passes
    bool capturing = false;          // the stream is being captured into the pipeline graph
    // CUDA-graph replay of the smj_run device pipeline (same tables, shapes and knobs as the previous call)
    u64 ws_gen = 0;                  // bumped whenever a workspace slot is (re)allocated
    u64 graph_key[16] = {};
    int graph_seen = 0;              // consecutive calls with this key: 1st runs eagerly, 2nd captures, later ones replay
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_launches = 0;      // kernels inside the captured pipeline
    int32_t **d_out_ptr = nullptr;   // device cells: [0] the output pointer the materialise kernel writes through, [1], [2] the input tables' addresses
};

enum SmjSlot {
    WS_T1 = 0, WS_T2,                // device copies of host input tables
    WS_PAIRS_A1, WS_PAIRS_B1,        // ping-pong (key,rowid) buffers, table 1
    WS_PAIRS_A2, WS_PAIRS_B2,        // table 2
    WS_SCRATCH,                      // zeroed per run: tile status words, tile counters, histograms, counts
    WS_PART,                         // merge-path partitions + run starts
    WS_MATCH,                        // matched (left rowid, right rowid)
    WS_TMP_ROWS, WS_TMP_ROWS2,       // staging for in-place sort / host outputs
    WS_XCHG_SEND1, WS_XCHG_SEND2, WS_XCHG_RECV1, WS_XCHG_RECV2, WS_SAMPLES,
    WS_MERGE_A, WS_MERGE_B, WS_RADIX, WS_MATCH_DENSE, WS_BLOOM,
    WS_DIST_RECV1, WS_DIST_RECV2,    // peer-mapped receive buffers of the key-range exchange (smj_dist.cu)
    WS_ROWSTORE1, WS_ROWSTORE2, WS_ZIPF,
};

// Device-side wait in front of a table's select kernel: go on once *flag >= *seq (null flag: no wait).
struct SmjWait { const u64 *flag = nullptr; const u64 *seq = nullptr; u32 *err = nullptr; };

// One smj_run device pipeline in flight on a context: prepare -> enqueue -> finish (smj_api.cu).
struct SmjRun {
    smj_config_t cfg;
    smj_table_t tb[2];
    const u64 *d_rows[2] = {nullptr, nullptr};   // device-resident row counts (tb[t].rows is then an upper bound)
    SmjWait wait[2];                             // per table: its select kernel waits for this device cell (smj_dist.cu)
    const int32_t *d_t[2] = {nullptr, nullptr};
    int c_out = 0;
    int64_t j_max = 0, launches0 = 0;
    size_t stiles[2] = {}, rb[2] = {}, jt = 0, off_sel = 0, off_radix = 0, off_join = 0, zero_bytes = 0;
    char *scr = nullptr;
    u64 *ping[2] = {}, *pong[2] = {};
    uint2 *mm = nullptr, *md = nullptr;
    int32_t *store[2] = {nullptr, nullptr};      // dense row stores (null: not used for this shape)
    u64 store_max_rows[2] = {0, 0};
    smj_table_t dev_out = {nullptr, 0, 0, 1};
    bool prepared = false, replayed = false;
    bool no_graph = false;                       // run the pipeline eagerly (a process driving several GPUs: smj_dist.cu)
};
int  smj_run_prepare(SmjCtx *c, const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, const u64 *const *d_rows, SmjRun *R);
int  smj_run_enqueue(SmjCtx *c, SmjRun *R);
int  smj_run_finish(SmjCtx *c, SmjRun *R, smj_table_t *out, smj_stats_t *stats);
void smj_run_abandon(SmjCtx *c, SmjRun *R);

void smj_preload_select(void);
void smj_preload_radix(void);
void smj_preload_join(void);
void smj_preload_partition(void);

int   smj_set_error(int code, const char *fmt, ...);
int   smj_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void *smj_ws(SmjCtx *c, int slot, size_t bytes);     // nullptr on failure (error already set)

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return smj_cuda_fail(e_, #x, __FILE__, __LINE__); } while (0)
#define SMJ_TRY(x)  do { int r_ = (x); if (r_ != SMJ_OK) return r_; } while (0)
#define KERNEL_CHECK(c) do { (c)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return smj_cuda_fail(e_, "kernel launch", __FILE__, __LINE__); } while (0)

// cudaEventRecord that also works while the stream is being captured into a graph (becomes an event-record node)
static inline cudaError_t smj_event_record(cudaEvent_t ev, cudaStream_t st)
{
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
        return cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal);
    return cudaEventRecord(ev, st);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef __CUDACC__
// Kernel launch with the programmatic-dependent-launch attribute (see PDL_ENTER in smj_dev.cuh); SMJ_PDL=0 turns it off.
bool smj_pdl_enabled(void);
bool smj_stage_events(void);
template <typename... KArgs, typename... Args>
static inline void smj_launch_on(SmjCtx *c, cudaStream_t st, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (smj_pdl_enabled() && !c->capturing) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline void smj_launch(SmjCtx *c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    // measured at the 10M x 10M config: eager launches 0.514 -> 0.497 ms per step with the attribute, the replayed graph
    // 0.489 -> 0.495 ms (kernel nodes of a graph already chain without a host round trip), so captures leave it off
    cfg.numAttrs = (smj_pdl_enabled() && !c->capturing) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through KERNEL_CHECK's cudaGetLastError
}
#endif

// ------------------------------------------------------------------ select (smj_select.cu)
// Scratch layout helpers: all tile-status arrays must be zero before the launch.
size_t smj_select_num_tiles(int64_t n);
// Evaluates `cell[sel_col] > sel_val` (or everything when select_all) over n rows of a row-major table,
// writes order-preserving (flipped key << 32 | rowid_base + row) pairs, the selected count, and (optionally)
// the 4 x 256 digit histogram of the flipped keys.
int smj_launch_select_pairs(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val,
                            int select_all, int key_col, u32 rowid_base, u64 *d_pairs,
                            u64 *d_tmp /*n pairs of scratch, or null: look-back kernel*/,
                            u64 *d_status /*[smj_select_num_tiles(n)], zero*/, u32 *d_tile_counter /*zero*/,
                            u32 *d_hist /*[4*256] zero, may be null*/, u64 *d_count /*out*/);

// Sort plan of one table, device-resident (lives in the zeroed scratch arena): the select kernel accumulates the
// surviving keys' range in kmin_inv (= ~min, so that zero means "none yet") and kmax; the scan kernel derives kmin and
// npass = ceil(bits(kmax - kmin) / 8).  The radix passes sort digits of (key - kmin) and passes >= npass exit at once.
struct SmjSortPlan { u32 kmin, npass, kmin_inv, kmax; };

// Semi-join reduction between select and sort (smj_run, zip mode): a row whose key no surviving row of the OTHER table
// carries can never be joined, and dropping it changes neither the matches nor their order, so it need not be sorted.
// Each table's select kernel sets one bit per surviving key in its own bitmap; the table that is selected second probes
// the first one's bitmap inside its select kernel, the first one is filtered by a pass over its pair slots.  A hash
// collision only keeps a row that the join then rejects.
struct SmjBloom {
    u32 *set = nullptr;           // bitmap this launch inserts into (or null)
    const u32 *probe = nullptr;   // bitmap of the other table to probe (or null)
    u32 shift = 0;                // 32 - log2(bits)
    u64 *sel_count = nullptr;     // += rows that passed the predicate (before the probe)
    u64 *kept_count = nullptr;    // += rows that also passed the probe (probing launches only)
};

// smj_run's select stage over both tables (smj_select.cu): see smj_launch_select_plan2.
struct SmjSelectJob {
    const int32_t *d_in; int64_t n; int cols, sel_col; int64_t sel_val; int select_all, key_col;
    u64 *buf[2];          // ping / pong pair arrays (n pairs each); the dense pairs go to buf[plan->npass & 1]
    u64 *slots;           // n pairs of scratch: the select kernel's per-tile slots
    u64 *d_status;        // [smj_select_num_tiles(n)] zeroed words: tile offsets + tile counts
    u32 *d_hist;          // [4 * 256] zero: digit histogram of key - kmin, passes < npass
    u64 *d_count;         // out: pairs that go on to the sort (survivors that passed the semi-join filter)
    SmjSortPlan *plan;    // zeroed
    u64 *d_sel_count;     // zeroed; out: rows that passed the predicate (smj_stats_t.rows_selected)
    u64 *d_kept_count;    // zeroed; out: of those, the rows whose key bit was set in the other table's bitmap
    const u64 *n_dev;     // device-resident row count (may be null); n is then the upper bound the buffers were sized for
    SmjWait wait;         // the table's select kernel starts with this wait (smj_dist.cu: the table is still arriving)
    const int32_t *const *d_in_ind;   // device cell holding the table's address (null: d_in is used): graph replay on new tables
    int32_t *store;       // dense row store of store_max_rows rows (null: none); *use_store (zeroed) = the device's decision
    u32 *use_store;
    u64 store_max_rows;
};
// returns 1 (nothing launched) when a table cannot take the TMA path
int smj_launch_select_plan2(SmjCtx *c, const SmjSelectJob job[2]);
bool smj_select_takes_tma(const int32_t *d_in, int cols);   // 16-byte aligned table of at most 32 columns
size_t smj_bloom_bytes(int64_t n0, int64_t n1);   // WS_BLOOM bytes the call above uses (0: no semi-join filter)

// ------------------------------------------------------------------ radix sort (smj_radix.cu)
size_t smj_radix_num_tiles(u32 n);
size_t smj_radix_status_words(u32 n);          // per pass
int smj_launch_radix_hist(SmjCtx *c, const u64 *d_pairs, u32 n, u32 *d_hist /*[4*256], zero*/);
int smj_launch_radix_scan(SmjCtx *c, const u32 *d_hist, u32 *d_bases /*[4*256]*/);
int smj_launch_radix_pass(SmjCtx *c, const u64 *d_in, u64 *d_out, const u64 *d_n /*device count or null*/, u32 n_max, int pass,
                          const u32 *d_bases_pass, u32 *d_status /*zero*/, u32 *d_status_next /*zeroed for the next pass*/,
                          u32 *d_tile_counter /*zero*/);
// Stable LSD sort by the key half of the first *d_n (<= n_max; d_n null: exactly n_max) pairs of buf_a, result in
// buf_a (four passes).  d_hist: the 4x256 digit histogram of those pairs (device).  No host synchronisation.
int smj_radix_sort_pairs(SmjCtx *c, u64 *buf_a, u64 *buf_b, const u64 *d_n, u32 n_max, const u32 *d_hist,
                         u32 *d_scratch /*smj_radix_scratch_bytes(n_max), zero*/);
// the same for one or two pair arrays in the same four launches (their tiles share the CTAs of each launch)
// d_plan (may be null, or hold nulls): device sort plans.  With a plan, problem i's unsorted pairs are expected in
// (plan->npass & 1 ? buf_b : buf_a)[i], d_hist[i] holds the digits of key - plan->kmin, and only npass passes do work;
// the result is in buf_a[i] either way.
int smj_radix_sort_pairs_n(SmjCtx *c, int nprob, u64 *const *buf_a, u64 *const *buf_b, const u64 *const *d_n, const u32 *n_max,
                           const u32 *const *d_hist, u32 *const *d_scratch, const SmjSortPlan *const *d_plan = nullptr);
size_t smj_radix_scratch_bytes(u32 n);         // bases + status(4 passes) + counters, all zero-initialised by caller

// ------------------------------------------------------------------ gather (smj_gather.cu)
// out[i][:] = in[lo32(pairs[i])][:]
int smj_launch_gather_rows(SmjCtx *c, const u64 *d_pairs, int64_t m, const int32_t *d_in, int cols, int32_t *d_out);
// two-source variant for smj_merge: rowid < split -> a[rowid], else b[rowid - split]
int smj_launch_gather_rows2(SmjCtx *c, const u64 *d_pairs, int64_t m, const int32_t *d_a, const int32_t *d_b,
                            u32 split, int cols, int32_t *d_out);

// ------------------------------------------------------------------ merge (smj_merge.cu)
size_t smj_merge_num_tiles(u64 total);
// merge two key-sorted pair arrays, a before b on equal keys
int smj_launch_merge_pairs(SmjCtx *c, const u64 *d_a, u32 na, const u64 *d_b, u32 nb, u64 *d_out, u32 *d_part);

// ------------------------------------------------------------------ join (smj_join.cu)
size_t smj_join_num_tiles(u64 total);
// zip mode: matches[i] = (left rowid, right rowid) in (key, left position) order; *d_count = number of matches.
// many mode with d_matches == nullptr: *d_count = sum over left rows of the right run length (count only).
// d_counts: device {m1, m2} (u64 each) or null, in which case m1_max / m2_max are the exact sizes.
// With tiles = smj_join_num_tiles(m1_max + m2_max): d_part holds 2*(tiles+1) u32, d_tile_count tiles u32 (zeroed by the
// caller), d_tile_off tiles u64, d_matches tiles * smj_join_tile_size() entries (tile t's matches start at t * tile size).
// zip mode: *d_count = number of matches (written by the scan); many mode: *d_count += total pair count (count only).
size_t smj_join_tile_size(void);
// d_tile_off needs smj_join_scan_blocks(tiles) more u64 words behind its `tiles` entries (block sums of the many-CTA scan)
size_t smj_join_scan_blocks(size_t tiles);
u32 smj_join_scan_chunk(void);
// d_dense (zip mode, may be null): the matches compacted in result order, min(m1_max, m2_max) entries of capacity.
int smj_launch_join_match(SmjCtx *c, const u64 *d_l, const u64 *d_r, const u64 *d_counts, u32 m1_max, u32 m2_max, int mode,
                          u32 *d_part, u32 *d_tile_count, u64 *d_tile_off, uint2 *d_matches, uint2 *d_dense, u64 *d_count);
// many-to-many: with mode = SMJ_JOIN_MANY, d_matches (if not null) receives one (first right position, right run length)
// entry per left element; smj_launch_join_many_expand turns them into the dense match list of `total` pairs.
size_t smj_join_many_scratch_bytes(u32 m1);
int smj_launch_join_many_expand(SmjCtx *c, const u64 *d_l, const u64 *d_r, const uint2 *d_runs, u32 m1, u64 total, char *d_scratch,
                                uint2 *d_dense);
// d_nj: device match count or null (then nj_max is exact)
// d_out_indirect (may be null): device cell holding the output pointer, read instead of d_out (graph replay)
// d_use_store / d_store1 / d_store2 (may be null): device flags [2] saying which table's row ids are dense indices into
// its row store (smj_select.cu, plan_compact_kernel), and the stores
int smj_launch_join_materialize(SmjCtx *c, const uint2 *d_dense, const u64 *d_nj, int64_t nj_max, const int32_t *d_t1, int c1,
                                const int32_t *d_t2, int c2, int key2, int32_t *d_out, int32_t *const *d_out_indirect = nullptr,
                                const u32 *d_use_store = nullptr, const int32_t *d_store1 = nullptr, const int32_t *d_store2 = nullptr,
                                const int32_t *const *d_tables_indirect = nullptr /* device [2]: the tables' addresses */);

// ------------------------------------------------------------------ key-range partitioning of rows (smj_partition.cu)
#define SMJ_MAX_G 8
struct SmjPartScratch { u32 *counts, *off32; u64 *blocksum, *bucket_total, *bucket_start; size_t tiles, ctas, bytes; };
// where the exchange kernel puts this rank's bucket b: behind base[b], row0[b] (+ the segment's offset inside the bucket) rows in
struct SmjPartitionDst {
    int32_t *base[SMJ_MAX_G];
    const u64 *row0;             // device [SMJ_MAX_G] (null: 0)
    const u32 *skip;             // device flag (may be null): non-zero = store nothing (receive overflow, see smj_dist.cu)
};
bool smj_partition_supported(const int32_t *d_in, int cols);
size_t smj_partition_tiles(int64_t n, int cols);
u32 smj_partition_tile_rows(int cols);
SmjPartScratch smj_partition_scratch(char *base, int64_t n, int cols);
size_t smj_partition_scratch_bytes(int64_t n, int cols);
int smj_launch_sample_rows(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int key_col, int S,
                           u32 *d_samples);
int smj_launch_select_partition(SmjCtx *c, cudaStream_t st, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val,
                                int key_col, const u32 *d_splitters, int G, int32_t *d_slots, char *d_scratch);
int smj_launch_partition_exchange(SmjCtx *c, cudaStream_t st, int64_t n, int cols, int sel_val_none, int G, const int32_t *d_slots,
                                  char *d_scratch, const SmjPartitionDst &D);
int smj_launch_splitters(SmjCtx *c, const u32 *d_samples, int n_samples, int G, u32 *d_splitters);

// ------------------------------------------------------------------ synth (smj_synth.cu)
int smj_launch_synth(SmjCtx *c, int32_t *d_out, int64_t row0, int64_t rows, int64_t total_rows, int cols,
                     int key_col, u64 seed, int kind, int64_t key_domain);
