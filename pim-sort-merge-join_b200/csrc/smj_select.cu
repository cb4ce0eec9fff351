// smj_select.cu -- predicate evaluation fused with warp-ballot/popc stream compaction into (key, row id) pairs, plus what
// the sort needs from the same pass: the surviving keys' range (sort plan), their digit histograms, and the semi-join
// key bitmaps that let smj_run drop rows the other table cannot match before they are sorted.
//
// Replaces the DPU select kernel (sort-merge-join/select.c:24-39 filter, :42-61 tasklet handshake prefix,
// :125-185 block loop) and cpu_app.c:81-112.  Same contract: keep rows with cell[sel_col] > sel_val
// (strict, signed), original order preserved.  Instead of compacting whole rows in place it emits
// (flipped key << 32 | row id) pairs: the payload stays where it is and is gathered once, after the join
// (late materialisation), so the select pass reads the table exactly once and writes 8 B per survivor.
//
// Kernels: select_pairs_kernel (decoupled look-back, any table shape; the fallback), select_tma_kernel (TMA-fed,
// warp-specialised, per-tile slots; the one smj_run uses), tile_scan / select_compact (slots -> dense pairs),
// plan_scan / bloom_filter / plan_compact (smj_run: both tables per launch, sort plans, semi-join filter).
#include "smj_internal.h"
#include "smj_dev.cuh"
#include <stdlib.h>

namespace {

constexpr int SEL_THREADS = 256;
constexpr int SEL_IPT = 8;                       // rows per thread per tile
constexpr int SEL_TILE = SEL_THREADS * SEL_IPT;  // 2048 rows
constexpr int SEL_WARPS = SEL_THREADS / 32;
static_assert((SEL_IPT * SEL_WARPS) % 32 == 0, "count matrix is scanned 32 entries at a time");

template <bool HIST>
__global__ void __launch_bounds__(SEL_THREADS)
select_pairs_kernel(const int32_t *__restrict__ in, int64_t n, int cols, int sel_col, int32_t sel_val, int select_all,
                    int key_col, u32 rowid_base, u64 *__restrict__ pairs, u64 *status, u32 *tile_counter, u32 *hist,
                    u64 *count, u32 num_tiles, u32 *err)
{
    __shared__ u32 s_cnt[SEL_IPT * SEL_WARPS];   // [j][warp] survivors of one warp-row, then its exclusive prefix
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    __shared__ u32 s_hist[HIST ? SMJ_KEY_PASSES * SMJ_RADIX : 1];

    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u32 lt = lanemask_lt();
    if (HIST)
        for (u32 i = tid; i < SMJ_KEY_PASSES * SMJ_RADIX; i += SEL_THREADS) s_hist[i] = 0;

    // Persistent CTAs pull tiles through an atomic ticket: tile t's predecessors have all been started,
    // which is what the look-back needs, and the histogram is flushed once per CTA instead of once per tile.
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= num_tiles) break;
        const int64_t tile_base = (int64_t)tile * SEL_TILE;

        int32_t key[SEL_IPT];
        u32 rank[SEL_IPT];
        u32 passmask = 0;
#pragma unroll
        for (int j = 0; j < SEL_IPT; j++) {
            const int64_t row = tile_base + j * SEL_THREADS + tid;
            int32_t sv = 0, kv = 0;
            const bool valid = row < n;
            if (valid) {
                const int32_t *r = in + row * cols;
                sv = __ldg(r + sel_col);
                kv = (key_col == sel_col) ? sv : __ldg(r + key_col);
            }
            const bool pass = valid && (select_all || sv > sel_val);
            key[j] = kv;
            const u32 b = __ballot_sync(FULL_MASK, pass);
            if (lane == 0) s_cnt[j * SEL_WARPS + w] = __popc(b);
            rank[j] = __popc(b & lt);
            passmask |= (pass ? 1u : 0u) << j;
        }
        __syncthreads();
        if (w == 0) {
            u32 run = 0;
#pragma unroll
            for (int b0 = 0; b0 < SEL_IPT * SEL_WARPS; b0 += 32) {
                const u32 v = s_cnt[b0 + lane];
                const u32 inc = warp_incl_scan(v);
                s_cnt[b0 + lane] = run + inc - v;
                run += __shfl_sync(FULL_MASK, inc, 31);
            }
            const u64 excl = lookback_warp(status, tile, (u64)run, err, SMJ_ERR_SPIN_SELECT);
            if (lane == 0) {
                s_base = excl;
                if (tile == num_tiles - 1) *count = excl + run;
            }
        }
        __syncthreads();
        const u64 base = s_base;
#pragma unroll
        for (int j = 0; j < SEL_IPT; j++) {
            if ((passmask >> j) & 1u) {
                const u64 pos = base + s_cnt[j * SEL_WARPS + w] + rank[j];
                const u64 p = make_pair(key[j], rowid_base + (u32)(tile_base + j * SEL_THREADS + tid));
                pairs[pos] = p;
                if (HIST) {
                    const u32 k = pair_key(p);
#pragma unroll
                    for (int d = 0; d < SMJ_KEY_PASSES; d++)
                        atomicAdd(&s_hist[d * SMJ_RADIX + ((k >> (d * SMJ_RADIX_BITS)) & (SMJ_RADIX - 1))], 1u);
                }
            }
        }
    }
    if (HIST) {
        __syncthreads();
        for (u32 i = tid; i < SMJ_KEY_PASSES * SMJ_RADIX; i += SEL_THREADS) {
            const u32 v = s_hist[i];
            if (v) atomicAdd(&hist[i], v);
        }
    }
}


// ------------------------------------------------------------------ TMA-fed, warp-specialised variant (the one smj_run uses)
// A tile is a CONTIGUOUS byte range of the row-major table, so one 1-D bulk copy (cp.async.bulk, SASS UBLKCP) per
// tile moves it into shared memory without touching registers.  Nine warps per CTA:
//   warp 8    producer: waits for a free ring stage and issues the bulk copy of the CTA's next tile (full/empty mbarriers);
//   warps 0-7 compute: predicate, ballot/popc ranks, digit histogram, and the tile's survivors written compacted
//             WITHIN the tile's own slot of a temporary array (slot t starts at t * tile_rows), plus the tile count.
// No CTA ever waits for another one: the first two versions of this kernel resolved each tile's global offset with
// a decoupled look-back inside the streaming loop, and ncu showed 70 % (single role) / 35 % (dedicated scan warp) of
// all stall samples waiting for that look-back's L2 round trips, with DRAM at 19 % of peak.  Here the offsets come
// from a separate scan over the tile counts (tile_scan_kernel, microseconds) and a compaction copy of the survivors
// (8 B per survivor, L2-resident at the 10M-row config), and the table scan itself is a pure stream.
#ifndef SMJ_SEL_STAGES
#define SMJ_SEL_STAGES 2
#endif
#ifndef SMJ_SEL_CTAS
#define SMJ_SEL_CTAS 2
#endif
constexpr int SEL_STAGES = SMJ_SEL_STAGES;
constexpr int SEL_STAGE_BYTES = 32768;
constexpr int SEL_MAX_COLS_TMA = SEL_STAGE_BYTES / 4 / SEL_THREADS;   // 32 columns: at least one row per thread
constexpr int SELW_THREADS = SEL_THREADS + 32;                         // + producer warp
constexpr size_t SELW_SMEM = (size_t)SEL_STAGES * SEL_STAGE_BYTES;

// MODE 0: pairs only; 1: + 4 x 256 digit histogram of the surviving keys; 2: + min / max of the surviving keys into
// *plan (the sort then runs only the passes the key RANGE needs, smj_radix.cu; the histogram of (key - min) digits is
// built by the compaction copy, once the minimum is known).
template <int MODE, bool PROBE = false>
__global__ void __launch_bounds__(SELW_THREADS, PROBE ? 2 : 3)   // shared memory allows 3 CTAs per SM; the launch uses SMJ_SEL_CTAS
select_tma_kernel(const int32_t *__restrict__ in, int64_t n, int cols, int ipt, int sel_col, int32_t sel_val, int select_all,
                  int key_col, u32 rowid_base, u64 *__restrict__ slots, u32 *__restrict__ tile_count, u32 *hist, u32 num_tiles,
                  SmjSortPlan *plan, SmjBloom bloom, const u64 *__restrict__ n_dev, const int32_t *const *__restrict__ in_ind)
{
    constexpr bool HIST = MODE == 1;
    PDL_ENTER();
    if (in_ind) in = *in_ind;   // the table's address comes from a device cell: the replayed pipeline graph serves any table of this shape
    if (n_dev) {   // the table is a receive buffer: its fill is device-resident, n / num_tiles are the upper bounds
        const u64 v = *n_dev;
        if (v < (u64)n) {
            n = (int64_t)v;
            num_tiles = (u32)((v + (u64)ipt * SEL_THREADS - 1) / ((u64)ipt * SEL_THREADS));
        }
    }
    extern __shared__ __align__(128) unsigned char sel_smem[];      // stage ring
    __shared__ __align__(8) u64 s_full[SEL_STAGES], s_empty[SEL_STAGES];
    __shared__ u32 s_cnt[2][SEL_IPT * SEL_WARPS];
    __shared__ u32 s_hist[HIST ? SMJ_KEY_PASSES * SMJ_RADIX : 1];

    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u32 tile_rows = (u32)ipt * SEL_THREADS;

    if (HIST)
        for (u32 i = tid; i < SMJ_KEY_PASSES * SMJ_RADIX; i += SELW_THREADS) s_hist[i] = 0;
    if (tid == 0) {
        for (int st = 0; st < SEL_STAGES; st++) { mbar_init(&s_full[st], 1); mbar_init(&s_empty[st], SEL_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    if (w == SEL_WARPS) {
        // ------------------------------------------------ producer (one lane); tiles are dealt round-robin
        if (lane != 0) return;
        const size_t row_bytes = (size_t)cols * 4;
        const u64 stream_policy = l2_policy_evict_first();
        u32 stage = 0, parity = 0;
        for (u32 t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            mbar_wait(&s_empty[stage], parity ^ 1u);      // passes at once the first time round the ring
            const int64_t row0 = (int64_t)t * tile_rows;
            const int64_t rows = (n - row0 < (int64_t)tile_rows) ? (n - row0) : (int64_t)tile_rows;
            const u32 bytes = (u32)(rows * row_bytes);
            const u32 b16 = bytes & ~15u;
            const unsigned char *src = reinterpret_cast<const unsigned char *>(in) + (size_t)row0 * row_bytes;
            unsigned char *dst = sel_smem + (size_t)stage * SEL_STAGE_BYTES;
            for (u32 b = b16; b < bytes; b += 4)          // ragged tail of the last tile (< 16 bytes)
                *reinterpret_cast<int32_t *>(dst + b) = *reinterpret_cast<const int32_t *>(src + b);
            if (b16) {
                mbar_expect_tx(&s_full[stage], b16);
                bulk_g2s_hint(dst, src, b16, &s_full[stage], stream_policy);
            } else {
                mbar_arrive(&s_full[stage]);
            }
            if (++stage == SEL_STAGES) { stage = 0; parity ^= 1u; }
        }
        return;
    }

    // ---------------------------------------------------- compute warps
    const u32 lt = lanemask_lt();
    const u64 keep_policy = l2_policy_evict_last();   // bitmap words
    u32 tmin = 0xffffffffu, tmax = 0u;   // MODE 2: this thread's surviving flipped keys
    u32 nsel = 0;                        // MODE 2: rows of this thread that passed the predicate (before the semi-join probe)
    u32 nkept = 0;                       // PROBE: pairs written (thread SEL_THREADS - 1 sums the tile totals)

    // what a thread keeps of a tile once the ring stage has been released
    struct TileRegs { int32_t key[SEL_IPT]; u32 probe[PROBE ? SEL_IPT : 1]; u32 passmask; u32 tile; };

    // predicate on the tile in ring stage `stage`; with PROBE also issues the loads of the other table's bitmap words
    auto load_tile = [&](u32 tile, u32 stage, TileRegs &T) {
        const int64_t tile_base = (int64_t)tile * tile_rows;
        const u32 rows_valid = (u32)((n - tile_base < (int64_t)tile_rows) ? (n - tile_base) : (int64_t)tile_rows);
        const int32_t *s_rows = reinterpret_cast<const int32_t *>(sel_smem + (size_t)stage * SEL_STAGE_BYTES);
        T.tile = tile;
        T.passmask = 0;
#pragma unroll
        for (int j = 0; j < SEL_IPT; j++) {
            if (j < ipt) {
                const u32 row = j * SEL_THREADS + tid;
                int32_t sv = 0, kv = 0;
                const bool valid = row < rows_valid;
                if (valid) {
                    sv = s_rows[row * cols + sel_col];
                    kv = (key_col == sel_col) ? sv : s_rows[row * cols + key_col];
                }
                const bool pass = valid && (select_all || sv > sel_val);
                T.key[j] = kv;
                T.passmask |= (pass ? 1u : 0u) << j;
                if (PROBE) T.probe[j] = pass ? ld_nc_hint(bloom.probe + (bloom_hash((u32)kv ^ 0x80000000u, bloom.shift) >> 5), keep_policy) : 0u;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);   // this warp holds its cells in registers: stage can be refilled
        if (MODE == 2) nsel += __popc(T.passmask);
    };

    // ranks, the tile's count, the survivors written compacted into the tile's slot (`it` = tiles finished so far)
    auto finish_tile = [&](const TileRegs &T, u32 it) {
        const int64_t tile_base = (int64_t)T.tile * tile_rows;
        u32 *cntbuf = s_cnt[it & 1u];
        u32 rank[SEL_IPT];
        u32 passmask = T.passmask;
#pragma unroll
        for (int j = 0; j < SEL_IPT; j++) {
            if (j < ipt) {
                if (PROBE) {   // no row of the other table hashes to this key's bit: the row cannot be joined
                    const u32 h = bloom_hash((u32)T.key[j] ^ 0x80000000u, bloom.shift);
                    if (!((T.probe[j] >> (h & 31u)) & 1u)) passmask &= ~(1u << j);
                }
                const u32 b = __ballot_sync(FULL_MASK, (passmask >> j) & 1u);
                if (lane == 0) cntbuf[j * SEL_WARPS + w] = __popc(b);
                rank[j] = __popc(b & lt);
            } else if (lane == 0) {
                cntbuf[j * SEL_WARPS + w] = 0;
            }
        }
        named_bar_sync(1, SEL_THREADS);                // counts complete (double-buffered: one barrier per tile)

        // every warp scans the 64 (row group, warp) counts itself: no second barrier, no broadcast
        const u32 v0 = cntbuf[2 * lane], v1 = cntbuf[2 * lane + 1];
        const u32 inc = warp_incl_scan(v0 + v1);
        const u32 ex0 = inc - (v0 + v1);               // exclusive prefix of entry 2*lane; entry 2*lane+1 adds v0
        if (tid == SEL_THREADS - 1) { tile_count[T.tile] = inc; nkept += inc; }   // lane 31: the tile total
        u64 *dst = slots + (size_t)tile_base;
#pragma unroll
        for (int j = 0; j < SEL_IPT; j++) {
            const u32 e = (u32)j * SEL_WARPS + w;      // warp-uniform entry index
            u32 off = __shfl_sync(FULL_MASK, ex0, e >> 1);
            const u32 add = __shfl_sync(FULL_MASK, v0, e >> 1);
            if (e & 1u) off += add;
            if ((passmask >> j) & 1u) {
                const u64 p = make_pair(T.key[j], rowid_base + (u32)(tile_base + j * SEL_THREADS + tid));
                dst[off + rank[j]] = p;
                if (HIST) {
                    const u32 k = pair_key(p);
#pragma unroll
                    for (int d = 0; d < SMJ_KEY_PASSES; d++)
                        atomicAdd(&s_hist[d * SMJ_RADIX + ((k >> (d * SMJ_RADIX_BITS)) & (SMJ_RADIX - 1))], 1u);
                }
                if (MODE == 2) {
                    const u32 k = pair_key(p);
                    tmin = k < tmin ? k : tmin;
                    tmax = k > tmax ? k : tmax;
                    if (bloom.set) {
                        const u32 h = bloom_hash(k, bloom.shift);
                        red_or_hint(bloom.set + (h >> 5), 1u << (h & 31u), keep_policy);
                    }
                }
            }
        }
    };

    {
        u32 stage = 0, parity = 0, it = 0;
        TileRegs prev;
        bool have_prev = false;
        for (u32 tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&s_full[stage], parity);
            TileRegs cur;
            load_tile(tile, stage, cur);
            if (PROBE) {
                // software pipeline: tile i's bitmap words travel while tile i-1 is ranked and written and tile i+1 is
                // waited for, so the probe's L2 round trip is off the per-tile critical path
                if (have_prev) finish_tile(prev, it++);
                prev = cur;
                have_prev = true;
            } else {
                finish_tile(cur, it++);
            }
            if (++stage == SEL_STAGES) { stage = 0; parity ^= 1u; }
        }
        if (PROBE && have_prev) finish_tile(prev, it++);
    }
    if (HIST) {
        named_bar_sync(1, SEL_THREADS);
        for (u32 i = tid; i < SMJ_KEY_PASSES * SMJ_RADIX; i += SEL_THREADS) {
            const u32 v = s_hist[i];
            if (v) atomicAdd(&hist[i], v);
        }
    }
    if (MODE == 2) {
        // the arena is zeroed, so the minimum is accumulated as the maximum of the complement
        const u32 wmin = __reduce_min_sync(FULL_MASK, tmin), wmax = __reduce_max_sync(FULL_MASK, tmax);
        if (lane == 0 && wmin <= wmax) {
            atomicMax(&plan->kmin_inv, ~wmin);
            atomicMax(&plan->kmax, wmax);
        }
        const u32 wsel = __reduce_add_sync(FULL_MASK, nsel);
        if (lane == 0 && wsel) atomicAdd(bloom.sel_count, (u64)wsel);
        if (PROBE && tid == SEL_THREADS - 1 && nkept) atomicAdd(bloom.kept_count, (u64)nkept);
    }
}

// A table that is a receive buffer may still be filling while the pipeline in front of it runs: another stream's exchange
// sets the table's arrival cell to the step's sequence number when every rank's rows have landed (smj_dist.cu).  This
// one-warp kernel polls the cell and only then lets the stream go on.  Both cells sit at fixed device addresses, so the
// kernel can be part of the replayed pipeline graph.  It must be a kernel of its own, and a small one: a select kernel
// that spun in its prologue held every SM (together with the CTAs its programmatic-launch successors parked behind it),
// and the exchange kernels it was waiting for could not be scheduled at all (seen at 250M x 50M rows per GPU).
// No griddepcontrol.launch_dependents here: the kernels behind it must not become resident before the wait is over.
__global__ void __launch_bounds__(32) smj_wait_kernel(const SmjWait wait)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) {
        const u64 want = *reinterpret_cast<const volatile u64 *>(wait.seq);
        unsigned long long t0 = 0;
        for (u32 spins = 0;; spins++) {
            u64 v;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(wait.flag) : "memory");
            if (v >= want) break;
            if ((spins & 1023u) == 1023u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 30ull * 1000 * 1000 * 1000) { atomicExch(wait.err, 5u); break; }
            }
        }
        __threadfence();
    }
}

// Exclusive scan of per-tile counts (one CTA): offsets[t] = sum of counts[0..t), *total = sum of all.
// Each thread owns a contiguous chunk, so the block-level part is one scan of 1024 partial sums.
constexpr int TS_THREADS = SCAN1_THREADS;
__global__ void __launch_bounds__(TS_THREADS) tile_scan_kernel(const u32 *__restrict__ counts, u32 num_tiles, u64 *offsets, u64 *total)
{
    __shared__ u32 s_stage[SCAN1_STAGE];
    __shared__ u64 s_w[TS_THREADS / 32];
    PDL_ENTER();
    const u64 t = scan1_counts(counts, num_tiles, offsets, s_stage, s_w);
    if (threadIdx.x == 0) *total = t;
}

// The same over many CTAs (scan_large_* in smj_dev.cuh) for the stage entry points (smj_select / smj_sort / smj_merge / smj_join on
// tables of more than SCAN1_STAGE tiles, ~16 M rows): the one-CTA scan walks its chunks from global memory beyond that size.
__global__ void __launch_bounds__(TS_THREADS) tile_blocksum_kernel(const u32 *__restrict__ counts, u32 num_tiles, u32 chunk, u64 *blocksum)
{
    __shared__ u64 s_w[TS_THREADS / 32];
    PDL_ENTER();
    scan_large_blocksum(counts, num_tiles, chunk, blockIdx.x, blocksum, s_w);
}
__global__ void __launch_bounds__(TS_THREADS) tile_apply_kernel(const u32 *__restrict__ counts, u32 num_tiles, u32 chunk, u64 *offsets,
                                                                const u64 *__restrict__ blocksum, u64 *total)
{
    __shared__ u32 s_stage[SCAN1_STAGE];
    __shared__ u64 s_w[TS_THREADS / 32];
    PDL_ENTER();
    const u64 t = scan_large_apply(counts, num_tiles, chunk, blockIdx.x, offsets, blocksum, s_stage, s_w);
    if (blockIdx.x + 1 == gridDim.x && threadIdx.x == 0) *total = t;   // the last block's running total is the grand total
}

// pairs[offsets[t] + i] = slots[t * tile_rows + i], i < counts[t]: the survivors in table order, contiguous.
// One WARP per tile: a tile is ~1-2 K pairs, and with 64 tiles in flight per SM instead of 8 (one per CTA) the
// dependent count/offset -> data round trips overlap.
__global__ void __launch_bounds__(256)
select_compact_kernel(const u64 *__restrict__ slots, const u32 *__restrict__ counts, const u64 *__restrict__ offsets, u32 num_tiles,
                      u32 tile_rows, u64 *__restrict__ pairs)
{
    PDL_ENTER();
    const u32 lane = threadIdx.x & 31u;
    const u32 warps = gridDim.x * (blockDim.x >> 5);
    for (u32 t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < num_tiles; t += warps) {
        const u32 cnt = counts[t];
        const u64 *src = slots + (size_t)t * tile_rows;
        u64 *dst = pairs + offsets[t];
        for (u32 i = lane; i < cnt; i += 32) dst[i] = src[i];
    }
}

// ---- smj_run's variant of the two kernels above: both tables per launch, and the sort plan.
// plan_scan_kernel (one CTA per table): tile offsets and the survivor count as in tile_scan_kernel, then the sort plan
// from the key range the select kernel left in *plan: LSD passes run over (key - kmin), so a table whose surviving keys
// span R values needs ceil(log2(R) / 8) passes instead of four (the reference draws its keys from [1, 3n],
// data/generate_data.py:9).  full_passes forces four passes from key 0 (A/B measurements, SMJ_FULL_PASSES=1).
// use_store / store_max_rows: the dense row store (see plan_compact_kernel) is used when the pairs that go on to the sort
// are at most store_max_rows (0: never) -- decided here, on the device, where the count is first known
struct PlanScanJob { const u32 *counts; u32 num_tiles; u64 *offsets; u64 *total; SmjSortPlan *plan; u32 *use_store; u64 store_max_rows; };
struct PlanScanArgs { PlanScanJob t[2]; int full_passes; };

__global__ void __launch_bounds__(TS_THREADS) plan_scan_kernel(const PlanScanArgs A)
{
    __shared__ u32 s_stage[SCAN1_STAGE];
    __shared__ u64 s_w[TS_THREADS / 32];
    PDL_ENTER();
    const PlanScanJob J = blockIdx.x ? A.t[1] : A.t[0];   // (a dynamic index would copy the parameters to local memory)
    u32 kmin_inv = 0, kmax = 0;
    if (threadIdx.x == 0) { kmin_inv = J.plan->kmin_inv; kmax = J.plan->kmax; }   // in flight during the scan
    const u64 total = scan1_counts(J.counts, J.num_tiles, J.offsets, s_stage, s_w);
    if (threadIdx.x == 0) {
        *J.total = total;
        u32 kmin = ~kmin_inv, npass = 0;
        if (total >= 2 && kmax > kmin) npass = (32u - (u32)__clz(kmax - kmin) + 7u) / 8u;
        if (total == 0) kmin = 0;
        if (A.full_passes) { kmin = 0; npass = SMJ_KEY_PASSES; }
        J.plan->kmin = kmin;
        J.plan->npass = npass;
        if (J.use_store) *J.use_store = (J.store_max_rows && total <= J.store_max_rows) ? 1u : 0u;
    }
}

// The same over many CTAs (tables with more tiles than one CTA stages at once: scan_large_* in smj_dev.cuh).  Blocks
// [0, nb0) belong to table 1, the others to table 2; a table without tiles has no block and keeps the zeroed plan and total.
struct PlanScanLargeArgs { PlanScanJob t[2]; u64 *blocksum[2]; u32 nb0, chunk; int full_passes; };

__global__ void __launch_bounds__(TS_THREADS) plan_blocksum_kernel(const PlanScanLargeArgs A)
{
    __shared__ u64 s_w[TS_THREADS / 32];
    PDL_ENTER();
    const bool tb = blockIdx.x >= A.nb0;
    const PlanScanJob J = tb ? A.t[1] : A.t[0];
    scan_large_blocksum(J.counts, J.num_tiles, A.chunk, tb ? blockIdx.x - A.nb0 : blockIdx.x, tb ? A.blocksum[1] : A.blocksum[0], s_w);
}

__global__ void __launch_bounds__(TS_THREADS) plan_apply_kernel(const PlanScanLargeArgs A)
{
    __shared__ u32 s_stage[SCAN1_STAGE];
    __shared__ u64 s_w[TS_THREADS / 32];
    PDL_ENTER();
    const bool tb = blockIdx.x >= A.nb0;
    const PlanScanJob J = tb ? A.t[1] : A.t[0];
    const u32 b = tb ? blockIdx.x - A.nb0 : blockIdx.x;
    const u32 nb = (J.num_tiles + A.chunk - 1) / A.chunk;
    const u64 total = scan_large_apply(J.counts, J.num_tiles, A.chunk, b, J.offsets, tb ? A.blocksum[1] : A.blocksum[0], s_stage, s_w);
    if (b + 1 == nb && threadIdx.x == 0) {   // the table's last block holds the grand total: survivor count and sort plan
        *J.total = total;
        const u32 kmax = J.plan->kmax;
        u32 kmin = ~J.plan->kmin_inv, npass = 0;
        if (total >= 2 && kmax > kmin) npass = (32u - (u32)__clz(kmax - kmin) + 7u) / 8u;
        if (total == 0) kmin = 0;
        if (A.full_passes) { kmin = 0; npass = SMJ_KEY_PASSES; }
        J.plan->kmin = kmin;
        J.plan->npass = npass;
        if (J.use_store) *J.use_store = (J.store_max_rows && total <= J.store_max_rows) ? 1u : 0u;
    }
}

// The compaction copy of both tables in one launch (one warp per tile, table 1's tiles first), fused with the digit
// histograms of (key - kmin) for the passes the plan runs.  The dense pairs land in buf[npass & 1]: pass p reads
// buf[(npass - p) & 1] and writes the other one, so the sorted pairs always end in buf[0].
struct PlanCompactJob { const u64 *slots; const u32 *counts; const u64 *offsets; u32 num_tiles, tile_rows; u64 *buf[2];
                        const SmjSortPlan *plan; u32 *hist; };
struct PlanCompactArgs { PlanCompactJob t[2]; };
#ifndef SMJ_PC_UNROLL
#define SMJ_PC_UNROLL 8
#endif
constexpr int PC_UNROLL = SMJ_PC_UNROLL;

__global__ void __launch_bounds__(256) plan_compact_kernel(const PlanCompactArgs A)
{
    __shared__ u32 s_hist[2][SMJ_KEY_PASSES * SMJ_RADIX];
    PDL_ENTER();
    const u32 tid = threadIdx.x, lane = tid & 31u;
    for (u32 i = tid; i < 2 * SMJ_KEY_PASSES * SMJ_RADIX; i += 256) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const u32 tiles0 = A.t[0].num_tiles, all_tiles = tiles0 + A.t[1].num_tiles;
    const u32 warps = gridDim.x * 8u;
    for (u32 g = blockIdx.x * 8u + (tid >> 5); g < all_tiles; g += warps) {
        const bool tb = g >= tiles0;
        const u32 t = tb ? g - tiles0 : g;
        const SmjSortPlan *plan = tb ? A.t[1].plan : A.t[0].plan;
        const u32 kmin = plan->kmin, npass = plan->npass;
        const u32 cnt = (tb ? A.t[1].counts : A.t[0].counts)[t];
        const u64 *src = (tb ? A.t[1].slots : A.t[0].slots) + (size_t)t * (tb ? A.t[1].tile_rows : A.t[0].tile_rows);
        const u64 off = (tb ? A.t[1].offsets : A.t[0].offsets)[t];
        u64 *dst = (tb ? ((npass & 1u) ? A.t[1].buf[1] : A.t[1].buf[0]) : ((npass & 1u) ? A.t[0].buf[1] : A.t[0].buf[0])) + off;
        u32 *h = s_hist[tb];
        // eight independent 256-byte loads in flight per warp before the first dependent store: the first version
        // (load, store, atomics per iteration, no __restrict__) exposed one DRAM round trip per 32 pairs
        for (u32 i0 = 0; i0 < cnt; i0 += 32 * PC_UNROLL) {
            u64 v[PC_UNROLL];
#pragma unroll
            for (int k = 0; k < PC_UNROLL; k++) {
                const u32 i = i0 + k * 32 + lane;
                v[k] = i < cnt ? __ldg(src + i) : 0ull;
            }
#pragma unroll
            for (int k = 0; k < PC_UNROLL; k++) {
                const u32 i = i0 + k * 32 + lane;
                if (i < cnt) {
                    dst[i] = v[k];
                    const u32 d = pair_key(v[k]) - kmin;
                    if (npass > 0) atomicAdd(&h[d & 255u], 1u);
                    if (npass > 1) atomicAdd(&h[SMJ_RADIX + ((d >> 8) & 255u)], 1u);
                    if (npass > 2) atomicAdd(&h[2 * SMJ_RADIX + ((d >> 16) & 255u)], 1u);
                    if (npass > 3) atomicAdd(&h[3 * SMJ_RADIX + (d >> 24)], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (u32 i = tid; i < 2 * SMJ_KEY_PASSES * SMJ_RADIX; i += 256) {
        const u32 v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&(i >= SMJ_KEY_PASSES * SMJ_RADIX ? A.t[1].hist : A.t[0].hist)[i % (SMJ_KEY_PASSES * SMJ_RADIX)], v);
    }
}

// Dense row store.  When few rows survive select + semi-join filter (use_store, decided on the device by the scan kernel:
// the survivors fit SMJ_ROWSTORE_MB and are at most a quarter of the table), each survivor's payload row is copied into
// store[dense index] and its pair's payload becomes that index -- still ascending with the row id, so the stable sort keeps
// the reference's order.  The join then gathers payload from a store of a few tens of MB that stays in the 126 MB L2,
// instead of 16-byte rows scattered over the whole table (ncu, round 1: 395 MB of DRAM traffic for 100 MB of algorithmic
// bytes in join_materialize_kernel, every gather pulling a whole 64..128-byte line; here the rows are read once more, in
// ascending order, and only those the filter kept).
struct RowStoreJob { u64 *buf[2]; const SmjSortPlan *plan; const u64 *count; const int32_t *table; int32_t *store; int cols; const u32 *use_store; };
struct RowStoreArgs { RowStoreJob t[2]; };
constexpr int RSK_ILP = 4;

__global__ void __launch_bounds__(256) rowstore_kernel(const RowStoreArgs A)
{
    PDL_ENTER();
    for (int tb = 0; tb < 2; tb++) {
        const RowStoreJob J = tb ? A.t[1] : A.t[0];
        if (!J.store || !J.use_store || !*J.use_store) continue;
        const u64 m = *J.count;
        u64 *pairs = (J.plan->npass & 1u) ? J.buf[1] : J.buf[0];      // where plan_compact_kernel put the dense pairs
        const int cols = J.cols;
        const bool vec = (cols % 4 == 0) && (((reinterpret_cast<uintptr_t>(J.table) | reinterpret_cast<uintptr_t>(J.store)) & 15) == 0);
        const u64 stride = (u64)gridDim.x * blockDim.x;
        for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < m; i0 += stride * RSK_ILP) {
            u64 p[RSK_ILP];
#pragma unroll
            for (int k = 0; k < RSK_ILP; k++) { const u64 i = i0 + k * stride; p[k] = i < m ? pairs[i] : 0ull; }
            if (vec && cols == 4) {       // one 16-byte row per pair: RSK_ILP independent row loads in flight per thread
                int4 r[RSK_ILP];
#pragma unroll
                for (int k = 0; k < RSK_ILP; k++) { const u64 i = i0 + k * stride; if (i < m) r[k] = __ldg(reinterpret_cast<const int4 *>(J.table) + pair_row(p[k])); }
#pragma unroll
                for (int k = 0; k < RSK_ILP; k++) { const u64 i = i0 + k * stride; if (i < m) reinterpret_cast<int4 *>(J.store)[i] = r[k]; }
            } else {
#pragma unroll
                for (int k = 0; k < RSK_ILP; k++) {
                    const u64 i = i0 + k * stride;
                    if (i < m) {
                        const int32_t *rs = J.table + (size_t)pair_row(p[k]) * cols;
                        int32_t *rd = J.store + (size_t)i * cols;
                        if (vec) for (int q = 0; q < cols / 4; q++) reinterpret_cast<int4 *>(rd)[q] = __ldg(reinterpret_cast<const int4 *>(rs) + q);
                        else for (int q = 0; q < cols; q++) rd[q] = __ldg(rs + q);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < RSK_ILP; k++) { const u64 i = i0 + k * stride; if (i < m) pairs[i] = (p[k] & 0xffffffff00000000ull) | (u64)(u32)i; }
        }
    }
}

// Semi-join filter of the table that was selected FIRST (its partner's bitmap did not exist yet): one warp per tile
// compacts the tile's pair slot in place, keeping the pairs whose key bit is set in the other table's bitmap.  A chunk
// of 8 x 32 pairs and its 8 bitmap words are in registers before the first store, and the write cursor never passes
// the chunk's first element, so no pair is overwritten before it has been read.
// The pass is skipped when more than 3/4 of the OTHER table's selected rows found their key in this table's bitmap:
// the two key sets then largely coincide and the filter would keep almost everything.
struct BloomFilterArgs { u64 *slots; u32 *counts; u32 num_tiles, tile_rows; const u32 *probe; u32 shift; const u64 *other_sel, *other_kept; };

__global__ void __launch_bounds__(256) bloom_filter_kernel(const BloomFilterArgs A)
{
    PDL_ENTER();
    if (*A.other_kept * 4 > *A.other_sel * 3) return;
    const u32 lane = threadIdx.x & 31u;
    const u32 lt = lanemask_lt();
    const u32 warps = gridDim.x * 8u;
    const u64 keep_policy = l2_policy_evict_last();
    for (u32 t = blockIdx.x * 8u + (threadIdx.x >> 5); t < A.num_tiles; t += warps) {
        const u32 cnt = A.counts[t];
        u64 *slot = A.slots + (size_t)t * A.tile_rows;
        u32 wpos = 0;
        // (prefetching the next chunk's pairs into a second register set was measured: 33 -> 37 us, fewer warps fit)
        for (u32 i0 = 0; i0 < cnt; i0 += 32 * PC_UNROLL) {
            u64 v[PC_UNROLL];
            u32 word[PC_UNROLL];
#pragma unroll
            for (int k = 0; k < PC_UNROLL; k++) {
                const u32 i = i0 + k * 32 + lane;
                v[k] = i < cnt ? slot[i] : 0ull;
            }
#pragma unroll
            for (int k = 0; k < PC_UNROLL; k++) {
                const u32 i = i0 + k * 32 + lane;
                word[k] = i < cnt ? ld_nc_hint(A.probe + (bloom_hash(pair_key(v[k]), A.shift) >> 5), keep_policy) : 0u;
            }
            __syncwarp();   // every lane holds its part of the chunk before any lane stores
#pragma unroll
            for (int k = 0; k < PC_UNROLL; k++) {
                const u32 i = i0 + k * 32 + lane;
                const bool keep = i < cnt && ((word[k] >> (bloom_hash(pair_key(v[k]), A.shift) & 31u)) & 1u);
                const u32 b = __ballot_sync(FULL_MASK, keep);
                if (keep) slot[wpos + __popc(b & lt)] = v[k];
                wpos += __popc(b);
            }
            __syncwarp();
        }
        if (lane == 0) A.counts[t] = wpos;
    }
}

}  // namespace

// rows per thread per tile for the TMA path: the largest power-of-two count (<= 8) whose tile fits one stage
static int select_ipt(int cols)
{
    int ipt = SEL_IPT;
    while (ipt > 1 && (size_t)ipt * SEL_THREADS * cols * 4 > SEL_STAGE_BYTES) ipt >>= 1;
    return ipt;
}
static bool select_use_tma(const int32_t *d_in, int cols)
{
    return cols <= SEL_MAX_COLS_TMA && (((uintptr_t)d_in) & 15) == 0;
}
bool smj_select_takes_tma(const int32_t *d_in, int cols) { return select_use_tma(d_in, cols); }

static int select_set_attrs(SmjCtx *c)   // function attributes are per device
{
    CUDA_TRY(cudaFuncSetAttribute(select_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELW_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(select_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELW_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(select_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELW_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(select_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SELW_SMEM));
    c->select_attr_set = true;
    return SMJ_OK;
}

// tiles one CTA scans; beyond it the many-CTA scans take over (tests force them at small sizes with a small chunk)
static u32 scan_chunk(void)
{
    static const u32 chunk = [] {
        const char *e = getenv("SMJ_SCAN_CHUNK");
        const long v = e ? atol(e) : 0;
        return (u32)((v >= 32 && v <= SCAN1_STAGE) ? v : SCAN1_STAGE);
    }();
    return chunk;
}

// scratch words (u64) per table: look-back status (fallback kernel) or tile offsets + tile counts (TMA path);
// the smallest tile either path uses is SEL_THREADS rows
size_t smj_select_num_tiles(int64_t n) { return 2 * (size_t)((n + SEL_THREADS - 1) / SEL_THREADS) + 2; }

int smj_launch_select_pairs(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val,
                            int select_all, int key_col, u32 rowid_base, u64 *d_pairs, u64 *d_tmp, u64 *d_status,
                            u32 *d_tile_counter, u32 *d_hist, u64 *d_count)
{
    if (n <= 0) return SMJ_OK;   // caller zeroed *d_count
    // cell is int32: cell > val is always true below INT32_MIN and never true from INT32_MAX up (cpu_app.c:88
    // compares int64 T against the int64 knob, with int32-valued cells).
    if (sel_val < (int64_t)INT32_MIN) select_all = 1;
    if (!select_all && sel_val >= (int64_t)INT32_MAX) return SMJ_OK;   // nothing passes; *d_count stays 0
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    if (d_tmp && select_use_tma(d_in, cols)) {
        const int ipt = select_ipt(cols);
        const int64_t tile_rows = (int64_t)ipt * SEL_THREADS;
        const u32 tiles = (u32)((n + tile_rows - 1) / tile_rows);
        const size_t smem = SELW_SMEM;
        if (!c->select_attr_set) {
            SMJ_TRY(select_set_attrs(c));
        }
        u64 *d_offsets = d_status;                                   // [tiles]
        u32 *d_counts = reinterpret_cast<u32 *>(d_status + tiles);   // [tiles]
        const u32 grid = tiles < (u32)(sms * SMJ_SEL_CTAS) ? tiles : (u32)(sms * SMJ_SEL_CTAS);
        if (d_hist)
            select_tma_kernel<1><<<grid, SELW_THREADS, smem, c->stream>>>(d_in, n, cols, ipt, sel_col, (int32_t)sel_val, select_all,
                                                                         key_col, rowid_base, d_tmp, d_counts, d_hist, tiles, nullptr, SmjBloom(), nullptr, nullptr);
        else
            select_tma_kernel<0><<<grid, SELW_THREADS, smem, c->stream>>>(d_in, n, cols, ipt, sel_col, (int32_t)sel_val, select_all,
                                                                         key_col, rowid_base, d_tmp, d_counts, nullptr, tiles, nullptr, SmjBloom(), nullptr, nullptr);
        KERNEL_CHECK(c);
        const u32 chunk = scan_chunk();
        if (tiles > chunk) {
            // spare words of the status region, behind [offsets u64 x tiles][counts u32 x tiles] (smj_select_num_tiles)
            u64 *d_blocksum = d_status + tiles + (tiles + 1) / 2;
            const u32 nb = (tiles + chunk - 1) / chunk;
            tile_blocksum_kernel<<<nb, TS_THREADS, 0, c->stream>>>(d_counts, tiles, chunk, d_blocksum);
            KERNEL_CHECK(c);
            tile_apply_kernel<<<nb, TS_THREADS, 0, c->stream>>>(d_counts, tiles, chunk, d_offsets, d_blocksum, d_count);
            KERNEL_CHECK(c);
        } else {
            tile_scan_kernel<<<1, TS_THREADS, 0, c->stream>>>(d_counts, tiles, d_offsets, d_count);
            KERNEL_CHECK(c);
        }
        const u32 cgrid = (tiles + 7) / 8 < (u32)(sms * 8) ? (tiles + 7) / 8 : (u32)(sms * 8);
        select_compact_kernel<<<cgrid, 256, 0, c->stream>>>(d_tmp, d_counts, d_offsets, tiles, (u32)tile_rows, d_pairs);
        KERNEL_CHECK(c);
        return SMJ_OK;
    }
    const u32 tiles = (u32)((n + SEL_TILE - 1) / SEL_TILE);
    const u32 grid = tiles < (u32)(sms * 8) ? tiles : (u32)(sms * 8);
    if (d_hist)
        select_pairs_kernel<true><<<grid, SEL_THREADS, 0, c->stream>>>(d_in, n, cols, sel_col, (int32_t)sel_val,
                                                                      select_all, key_col, rowid_base, d_pairs, d_status,
                                                                      d_tile_counter, d_hist, d_count, tiles, c->d_err);
    else
        select_pairs_kernel<false><<<grid, SEL_THREADS, 0, c->stream>>>(d_in, n, cols, sel_col, (int32_t)sel_val,
                                                                       select_all, key_col, rowid_base, d_pairs, d_status,
                                                                       d_tile_counter, nullptr, d_count, tiles, c->d_err);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

#ifndef SMJ_BLOOM_BITS_PER_ROW
#define SMJ_BLOOM_BITS_PER_ROW 4
#endif
// log2 of the bits per semi-join bitmap, or 0 when the filter is not used for these table sizes
static int bloom_log2_bits(int64_t n0, int64_t n1)
{
    const int64_t n_first = n0 <= n1 ? n0 : n1;
    if (n0 <= 0 || n1 <= 0 || 4 * n_first > (1ll << 29)) return 0;
    int lb = 16;
    while (lb < 28 && (1ll << lb) < SMJ_BLOOM_BITS_PER_ROW * n_first) lb++;
    return lb;
}
size_t smj_bloom_bytes(int64_t n0, int64_t n1)
{
    const int lb = bloom_log2_bits(n0, n1);
    return lb ? 2 * ((size_t)1 << (lb - 3)) : 0;
}

// smj_run's select stage, both tables: select (key min / max instead of histograms) per table, ONE scan launch (tile
// offsets, survivor counts, sort plans), ONE compaction launch (dense pairs into the buffer the plan names + digit
// histograms of key - kmin).  Returns 1 without launching anything when a table cannot take the TMA path (the caller
// then uses smj_launch_select_pairs and a plan-less four-pass sort).
int smj_launch_select_plan2(SmjCtx *c, const SmjSelectJob job[2])
{
    for (int t = 0; t < 2; t++)
        if (job[t].n > 0 && !select_use_tma(job[t].d_in, job[t].cols)) return 1;
    static const int full_passes = (getenv("SMJ_FULL_PASSES") && atoi(getenv("SMJ_FULL_PASSES")) != 0) ? 1 : 0;
    static const bool semijoin_on = !(getenv("SMJ_SEMIJOIN") && atoi(getenv("SMJ_SEMIJOIN")) == 0);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    if (!c->select_attr_set) SMJ_TRY(select_set_attrs(c));

    // Semi-join bitmaps: the smaller table goes first (it is the one that needs the extra filter pass), 4 bits per input
    // row of it, at most 2^28 bits (32 MB) per table so that both bitmaps stay L2-resident while they are probed.
    const int first = job[0].n <= job[1].n ? 0 : 1, second = first ^ 1;
    const int lb = bloom_log2_bits(job[0].n, job[1].n);
    const bool semijoin = semijoin_on && lb > 0;
    u32 *bm[2] = {nullptr, nullptr};
    if (semijoin) {
        const size_t words = (size_t)1 << (lb - 5);
        u32 *base = (u32 *)smj_ws(c, WS_BLOOM, 2 * words * 4);   // (already sized by the caller: smj_bloom_bytes)
        if (!base) return SMJ_ENOMEM;
        bm[0] = base; bm[1] = base + words;
        CUDA_TRY(cudaMemsetAsync(base, 0, 2 * words * 4, c->stream));
    }

    PlanScanArgs SA = {};
    PlanCompactArgs CA = {};
    RowStoreArgs RA = {};
    bool any_store = false;
    SA.full_passes = full_passes;
    u32 all_tiles = 0;
    u32 tiles_of[2] = {0, 0};
    u32 *counts_of[2] = {nullptr, nullptr};
    int64_t tile_rows_of[2] = {0, 0};
    const int order[2] = {first, second};
    for (int o = 0; o < 2; o++) {
        const int t = order[o];
        const SmjSelectJob &J = job[t];
        int select_all = J.select_all;
        int64_t sel_val = J.sel_val;
        // cell is int32: cell > val is always true below INT32_MIN and never true from INT32_MAX up (cpu_app.c:88)
        if (sel_val < (int64_t)INT32_MIN) select_all = 1;
        const bool none = J.n <= 0 || (!select_all && sel_val >= (int64_t)INT32_MAX);
        const int ipt = select_ipt(J.cols);
        const int64_t tile_rows = (int64_t)ipt * SEL_THREADS;
        const u32 tiles = none ? 0u : (u32)((J.n + tile_rows - 1) / tile_rows);
        u64 *d_offsets = J.d_status;                                   // [tiles]
        u32 *d_counts = reinterpret_cast<u32 *>(J.d_status + tiles);   // [tiles]
        if (tiles) {
            SmjBloom B;
            B.sel_count = J.d_sel_count;
            B.kept_count = J.d_kept_count;
            if (semijoin) {
                B.set = bm[t];
                B.probe = (t == second) ? bm[first] : nullptr;
                B.shift = 32u - (u32)lb;
            }
            const u32 grid = tiles < (u32)(sms * SMJ_SEL_CTAS) ? tiles : (u32)(sms * SMJ_SEL_CTAS);
            if (J.wait.flag) {   // the table is still arriving on another stream (smj_dist.cu): a one-warp kernel waits for it
                SmjWait W = J.wait;
                W.err = c->d_err;
                smj_launch(c, smj_wait_kernel, 1, 32, 0, W);
                KERNEL_CHECK(c);
            }
            if (B.probe)
                smj_launch(c, select_tma_kernel<2, true>, grid, SELW_THREADS, SELW_SMEM, J.d_in, J.n, J.cols, ipt, J.sel_col, (int32_t)sel_val, select_all,
                           J.key_col, 0u, J.slots, d_counts, (u32 *)nullptr, tiles, J.plan, B, J.n_dev, J.d_in_ind);
            else
                smj_launch(c, select_tma_kernel<2, false>, grid, SELW_THREADS, SELW_SMEM, J.d_in, J.n, J.cols, ipt, J.sel_col, (int32_t)sel_val, select_all,
                           J.key_col, 0u, J.slots, d_counts, (u32 *)nullptr, tiles, J.plan, B, J.n_dev, J.d_in_ind);
            KERNEL_CHECK(c);
        }
        tiles_of[t] = tiles; counts_of[t] = d_counts; tile_rows_of[t] = tile_rows;
        SA.t[t] = {d_counts, tiles, d_offsets, J.d_count, J.plan, J.use_store, J.store ? J.store_max_rows : 0};
        CA.t[t] = {J.slots, d_counts, d_offsets, tiles, (u32)tile_rows, {J.buf[0], J.buf[1]}, J.plan, J.d_hist};
        RA.t[t] = {{J.buf[0], J.buf[1]}, J.plan, J.d_count, J.d_in, J.store, J.cols, J.use_store};
        any_store = any_store || (J.store != nullptr && tiles > 0);
        all_tiles += tiles;
    }
    if (semijoin && tiles_of[first]) {
        BloomFilterArgs FA = {job[first].slots, counts_of[first], tiles_of[first], (u32)tile_rows_of[first], bm[second], 32u - (u32)lb,
                              job[second].d_sel_count, job[second].d_kept_count};
        const u32 fgrid = (tiles_of[first] + 7) / 8 < (u32)(sms * 8) ? (tiles_of[first] + 7) / 8 : (u32)(sms * 8);
        smj_launch(c, bloom_filter_kernel, fgrid, 256, 0, FA);
        KERNEL_CHECK(c);
    }
    // tile offsets, survivor counts and sort plans: one CTA per table, or many when a table has more tiles than one CTA stages
    const u32 chunk = scan_chunk();
    if (tiles_of[0] > chunk || tiles_of[1] > chunk) {
        PlanScanLargeArgs LA = {};
        LA.chunk = chunk;
        LA.full_passes = full_passes;
        u32 nb[2];
        for (int t = 0; t < 2; t++) {
            LA.t[t] = SA.t[t];
            nb[t] = (tiles_of[t] + chunk - 1) / chunk;
            // spare words of the table's status region, behind [offsets u64 x tiles][counts u32 x tiles] (smj_select_num_tiles)
            LA.blocksum[t] = job[t].d_status + tiles_of[t] + (tiles_of[t] + 1) / 2;
        }
        LA.nb0 = nb[0];
        smj_launch(c, plan_blocksum_kernel, nb[0] + nb[1], TS_THREADS, 0, LA);
        KERNEL_CHECK(c);
        smj_launch(c, plan_apply_kernel, nb[0] + nb[1], TS_THREADS, 0, LA);
        KERNEL_CHECK(c);
    } else {
        smj_launch(c, plan_scan_kernel, 2, TS_THREADS, 0, SA);
        KERNEL_CHECK(c);
    }
    if (all_tiles) {
        const u32 cgrid = (all_tiles + 7) / 8 < (u32)(sms * 8) ? (all_tiles + 7) / 8 : (u32)(sms * 8);
        smj_launch(c, plan_compact_kernel, cgrid, 256, 0, CA);
        KERNEL_CHECK(c);
    }
    if (any_store) {
        smj_launch(c, rowstore_kernel, (u32)(sms * 8), 256, 0, RA);
        KERNEL_CHECK(c);
    }
    return SMJ_OK;
}

// Loads this file's pipeline kernels on the current device.  CUDA loads a kernel lazily at its first launch, and that load can
// wait for other GPUs' running kernels when peer access is enabled; a process that drives several GPUs (smj_dist.cu) must
// not meet such a load while another rank's kernel spins on this rank's flags, so it loads everything up front.
void smj_preload_select(void)
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, select_pairs_kernel<true>);
    cudaFuncGetAttributes(&a, select_pairs_kernel<false>);
    cudaFuncGetAttributes(&a, select_tma_kernel<0>);
    cudaFuncGetAttributes(&a, select_tma_kernel<1>);
    cudaFuncGetAttributes(&a, select_tma_kernel<2>);
    cudaFuncGetAttributes(&a, select_tma_kernel<2, true>);
    cudaFuncGetAttributes(&a, tile_scan_kernel);
    cudaFuncGetAttributes(&a, tile_blocksum_kernel);
    cudaFuncGetAttributes(&a, tile_apply_kernel);
    cudaFuncGetAttributes(&a, select_compact_kernel);
    cudaFuncGetAttributes(&a, plan_scan_kernel);
    cudaFuncGetAttributes(&a, plan_blocksum_kernel);
    cudaFuncGetAttributes(&a, plan_apply_kernel);
    cudaFuncGetAttributes(&a, plan_compact_kernel);
    cudaFuncGetAttributes(&a, bloom_filter_kernel);
    cudaFuncGetAttributes(&a, rowstore_kernel);
    cudaFuncGetAttributes(&a, smj_wait_kernel);
    cudaGetLastError();
}
