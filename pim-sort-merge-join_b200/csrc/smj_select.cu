// smj_select.cu -- predicate evaluation fused with warp-ballot/popc stream compaction, a single-pass
// decoupled look-back scan and (optionally) the 4 x 256 radix digit histogram of the surviving keys.
//
// Replaces the DPU select kernel (sort-merge-join/select.c:24-39 filter, :42-61 tasklet handshake prefix,
// :125-185 block loop) and cpu_app.c:81-112.  Same contract: keep rows with cell[sel_col] > sel_val
// (strict, signed), original order preserved.  Instead of compacting whole rows in place it emits
// (flipped key << 32 | row id) pairs: the payload stays where it is and is gathered once, after the sort
// (late materialisation), so the select pass reads the table exactly once and writes 8 B per survivor.
#include "smj_internal.h"
#include "smj_dev.cuh"

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

}  // namespace

size_t smj_select_num_tiles(int64_t n) { return (size_t)((n + SEL_TILE - 1) / SEL_TILE); }

int smj_launch_select_pairs(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val,
                            int select_all, int key_col, u32 rowid_base, u64 *d_pairs, u64 *d_status,
                            u32 *d_tile_counter, u32 *d_hist, u64 *d_count)
{
    if (n <= 0) return SMJ_OK;   // caller zeroed *d_count
    // cell is int32: cell > val is always true below INT32_MIN and never true from INT32_MAX up (cpu_app.c:88
    // compares int64 T against the int64 knob, with int32-valued cells).
    if (sel_val < (int64_t)INT32_MIN) select_all = 1;
    if (!select_all && sel_val >= (int64_t)INT32_MAX) return SMJ_OK;   // nothing passes; *d_count stays 0
    const u32 tiles = (u32)smj_select_num_tiles(n);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
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
