// smj_partition.cu -- select fused with key-range partitioning of whole rows, for the multi-GPU path.
//
// Replaces, for G GPUs, what the reference does on the host between its stages: rows are split over devices by
// position (sort-merge-join/app.c:155-218) and, for the join, re-split by key range with a binary search per chunk
// (app.c:585-633).  Here every rank filters ITS row block (cell > val, cpu_app.c:81-112) and, in the same streaming
// pass, groups the surviving ROWS by destination rank (bucket b holds keys in [splitter[b-1], splitter[b])), keeping
// original row order inside each bucket -- so the buckets can be sent as they are (one grouped ncclSend/ncclRecv),
// the receiver's runs arrive in source-rank order, and a stable sort by key on the receiver reproduces the
// reference's stable order.  All passes are sequential streams; no row is gathered at random before it travels.
//
//   select_partition_kernel : TMA-fed like select_tma_kernel (1 producer warp, 8 compute warps).  Per tile: predicate,
//                             bucket id (<= 7 compares), per-(bucket,row group,warp) ballot counts, one 512-entry
//                             scan, rows written to the tile's own slot ordered by bucket, G counts per tile.
//   partition_blocksum / partition_offsets_kernel: offsets of every (bucket, tile) segment + bucket totals, many CTAs.
//   partition_exchange_kernel: one warp per tile routes the tile's survivors to their buckets' destinations.
//   sample_rows_kernel      : regular row samples (predicate applied) from which the splitters are derived.
#include "smj_internal.h"
#include "smj_dev.cuh"
#include <stdlib.h>

namespace {

constexpr int PT_THREADS = 256;
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_IPT = 8;
constexpr int PT_STAGES = 2;
constexpr int PT_STAGE_BYTES = 32768;
constexpr int PT_MAX_G = SMJ_MAX_G;
constexpr int PTW_THREADS = PT_THREADS + 32;
constexpr int PT_ENTRIES = PT_MAX_G * PT_IPT * PT_WARPS;   // 512 = 2 per compute thread
constexpr size_t PTW_SMEM = (size_t)PT_STAGES * PT_STAGE_BYTES;
constexpr int PT_MAX_COLS = PT_STAGE_BYTES / 4 / PT_THREADS;   // 32

__global__ void __launch_bounds__(PTW_THREADS, 3)
select_partition_kernel(const int32_t *__restrict__ in, int64_t n, int cols, int ipt, int sel_col, int32_t sel_val, int select_all,
                        int key_col, const u32 *__restrict__ splitters, int G, int32_t *__restrict__ slots,
                        u32 *__restrict__ tile_counts /*[tiles][PT_MAX_G]*/, u32 num_tiles)
{
    extern __shared__ __align__(128) unsigned char pt_smem[];
    __shared__ __align__(8) u64 s_full[PT_STAGES], s_empty[PT_STAGES];
    __shared__ u32 s_cnt[PT_ENTRIES], s_off[PT_ENTRIES + 1];
    __shared__ u32 s_wtot[PT_WARPS];
    __shared__ u32 s_split[PT_MAX_G];

    PDL_ENTER();
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u32 tile_rows = (u32)ipt * PT_THREADS;
    if (tid < (u32)PT_MAX_G) s_split[tid] = (tid < (u32)(G - 1)) ? splitters[tid] : 0xffffffffu;
    if (tid == 0) {
        for (int st = 0; st < PT_STAGES; st++) { mbar_init(&s_full[st], 1); mbar_init(&s_empty[st], PT_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    if (w == PT_WARPS) {   // ---------------- producer (one lane)
        if (lane != 0) return;
        const size_t row_bytes = (size_t)cols * 4;
        const u64 stream_policy = l2_policy_evict_first();
        u32 stage = 0, parity = 0;
        for (u32 t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            mbar_wait(&s_empty[stage], parity ^ 1u);
            const int64_t row0 = (int64_t)t * tile_rows;
            const int64_t rows = (n - row0 < (int64_t)tile_rows) ? (n - row0) : (int64_t)tile_rows;
            const u32 bytes = (u32)(rows * row_bytes);
            const u32 b16 = bytes & ~15u;
            const unsigned char *src = reinterpret_cast<const unsigned char *>(in) + (size_t)row0 * row_bytes;
            unsigned char *dst = pt_smem + (size_t)stage * PT_STAGE_BYTES;
            for (u32 b = b16; b < bytes; b += 4)
                *reinterpret_cast<int32_t *>(dst + b) = *reinterpret_cast<const int32_t *>(src + b);
            if (b16) {
                mbar_expect_tx(&s_full[stage], b16);
                bulk_g2s_hint(dst, src, b16, &s_full[stage], stream_policy);
            } else {
                mbar_arrive(&s_full[stage]);
            }
            if (++stage == PT_STAGES) { stage = 0; parity ^= 1u; }
        }
        return;
    }

    // ---------------------------------------- compute warps
    const u32 lt = lanemask_lt();
    const bool vec = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(slots) & 15) == 0);
    u32 sp[PT_MAX_G - 1];
#pragma unroll
    for (int q = 0; q < PT_MAX_G - 1; q++) sp[q] = s_split[q];
    u32 stage = 0, parity = 0;
    for (u32 tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&s_full[stage], parity);
        const int64_t tile_base = (int64_t)tile * tile_rows;
        const u32 rows_valid = (u32)((n - tile_base < (int64_t)tile_rows) ? (n - tile_base) : (int64_t)tile_rows);
        const int32_t *s_rows = reinterpret_cast<const int32_t *>(pt_smem + (size_t)stage * PT_STAGE_BYTES);

        u32 rank[PT_IPT / 4] = {};   // 8 bits per row group: position among the rows of the same (bucket, row group, warp)
        u32 bucket = 0;              // 4 bits per row group; PT_MAX_G = dropped
        for (u32 i = tid; i < (u32)PT_ENTRIES; i += PT_THREADS) s_cnt[i] = 0;
        named_bar_sync(1, PT_THREADS);
#pragma unroll
        for (int j = 0; j < PT_IPT; j++) {
            u32 myb = PT_MAX_G;
            if (j < ipt) {
                const u32 row = j * PT_THREADS + tid;
                bool pass = false;
                u32 b = 0;
                if (row < rows_valid) {
                    const int32_t sv = s_rows[row * cols + sel_col];
                    const int32_t kv = (key_col == sel_col) ? sv : s_rows[row * cols + key_col];
                    pass = select_all || sv > sel_val;
                    const u32 fk = (u32)kv ^ 0x80000000u;
#pragma unroll
                    for (int q = 0; q < PT_MAX_G - 1; q++) b += (fk >= sp[q]) ? 1u : 0u;   // unused splitters are 0xffffffff
                    if (b > (u32)(G - 1)) b = (u32)(G - 1);   // a key equal to 0xffffffff
                }
                // ranks inside every bucket: one ballot per bucket (independent of each other; the first version ran one
                // 64-bit shuffle scan per row group, a dependent chain of five steps, and the kernel took 72 us per 160 MB table)
                u32 mine_mask = 0, cnt_lane = 0;
#pragma unroll
                for (int q = 0; q < PT_MAX_G; q++) {
                    if (q < G) {
                        const u32 m = __ballot_sync(FULL_MASK, pass && b == (u32)q);
                        if (b == (u32)q) mine_mask = m;
                        if (lane == (u32)q) cnt_lane = __popc(m);
                    }
                }
                if (pass) { rank[j >> 2] |= (u32)__popc(mine_mask & lt) << (8 * (j & 3)); myb = b; }
                if (lane < (u32)G && cnt_lane) s_cnt[(lane * PT_IPT + j) * PT_WARPS + w] = cnt_lane;
            }
            bucket |= myb << (4 * j);
        }
        named_bar_sync(1, PT_THREADS);
        {   // exclusive scan of the 512 counts (bucket-major, then row group, then warp == row order inside a bucket)
            const u32 v0 = s_cnt[2 * tid], v1 = s_cnt[2 * tid + 1];
            const u32 inc = warp_incl_scan(v0 + v1);
            if (lane == 31) s_wtot[w] = inc;
            named_bar_sync(1, PT_THREADS);
            u32 base = inc - (v0 + v1);
            for (u32 ww = 0; ww < w; ww++) base += s_wtot[ww];
            s_off[2 * tid] = base;
            s_off[2 * tid + 1] = base + v0;
            if (tid == PT_THREADS - 1) s_off[PT_ENTRIES] = base + v0 + v1;
        }
        named_bar_sync(1, PT_THREADS);
        if (tid < (u32)PT_MAX_G) {
            const u32 lo = s_off[tid * PT_IPT * PT_WARPS], hi = s_off[(tid + 1) * PT_IPT * PT_WARPS];
            tile_counts[(size_t)tile * PT_MAX_G + tid] = hi - lo;
        }
        int32_t *dst_tile = slots + (size_t)tile_base * cols;
#pragma unroll
        for (int j = 0; j < PT_IPT; j++) {
            const u32 myb = (bucket >> (4 * j)) & 15u;
            if (myb < (u32)PT_MAX_G) {
                const u32 row = j * PT_THREADS + tid;
                const u32 pos = s_off[(myb * PT_IPT + j) * PT_WARPS + w] + ((rank[j >> 2] >> (8 * (j & 3))) & 255u);
                const int32_t *src = s_rows + row * cols;
                int32_t *dst = dst_tile + (size_t)pos * cols;
                if (vec) {
                    for (int q = 0; q < cols / 4; q++)
                        reinterpret_cast<int4 *>(dst)[q] = reinterpret_cast<const int4 *>(src)[q];
                } else {
                    for (int q = 0; q < cols; q++) dst[q] = src[q];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);   // rows copied out: the stage can be refilled
        if (++stage == PT_STAGES) { stage = 0; parity ^= 1u; }
        // s_cnt / s_off are rewritten only after the next tile's first named barrier, which every warp reaches after
        // finishing the reads above
    }
}

// ---- offsets of every (tile, bucket) segment, over many CTAs.
// off32[t][b] = rows of bucket b in this rank's tiles before tile t (segments of one bucket in tile order = original row
// order); bucket_total[b] = rows of bucket b on this rank; bucket_start[b] = first row of bucket b in a bucket-major send
// buffer, bucket_start[G] = all survivors.  Two launches of ceil(tiles / PS_CHUNK) CTAs: CTA c first sums the eight bucket
// counts of its PS_CHUNK tiles; then every CTA adds up the sums of the CTAs before it itself (a few hundred words at the
// 2 B-row config, no CTA waits for another one) and scans its own chunk.  The first version was one CTA walking
// tiles x G counts from global memory (11 us at 10 M rows; the same pattern cost 610 us per 488 K counts in the select scan).
constexpr int PS_THREADS = 256;
constexpr int PS_TPT = 8;                          // consecutive tiles per thread
constexpr int PS_CHUNK = PS_THREADS * PS_TPT;      // 2048 tiles per CTA

__device__ __forceinline__ void ps_load_tiles(const u32 *__restrict__ tile_counts, u32 num_tiles, u32 t0, u32 (&v)[PS_TPT][PT_MAX_G])
{
#pragma unroll
    for (int k = 0; k < PS_TPT; k++) {
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (t0 + k < num_tiles) {
            a = __ldg(reinterpret_cast<const uint4 *>(tile_counts + (size_t)(t0 + k) * PT_MAX_G));
            b = __ldg(reinterpret_cast<const uint4 *>(tile_counts + (size_t)(t0 + k) * PT_MAX_G) + 1);
        }
        v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w;
        v[k][4] = b.x; v[k][5] = b.y; v[k][6] = b.z; v[k][7] = b.w;
    }
}

__global__ void __launch_bounds__(PS_THREADS)
partition_blocksum_kernel(const u32 *__restrict__ tile_counts, u32 num_tiles, u64 *__restrict__ blocksum /*[ctas][8]*/)
{
    __shared__ u32 s_w[PS_THREADS / 32][PT_MAX_G];
    PDL_ENTER();
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    u32 v[PS_TPT][PT_MAX_G];
    ps_load_tiles(tile_counts, num_tiles, blockIdx.x * PS_CHUNK + tid * PS_TPT, v);
#pragma unroll
    for (int b = 0; b < PT_MAX_G; b++) {
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < PS_TPT; k++) s += v[k][b];
        s = __reduce_add_sync(FULL_MASK, s);
        if (lane == 0) s_w[w][b] = s;
    }
    __syncthreads();
    if (tid < (u32)PT_MAX_G) {
        u64 t = 0;
#pragma unroll
        for (int ww = 0; ww < PS_THREADS / 32; ww++) t += s_w[ww][tid];
        blocksum[(size_t)blockIdx.x * PT_MAX_G + tid] = t;
    }
}

__global__ void __launch_bounds__(PS_THREADS)
partition_offsets_kernel(const u32 *__restrict__ tile_counts, u32 num_tiles, int G, const u64 *__restrict__ blocksum, u32 *__restrict__ off32,
                         u64 *__restrict__ bucket_total /*[8]*/, u64 *__restrict__ bucket_start /*[G+1]*/)
{
    __shared__ u64 s_red[PS_THREADS / 32][2][PT_MAX_G];
    __shared__ u32 s_w[PS_THREADS / 32][PT_MAX_G];
    __shared__ u64 s_base[PT_MAX_G];
    PDL_ENTER();
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    // sums of the CTAs before this one (base) and of all CTAs (total), per bucket
    u64 before[PT_MAX_G] = {}, total[PT_MAX_G] = {};
    for (u32 cta = tid; cta < gridDim.x; cta += PS_THREADS) {
#pragma unroll
        for (int b = 0; b < PT_MAX_G; b++) {
            const u64 x = blocksum[(size_t)cta * PT_MAX_G + b];
            total[b] += x;
            if (cta < blockIdx.x) before[b] += x;
        }
    }
#pragma unroll
    for (int b = 0; b < PT_MAX_G; b++) {
        const u64 x = warp_sum(before[b]), y = warp_sum(total[b]);
        if (lane == 0) { s_red[w][0][b] = x; s_red[w][1][b] = y; }
    }
    u32 v[PS_TPT][PT_MAX_G];
    const u32 t0 = blockIdx.x * PS_CHUNK + tid * PS_TPT;
    ps_load_tiles(tile_counts, num_tiles, t0, v);
    u32 mine[PT_MAX_G], incl[PT_MAX_G];
#pragma unroll
    for (int b = 0; b < PT_MAX_G; b++) {
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < PS_TPT; k++) s += v[k][b];
        mine[b] = s;
        incl[b] = warp_incl_scan(s);
        if (lane == 31) s_w[w][b] = incl[b];
    }
    __syncthreads();
    if (tid < (u32)PT_MAX_G) {
        u64 x = 0, y = 0;
#pragma unroll
        for (int ww = 0; ww < PS_THREADS / 32; ww++) { x += s_red[ww][0][tid]; y += s_red[ww][1][tid]; }
        s_base[tid] = x;
        if (blockIdx.x == 0) bucket_total[tid] = y;
        s_red[0][1][tid] = y;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
        u64 run = 0;
        for (int b = 0; b < G; b++) { bucket_start[b] = run; run += s_red[0][1][b]; }
        bucket_start[G] = run;
    }
#pragma unroll
    for (int b = 0; b < PT_MAX_G; b++) {
        u32 run = (u32)s_base[b] + incl[b] - mine[b];
        for (u32 ww = 0; ww < w; ww++) run += s_w[ww][b];
#pragma unroll
        for (int k = 0; k < PS_TPT; k++) { const u32 c = v[k][b]; v[k][b] = run; run += c; }
    }
#pragma unroll
    for (int k = 0; k < PS_TPT; k++)
        if (t0 + k < num_tiles) {
            uint4 *o = reinterpret_cast<uint4 *>(off32 + (size_t)(t0 + k) * PT_MAX_G);
            o[0] = make_uint4(v[k][0], v[k][1], v[k][2], v[k][3]);
            o[1] = make_uint4(v[k][4], v[k][5], v[k][6], v[k][7]);
        }
}

// The exchange.  A tile's survivors sit bucket-ordered and contiguous at the start of the tile's slot; segment (t, b) goes to
// dst.base[b] + (row0[b] + off32[t][b]) rows -- dst.base[b] is the local send buffer (grouped ncclSend path), the rank's own
// receive buffer (its own bucket) or a PEER GPU's receive buffer (peer mapping: the stores travel over NVLink from the SMs,
// so the compaction IS the exchange).  One warp per tile: the tile's 8 counts and offsets arrive with two 32-byte loads,
// then the warp walks the slot as a flat array of cells (16-byte words when rows are whole 16-byte multiples), PX_UNROLL
// independent loads in flight per lane before the first store, each word routed to its bucket by comparing its index with
// the bucket boundaries held in registers.  (The first version took one warp per SEGMENT: a dependent chain of count ->
// offset -> data loads for every ~2 KB piece, 0.30 ms for 122 MB at 8 GPUs.)
constexpr int PX_UNROLL = 4;

template <typename W>   // W = int4 (rows are multiples of 16 bytes and every pointer is 16-byte aligned) or int32_t
__global__ void __launch_bounds__(256)
partition_exchange_kernel(const int32_t *__restrict__ slots, const u32 *__restrict__ tile_counts, const u32 *__restrict__ off32,
                          u32 num_tiles, int G, u32 tile_rows, int cols, const SmjPartitionDst D)
{
    PDL_ENTER();
    if (D.skip && *D.skip) return;
    constexpr u32 CPW = sizeof(W) / 4;               // cells per word
    const u32 wpr = (u32)cols / CPW;                 // words per row
    const u32 lane = threadIdx.x & 31u;
    const u32 warps = gridDim.x * (blockDim.x >> 5);
    for (u32 t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < num_tiles; t += warps) {
        u32 cnt = 0, off = 0;
        u64 r0 = 0;
        if (lane < (u32)G) {
            cnt = tile_counts[(size_t)t * PT_MAX_G + lane];
            off = off32[(size_t)t * PT_MAX_G + lane];
            r0 = D.row0 ? D.row0[lane] : 0ull;
        }
        const u32 incl = warp_incl_scan(cnt);
        const u32 total_w = __shfl_sync(FULL_MASK, incl, PT_MAX_G - 1) * wpr;
        if (total_w == 0) continue;
        // per bucket: first word of its segment inside the slot, and where that word goes
        u32 start_w[PT_MAX_G];
        W *dstp[PT_MAX_G];
#pragma unroll
        for (int b = 0; b < PT_MAX_G; b++) {
            start_w[b] = (__shfl_sync(FULL_MASK, incl, b) - __shfl_sync(FULL_MASK, cnt, b)) * wpr;
            const u64 row = __shfl_sync(FULL_MASK, r0, b) + (u64)__shfl_sync(FULL_MASK, off, b);
            dstp[b] = reinterpret_cast<W *>(D.base[b]) + row * wpr;
        }
        const W *src = reinterpret_cast<const W *>(slots + (size_t)t * tile_rows * cols);
        for (u32 i0 = 0; i0 < total_w; i0 += 32 * PX_UNROLL) {
            W v[PX_UNROLL];
#pragma unroll
            for (int k = 0; k < PX_UNROLL; k++) {
                const u32 i = i0 + k * 32 + lane;
                if (i < total_w) v[k] = __ldcs(src + i);   // read once
            }
#pragma unroll
            for (int k = 0; k < PX_UNROLL; k++) {
                const u32 i = i0 + k * 32 + lane;
                if (i < total_w) {
                    W *d = dstp[0];
                    u32 s = 0;
#pragma unroll
                    for (int b = 1; b < PT_MAX_G; b++)
                        if (b < G && i >= start_w[b]) { d = dstp[b]; s = start_w[b]; }
                    d[i - s] = v[k];
                }
            }
        }
    }
    // Every thread waits until its own stores -- most of them into peer memory, posted over NVLink -- have been performed
    // system-wide before the kernel may count as finished: the arrival flag that the next kernel on the stream sends must not
    // overtake a row still in flight.  (Without it, 14 of 3.33 M joined rows were missing at 250M x 50M rows per GPU on two
    // GPUs: the largest exchange that had been run; every smaller one had passed.)
    __threadfence_system();
}

// samples[i] = flipped key of row floor((2i+1) n / 2S) if it passes the predicate, else 0xffffffff
__global__ void sample_rows_kernel(const int32_t *__restrict__ in, int64_t n, int cols, int sel_col, int32_t sel_val, int select_all,
                                   int key_col, int S, u32 *samples)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    u32 v = 0xffffffffu;
    if (n > 0) {
        int64_t pos = (int64_t)(((unsigned long long)(2 * i + 1) * (unsigned long long)n) / (unsigned long long)(2 * S));
        if (pos >= n) pos = n - 1;
        const int32_t *r = in + pos * cols;
        if (select_all || r[sel_col] > sel_val) v = (u32)r[key_col] ^ 0x80000000u;
    }
    samples[i] = v;
}

// Device twin of smj_plan_splitters (smj_dist.cu): sort the gathered samples (one CTA, bitonic in shared memory), drop the
// "none" marks (0xffffffff sorts last), splitter b = sample at quantile b/G.  Every rank runs it on the same gathered
// samples, so every rank gets the same splitters without a host round trip.
constexpr int SPL_THREADS = 1024;
__global__ void __launch_bounds__(SPL_THREADS) splitters_kernel(const u32 *__restrict__ samples, int n_samples, int n_pow2, int G, u32 *splitters)
{
    extern __shared__ u32 s_s[];
    __shared__ int s_valid;
    const int tid = threadIdx.x;
    if (tid == 0) s_valid = 0;
    for (int i = tid; i < n_pow2; i += SPL_THREADS) s_s[i] = i < n_samples ? samples[i] : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pow2; i += SPL_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const u32 a = s_s[i], b = s_s[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_s[i] = b; s_s[p] = a; }
                }
            }
            __syncthreads();
        }
    int local = 0;
    for (int i = tid; i < n_pow2; i += SPL_THREADS) local += s_s[i] != 0xffffffffu;
    atomicAdd(&s_valid, local);
    __syncthreads();
    const int nv = s_valid;
    if (tid >= 1 && tid < G) {
        u32 v = 0xffffffffu;
        if (nv > 0) {
            long long pos = (long long)tid * nv / G;
            if (pos >= nv) pos = nv - 1;
            v = s_s[pos];
        }
        splitters[tid - 1] = v;
    }
}

int pt_ipt(int cols)
{
    int ipt = PT_IPT;
    while (ipt > 1 && (size_t)ipt * PT_THREADS * cols * 4 > PT_STAGE_BYTES) ipt >>= 1;
    return ipt;
}

}  // namespace

bool smj_partition_supported(const int32_t *d_in, int cols) { return cols <= PT_MAX_COLS && ((uintptr_t)d_in & 15) == 0; }
size_t smj_partition_tiles(int64_t n, int cols) { const int64_t tr = (int64_t)pt_ipt(cols) * PT_THREADS; return (size_t)((n + tr - 1) / tr); }
u32 smj_partition_tile_rows(int cols) { return (u32)(pt_ipt(cols) * PT_THREADS); }

// scratch of one table: [tile_counts u32 tiles*8][off32 u32 tiles*8][blocksum u64 ctas*8][bucket_total u64 8][bucket_start u64 9]
SmjPartScratch smj_partition_scratch(char *base, int64_t n, int cols)
{
    SmjPartScratch s;
    s.tiles = smj_partition_tiles(n, cols);
    s.ctas = (s.tiles + PS_CHUNK - 1) / PS_CHUNK;
    const size_t tb = align_up(s.tiles * PT_MAX_G * 4, 256);
    s.counts = (u32 *)base;
    s.off32 = (u32 *)(base + tb);
    s.blocksum = (u64 *)(base + 2 * tb);
    s.bucket_total = s.blocksum + (s.ctas ? s.ctas : 1) * PT_MAX_G;
    s.bucket_start = s.bucket_total + PT_MAX_G;
    s.bytes = 2 * tb + ((s.ctas ? s.ctas : 1) * PT_MAX_G + PT_MAX_G + PT_MAX_G + 1) * 8 + 256;
    return s;
}
size_t smj_partition_scratch_bytes(int64_t n, int cols) { return smj_partition_scratch(nullptr, n, cols).bytes; }

int smj_launch_sample_rows(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int key_col, int S,
                           u32 *d_samples)
{
    int select_all = sel_val < (int64_t)INT32_MIN;
    int32_t sv = (int32_t)sel_val;
    if (!select_all && sel_val >= (int64_t)INT32_MAX) { n = 0; }   // nothing passes: all samples "none"
    sample_rows_kernel<<<(S + 255) / 256, 256, 0, c->stream>>>(d_in, n, cols, sel_col, sv, select_all, key_col, S, d_samples);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// Stage 1 (on stream st): rows of d_in that pass the predicate, grouped by destination bucket inside each tile's slot (d_slots:
// n*cols cells of scratch), then the per-tile counts, every segment's offset inside its bucket, the bucket totals and
// the bucket starts (d_scratch: smj_partition_scratch).
int smj_launch_select_partition(SmjCtx *c, cudaStream_t st, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val,
                                int key_col, const u32 *d_splitters, int G, int32_t *d_slots, char *d_scratch)
{
    if (G < 1 || G > PT_MAX_G) return smj_set_error(SMJ_EINVAL, "select_partition: %d buckets (max %d)", G, PT_MAX_G);
    int select_all = sel_val < (int64_t)INT32_MIN;
    if (!select_all && sel_val >= (int64_t)INT32_MAX) n = 0;
    const SmjPartScratch S = smj_partition_scratch(d_scratch, n > 0 ? n : 0, cols);
    if (n <= 0) { CUDA_TRY(cudaMemsetAsync(S.bucket_total, 0, (size_t)(2 * PT_MAX_G + 1) * 8, st)); return SMJ_OK; }
    static bool attr_set[16] = {};
    if (!attr_set[c->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(select_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PTW_SMEM));
        attr_set[c->device & 15] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const int ipt = pt_ipt(cols);
    // CTAs per SM (SMJ_PT_CTAS, default 3 = what the registers allow).  This kernel's persistent CTAs hold their SM for the whole
    // pass, and with three of them the OTHER table's exchange kernel, launched on the second stream to overlap with this
    // pass, finds no room until this kernel drains; two leave room (measured at two GPUs: first arrival 148 instead of 162 us
    // after the first pass, but the pass itself 98 instead of 86 us -- a wash, so the default stays at three).
    static const int pt_ctas = getenv("SMJ_PT_CTAS") ? atoi(getenv("SMJ_PT_CTAS")) : 3;
    const u32 per_sm = (u32)(pt_ctas >= 1 && pt_ctas <= 3 ? pt_ctas : 3);
    const u32 grid = S.tiles < (size_t)(sms * per_sm) ? (u32)S.tiles : (u32)(sms * per_sm);
    // (the chain's kernels are launched with programmatic stream serialization: each becomes resident while its predecessor
    // drains and starts with griddepcontrol.wait, which takes the launch latency out of a chain of seven kernels per table)
    smj_launch_on(c, st, select_partition_kernel, grid, PTW_THREADS, PTW_SMEM, d_in, n, cols, ipt, sel_col, (int32_t)sel_val, select_all, key_col,
                  d_splitters, G, d_slots, S.counts, (u32)S.tiles);
    KERNEL_CHECK(c);
    smj_launch_on(c, st, partition_blocksum_kernel, (u32)S.ctas, PS_THREADS, 0, (const u32 *)S.counts, (u32)S.tiles, S.blocksum);
    KERNEL_CHECK(c);
    smj_launch_on(c, st, partition_offsets_kernel, (u32)S.ctas, PS_THREADS, 0, (const u32 *)S.counts, (u32)S.tiles, G, (const u64 *)S.blocksum, S.off32,
                  S.bucket_total, S.bucket_start);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// Stage 2 (on stream st): every (tile, bucket) segment to its place behind D.base[bucket] (see partition_exchange_kernel).
int smj_launch_partition_exchange(SmjCtx *c, cudaStream_t st, int64_t n, int cols, int sel_val_none, int G, const int32_t *d_slots,
                                  char *d_scratch, const SmjPartitionDst &D)
{
    if (n <= 0 || sel_val_none) return SMJ_OK;
    const SmjPartScratch S = smj_partition_scratch(d_scratch, n, cols);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const u32 cgrid = (u32)((S.tiles + 7) / 8 < (size_t)sms * 8 ? (S.tiles + 7) / 8 : (size_t)sms * 8);
    uintptr_t al = (uintptr_t)d_slots;
    for (int b = 0; b < G; b++) al |= (uintptr_t)D.base[b];
    if (cols % 4 == 0 && (al & 15) == 0)
        smj_launch_on(c, st, partition_exchange_kernel<int4>, cgrid, 256, 0, d_slots, (const u32 *)S.counts, (const u32 *)S.off32, (u32)S.tiles, G,
                      smj_partition_tile_rows(cols), cols, D);
    else
        smj_launch_on(c, st, partition_exchange_kernel<int32_t>, cgrid, 256, 0, d_slots, (const u32 *)S.counts, (const u32 *)S.off32, (u32)S.tiles, G,
                      smj_partition_tile_rows(cols), cols, D);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// splitters[0..G-2] from n_samples gathered samples (<= 16384), on the device
int smj_launch_splitters(SmjCtx *c, const u32 *d_samples, int n_samples, int G, u32 *d_splitters)
{
    if (G <= 1) return SMJ_OK;
    int p2 = 1;
    while (p2 < n_samples) p2 <<= 1;
    if (p2 > 16384) return smj_set_error(SMJ_EINVAL, "smj_launch_splitters: %d samples (max 16384)", n_samples);
    static bool attr_set[16] = {};
    if (!attr_set[c->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(splitters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
        attr_set[c->device & 15] = true;
    }
    splitters_kernel<<<1, SPL_THREADS, (size_t)p2 * 4, c->stream>>>(d_samples, n_samples, p2, G, d_splitters);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// Loads this file's pipeline kernels on the current device.  CUDA loads a kernel lazily at its first launch, and that load can
// wait for other GPUs' running kernels when peer access is enabled; a process that drives several GPUs (smj_dist.cu) must
// not meet such a load while another rank's kernel spins on this rank's flags, so it loads everything up front.
void smj_preload_partition(void)
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, select_partition_kernel);
    cudaFuncGetAttributes(&a, partition_blocksum_kernel);
    cudaFuncGetAttributes(&a, partition_offsets_kernel);
    cudaFuncGetAttributes(&a, partition_exchange_kernel<int4>);
    cudaFuncGetAttributes(&a, partition_exchange_kernel<int32_t>);
    cudaFuncGetAttributes(&a, sample_rows_kernel);
    cudaFuncGetAttributes(&a, splitters_kernel);
    cudaGetLastError();
}
