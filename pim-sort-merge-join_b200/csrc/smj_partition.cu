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
//   partition_scan_kernel   : offsets of every (bucket, tile) segment in the send buffer + bucket starts.
//   partition_compact_kernel: one warp per segment copies it to its place.
//   sample_rows_kernel      : regular row samples (predicate applied) from which the splitters are derived.
#include "smj_internal.h"
#include "smj_dev.cuh"

namespace {

constexpr int PT_THREADS = 256;
constexpr int PT_WARPS = PT_THREADS / 32;
constexpr int PT_IPT = 8;
constexpr int PT_STAGES = 2;
constexpr int PT_STAGE_BYTES = 32768;
constexpr int PT_MAX_G = 8;
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
                bulk_g2s(dst, src, b16, &s_full[stage]);
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
                // ranks inside every bucket at once: each lane adds 1 to the byte of its bucket, one 64-bit warp scan
                // (a row group holds at most 32 rows per warp, so a byte per bucket is enough)
                const u64 mine = pass ? (1ull << (8 * b)) : 0ull;
                u64 inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u64 up = __shfl_up_sync(FULL_MASK, inc, o);
                    if (lane >= (u32)o) inc += up;
                }
                const u64 tot = __shfl_sync(FULL_MASK, inc, 31);
                if (pass) { rank[j >> 2] |= (u32)(((inc - mine) >> (8 * b)) & 255ull) << (8 * (j & 3)); myb = b; }
                if (lane < (u32)G) {
                    const u32 cnt = (u32)((tot >> (8 * lane)) & 255ull);
                    if (cnt) s_cnt[(lane * PT_IPT + j) * PT_WARPS + w] = cnt;
                }
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

// off[t][b] = destination row of segment (tile t, bucket b) in the send buffer, segments ordered bucket-major then by
// tile (= original row order inside a bucket); bucket_start[b] = first row of bucket b, bucket_start[G] = total.
constexpr int PS_THREADS = 1024;
__global__ void __launch_bounds__(PS_THREADS)
partition_scan_kernel(const u32 *__restrict__ tile_counts, u32 num_tiles, int G, u64 *__restrict__ off, u64 *__restrict__ bucket_start)
{
    __shared__ u64 s_w[PS_THREADS / 32];
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u64 total_items = (u64)num_tiles * (u64)G;
    const u64 chunk = (total_items + PS_THREADS - 1) / PS_THREADS;
    const u64 lo = (u64)tid * chunk < total_items ? (u64)tid * chunk : total_items;
    const u64 hi = lo + chunk < total_items ? lo + chunk : total_items;
    u64 sum = 0;
    for (u64 i = lo; i < hi; i++) { const u32 b = (u32)(i / num_tiles), t = (u32)(i % num_tiles); sum += tile_counts[(size_t)t * PT_MAX_G + b]; }
    u64 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(FULL_MASK, inc, o);
        if (lane >= (u32)o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    u64 run = inc - sum;
    for (u32 ww = 0; ww < w; ww++) run += s_w[ww];
    if (tid == PS_THREADS - 1) bucket_start[G] = run + sum;
    for (u64 i = lo; i < hi; i++) {
        const u32 b = (u32)(i / num_tiles), t = (u32)(i % num_tiles);
        if (t == 0) bucket_start[b] = run;
        off[(size_t)t * PT_MAX_G + b] = run;
        run += tile_counts[(size_t)t * PT_MAX_G + b];
    }
}

__global__ void __launch_bounds__(256)
partition_compact_kernel(const int32_t *__restrict__ slots, const u32 *__restrict__ tile_counts, const u64 *__restrict__ off,
                         const u64 *__restrict__ bucket_start, u32 num_tiles, int G, u32 tile_rows, int cols,
                         int32_t *__restrict__ send, int32_t *const *__restrict__ dst_by_bucket)
{
    // dst_by_bucket (may be null): where the first row of this rank's bucket b goes -- the local receive buffer for the
    // rank's own bucket, a PEER GPU's receive buffer (CUDA-IPC mapping, stores travel over NVLink) for the others.
    // With it the compaction IS the exchange: no send buffer, no separate copy.
    const u32 lane = threadIdx.x & 31u;
    const u64 warps = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 segs = (u64)num_tiles * (u64)G;
    for (u64 s = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < segs; s += warps) {
        const u32 t = (u32)(s / (u64)G), b = (u32)(s % (u64)G);
        const u32 *tc = tile_counts + (size_t)t * PT_MAX_G;
        const u32 cnt = tc[b];
        if (cnt == 0) continue;
        u32 before = 0;
        for (u32 q = 0; q < b; q++) before += tc[q];
        const int32_t *src = slots + ((size_t)t * tile_rows + before) * cols;
        int32_t *dst = dst_by_bucket ? dst_by_bucket[b] + (off[(size_t)t * PT_MAX_G + b] - bucket_start[b]) * (u64)cols
                                     : send + off[(size_t)t * PT_MAX_G + b] * (u64)cols;
        const u32 cells = cnt * (u32)cols;
        if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            const u32 v = cells >> 2;
            for (u32 i = lane; i < v; i += 32) reinterpret_cast<int4 *>(dst)[i] = reinterpret_cast<const int4 *>(src)[i];
            for (u32 i = (v << 2) + lane; i < cells; i += 32) dst[i] = src[i];
        } else {
            for (u32 i = lane; i < cells; i += 32) dst[i] = src[i];
        }
    }
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
// scratch: [tile_counts u32 tiles*8][off u64 tiles*8][bucket_start u64 16]
size_t smj_partition_scratch_bytes(int64_t n, int cols) { return smj_partition_tiles(n, cols) * PT_MAX_G * 12 + 256; }

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

// Stage 1: rows of d_in that pass the predicate, grouped by destination bucket inside each tile's slot (d_slots:
// n*cols cells of scratch), with the per-tile counts, every segment's offset and the bucket starts (G+1 u64 row
// offsets, returned in *d_bucket_start, inside d_scratch of smj_partition_scratch_bytes).
int smj_launch_select_partition(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int key_col,
                                const u32 *d_splitters, int G, int32_t *d_slots, char *d_scratch, u64 **d_bucket_start)
{
    const size_t tiles = smj_partition_tiles(n, cols);
    u32 *d_counts = (u32 *)d_scratch;
    u64 *d_off = (u64 *)(d_scratch + align_up(tiles * PT_MAX_G * 4, 8));
    u64 *d_bs = d_off + tiles * PT_MAX_G;
    *d_bucket_start = d_bs;
    if (G < 1 || G > PT_MAX_G) return smj_set_error(SMJ_EINVAL, "select_partition: %d buckets (max %d)", G, PT_MAX_G);
    int select_all = sel_val < (int64_t)INT32_MIN;
    if (!select_all && sel_val >= (int64_t)INT32_MAX) n = 0;
    if (n <= 0) { CUDA_TRY(cudaMemsetAsync(d_bs, 0, (size_t)(G + 1) * 8, c->stream)); return SMJ_OK; }
    static bool attr_set[16] = {};
    if (!attr_set[c->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(select_partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PTW_SMEM));
        attr_set[c->device & 15] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const int ipt = pt_ipt(cols);
    const u32 grid = tiles < (size_t)(sms * 3) ? (u32)tiles : (u32)(sms * 3);
    select_partition_kernel<<<grid, PTW_THREADS, PTW_SMEM, c->stream>>>(d_in, n, cols, ipt, sel_col, (int32_t)sel_val, select_all, key_col,
                                                                       d_splitters, G, d_slots, d_counts, (u32)tiles);
    KERNEL_CHECK(c);
    partition_scan_kernel<<<1, PS_THREADS, 0, c->stream>>>(d_counts, (u32)tiles, G, d_off, d_bs);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// Stage 2: every (tile, bucket) segment to its place: in d_send (buckets contiguous, for ncclSend) when d_dst_by_bucket is
// null, else straight into the destination ranks' receive buffers (device array of G pointers).
int smj_launch_partition_compact(SmjCtx *c, int64_t n, int cols, int sel_val_none, int G, const int32_t *d_slots, char *d_scratch,
                                 int32_t *d_send, int32_t *const *d_dst_by_bucket)
{
    if (n <= 0 || sel_val_none) return SMJ_OK;
    const size_t tiles = smj_partition_tiles(n, cols);
    const u32 *d_counts = (const u32 *)d_scratch;
    const u64 *d_off = (const u64 *)(d_scratch + align_up(tiles * PT_MAX_G * 4, 8));
    const u64 *d_bs = d_off + tiles * PT_MAX_G;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const u64 segs = (u64)tiles * G;
    const u32 cgrid = (u32)((segs + 7) / 8 < (u64)sms * 8 ? (segs + 7) / 8 : (u64)sms * 8);
    partition_compact_kernel<<<cgrid, 256, 0, c->stream>>>(d_slots, d_counts, d_off, d_bs, (u32)tiles, G, (u32)(pt_ipt(cols) * PT_THREADS), cols,
                                                           d_send, d_dst_by_bucket);
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
