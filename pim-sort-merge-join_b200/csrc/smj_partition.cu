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
//   partition_count_kernel  : first pass over the table: predicate, bucket id, G survivor counts per 256-row warp tile.
//   partition_route_kernel  : second pass, once the offsets and the peers' counts are known: every surviving row is stored
//                             straight to its destination (send buffer, own or PEER receive buffer).
//   partition_blocksum / partition_offsets_kernel: offsets of every (bucket, tile) segment + bucket totals, many CTAs.
//   sample_rows_kernel      : regular row samples (predicate applied) from which the splitters are derived.
#include "smj_internal.h"
#include "smj_dev.cuh"
#include <stdlib.h>

namespace {

constexpr int PT_MAX_G = SMJ_MAX_G;
constexpr int PT_MAX_COLS = 32;   // (the local pipeline's TMA select takes tables of up to 32 columns)

// bucket of a flipped key: the number of splitters <= key (unused splitters are 0xffffffff), at most G - 1
__device__ __forceinline__ u32 pt_bucket(u32 fk, const u32 (&sp)[PT_MAX_G - 1], int G)
{
    u32 b = 0;
#pragma unroll
    for (int q = 0; q < PT_MAX_G - 1; q++) b += (q < G - 1 && fk >= sp[q]) ? 1u : 0u;
    return b;
}

// The partition is TWO streaming passes over the table and NO intermediate copy of the rows.  The unit of work is a WARP
// TILE: 256 consecutive rows (eight groups of 32, lane = row inside a group), so a warp needs nobody else -- no shared-memory
// ring, no barrier -- and an SM runs as many of them as its registers hold:
//   count pass: predicate + bucket per row, the tile's survivor count per bucket (one ballot per bucket and row group, lane q
//               keeps bucket q's count) -> tile_counts;
//   route pass: once the scan has given every (tile, bucket) segment its offset and the count exchange has said where this
//               rank's bucket b starts at rank b, the same evaluation again; one ballot per bucket ranks a group's rows, lane q
//               keeps bucket q's cursor, and every lane stores its own row straight behind D.base[bucket] -- the local send
//               buffer (ncclSend path), the rank's own receive buffer, or a PEER GPU's receive buffer (peer mapping: NVLink
//               stores from the SMs, the compaction IS the exchange).  Rows of one bucket leave in original order.
// 16-byte rows (4 columns) are loaded whole, all eight groups of a tile in flight before the first one is routed; other
// shapes load the select and key cells first and copy each surviving row cell by cell (16 bytes at a time when they can).
// History (profiles/r02_partition_history.md): slot-writing passes fed by a TMA ring -- bucket-ordered slots with a 512-entry
// scan per tile, then order-preserving slots with routing in a second kernel, then the two passes of today but on 2048-row
// CTA tiles -- all ran at 40-110 us per 160 MB table: two or three CTAs of eight warps per SM could not hide the per-tile
// dependency chain (ring wait, barrier, ballots, cursor shuffles, stores).
// STAGED (route pass, 16-byte rows, three or more buckets): the rows of every bucket are collected in a 512-byte
// shared-memory buffer per warp and leave as one coalesced store; without it a group of 32 rows leaves as G runs of ~32 / G
// rows that start anywhere (partial lines over NVLink).
constexpr int PW_GROUPS = 8;                       // row groups of 32 per warp tile
constexpr int PW_TILE = PW_GROUPS * 32;            // 256 rows
constexpr int PW_THREADS = 256;
constexpr int PW_WARPS = PW_THREADS / 32;
constexpr size_t PW_SMEM_STAGED = (size_t)PW_WARPS * PT_MAX_G * 32 * 16;   // 32 KB

struct PartPassArgs {
    const int32_t *in; int64_t n; int cols, sel_col; int32_t sel_val; int select_all, key_col; const u32 *splitters; int G;
    u32 *tile_counts;            // count pass: out [tiles][PT_MAX_G]
    const u32 *off32;            // route pass: offsets of the (tile, bucket) segments inside their buckets
    u32 num_tiles;
};

// select and key cells of the tile's rows (COLS4: the whole 16-byte rows), predicate and bucket per row group
template <bool COLS4>
__device__ __forceinline__ void pw_load_tile(const PartPassArgs &A, u32 tile, u32 lane, const u32 (&sp)[PT_MAX_G - 1], int4 (&r4)[COLS4 ? PW_GROUPS : 1],
                                             u32 &passmask, u32 &bpack)
{
    const int64_t row0 = (int64_t)tile * PW_TILE + lane;
    int32_t sv[PW_GROUPS], kv[PW_GROUPS];
#pragma unroll
    for (int j = 0; j < PW_GROUPS; j++) {
        const int64_t row = row0 + j * 32;
        sv[j] = 0; kv[j] = 0;
        if (row < A.n) {
            if (COLS4) {
                const int4 r = __ldcs(reinterpret_cast<const int4 *>(A.in) + row);   // read once per pass
                r4[COLS4 ? j : 0] = r;
                sv[j] = A.sel_col == 0 ? r.x : A.sel_col == 1 ? r.y : A.sel_col == 2 ? r.z : r.w;
                kv[j] = A.key_col == 0 ? r.x : A.key_col == 1 ? r.y : A.key_col == 2 ? r.z : r.w;
            } else {
                const int32_t *p = A.in + row * A.cols;
                sv[j] = __ldg(p + A.sel_col);
                kv[j] = (A.key_col == A.sel_col) ? sv[j] : __ldg(p + A.key_col);
            }
        }
    }
    passmask = 0; bpack = 0;
#pragma unroll
    for (int j = 0; j < PW_GROUPS; j++) {
        const bool pass = (row0 + j * 32 < A.n) && (A.select_all || sv[j] > A.sel_val);
        passmask |= (pass ? 1u : 0u) << j;
        bpack |= pt_bucket((u32)kv[j] ^ 0x80000000u, sp, A.G) << (3 * j);
    }
}

__global__ void __launch_bounds__(PW_THREADS)
partition_count_kernel(const PartPassArgs A, int cols4)
{
    PDL_ENTER();
    const u32 lane = threadIdx.x & 31u;
    const u32 warps = gridDim.x * PW_WARPS;
    u32 sp[PT_MAX_G - 1];
#pragma unroll
    for (int q = 0; q < PT_MAX_G - 1; q++) sp[q] = (q < A.G - 1) ? A.splitters[q] : 0xffffffffu;
    for (u32 t = blockIdx.x * PW_WARPS + (threadIdx.x >> 5); t < A.num_tiles; t += warps) {
        u32 passmask, bpack, bcnt = 0;
        if (cols4) { int4 r4[PW_GROUPS]; pw_load_tile<true>(A, t, lane, sp, r4, passmask, bpack); }
        else { int4 r1[1]; pw_load_tile<false>(A, t, lane, sp, r1, passmask, bpack); }
#pragma unroll
        for (int j = 0; j < PW_GROUPS; j++) {
            const bool pass = (passmask >> j) & 1u;
            const u32 b = (bpack >> (3 * j)) & 7u;
#pragma unroll
            for (int q = 0; q < PT_MAX_G; q++) {
                if (q < A.G) {   // warp-uniform
                    const u32 mq = __ballot_sync(FULL_MASK, pass && b == (u32)q);
                    if (lane == (u32)q) bcnt += __popc(mq);
                }
            }
        }
        if (lane < (u32)PT_MAX_G) A.tile_counts[(size_t)t * PT_MAX_G + lane] = bcnt;
    }
}

template <bool COLS4, bool STAGED>
__global__ void __launch_bounds__(PW_THREADS)
partition_route_kernel(const PartPassArgs A, const SmjPartitionDst D)
{
    extern __shared__ __align__(16) int4 pw_stage[];
    PDL_ENTER();
    if (D.skip && *D.skip) return;
    const int cols = A.cols, G = A.G;
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const u32 lt = lanemask_lt();
    const u32 warps = gridDim.x * PW_WARPS;
    const u64 row_bytes = (u64)cols * 4;
    u32 sp[PT_MAX_G - 1];
#pragma unroll
    for (int q = 0; q < PT_MAX_G - 1; q++) sp[q] = (q < G - 1) ? A.splitters[q] : 0xffffffffu;
    u64 base_q = 0;        // lane q: byte address where this rank's bucket q starts counting rows
    if (lane < (u32)G) {
        int32_t *bp = D.base[0];
#pragma unroll
        for (int q = 1; q < PT_MAX_G; q++) if (lane == (u32)q) bp = D.base[q];
        base_q = reinterpret_cast<u64>(bp) + (D.row0 ? D.row0[lane] : 0ull) * row_bytes;
    }
    const bool vec = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.in) & 15) == 0);   // (the launcher checks the destinations)
    int4 *stage_buf = STAGED ? pw_stage + (size_t)w * PT_MAX_G * 32 : nullptr;   // [bucket][32 rows]
    for (u32 t = blockIdx.x * PW_WARPS + w; t < A.num_tiles; t += warps) {
        u64 cur_q = 0;     // lane q: byte address of bucket q's next row from this tile
        u32 fill_q = 0;    // STAGED, lane q: rows of bucket q waiting in the warp's buffer
        if (lane < (u32)G) cur_q = base_q + (u64)A.off32[(size_t)t * PT_MAX_G + lane] * row_bytes;
        int4 r4[COLS4 ? PW_GROUPS : 1];
        u32 passmask, bpack;
        pw_load_tile<COLS4>(A, t, lane, sp, r4, passmask, bpack);
#pragma unroll
        for (int j = 0; j < PW_GROUPS; j++) {
            const bool pass = (passmask >> j) & 1u;
            const u32 b = (bpack >> (3 * j)) & 7u;
            if (__ballot_sync(FULL_MASK, pass) == 0u) continue;   // warp-uniform: nothing survives in this group
            u32 mine = 0, add = 0;
#pragma unroll
            for (int q = 0; q < PT_MAX_G; q++) {
                if (q < G) {   // warp-uniform
                    const u32 mq = __ballot_sync(FULL_MASK, pass && b == (u32)q);
                    if (b == (u32)q) mine = mq;
                    if (lane == (u32)q) add = __popc(mq);
                }
            }
            if (STAGED) {
                const u32 pos = __shfl_sync(FULL_MASK, fill_q, b) + __popc(mine & lt);   // slot in bucket b's buffer (may reach 62)
                if (pass && pos < 32u) stage_buf[b * 32 + pos] = r4[COLS4 ? j : 0];
                __syncwarp();
                u32 full = __ballot_sync(FULL_MASK, lane < (u32)G && fill_q + add >= 32u);   // full buffers leave as 512-byte stores
                while (full) {
                    const int q = __ffs(full) - 1;
                    full &= full - 1;
                    const u64 dst = __shfl_sync(FULL_MASK, cur_q, q);
                    reinterpret_cast<int4 *>(dst)[lane] = stage_buf[q * 32 + lane];
                    if (lane == (u32)q) cur_q += 512ull;
                }
                __syncwarp();
                if (pass && pos >= 32u) stage_buf[b * 32 + (pos - 32u)] = r4[COLS4 ? j : 0];   // rows that did not fit: into the emptied buffer
                if (lane < (u32)G) { fill_q += add; if (fill_q >= 32u) fill_q -= 32u; }
                __syncwarp();
            } else {
                const u64 dst_b = __shfl_sync(FULL_MASK, cur_q, b);     // bucket b's cursor lives in lane b
                cur_q += (u64)add * row_bytes;
                if (pass) {
                    const u64 dst = dst_b + (u64)__popc(mine & lt) * row_bytes;
                    if (COLS4) {
                        *reinterpret_cast<int4 *>(dst) = r4[COLS4 ? j : 0];
                    } else {
                        const int32_t *rs = A.in + ((int64_t)t * PW_TILE + j * 32 + lane) * cols;
                        int32_t *rd = reinterpret_cast<int32_t *>(dst);
                        if (vec) for (int q = 0; q < cols / 4; q++) reinterpret_cast<int4 *>(rd)[q] = __ldcs(reinterpret_cast<const int4 *>(rs) + q);
                        else for (int q = 0; q < cols; q++) rd[q] = __ldcs(rs + q);
                    }
                }
            }
        }
        if (STAGED) {   // what is left of the tile
#pragma unroll
            for (int q = 0; q < PT_MAX_G; q++) {
                if (q < G) {
                    const u32 f = __shfl_sync(FULL_MASK, fill_q, q);
                    const u64 dst = __shfl_sync(FULL_MASK, cur_q, q);
                    if (lane < f) reinterpret_cast<int4 *>(dst)[lane] = stage_buf[q * 32 + lane];
                }
            }
            __syncwarp();
        }
    }
    // Every thread waits until its own stores -- most of them into peer memory, posted over NVLink -- have been performed
    // system-wide before the kernel may count as finished: the arrival flag that the next kernel on the stream sends must not
    // overtake a row still in flight.  (Without it, 14 of 3.33 M joined rows were missing at 250M x 50M rows per GPU on two
    // GPUs: the largest exchange that had been run; every smaller one had passed.)
    __threadfence_system();
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

}  // namespace

bool smj_partition_supported(const int32_t *d_in, int cols) { return cols <= PT_MAX_COLS && ((uintptr_t)d_in & 15) == 0; }
size_t smj_partition_tiles(int64_t n, int cols) { (void)cols; return (size_t)((n + PW_TILE - 1) / PW_TILE); }
u32 smj_partition_tile_rows(int cols) { (void)cols; return (u32)PW_TILE; }

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

static PartPassArgs part_args(const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int key_col, const u32 *d_splitters, int G,
                              const SmjPartScratch &S)
{
    PartPassArgs A = {};
    A.in = d_in; A.n = n; A.cols = cols; A.sel_col = sel_col; A.sel_val = (int32_t)sel_val;
    A.select_all = sel_val < (int64_t)INT32_MIN; A.key_col = key_col; A.splitters = d_splitters; A.G = G;
    A.tile_counts = S.counts; A.off32 = S.off32; A.num_tiles = (u32)S.tiles;
    return A;
}

// CTAs per SM of the two passes.  Both kernels are persistent (a warp takes tiles with a grid stride), and the count pass of
// one table is meant to run NEXT TO the route pass of the other (second stream): each takes about half an SM's warps and
// registers, so neither waits for the other to drain.
static u32 part_grid(SmjCtx *c, size_t tiles, int per_sm)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const size_t ctas = (tiles + PW_WARPS - 1) / PW_WARPS;
    return (u32)(ctas < (size_t)sms * per_sm ? ctas : (size_t)sms * per_sm);
}

// Pass 1 (on stream st): survivors of every warp tile per destination bucket, then every (tile, bucket) segment's offset
// inside its bucket, the bucket totals and the bucket starts (d_scratch: smj_partition_scratch).
int smj_launch_select_partition(SmjCtx *c, cudaStream_t st, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val,
                                int key_col, const u32 *d_splitters, int G, char *d_scratch)
{
    if (G < 1 || G > PT_MAX_G) return smj_set_error(SMJ_EINVAL, "select_partition: %d buckets (max %d)", G, PT_MAX_G);
    const int select_all = sel_val < (int64_t)INT32_MIN;
    if (!select_all && sel_val >= (int64_t)INT32_MAX) n = 0;
    const SmjPartScratch S = smj_partition_scratch(d_scratch, n > 0 ? n : 0, cols);
    if (n <= 0) { CUDA_TRY(cudaMemsetAsync(S.bucket_total, 0, (size_t)(2 * PT_MAX_G + 1) * 8, st)); return SMJ_OK; }
    static const int count_ctas = getenv("SMJ_PT_COUNT_CTAS") ? atoi(getenv("SMJ_PT_COUNT_CTAS")) : 4;
    const int cols4 = (cols == 4 && ((uintptr_t)d_in & 15) == 0) ? 1 : 0;
    // (the chain's kernels are launched with programmatic stream serialization: each becomes resident while its predecessor
    // drains and starts with griddepcontrol.wait, which takes the launch latency out of a chain of seven kernels per table)
    smj_launch_on(c, st, partition_count_kernel, part_grid(c, S.tiles, count_ctas), PW_THREADS, 0,
                  part_args(d_in, n, cols, sel_col, sel_val, key_col, d_splitters, G, S), cols4);
    KERNEL_CHECK(c);
    smj_launch_on(c, st, partition_blocksum_kernel, (u32)S.ctas, PS_THREADS, 0, (const u32 *)S.counts, (u32)S.tiles, S.blocksum);
    KERNEL_CHECK(c);
    smj_launch_on(c, st, partition_offsets_kernel, (u32)S.ctas, PS_THREADS, 0, (const u32 *)S.counts, (u32)S.tiles, G, (const u64 *)S.blocksum, S.off32,
                  S.bucket_total, S.bucket_start);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// Pass 2 (on stream st): the table is streamed again and every surviving row goes straight to its place behind
// D.base[bucket] (see partition_route_kernel).
int smj_launch_partition_exchange(SmjCtx *c, cudaStream_t st, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int key_col,
                                  const u32 *d_splitters, int G, char *d_scratch, const SmjPartitionDst &D)
{
    const int select_all = sel_val < (int64_t)INT32_MIN;
    if (n <= 0 || (!select_all && sel_val >= (int64_t)INT32_MAX)) return SMJ_OK;
    const SmjPartScratch S = smj_partition_scratch(d_scratch, n, cols);
    uintptr_t al = 0;
    for (int b = 0; b < G; b++) al |= (uintptr_t)D.base[b];
    if (cols % 4 == 0 && (al & 15) != 0) return smj_set_error(SMJ_EINVAL, "partition route pass: destinations must be 16-byte aligned");
    static const int stage_min_g = getenv("SMJ_DIST_STAGE_MIN_G") ? atoi(getenv("SMJ_DIST_STAGE_MIN_G")) : 3;
    static const int route_ctas = getenv("SMJ_PT_ROUTE_CTAS") ? atoi(getenv("SMJ_PT_ROUTE_CTAS")) : 2;
    const PartPassArgs A = part_args(d_in, n, cols, sel_col, sel_val, key_col, d_splitters, G, S);
    const u32 grid = part_grid(c, S.tiles, route_ctas);
    const bool cols4 = cols == 4 && ((uintptr_t)d_in & 15) == 0;
    if (cols4 && G >= stage_min_g)
        smj_launch_on(c, st, partition_route_kernel<true, true>, grid, PW_THREADS, PW_SMEM_STAGED, A, D);
    else if (cols4)
        smj_launch_on(c, st, partition_route_kernel<true, false>, grid, PW_THREADS, 0, A, D);
    else
        smj_launch_on(c, st, partition_route_kernel<false, false>, grid, PW_THREADS, 0, A, D);
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
    cudaFuncGetAttributes(&a, partition_count_kernel);
    cudaFuncGetAttributes(&a, partition_route_kernel<true, true>);
    cudaFuncGetAttributes(&a, partition_route_kernel<true, false>);
    cudaFuncGetAttributes(&a, partition_route_kernel<false, false>);
    cudaFuncGetAttributes(&a, partition_blocksum_kernel);
    cudaFuncGetAttributes(&a, partition_offsets_kernel);
    cudaFuncGetAttributes(&a, sample_rows_kernel);
    cudaFuncGetAttributes(&a, splitters_kernel);
    cudaGetLastError();
}
