// smj_radix.cu -- LSD "onesweep" radix sort of (flipped key << 32 | row id) pairs by the key half.
//
// Replaces the per-DPU sort (sort-merge-join/sort_dpu.c:157-187 insertion sort per tasklet, :251-323 tasklet
// merge tree) and cpu_app.c:172-202.  The reference sort is a STABLE ascending insertion sort; every pass
// below is stable, so LSD over the key digits gives the identical order (ties stay in row-id order).
//
// One kernel per 8-bit digit: a CTA takes a tile through an atomic ticket, ranks its items with
// warp-match (__match_any_sync) against per-warp digit counters in shared memory, publishes the tile's
// 256 digit counts, resolves the global offset of each digit with a per-digit decoupled look-back over the
// earlier tiles (chained scan), reorders the tile in shared memory and writes each digit's run contiguously.
// Digit histograms for all passes come from the select kernel (or radix_hist_kernel) up front.
#include "smj_internal.h"
#include "smj_dev.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_THREADS * RS_IPT;   // 4096 pairs = 32 KB
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr u32 RS_FLAG_LOCAL = 1u << 30, RS_FLAG_INCL = 2u << 30, RS_VAL_MASK = (1u << 30) - 1;
static_assert(RS_THREADS >= SMJ_RADIX, "one thread per digit bin");

constexpr size_t RS_SMEM = (size_t)RS_TILE * 8 + (size_t)RS_WARPS * SMJ_RADIX * 4 + 2 * SMJ_RADIX * 4 + 32 * 4;

__global__ void __launch_bounds__(RS_THREADS)
radix_pass_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u32 n, int shift,
                  const u32 *__restrict__ bin_base, u32 *status, u32 *tile_counter, u32 *err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_items = reinterpret_cast<u64 *>(smem_raw);              // RS_TILE, tile in locally sorted order
    u32 *s_wcnt = reinterpret_cast<u32 *>(s_items + RS_TILE);      // [warp][256] counts -> exclusive over warps
    u32 *s_lbase = s_wcnt + RS_WARPS * SMJ_RADIX;                  // [256] first local slot of each digit
    u32 *s_goff = s_lbase + SMJ_RADIX;                             // [256] global slot of local slot 0 of the digit
    u32 *s_wsum = s_goff + SMJ_RADIX;                              // warp totals of the bin scan
    __shared__ u32 s_tile;

    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u32 lt = lanemask_lt();
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (u32 i = tid; i < RS_WARPS * SMJ_RADIX; i += RS_THREADS) s_wcnt[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 base = tile * RS_TILE;
    const u32 valid = (n - base < (u32)RS_TILE) ? (n - base) : (u32)RS_TILE;

    // warp-striped load: element order inside the tile is (warp, j, lane) == ascending index
    u64 item[RS_IPT];
    const u32 wbase = base + w * 32 * RS_IPT + lane;
#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        const u32 idx = wbase + j * 32;
        item[j] = (idx < n) ? in[idx] : ~0ull;   // padding sorts last (digit 255, after every real 255)
    }

    u32 rank[RS_IPT];
    u32 *my_cnt = s_wcnt + w * SMJ_RADIX;
#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        const u32 d = (u32)(item[j] >> shift) & (SMJ_RADIX - 1);
        const u32 peers = __match_any_sync(FULL_MASK, d);
        const u32 leader = __ffs(peers) - 1;
        u32 pre = 0;
        if (lane == leader) {
            pre = my_cnt[d];
            my_cnt[d] = pre + __popc(peers);
        }
        pre = __shfl_sync(FULL_MASK, pre, leader);
        rank[j] = pre + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // one thread per digit: totals over warps, publish, local scan, look-back
    u32 cnt_pad = 0, cnt = 0;
    if (tid < SMJ_RADIX) {
        u32 sum = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            const u32 t = s_wcnt[ww * SMJ_RADIX + tid];
            s_wcnt[ww * SMJ_RADIX + tid] = sum;
            sum += t;
        }
        cnt_pad = sum;
        cnt = sum;
        if (tid == SMJ_RADIX - 1) cnt -= (u32)RS_TILE - valid;   // do not count the padding
        st_relaxed(&status[(size_t)tile * SMJ_RADIX + tid], (tile == 0 ? RS_FLAG_INCL : RS_FLAG_LOCAL) | cnt);
    }
    const u32 inc = warp_incl_scan(cnt_pad);
    if (lane == 31) s_wsum[w] = inc;
    __syncthreads();
    if (tid < SMJ_RADIX) {
        u32 wp = 0;
        for (u32 ww = 0; ww < w; ww++) wp += s_wsum[ww];
        const u32 excl_local = wp + inc - cnt_pad;
        s_lbase[tid] = excl_local;

        u32 excl = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            u32 spins = 0;
            while (true) {
                const u32 v = ld_relaxed(&status[(size_t)t * SMJ_RADIX + tid]);
                const u32 flag = v >> 30;
                if (flag == 0) {
                    if (++spins > SMJ_SPIN_LIMIT) { atomicExch(err, SMJ_ERR_SPIN_RADIX); break; }
                    continue;
                }
                excl += v & RS_VAL_MASK;
                if (flag == 2 || t == 0) break;
                t--;
            }
            st_relaxed(&status[(size_t)tile * SMJ_RADIX + tid], RS_FLAG_INCL | ((excl + cnt) & RS_VAL_MASK));
        }
        s_goff[tid] = bin_base[tid] + excl - excl_local;   // mod 2^32: added to a local slot >= excl_local
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        const u32 d = (u32)(item[j] >> shift) & (SMJ_RADIX - 1);
        s_items[s_lbase[d] + my_cnt[d] + rank[j]] = item[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_IPT; k++) {
        const u32 idx = tid + k * RS_THREADS;
        if (idx < valid) {
            const u64 it = s_items[idx];
            const u32 d = (u32)(it >> shift) & (SMJ_RADIX - 1);
            out[s_goff[d] + idx] = it;
        }
    }
}

// Stand-alone digit histogram (used when the pairs did not come out of the select kernel).
__global__ void __launch_bounds__(256) radix_hist_kernel(const u64 *__restrict__ pairs, u32 n, u32 *hist)
{
    __shared__ u32 s_hist[SMJ_KEY_PASSES * SMJ_RADIX];
    for (u32 i = threadIdx.x; i < SMJ_KEY_PASSES * SMJ_RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u32 k = pair_key(pairs[i]);
#pragma unroll
        for (int d = 0; d < SMJ_KEY_PASSES; d++)
            atomicAdd(&s_hist[d * SMJ_RADIX + ((k >> (d * SMJ_RADIX_BITS)) & (SMJ_RADIX - 1))], 1u);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < SMJ_KEY_PASSES * SMJ_RADIX; i += blockDim.x) {
        const u32 v = s_hist[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// bases[p][b] = number of keys whose digit p is < b (exclusive scan per pass); one CTA of 4 x 256 threads.
__global__ void __launch_bounds__(SMJ_KEY_PASSES * SMJ_RADIX) radix_scan_kernel(const u32 *__restrict__ hist, u32 *bases)
{
    __shared__ u32 s_w[SMJ_KEY_PASSES * SMJ_RADIX / 32];
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u32 v = hist[tid];
    const u32 inc = warp_incl_scan(v);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    u32 wp = 0;
    for (u32 ww = (w & ~7u); ww < w; ww++) wp += s_w[ww];   // 8 warps per pass
    bases[tid] = wp + inc - v;
}

}  // namespace

size_t smj_radix_num_tiles(u32 n) { return ((size_t)n + RS_TILE - 1) / RS_TILE; }
size_t smj_radix_status_words(u32 n) { return smj_radix_num_tiles(n) * SMJ_RADIX; }

size_t smj_radix_scratch_bytes(u32 n)
{
    // [bases 4*256][counters 4 (+pad to 16)][status 4 passes]
    return (size_t)(SMJ_KEY_PASSES * SMJ_RADIX + 16) * 4 + SMJ_KEY_PASSES * smj_radix_status_words(n) * 4;
}

int smj_launch_radix_hist(SmjCtx *c, const u64 *d_pairs, u32 n, u32 *d_hist)
{
    if (n == 0) return SMJ_OK;
    u32 grid = (n + 256 * 16 - 1) / (256 * 16);
    if (grid > 148 * 8) grid = 148 * 8;
    radix_hist_kernel<<<grid, 256, 0, c->stream>>>(d_pairs, n, d_hist);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

int smj_launch_radix_scan(SmjCtx *c, const u32 *d_hist, u32 *d_bases)
{
    radix_scan_kernel<<<1, SMJ_KEY_PASSES * SMJ_RADIX, 0, c->stream>>>(d_hist, d_bases);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

int smj_launch_radix_pass(SmjCtx *c, const u64 *d_in, u64 *d_out, u32 n, int pass, const u32 *d_bases_pass,
                          u32 *d_status, u32 *d_tile_counter)
{
    if (!c->radix_attr_set) {   // function attributes are per device
        CUDA_TRY(cudaFuncSetAttribute(radix_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM));
        c->radix_attr_set = true;
    }
    const u32 tiles = (u32)smj_radix_num_tiles(n);
    radix_pass_kernel<<<tiles, RS_THREADS, RS_SMEM, c->stream>>>(d_in, d_out, n, 32 + pass * SMJ_RADIX_BITS,
                                                                 d_bases_pass, d_status, d_tile_counter, c->d_err);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

int smj_radix_sort_pairs(SmjCtx *c, u64 *buf_a, u64 *buf_b, u32 n, const u32 *d_hist, const u32 *h_hist,
                         u32 *d_scratch, u64 **d_sorted)
{
    // d_scratch: smj_radix_scratch_bytes(n), zeroed by the caller.
    *d_sorted = buf_a;
    if (n < 2) return SMJ_OK;
    if (n > SMJ_MAX_SORT_ROWS) return smj_set_error(SMJ_ETOOBIG, "radix sort of %u pairs exceeds 2^30 - 1", n);
    u32 *d_bases = d_scratch;
    u32 *d_counters = d_scratch + SMJ_KEY_PASSES * SMJ_RADIX;
    u32 *d_status = d_counters + 16;
    const size_t words = smj_radix_status_words(n);
    SMJ_TRY(smj_launch_radix_scan(c, d_hist, d_bases));
    u64 *src = buf_a, *dst = buf_b;
    for (int p = 0; p < SMJ_KEY_PASSES; p++) {
        bool trivial = false;   // every key has the same digit: the pass would be the identity permutation
        for (int b = 0; b < SMJ_RADIX; b++)
            if (h_hist[p * SMJ_RADIX + b] == n) { trivial = true; break; }
        if (trivial) continue;
        const bool timed = c->pass_count < SmjCtx::kMaxTimedPasses;
        if (timed) CUDA_TRY(cudaEventRecord(c->pass_ev[2 * c->pass_count], c->stream));
        SMJ_TRY(smj_launch_radix_pass(c, src, dst, n, p, d_bases + p * SMJ_RADIX, d_status + p * words, d_counters + p));
        if (timed) {
            CUDA_TRY(cudaEventRecord(c->pass_ev[2 * c->pass_count + 1], c->stream));
            c->pass_items[c->pass_count] = n;
            c->pass_count++;
        }
        u64 *t = src; src = dst; dst = t;
    }
    *d_sorted = src;
    return SMJ_OK;
}
