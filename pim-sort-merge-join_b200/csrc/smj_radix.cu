// smj_radix.cu -- LSD "onesweep" radix sort of (flipped key << 32 | row id) pairs by the key half.
//
// Replaces the per-DPU sort (sort-merge-join/sort_dpu.c:157-187 insertion sort per tasklet, :251-323 tasklet
// merge tree) and cpu_app.c:172-202.  The reference sort is a STABLE ascending insertion sort; every pass
// below is stable, so LSD over the key digits gives the identical order (ties stay in row-id order).
//
// One kernel launch per 8-bit digit of (key - smallest key) -- a device-resident sort plan says how many of the four
// the key range needs, the others exit at once: a CTA takes a tile through an atomic ticket, counts digits per warp in
// shared memory, publishes the tile's 256 digit counts, ranks its items against shared-memory peer masks (atomicOr, not
// __match_any_sync: see below), reorders the tile in shared memory, resolves the global offset of each digit with a
// batched decoupled look-back over the earlier tiles and writes each digit's run contiguously.
// Digit histograms for all passes come from the compaction copy (smj_select.cu) or radix_hist_kernel up front.
#include "smj_internal.h"
#include "smj_dev.cuh"
#include <stdlib.h>

// Optional per-phase cycle accounting for tools/radix_lab.cu (compiled out of libsmj.so).
#ifdef SMJ_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_INIT() long long ph_t = clock64()
#define PHASE(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(t_ - ph_t)); ph_t = t_; } } while (0)
#else
#define PHASE_INIT() do { } while (0)
#define PHASE(i) do { } while (0)
#endif

namespace {

#ifndef SMJ_RS_THREADS
#define SMJ_RS_THREADS 512
#endif
constexpr int RS_THREADS = SMJ_RS_THREADS;   // 512: two CTAs of 8192 pairs per SM; 256: four of 4096 (tools/build_variant.sh rs256)
constexpr int RS_IPT = 16;
constexpr int RS_CTAS_PER_SM = 1024 / RS_THREADS;
constexpr int RS_TILE = RS_THREADS * RS_IPT;   // 8192 pairs = 64 KB
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr u32 RS_FLAG_LOCAL = 1u << 30, RS_FLAG_INCL = 2u << 30, RS_VAL_MASK = (1u << 30) - 1;
static_assert(RS_THREADS >= SMJ_RADIX, "one thread per digit bin");
constexpr int RS_LB = 8;   // predecessors examined per look-back round

// shared memory: reordered tile | per-warp digit counters | per-warp peer masks | per-digit offsets | scan partials
constexpr size_t RS_SMEM = (size_t)RS_TILE * 8 + (size_t)RS_WARPS * SMJ_RADIX * 4 + (size_t)RS_WARPS * SMJ_RADIX * 4 +
                           SMJ_RADIX * 4 + 32 * 4;

// digit `pass` (0..3) of the key half of a pair, relative to the table's smallest key (sort plan): one IADD + one PRMT
__device__ __forceinline__ u32 pair_digit(u64 p, u32 sel, u32 kmin) { return __byte_perm((u32)(p >> 32) - kmin, 0u, sel); }

// Why not __match_any_sync: MATCH.ANY runs on the SM-wide ADU pipe at ~61 cycles per warp instruction on sm_100
// (profiles/r01_ubench_primitives.txt) and made the first version of this kernel ADU-bound (54 % pipe utilisation,
// 7 % DRAM).  A shared-memory atomicOr of the lane bit into a per-warp, per-digit mask word gives the same peer mask
// at ~2.7 cycles per warp instruction.
//
// Persistent CTAs of 16 warps take 8192-pair tiles through an atomic ticket (a tile's predecessors have all been
// started, which is what the look-back needs).  Per tile: per-warp digit counts (shared-memory atomics) -> publish
// the tile's 256 counts EARLY -> rank and reorder into shared memory -> batched look-back -> coalesced copy-out while
// the next tile's loads are already in flight.
template <bool FULL>
__device__ __forceinline__ void radix_count_tile(const u64 (&item)[RS_IPT], u32 sel, u32 kmin, u32 *my_cnt, u32 rel0, u32 valid, u32 ipt)
{
#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        const u32 d = pair_digit(item[j], sel, kmin);
        if (FULL || ((u32)j < ipt && rel0 + j * 32 < valid)) atomicAdd(&my_cnt[d], 1u);
    }
}

template <bool FULL>
__device__ __forceinline__ void radix_rank_tile(const u64 (&item)[RS_IPT], u32 sel, u32 kmin, u32 *my_cnt, u32 *my_mask, u64 *s_items,
                                                u32 rel0, u32 valid, u32 lane, u32 lt, u32 ipt)
{
#pragma unroll
    for (int j = 0; j < RS_IPT; j++) {
        if (!FULL && (u32)j >= ipt) break;   // warp-uniform: this launch's tiles hold ipt items per thread
        const u32 d = pair_digit(item[j], sel, kmin);
        const bool ok = FULL || rel0 + j * 32 < valid;
        u32 *mk = my_mask + d;
        if (ok) atomicOr(mk, 1u << lane);
        __syncwarp();
        const u32 peers = *mk;
        const u32 pre = my_cnt[d];
        __syncwarp();
        if (ok) {
            if ((peers >> lane) == 1u) {   // highest peer lane: reset the mask, advance the running slot
                *mk = 0;
                my_cnt[d] = pre + __popc(peers);
            }
            s_items[pre + __popc(peers & lt)] = item[j];
        }
        __syncwarp();                      // the reset lands before the next row's atomicOr
    }
}

// One sort problem of a pass (one table's pair array).  A launch covers up to two of them: their tiles share one
// ticket space (table 1's tiles first), so the CTAs run 2x as many tiles per launch and the wave-quantisation loss
// (2.06 tiles per CTA = 3 tile times at the 10M-row config) is halved; every problem keeps its own look-back chain.
// Pass p of a problem with npass planned passes reads buf[(npass - p) & 1] and writes the other buffer, so the last
// pass always lands in buf[0]; passes >= npass have no tiles.  Without a plan: four passes from key 0.
struct RadixProblem {
    u64 *buf[2];
    const u64 *n_dev;      // device count or null
    u32 n_max;
    const u32 *bin_base;   // [256] first output slot of each digit
    u32 *status, *status_next;
    const SmjSortPlan *plan;   // device, or null
    u32 cap_tiles;             // tiles the status arrays hold
};
struct RadixLaunch { RadixProblem p[2]; int nprob; };

__global__ void __launch_bounds__(RS_THREADS, RS_CTAS_PER_SM)
radix_pass_kernel(const RadixLaunch L, int pass, u32 *tile_counter, u32 *err, int dyn_tiles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_items = reinterpret_cast<u64 *>(smem_raw);              // RS_TILE, tile in digit order
    u32 *s_wcnt = reinterpret_cast<u32 *>(s_items + RS_TILE);      // [warp][256] counts -> running local slot
    u32 *s_mask = s_wcnt + RS_WARPS * SMJ_RADIX;                   // [warp][256] peer masks (self-resetting)
    u32 *s_goff = s_mask + RS_WARPS * SMJ_RADIX;                   // [256] global slot minus local slot per digit
    u32 *s_wsum = s_goff + SMJ_RADIX;                              // warp totals of the bin scan
    __shared__ u32 s_tile[2];
    __shared__ const u64 *s_in[2];   // per problem: this pass's source and destination buffers, smallest key
    __shared__ u64 *s_out[2];
    __shared__ u32 s_kmin[2];

    PDL_ENTER();
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u32 lt = lanemask_lt();
    const u32 sel = 0x4440u | (u32)pass;
    // sizes of both problems (device-resident counts), and the shared ticket space [0, tiles0) [tiles0, tiles0 + tiles1)
    u32 n0, n1 = 0;
    u32 raw0 = 0, raw1 = 0;   // pairs of each problem whatever its pass count: the tile size must be the same in every pass
    {
        u32 np0 = SMJ_KEY_PASSES, np1 = SMJ_KEY_PASSES, km0 = 0, km1 = 0;
        const u64 v = L.p[0].n_dev ? *L.p[0].n_dev : (u64)L.p[0].n_max;
        n0 = v < (u64)L.p[0].n_max ? (u32)v : L.p[0].n_max;
        raw0 = n0;
        if (L.p[0].plan) { np0 = L.p[0].plan->npass; km0 = L.p[0].plan->kmin; }
        if ((u32)pass >= np0) n0 = 0;
        if (L.nprob > 1) {
            const u64 v1 = L.p[1].n_dev ? *L.p[1].n_dev : (u64)L.p[1].n_max;
            n1 = v1 < (u64)L.p[1].n_max ? (u32)v1 : L.p[1].n_max;
            raw1 = n1;
            if (L.p[1].plan) { np1 = L.p[1].plan->npass; km1 = L.p[1].plan->kmin; }
            if ((u32)pass >= np1) n1 = 0;
        }
        if (tid == 0) {
            const bool odd0 = (np0 - (u32)pass) & 1u, odd1 = (np1 - (u32)pass) & 1u;
            s_in[0] = odd0 ? L.p[0].buf[1] : L.p[0].buf[0];
            s_out[0] = odd0 ? L.p[0].buf[0] : L.p[0].buf[1];
            s_in[1] = odd1 ? L.p[1].buf[1] : L.p[1].buf[0];
            s_out[1] = odd1 ? L.p[1].buf[0] : L.p[1].buf[1];
            s_kmin[0] = km0;
            s_kmin[1] = km1;
        }
    }
    // Tile size of THIS launch: ipt items per thread (1..16), tile = 512 * ipt pairs.  A fixed 8192-pair tile leaves the
    // last wave mostly idle when the pairs are few (3.3 M pairs = 407 tiles on 296 resident CTAs = two tile times for 1.4
    // waves of work, profiles/r01_ncu_full_final.txt); here the pairs of both problems are cut into k whole waves of
    // equal tiles, k = the waves 8192-pair tiles would need.  The status arrays hold cap tiles per problem, so the tile
    // only shrinks as far as every problem's tile count still fits.
    u32 ipt = RS_IPT;
    {
        const u64 total = (u64)raw0 + raw1;   // (a pass re-zeroes the other status array only for the tiles it has itself)
        const u64 slots = gridDim.x;
        const u64 k = (total + slots * RS_TILE - 1) / (slots * RS_TILE);
        // (measured at 3.3 M pairs, k = 2: 592 tiles of 5632 pairs ran 9 % SLOWER than 407 of 8192 -- the predicated path and the
        // per-tile fixed costs, look-back and digit scan, outweigh the idle half wave -- so only single-wave launches shrink)
        if (k == 1 && dyn_tiles) {
            ipt = (u32)((total + k * slots * RS_THREADS - 1) / (k * slots * RS_THREADS));
            if (ipt < 1) ipt = 1;
            if (ipt > (u32)RS_IPT) ipt = RS_IPT;
            while (ipt < (u32)RS_IPT &&
                   (((u64)raw0 + RS_THREADS * ipt - 1) / (RS_THREADS * ipt) > (u64)L.p[0].cap_tiles ||
                    (L.nprob > 1 && ((u64)raw1 + RS_THREADS * ipt - 1) / (RS_THREADS * ipt) > (u64)L.p[1].cap_tiles)))
                ipt++;
        }
    }
    const u32 tile_items = ipt * RS_THREADS;
    const u32 tiles0 = (n0 + tile_items - 1) / tile_items;
    const u32 all_tiles = tiles0 + (n1 + tile_items - 1) / tile_items;

    if (tid == 0) s_tile[0] = atomicAdd(tile_counter, 1u);
    for (u32 i = tid; i < RS_WARPS * SMJ_RADIX; i += RS_THREADS) s_mask[i] = 0;
    __syncthreads();
    u32 ticket = s_tile[0];
    int par = 0;

    // warp-striped layout: element order inside the tile is (warp, j, lane) == ascending index
    const u32 rel0 = w * 32 * ipt + lane;   // tile-relative index of item[0]
    u64 item[RS_IPT];
    if (ticket < all_tiles) {
        const bool second = ticket >= tiles0;
        const u64 *src_in = s_in[second];
        const u32 nn = second ? n1 : n0;
        const u32 g0 = (second ? ticket - tiles0 : ticket) * tile_items + rel0;
#pragma unroll
        for (int j = 0; j < RS_IPT; j++) item[j] = ((u32)j < ipt && g0 + j * 32 < nn) ? src_in[g0 + j * 32] : 0ull;
    }

    u32 *my_cnt = s_wcnt + w * SMJ_RADIX;
    u32 *my_mask = s_mask + w * SMJ_RADIX;
    PHASE_INIT();
    while (ticket < all_tiles) {
        const bool second = ticket >= tiles0;
        const RadixProblem &P = second ? L.p[1] : L.p[0];
        const u32 n = second ? n1 : n0;
        const u32 tile = second ? ticket - tiles0 : ticket;
        u64 *__restrict__ out = s_out[second];
        const u32 kmin = s_kmin[second];
        const u32 *__restrict__ bin_base = P.bin_base;
        u32 *status = P.status, *status_next = P.status_next;
        const u32 base = tile * tile_items;
        const u32 valid = (n - base < tile_items) ? (n - base) : tile_items;
        const bool full = valid == (u32)RS_TILE;   // (only 8192-pair tiles take the unpredicated path)

#pragma unroll
        for (int i = 0; i < RS_WARPS * SMJ_RADIX / RS_THREADS; i++) s_wcnt[i * RS_THREADS + tid] = 0;
        __syncthreads();
        PHASE(0);   // zero counters (+ wait for this tile's loads to be issued)

        // ---- early counts: per-warp digit histogram
        if (full) radix_count_tile<true>(item, sel, kmin, my_cnt, rel0, valid, ipt);
        else radix_count_tile<false>(item, sel, kmin, my_cnt, rel0, valid, ipt);
        __syncthreads();
        PHASE(1);   // load latency + count

        // ---- one thread per digit: totals over warps, publish the tile aggregate, scan digits
        u32 cnt = 0;
        if (tid < SMJ_RADIX) {
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ww++) cnt += s_wcnt[ww * SMJ_RADIX + tid];
            st_relaxed(&status[(size_t)tile * SMJ_RADIX + tid], (tile == 0 ? RS_FLAG_INCL : RS_FLAG_LOCAL) | cnt);
            status_next[(size_t)tile * SMJ_RADIX + tid] = 0;   // the next pass (next kernel) reuses the other array
            const u32 inc = warp_incl_scan(cnt);
            if (lane == 31) s_wsum[w] = inc;
            s_goff[tid] = inc - cnt;                           // exclusive within this warp's 32 digits
        }
        __syncthreads();
        if (tid < SMJ_RADIX) {
            u32 run = s_goff[tid];                             // -> exclusive over all digits: first local slot
            for (u32 ww = 0; ww < w; ww++) run += s_wsum[ww];
            s_goff[tid] = run;                                 // local base, turned into (global - local) after the look-back
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ww++) {
                const u32 c = s_wcnt[ww * SMJ_RADIX + tid];
                s_wcnt[ww * SMJ_RADIX + tid] = run;
                run += c;
            }
        }
        __syncthreads();
        PHASE(2);   // digit totals, publish, scan

        // ---- rank (stable: lanes in order, rows in order, warps in order) and reorder into shared memory
        if (full) radix_rank_tile<true>(item, sel, kmin, my_cnt, my_mask, s_items, rel0, valid, lane, lt, ipt);
        else radix_rank_tile<false>(item, sel, kmin, my_cnt, my_mask, s_items, rel0, valid, lane, lt, ipt);

        PHASE(3);   // rank + reorder (thread 0's view)
        // ---- next ticket, then this tile's look-back (predecessors published before they started ranking)
        if (tid == RS_THREADS - 1) s_tile[par ^ 1] = atomicAdd(tile_counter, 1u);
        if (tid < SMJ_RADIX) {
            u32 excl = 0;
            if (tile > 0) {
                // Batched look-back: RS_LB predecessors per round trip.  When a whole wave of tiles starts together
                // none of them has an inclusive prefix yet and tile k needs ~sqrt(2k / RS_LB) rounds, each an L2
                // round trip; one predecessor per round (the first version) cost ~30 rounds for the last tiles.
                int t = (int)tile - 1;
                u32 spins = 0;
                bool done = false;
                while (!done) {
                    u32 v[RS_LB];
#pragma unroll
                    for (int r = 0; r < RS_LB; r++)
                        v[r] = (t - r >= 0) ? ld_relaxed(&status[(size_t)(t - r) * SMJ_RADIX + tid]) : RS_FLAG_INCL;
                    int used = 0;
#pragma unroll
                    for (int r = 0; r < RS_LB; r++) {
                        if (!done && used == r) {
                            const u32 flag = v[r] >> 30;
                            if (flag != 0) {
                                excl += v[r] & RS_VAL_MASK;
                                used = r + 1;
                                if (flag == 2) done = true;
                            }
                        }
                    }
                    t -= used;
                    if (!done && used < RS_LB && ++spins > SMJ_SPIN_LIMIT) { atomicExch(err, SMJ_ERR_SPIN_RADIX); break; }
                }
                st_relaxed(&status[(size_t)tile * SMJ_RADIX + tid], RS_FLAG_INCL | ((excl + cnt) & RS_VAL_MASK));
            }
            s_goff[tid] = bin_base[tid] + excl - s_goff[tid];   // mod 2^32: added to a local slot >= the local base
        }
        PHASE(4);   // look-back (thread 0's digit)
        __syncthreads();
        PHASE(5);   // wait for the other warps (rank stragglers, other digits' look-back)

        // ---- issue the next tile's loads, then copy this tile out while they are in flight
        const u32 next = s_tile[par ^ 1];
        par ^= 1;
        if (next < all_tiles) {
            const bool nsecond = next >= tiles0;
            const u64 *src_in = s_in[nsecond];
            const u32 nn = nsecond ? n1 : n0;
            const u32 g0 = (nsecond ? next - tiles0 : next) * tile_items + rel0;
#pragma unroll
            for (int j = 0; j < RS_IPT; j++) item[j] = ((u32)j < ipt && g0 + j * 32 < nn) ? src_in[g0 + j * 32] : 0ull;
        }
#pragma unroll
        for (int k = 0; k < RS_IPT; k++) {
            const u32 idx = tid + k * RS_THREADS;
            if (!full && (u32)k >= ipt) break;
            if (full || idx < valid) {
                const u64 it = s_items[idx];
                out[s_goff[pair_digit(it, sel, kmin)] + idx] = it;
            }
        }
        PHASE(6);   // next loads issued + copy-out
        ticket = next;
        // the __syncthreads after the counter reset at the loop top orders these reads before the next reorder
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
struct RadixScanArgs { const u32 *hist[2]; u32 *bases[2]; };
__global__ void __launch_bounds__(SMJ_KEY_PASSES * SMJ_RADIX) radix_scan_kernel(const RadixScanArgs A)
{
    __shared__ u32 s_w[SMJ_KEY_PASSES * SMJ_RADIX / 32];
    PDL_ENTER();
    const u32 *__restrict__ hist = blockIdx.x ? A.hist[1] : A.hist[0];
    u32 *bases = blockIdx.x ? A.bases[1] : A.bases[0];
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

// tiles the status arrays of an n-pair sort hold: 8192-pair tiles, plus room for one wave of smaller ones (see radix_pass_kernel)
constexpr u32 RS_EXTRA_TILES = 320;
size_t smj_radix_num_tiles(u32 n) { return ((size_t)n + RS_TILE - 1) / RS_TILE; }
static size_t radix_cap_tiles(u32 n) { return smj_radix_num_tiles(n) + RS_EXTRA_TILES; }
size_t smj_radix_status_words(u32 n) { return radix_cap_tiles(n) * SMJ_RADIX; }

size_t smj_radix_scratch_bytes(u32 n)
{
    // [bases 4*256][counters 4 (+pad to 16)][status ping][status pong]; the caller zeroes everything once per sort,
    // each pass re-zeroes the other status array for its successor
    return (size_t)(SMJ_KEY_PASSES * SMJ_RADIX + 16) * 4 + 2 * smj_radix_status_words(n) * 4;
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
    RadixScanArgs A = {{d_hist, nullptr}, {d_bases, nullptr}};
    radix_scan_kernel<<<1, SMJ_KEY_PASSES * SMJ_RADIX, 0, c->stream>>>(A);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

static int launch_radix_pass(SmjCtx *c, const RadixLaunch &L, int pass, u32 *d_tile_counter)
{
    if (!c->radix_attr_set) {   // function attributes are per device
        CUDA_TRY(cudaFuncSetAttribute(radix_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM));
        c->radix_attr_set = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    static const int dyn_tiles = !(getenv("SMJ_RADIX_DYN_TILES") && atoi(getenv("SMJ_RADIX_DYN_TILES")) == 0);
    size_t tiles = 0;   // the smallest tile is one item per thread
    for (int i = 0; i < L.nprob; i++) tiles += dyn_tiles ? ((size_t)L.p[i].n_max + RS_THREADS - 1) / RS_THREADS : smj_radix_num_tiles(L.p[i].n_max);
    if (tiles == 0) return SMJ_OK;
    const u32 grid = tiles < (size_t)(sms * RS_CTAS_PER_SM) ? (u32)tiles : (u32)(sms * RS_CTAS_PER_SM);
    smj_launch(c, radix_pass_kernel, grid, RS_THREADS, RS_SMEM, L, pass, d_tile_counter, c->d_err, dyn_tiles);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

int smj_launch_radix_pass(SmjCtx *c, const u64 *d_in, u64 *d_out, const u64 *d_n, u32 n_max, int pass,
                          const u32 *d_bases_pass, u32 *d_status, u32 *d_status_next, u32 *d_tile_counter)
{
    RadixLaunch L = {};
    L.nprob = 1;
    // plan-less pass p reads buf[p & 1]
    if (pass & 1) L.p[0] = {{d_out, const_cast<u64 *>(d_in)}, d_n, n_max, d_bases_pass, d_status, d_status_next, nullptr, (u32)radix_cap_tiles(n_max)};
    else L.p[0] = {{const_cast<u64 *>(d_in), d_out}, d_n, n_max, d_bases_pass, d_status, d_status_next, nullptr, (u32)radix_cap_tiles(n_max)};
    return launch_radix_pass(c, L, pass, d_tile_counter);
}

// Stable LSD sort of one or two pair arrays at once (problem i: the first *d_n[i] (<= n_max[i]) pairs of buf_a[i], buf_b[i]
// its ping-pong partner, d_hist[i] its 4x256 digit histogram, d_scratch[i] smj_radix_scratch_bytes(n_max[i]) zeroed bytes).
// Four passes, so each result is back in buf_a[i].  Nothing here waits for the host: the pair counts are read on the
// device, and a digit every key shares costs one identity-permutation pass instead of a host round trip.
int smj_radix_sort_pairs_n(SmjCtx *c, int nprob, u64 *const *buf_a, u64 *const *buf_b, const u64 *const *d_n, const u32 *n_max,
                           const u32 *const *d_hist, u32 *const *d_scratch, const SmjSortPlan *const *d_plan)
{
    if (nprob < 1 || nprob > 2) return smj_set_error(SMJ_EINVAL, "radix sort of %d arrays at once (1 or 2)", nprob);
    u32 *d_bases[2], *d_counters[2], *d_status[2][2];
    RadixScanArgs SA = {};
    int live = 0, idx[2];
    for (int i = 0; i < nprob; i++) {
        if (n_max[i] > SMJ_MAX_SORT_ROWS) return smj_set_error(SMJ_ETOOBIG, "radix sort of %u pairs exceeds 2^30 - 1", n_max[i]);
        if (n_max[i] < 2) continue;
        d_bases[live] = d_scratch[i];
        d_counters[live] = d_scratch[i] + SMJ_KEY_PASSES * SMJ_RADIX;
        d_status[live][0] = d_counters[live] + 16;
        d_status[live][1] = d_counters[live] + 16 + smj_radix_status_words(n_max[i]);
        SA.hist[live] = d_hist[i];
        SA.bases[live] = d_bases[live];
        idx[live++] = i;
    }
    if (live == 0) return SMJ_OK;
    smj_launch(c, radix_scan_kernel, live, SMJ_KEY_PASSES * SMJ_RADIX, 0, SA);   // one CTA per problem
    KERNEL_CHECK(c);
    // one event pair around the four back-to-back passes (per-pass event records cost more stream time than they measure)
    const bool timed = smj_stage_events() && c->pass_count < SmjCtx::kMaxTimedPasses;
    if (timed) CUDA_TRY(smj_event_record(c->pass_ev[2 * c->pass_count], c->stream));
    for (int p = 0; p < SMJ_KEY_PASSES; p++) {
        RadixLaunch L = {};
        L.nprob = live;
        for (int k = 0; k < live; k++) {
            const int i = idx[k];
            L.p[k] = {{buf_a[i], buf_b[i]}, d_n[i], n_max[i], d_bases[k] + p * SMJ_RADIX, d_status[k][p & 1], d_status[k][(p + 1) & 1],
                      d_plan ? d_plan[i] : nullptr, (u32)radix_cap_tiles(n_max[i])};
        }
        SMJ_TRY(launch_radix_pass(c, L, p, d_counters[0] + p));   // the shared ticket counter lives in the first problem's scratch
    }
    if (timed) {
        CUDA_TRY(smj_event_record(c->pass_ev[2 * c->pass_count + 1], c->stream));
        u64 items = 0;
        for (int k = 0; k < live; k++) items += n_max[idx[k]];
        c->pass_items[c->pass_count] = (u32)(items > 0xffffffffull ? 0xffffffffull : items);
        c->pass_count++;
    }
    return SMJ_OK;
}

int smj_radix_sort_pairs(SmjCtx *c, u64 *buf_a, u64 *buf_b, const u64 *d_n, u32 n_max, const u32 *d_hist, u32 *d_scratch)
{
    return smj_radix_sort_pairs_n(c, 1, &buf_a, &buf_b, &d_n, &n_max, &d_hist, &d_scratch);
}

// Loads this file's pipeline kernels on the current device.  CUDA loads a kernel lazily at its first launch, and that load can
// wait for other GPUs' running kernels when peer access is enabled; a process that drives several GPUs (smj_dist.cu) must
// not meet such a load while another rank's kernel spins on this rank's flags, so it loads everything up front.
void smj_preload_radix(void)
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, radix_pass_kernel);
    cudaFuncGetAttributes(&a, radix_hist_kernel);
    cudaFuncGetAttributes(&a, radix_scan_kernel);
    cudaGetLastError();
}
