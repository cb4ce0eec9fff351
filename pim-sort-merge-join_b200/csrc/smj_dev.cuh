// smj_dev.cuh -- device-side primitives shared by the kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define FULL_MASK 0xffffffffu
// Bounded spin for every decoupled look-back: a bug must surface as SMJ_EINTERNAL, never as a hung GPU.
#define SMJ_SPIN_LIMIT (1u << 24)
#define SMJ_ERR_SPIN_SELECT 1u
#define SMJ_ERR_SPIN_RADIX  2u
#define SMJ_ERR_SPIN_JOIN   3u
#define SMJ_ERR_UNSORTED    6u   // smj_join / smj_join_count / smj_merge: an input is not sorted by its key column

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Look-back status words carry flag and value in ONE word, so relaxed gpu-scope accesses are sufficient:
// a reader either sees the old word (flag 0) or the complete new one.
__device__ __forceinline__ u32 ld_relaxed(const u32 *p)
{
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u32 *p, u32 v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(FULL_MASK, v, o);
        if (lane_id() >= (u32)o) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_sum(u64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// (flipped key << 32) | rowid.  Flipping the sign bit makes signed int32 order == unsigned order.
__device__ __forceinline__ u64 make_pair(int32_t key, u32 rowid) { return ((u64)((u32)key ^ 0x80000000u) << 32) | rowid; }
__device__ __forceinline__ u32 pair_key(u64 p) { return (u32)(p >> 32); }
__device__ __forceinline__ u32 pair_row(u64 p) { return (u32)p; }

// ---- searches over key-sorted pair arrays (global or shared memory)
// Elements taken from A among the first `diag` of merge(A,B), A first on equal keys.
__device__ __forceinline__ u32 merge_path(const u64 *A, u32 na, const u64 *B, u32 nb, u32 diag)
{
    u32 lo = diag > nb ? diag - nb : 0u, hi = diag < na ? diag : na;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (pair_key(A[mid]) <= pair_key(B[diag - 1 - mid])) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ u32 lower_bound_key(const u64 *A, u32 lo, u32 hi, u32 k)
{
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (pair_key(A[mid]) < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ u32 upper_bound_key(const u64 *A, u32 lo, u32 hi, u32 k)
{
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (pair_key(A[mid]) <= k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}


// ---- 64-bit tile status for scans with one value per tile (select, join): flag in the top 2 bits.
#define ST64_EMPTY 0ull
#define ST64_LOCAL (1ull << 62)
#define ST64_INCL  (2ull << 62)
#define ST64_VAL(x) ((x) & ((1ull << 62) - 1))
#define ST64_FLAG(x) ((x) >> 62)

// Warp-parallel decoupled look-back (one warp of the CTA calls this; all 32 lanes).
// Publishes this tile's aggregate, returns the exclusive prefix of all earlier tiles and publishes the
// inclusive prefix.  status[] must have been zeroed.  Tiles are numbered by an atomic ticket so every
// predecessor is already resident or finished (forward progress without relying on block scheduling order).
__device__ __forceinline__ u64 lookback_warp(u64 *status, u32 tile, u64 aggregate, u32 *err, u32 err_code)
{
    const u32 lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_relaxed(&status[0], ST64_INCL | aggregate);
        return 0;
    }
    if (lane == 0) st_relaxed(&status[tile], ST64_LOCAL | aggregate);
    u64 excl = 0;
    int look = (int)tile - 1;
    u32 spins = 0;
    bool bail = false;
    while (!bail) {
        int idx = look - (int)lane;
        u64 st = (idx >= 0) ? ld_relaxed(&status[idx]) : ST64_INCL;
        while (__any_sync(FULL_MASK, ST64_FLAG(st) == 0)) {
            if (ST64_FLAG(st) == 0) st = ld_relaxed(&status[idx]);
            if (++spins > SMJ_SPIN_LIMIT) {   // warp-uniform: every lane counts the same iterations
                if (lane == 0) atomicExch(err, err_code);
                bail = true;                   // still publish below so successors do not cascade the wait
                break;
            }
        }
        if (bail) break;
        u32 incl_mask = __ballot_sync(FULL_MASK, ST64_FLAG(st) == 2);
        int first = incl_mask ? (__ffs(incl_mask) - 1) : 31;
        u64 v = ((int)lane <= first) ? ST64_VAL(st) : 0ull;
        excl += warp_sum(v);
        if (incl_mask) break;
        look -= 32;
    }
    if (lane == 0) st_relaxed(&status[tile], ST64_INCL | (excl + aggregate));
    return excl;
}


// ---- one-CTA exclusive scan of per-tile counts (1024 threads): offsets[t] = sum of counts[0..t), returns the total to
// every thread.  Up to SCAN1_STAGE counts are staged in shared memory with eight independent coalesced loads in flight
// per thread (one L2/DRAM round trip), each thread then scans its contiguous chunk out of shared memory; the first
// version walked its chunk straight from global memory, one dependent round trip per element.
#define SCAN1_THREADS 1024
#define SCAN1_STAGE 8192
__device__ __forceinline__ u64 scan1_counts(const u32 *__restrict__ counts, u32 num_tiles, u64 *offsets, u32 *s_stage /*[SCAN1_STAGE]*/,
                                            u64 *s_w /*[32]*/, u64 base = 0 /*added to every offset*/)
{
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const bool staged = num_tiles <= SCAN1_STAGE;
    if (staged) {
        u32 v[SCAN1_STAGE / SCAN1_THREADS];
#pragma unroll
        for (int k = 0; k < SCAN1_STAGE / SCAN1_THREADS; k++) {
            const u32 i = k * SCAN1_THREADS + tid;
            v[k] = i < num_tiles ? counts[i] : 0u;
        }
#pragma unroll
        for (int k = 0; k < SCAN1_STAGE / SCAN1_THREADS; k++) s_stage[k * SCAN1_THREADS + tid] = v[k];
        __syncthreads();
    }
    const u32 chunk = (num_tiles + SCAN1_THREADS - 1) / SCAN1_THREADS;
    const u32 lo = tid * chunk < num_tiles ? tid * chunk : num_tiles;
    const u32 hi = lo + chunk < num_tiles ? lo + chunk : num_tiles;
    u64 sum = 0;
    if (staged) for (u32 i = lo; i < hi; i++) sum += s_stage[i];
    else for (u32 i = lo; i < hi; i++) sum += counts[i];
    u64 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(FULL_MASK, inc, o);
        if (lane >= (u32)o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    u64 run = base + inc - sum, total = 0;
#pragma unroll
    for (u32 ww = 0; ww < SCAN1_THREADS / 32; ww++) {
        const u64 x = s_w[ww];
        if (ww < w) run += x;
        total += x;
    }
    if (staged) for (u32 i = lo; i < hi; i++) { offsets[i] = run; run += s_stage[i]; }
    else for (u32 i = lo; i < hi; i++) { offsets[i] = run; run += counts[i]; }
    return total;
}

// ---- the same scan over many CTAs, for tables with more tiles than one CTA stages at once (ncu at the 500M-row config:
// the one-CTA scan of 488 K tile counts took 610 us, 12 % of the step, walking its chunks straight from global memory).
// Two launches of ceil(num_tiles / chunk) CTAs, chunk <= SCAN1_STAGE: block b first sums its chunk into blocksum[b];
// then every block adds up the sums of the blocks before it itself (a few hundred words at most, no block waits for
// another one) and scans its chunk out of shared memory.  scan_large_apply returns the running total through the
// block's chunk to every thread: the grand total on the last block.
__device__ __forceinline__ void scan_large_blocksum(const u32 *__restrict__ counts, u32 num_tiles, u32 chunk, u32 b, u64 *blocksum,
                                                    u64 *s_w /*[32]*/)
{
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u64 lo64 = (u64)b * chunk;
    const u32 lo = lo64 < (u64)num_tiles ? (u32)lo64 : num_tiles;
    const u32 n = num_tiles - lo < chunk ? num_tiles - lo : chunk;
    u64 sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN1_STAGE / SCAN1_THREADS; k++) {
        const u32 i = k * SCAN1_THREADS + tid;
        if (i < n) sum += counts[lo + i];
    }
    sum = warp_sum(sum);
    if (lane == 0) s_w[w] = sum;
    __syncthreads();
    if (tid == 0) {
        u64 t = 0;
#pragma unroll
        for (u32 ww = 0; ww < SCAN1_THREADS / 32; ww++) t += s_w[ww];
        blocksum[b] = t;
    }
}
__device__ __forceinline__ u64 scan_large_apply(const u32 *__restrict__ counts, u32 num_tiles, u32 chunk, u32 b, u64 *offsets,
                                                const u64 *__restrict__ blocksum, u32 *s_stage /*[SCAN1_STAGE]*/, u64 *s_w /*[32]*/)
{
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    u64 part = 0;
    for (u32 i = tid; i < b; i += SCAN1_THREADS) part += blocksum[i];
    part = warp_sum(part);
    if (lane == 0) s_w[w] = part;
    __syncthreads();
    u64 base = 0;
#pragma unroll
    for (u32 ww = 0; ww < SCAN1_THREADS / 32; ww++) base += s_w[ww];
    __syncthreads();                                   // scan1_counts reuses s_w
    const u64 lo64 = (u64)b * chunk;
    const u32 lo = lo64 < (u64)num_tiles ? (u32)lo64 : num_tiles;
    const u32 n = num_tiles - lo < chunk ? num_tiles - lo : chunk;
    return base + scan1_counts(counts + lo, n, offsets + lo, s_stage, s_w, base);
}

// ---- semi-join key bitmaps (smj_select.cu): bit bloom_hash(key) of a table's bitmap is set iff some surviving row of
// that table hashes there.  One multiplicative hash, 2^(32 - shift) bits.
__device__ __forceinline__ u32 bloom_hash(u32 flipped_key, u32 shift) { return (flipped_key * 0x9E3779B1u) >> shift; }

// ---- programmatic dependent launch (sm_90+): every kernel of the smj_run pipeline starts with PDL_ENTER().  When the
// launch carries cudaLaunchAttributeProgrammaticStreamSerialization, `wait` blocks until the preceding kernel has
// completed and its writes are visible, and `launch_dependents` lets the NEXT kernel's CTAs become resident (and park
// at their own wait) as this kernel's CTAs retire, so launch latency and ramp-up overlap the tail.  Because the trigger
// comes after the wait, at most two kernels are ever co-resident.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#define PDL_ENTER() pdl_enter()

// ---- mbarrier / bulk-copy (TMA) / named-barrier primitives for the warp-specialised pipelines (sm_90+ PTX)
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar)   // release at CTA scope
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity)   // acquire at CTA scope
{
    u32 ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk copy global -> shared (SASS UBLKCP); dst, src and bytes are multiples of 16; completion is counted in
// bytes on `bar` (pair it with mbar_expect_tx).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// ---- L2 residency hints.  The table scan is read exactly once: its bulk copies carry an evict_first policy so that the
// 160 MB+ stream does not push the pair slots the same kernel writes, or the semi-join bitmaps it probes, out of the
// 126 MB L2; bitmap words are read and set with evict_last.
__device__ __forceinline__ u64 l2_policy_evict_first()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_last()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar, u64 policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ u32 ld_nc_hint(const u32 *p, u64 policy)
{
    u32 v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void red_or_hint(u32 *p, u32 v, u64 policy)
{
    asm volatile("red.global.or.b32.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(v), "l"(policy) : "memory");
}

// barrier among a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(u32 id, u32 nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
