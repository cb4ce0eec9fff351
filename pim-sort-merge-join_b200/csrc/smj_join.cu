// smj_join.cu -- merge join of two key-sorted (flipped key << 32 | row id) arrays.
//
// Replaces the DPU join kernel (sort-merge-join/join.c:99-118 binary-search range split per tasklet, :153-177
// count pass, :185-191 prefix, :205-248 write pass), the host range split (app.c:585-633) and cpu_app.c:204-266.
//
// Semantics are the reference's "zipper" (cpu_app.c:213-218): on equal keys BOTH cursors advance, so a key
// occurring cL times on the left and cR times on the right yields min(cL,cR) rows, the i-th left duplicate
// paired with the i-th right duplicate.  In closed form, for left position i with key k:
//     t  = i - lower_bound(KL, k)          rank inside its left run
//     j  = lower_bound(KR, k) + t          candidate partner
//     emit (i, j)  iff  j < m2 and KR[j] == k
// Work is split into equal tiles of the merged sequence by merge-path co-ranking (left first on ties), so
// lower_bound(KR, k) is simply the right cursor when the left element is consumed; duplicates that straddle
// a tile are resolved with lower-bound searches (one global search per tile, precomputed by the partition
// kernel).  Count -> scan -> write: each tile writes its (left row id, right row id) matches into its own slot
// and its count into tile_count[]; a one-CTA scan turns the counts into output offsets and a compaction copy makes
// the dense match list (no CTA waits for another one); a last kernel materialises the joined rows (all left
// columns, then right columns except key2: cpu_app.c:240-251) with coalesced stores.
#include "smj_internal.h"
#include "smj_dev.cuh"
#include <stdlib.h>

namespace {

#ifndef SMJ_JN_GRID
#define SMJ_JN_GRID 4
#endif
constexpr int JN_THREADS = 256;
constexpr int JN_VT = 8;
constexpr int JN_TILE = JN_THREADS * JN_VT;   // merged elements per tile
constexpr int JN_WARPS = JN_THREADS / 32;

// m1/m2 live in device memory when the join follows select+sort without a host round trip.
__device__ __forceinline__ void load_counts(const u64 *counts, u32 m1_max, u32 m2_max, u32 &m1, u32 &m2)
{
    m1 = m1_max; m2 = m2_max;
    if (counts) {
        const u64 a = counts[0], b = counts[1];
        m1 = a < (u64)m1_max ? (u32)a : m1_max;
        m2 = b < (u64)m2_max ? (u32)b : m2_max;
    }
}

// One warp per tile boundary: part[t] = left elements before merged position t*TILE (left first on ties),
// found with a 32-ary search on the merge-path diagonal (5 dependent round trips for 2^25 elements instead of
// 25); runstart[t] = first left position holding the key of L[part[t]] (the only left run that can begin
// before tile t).
__global__ void __launch_bounds__(256)
join_partition_kernel(const u64 *__restrict__ L, const u64 *__restrict__ R, const u64 *__restrict__ counts, u32 m1_max,
                      u32 m2_max, u32 *part, u32 *runstart)
{
    PDL_ENTER();
    u32 m1, m2;
    load_counts(counts, m1_max, m2_max, m1, m2);
    const u64 total = (u64)m1 + m2;
    const u32 num_tiles = (u32)((total + JN_TILE - 1) / JN_TILE);
    const u32 t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const u32 lane = lane_id();
    if (t > num_tiles) return;
    u64 d64 = (u64)t * JN_TILE;
    if (d64 > total) d64 = total;
    const u32 diag = (u32)d64;
    u32 lo = diag > m2 ? diag - m2 : 0u, hi = diag < m1 ? diag : m1;
    while (lo < hi) {   // P(mid) = key(L[mid]) <= key(R[diag-1-mid]) is true on a prefix of [lo, hi)
        const u32 step = (hi - lo + 31u) >> 5;
        const u32 mid = lo + lane * step;
        bool p = false;
        if (mid < hi) p = pair_key(L[mid]) <= pair_key(R[diag - 1 - mid]);
        const u32 c = __popc(__ballot_sync(FULL_MASK, p));
        if (c == 0) { hi = lo; break; }
        const u32 nlo = lo + (c - 1) * step + 1;
        const u64 nhi = (u64)lo + (u64)c * step;
        if (c < 32 && nhi < hi) hi = (u32)nhi;
        lo = nlo;
    }
    const u32 a = lo;
    if (lane == 0) {
        part[t] = a;
        if (t < num_tiles) {
            u32 rs = a;
            if (a > 0 && a < m1) {
                const u32 k = pair_key(L[a]);
                if (pair_key(L[a - 1]) == k) {
                    // gallop backwards to bracket the start of the run: a plain lower bound over [0, a) costs ~28 dependent
                    // loads whatever the run length (ncu at 200M x 200M with ~10 rows per key: 652 us for this kernel)
                    u32 off = 2;
                    while (off <= a && pair_key(L[a - off]) == k) off <<= 1;
                    const u32 lo = off <= a ? a - off + 1 : 0u;    // L[lo - 1] < k (or lo == 0); L[a - off / 2] == k
                    rs = lower_bound_key(L, lo, a - (off >> 1), k);
                }
            }
            runstart[t] = rs;
        }
    }
}

__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Tiles of JN_TILE merged elements, dealt round-robin to persistent CTAs.  No CTA waits for another: a tile's matches
// go to its own slot of `matches` (slot t starts at t * JN_TILE) with the count in tile_count[t]; the output offsets
// come from tile_scan_kernel afterwards.  (The first version resolved them in-kernel with a decoupled look-back and
// spent 37 % of its stall samples behind that barrier.)  The next tile's elements are fetched with cp.async into the
// other half of a double buffer while the current tile is merged.
template <int MODE>
__global__ void __launch_bounds__(JN_THREADS)
join_match_kernel(const u64 *__restrict__ L, const u64 *__restrict__ R, const u64 *__restrict__ counts, u32 m1_max,
                  u32 m2_max, const u32 *__restrict__ part, const u32 *__restrict__ runstart, uint2 *__restrict__ matches,
                  u32 *__restrict__ tile_count, u64 *count, uint2 *__restrict__ many_runs, u32 *err)
{
    PDL_ENTER();
    u32 m1, m2;
    load_counts(counts, m1_max, m2_max, m1, m2);
    const u32 num_tiles = (m1 == 0 || m2 == 0) ? 0u : (u32)(((u64)m1 + m2 + JN_TILE - 1) / JN_TILE);
    constexpr int BUF = JN_TILE + 4;                      // [prev left element][left segment][right segment + halo]
    __shared__ __align__(16) u64 s_buf[2][BUF];
    __shared__ u32 s_wsum[JN_WARPS];
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u64 total = (u64)m1 + m2;
    u64 many_total = 0;

    // geometry of a tile from its partition entries
    struct Geo { u32 a0, na, b0, nb, nbh, rs; };
    auto geo = [&](u32 tile, u32 a0, u32 a1, u32 rs) {
        Geo g;
        const u64 d0 = (u64)tile * JN_TILE;
        const u64 d1 = (d0 + JN_TILE < total) ? d0 + JN_TILE : total;
        g.a0 = a0; g.na = a1 - a0;
        g.b0 = (u32)(d0 - a0);
        const u32 b1 = (u32)(d1 - a1);
        g.nb = b1 - g.b0;
        g.nbh = g.nb + (b1 < m2 ? 1u : 0u);               // right segment plus one halo element
        g.rs = rs;
        // Inputs that are NOT sorted by their keys (a caller's mistake: smj_join / smj_join_count take sorted tables) give
        // co-ranks that do not increase from tile to tile; such a tile is skipped and reported instead of being fetched
        // past the shared-memory buffer.
        if (a1 < a0 || g.na > (u32)JN_TILE || b1 < g.b0 || g.nb > (u32)JN_TILE || g.na + g.nb > (u32)JN_TILE) {
            g.na = 0; g.nb = 0; g.nbh = 0;
            if (threadIdx.x == 0) atomicExch(err, SMJ_ERR_UNSORTED);
        }
        return g;
    };
    auto fetch = [&](const Geo &g, u64 *buf) {            // buf[0] = L[a0-1] (if any), buf[1..] = left, then right
        const u32 lead = g.a0 > 0 ? 1u : 0u;
        const u64 *srcL = L + g.a0 - lead;
        for (u32 i = tid; i < g.na + lead; i += JN_THREADS) cp_async8(buf + 1 - lead + i, srcL + i);
        const u64 *srcR = R + g.b0;
        u64 *dstR = buf + 1 + g.na;
        for (u32 i = tid; i < g.nbh; i += JN_THREADS) cp_async8(dstR + i, srcR + i);
        cp_async_commit();
    };

    u32 tile = blockIdx.x;
    u32 cur = 0;
    Geo g = {}, gn = {};
    if (tile < num_tiles) {
        g = geo(tile, part[tile], part[tile + 1], runstart[tile]);
        fetch(g, s_buf[0]);
        const u32 nt = tile + gridDim.x;
        if (nt < num_tiles) gn = geo(nt, part[nt], part[nt + 1], runstart[nt]);
    }
    while (tile < num_tiles) {
        cp_async_wait_all();
        __syncthreads();                                  // tile data visible; the other buffer is free again
        const u32 next = tile + gridDim.x;
        if (next < num_tiles) fetch(gn, s_buf[cur ^ 1]);
        const Geo gc = g;
        g = gn;
        {
            const u32 n2 = next + gridDim.x;              // partition entries two tiles ahead (plain loads, consumed next round)
            if (n2 < num_tiles) gn = geo(n2, part[n2], part[n2 + 1], runstart[n2]);
        }
        const u32 a0 = gc.a0, na = gc.na, b0 = gc.b0, nb = gc.nb, nbh = gc.nbh, tile_rs = gc.rs;
        u64 *buf = s_buf[cur];
        const u64 *sA = buf + 1, *sB = buf + 1 + na;
        const bool has_prev = a0 > 0;
        const u32 prevk = has_prev ? pair_key(buf[0]) : 0u;

        const u32 ntile = na + nb;
        const u32 diag = (tid * JN_VT < ntile) ? tid * JN_VT : ntile;
        u32 a = merge_path(sA, na, sB, nb, diag);
        u32 b = diag - a;
        const u32 steps = (ntile - diag < (u32)JN_VT) ? ntile - diag : (u32)JN_VT;
        u64 va = a < na ? sA[a] : 0ull, vb = b < nb ? sB[b] : 0ull;
        uint2 mt[JN_VT];
        u32 mmask = 0;
        bool have = false;
        u32 runk = 0, run_val = 0;   // zip: rank of the current left element inside its key run; many: the right run length
#pragma unroll
        for (int st = 0; st < JN_VT; st++) {
            if (st < (int)steps) {
                const u32 ka = pair_key(va), kb = pair_key(vb);
                const bool takeA = (b >= nb) || (a < na && ka <= kb);
                if (takeA) {
                    const u32 k = ka;
                    if (MODE == SMJ_JOIN_ZIP) {
                        if (have) {
                            run_val = (runk == k) ? run_val + 1 : 0u;   // the previous left element was this thread's
                        } else {                                        // first left element of this thread
                            u32 rs;
                            if (a == 0) rs = (has_prev && prevk == k) ? tile_rs : a0;
                            else if (pair_key(sA[a - 1]) != k) rs = a0 + a;
                            else {
                                const u32 lo = lower_bound_key(sA, 0, a, k);
                                rs = (lo == 0 && has_prev && prevk == k) ? tile_rs : a0 + lo;
                            }
                            run_val = (a0 + a) - rs;
                            have = true;
                        }
                        runk = k;
                        // partner: the run_val-th right element at or after the right cursor, if its key is k
                        const u32 jj = b + run_val;                     // tile-relative, halo included
                        u64 rp = vb;
                        bool ok = b < nb;
                        if (run_val != 0u || !ok) {
                            ok = true;
                            if (jj < nbh) rp = sB[jj];
                            else if ((u64)b0 + jj < (u64)m2) rp = R[(u64)b0 + jj];
                            else ok = false;
                        }
                        if (ok && pair_key(rp) == k) {
                            mt[st] = make_uint2(pair_row(va), pair_row(rp));
                            mmask |= 1u << st;
                        }
                    } else {
                        if (!(have && runk == k)) {
                            const u32 e = upper_bound_key(sB, b, nbh, k);
                            u32 ub = b0 + e;
                            if (e == nbh && ub < m2 && pair_key(R[ub]) == k) ub = upper_bound_key(R, ub, m2, k);
                            run_val = ub - (b0 + b); runk = k; have = true;
                        }
                        many_total += run_val;
                        // per left element: (first right position of its key, right run length) for the expand step
                        if (many_runs) many_runs[a0 + a] = make_uint2(b0 + b, run_val);
                    }
                    a++;
                    va = a < na ? sA[a] : 0ull;
                } else {
                    b++;
                    vb = b < nb ? sB[b] : 0ull;
                }
            }
        }
        if (MODE == SMJ_JOIN_ZIP) {
            const u32 cnt = __popc(mmask);
            const u32 inc = warp_incl_scan(cnt);
            if (lane == 31) s_wsum[w] = inc;
            __syncthreads();   // also: every thread is done reading sA/sB, the buffer may be reused for staging
            u32 wp = 0, tile_total = 0;
#pragma unroll
            for (int ww = 0; ww < JN_WARPS; ww++) {
                const u32 v = s_wsum[ww];
                if (ww < (int)w) wp += v;
                tile_total += v;
            }
            if (tid == 0) tile_count[tile] = tile_total;
            uint2 *s_out = reinterpret_cast<uint2 *>(buf);
            u32 o = wp + inc - cnt;
#pragma unroll
            for (int st = 0; st < JN_VT; st++)
                if ((mmask >> st) & 1u) s_out[o++] = mt[st];
            __syncthreads();
            uint2 *dst = matches + (size_t)tile * JN_TILE;
            for (u32 i = tid; i < tile_total; i += JN_THREADS) dst[i] = s_out[i];
        }
        tile = next;
        cur ^= 1;
    }
    if (MODE != SMJ_JOIN_ZIP) {
        many_total = warp_sum(many_total);
        if (lane == 0 && many_total) atomicAdd(count, many_total);
    }
}

// Exclusive scan of per-tile counts (one CTA): offsets[t] = sum of counts[0..t), *total = sum of all.
constexpr int JS_THREADS = SCAN1_THREADS;
__global__ void __launch_bounds__(JS_THREADS)
join_scan_kernel(const u32 *__restrict__ tile_count, u32 num_tiles, u64 *offsets, u64 *total, const u64 *__restrict__ counts,
                 u32 m1_max, u32 m2_max)
{
    __shared__ u32 s_stage[SCAN1_STAGE];
    __shared__ u64 s_w[JS_THREADS / 32];
    PDL_ENTER();
    // the tiles the match kernel actually used (device-resident sizes), not the host's upper bound
    u32 m1, m2;
    load_counts(counts, m1_max, m2_max, m1, m2);
    const u32 used = (m1 == 0 || m2 == 0) ? 0u : (u32)(((u64)m1 + m2 + JN_TILE - 1) / JN_TILE);
    if (used < num_tiles) num_tiles = used;
    const u64 t = scan1_counts(tile_count, num_tiles, offsets, s_stage, s_w);
    if (threadIdx.x == 0) *total = t;
}

// The same over many CTAs (more tiles than one CTA stages at once: scan_large_* in smj_dev.cuh); blocksum: one u64 per CTA.
__global__ void __launch_bounds__(JS_THREADS)
join_blocksum_kernel(const u32 *__restrict__ tile_count, u32 num_tiles, u32 chunk, u64 *blocksum, const u64 *__restrict__ counts,
                     u32 m1_max, u32 m2_max)
{
    __shared__ u64 s_w[JS_THREADS / 32];
    PDL_ENTER();
    u32 m1, m2;
    load_counts(counts, m1_max, m2_max, m1, m2);
    const u32 used = (m1 == 0 || m2 == 0) ? 0u : (u32)(((u64)m1 + m2 + JN_TILE - 1) / JN_TILE);
    if (used < num_tiles) num_tiles = used;
    scan_large_blocksum(tile_count, num_tiles, chunk, blockIdx.x, blocksum, s_w);
}

__global__ void __launch_bounds__(JS_THREADS)
join_apply_kernel(const u32 *__restrict__ tile_count, u32 num_tiles, u32 chunk, u64 *offsets, const u64 *__restrict__ blocksum, u64 *total,
                  const u64 *__restrict__ counts, u32 m1_max, u32 m2_max)
{
    __shared__ u32 s_stage[SCAN1_STAGE];
    __shared__ u64 s_w[JS_THREADS / 32];
    PDL_ENTER();
    u32 m1, m2;
    load_counts(counts, m1_max, m2_max, m1, m2);
    const u32 used = (m1 == 0 || m2 == 0) ? 0u : (u32)(((u64)m1 + m2 + JN_TILE - 1) / JN_TILE);
    if (used < num_tiles) num_tiles = used;
    const u64 t = scan_large_apply(tile_count, num_tiles, chunk, blockIdx.x, offsets, blocksum, s_stage, s_w);
    if (blockIdx.x + 1 == gridDim.x && threadIdx.x == 0) *total = t;   // the last block's running total is the grand total
}

// dense[offsets[t] + i] = slots[t * JN_TILE + i], i < tile_count[t]: the matches in result order, contiguous.
__global__ void __launch_bounds__(256)
join_compact_kernel(const uint2 *__restrict__ slots, const u32 *__restrict__ tile_count, const u64 *__restrict__ tile_off, u32 num_tiles,
                    uint2 *__restrict__ dense, const u64 *__restrict__ counts, u32 m1_max, u32 m2_max, u64 dense_cap)
{
    PDL_ENTER();
    {
        u32 m1, m2;
        load_counts(counts, m1_max, m2_max, m1, m2);
        const u32 used = (m1 == 0 || m2 == 0) ? 0u : (u32)(((u64)m1 + m2 + JN_TILE - 1) / JN_TILE);
        if (used < num_tiles) num_tiles = used;
    }
    const u32 lane = threadIdx.x & 31u;
    const u32 warps = gridDim.x * (blockDim.x >> 5);
    for (u32 t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < num_tiles; t += warps) {   // one warp per tile
        const u32 cnt = tile_count[t];
        const uint2 *src = slots + (size_t)t * JN_TILE;
        const u64 off = tile_off[t];
        uint2 *dst = dense + off;
        // (sorted inputs give at most min(m1, m2) matches, which is what `dense` holds; unsorted ones can give more: never past the end)
        for (u32 i = lane; i < cnt; i += 32) if (off + i < dense_cap) dst[i] = src[i];
    }
}

#ifndef SMJ_MT_ILP
#define SMJ_MT_ILP 8
#endif
#ifndef SMJ_MT_RPB
#define SMJ_MT_RPB 1024
#endif
#ifndef SMJ_MT_GRID
#define SMJ_MT_GRID 6
#endif
constexpr int MT_THREADS = 256;
constexpr int MT_ILP = SMJ_MT_ILP;   // gather tasks in flight per thread
constexpr int MT_SMEM_CELLS = 8192;   // 32 KB staging: rows per block = MT_SMEM_CELLS / c_out (<= 1024)

// out[o] = t1[matches[o].x][0..c1) ++ t2[matches[o].y][c != key2]   (cpu_app.c:240-251, join.c:214-229)
// One thread gathers one source row (128-bit loads when the row is a whole number of 16-byte words and aligned,
// i.e. the 4- and 8-column configs), rows are assembled in shared memory and leave as one contiguous,
// fully coalesced block.  The row count comes from device memory so no host round trip precedes the launch.
template <bool VEC>
__global__ void __launch_bounds__(MT_THREADS)
join_materialize_kernel(const uint2 *__restrict__ matches, const u64 *__restrict__ nj_dev, int64_t nj_max,
                        const int32_t *__restrict__ t1, int c1, const int32_t *__restrict__ t2, int c2, int key2,
                        int32_t *__restrict__ out_direct, int32_t *const *__restrict__ out_indirect, int rows_per_block,
                        const u32 *__restrict__ use_store, const int32_t *__restrict__ store1, const int32_t *__restrict__ store2,
                        const int32_t *const *__restrict__ tables_ind)
{
    __shared__ __align__(16) int32_t s_out[MT_SMEM_CELLS];
    PDL_ENTER();
    if (tables_ind) { t1 = tables_ind[0]; t2 = tables_ind[1]; }   // (graph replay on new tables of the same shape)
    if (use_store) {   // the match list's row ids of a table are dense indices into its row store when the flag is set
        if (use_store[0]) t1 = store1;
        if (use_store[1]) t2 = store2;
    }
    int32_t *__restrict__ out = out_indirect ? *out_indirect : out_direct;
    int64_t nj = nj_max;
    if (nj_dev) { const u64 v = *nj_dev; nj = v < (u64)nj_max ? (int64_t)v : nj_max; }
    const int c_out = c1 + c2 - 1;
    const int64_t nblocks = (nj + rows_per_block - 1) / rows_per_block;
    {
      int32_t *out_slot = out;
      for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t row0 = blk * rows_per_block;
        const int nrows = (int)((nj - row0 < rows_per_block) ? (nj - row0) : rows_per_block);
        __syncthreads();   // previous block's copy-out is done with s_out
        // 2 * nrows gather tasks: task i < nrows is the left row of match i, the others the right rows.  MT_ILP tasks
        // per thread are in flight at once (match entries first, then the dependent row loads): the random 16..32-byte
        // row reads are latency-bound, so memory-level parallelism per thread is what sets the rate.
        for (int task0 = threadIdx.x; task0 < 2 * nrows; task0 += MT_ILP * MT_THREADS) {
            uint2 m[MT_ILP];
            const int32_t *src[MT_ILP];
            int32_t *dst[MT_ILP];
            int cc[MT_ILP];
            bool right[MT_ILP], live[MT_ILP];
#pragma unroll
            for (int u = 0; u < MT_ILP; u++) {
                const int task = task0 + u * MT_THREADS;
                live[u] = task < 2 * nrows;
                right[u] = task >= nrows;
                const int r = right[u] ? task - nrows : task;
                m[u] = live[u] ? matches[row0 + r] : make_uint2(0u, 0u);
                dst[u] = s_out + r * c_out + (right[u] ? c1 : 0);
            }
#pragma unroll
            for (int u = 0; u < MT_ILP; u++) {
                cc[u] = right[u] ? c2 : c1;
                src[u] = right[u] ? t2 + (size_t)m[u].y * c2 : t1 + (size_t)m[u].x * c1;
            }
            if (VEC) {
                int4 v[MT_ILP];
#pragma unroll
                for (int u = 0; u < MT_ILP; u++)
                    if (live[u]) v[u] = __ldg(reinterpret_cast<const int4 *>(src[u]));
#pragma unroll
                for (int u = 0; u < MT_ILP; u++) {
                    if (!live[u]) continue;
                    const int skip = right[u] ? key2 : -1;      // right rows drop column key2
                    for (int q = 0;;) {
                        const int32_t vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int col = 4 * q + e;
                            if (col != skip) dst[u][col - ((skip >= 0 && col > skip) ? 1 : 0)] = vv[e];
                        }
                        if (++q >= cc[u] / 4) break;
                        v[u] = __ldg(reinterpret_cast<const int4 *>(src[u]) + q);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < MT_ILP; u++) {
                    if (!live[u]) continue;
                    const int skip = right[u] ? key2 : -1;
                    for (int q = 0; q < cc[u]; q++)
                        if (q != skip) dst[u][q - ((skip >= 0 && q > skip) ? 1 : 0)] = __ldg(src[u] + q);
                }
            }
        }
        __syncthreads();
        const int ncell = nrows * c_out;
        int32_t *o = out_slot + row0 * c_out;
        for (int cell = threadIdx.x; cell < ncell; cell += MT_THREADS) o[cell] = s_out[cell];
      }
    }
}

}  // namespace

// ---- many-to-many expansion (SMJ_JOIN_MANY): runs[i] = (first right position, right run length) of left element i.
namespace {
constexpr int EX_BLOCK = 4096;   // left elements per scan block

// block_sum[b] = sum of run lengths of left elements [b*EX_BLOCK, (b+1)*EX_BLOCK)
__global__ void __launch_bounds__(256) many_blocksum_kernel(const uint2 *__restrict__ runs, u32 m1, u64 *block_sum)
{
    __shared__ u64 s_w[8];
    const u32 base = blockIdx.x * EX_BLOCK;
    u64 sum = 0;
    for (u32 i = threadIdx.x; i < (u32)EX_BLOCK; i += 256) if (base + i < m1) sum += runs[base + i].y;
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (int w = 0; w < 8; w++) t += s_w[w]; block_sum[blockIdx.x] = t; }
}

// exclusive scan of the block sums (one CTA, chunk per thread) and the grand total
__global__ void __launch_bounds__(1024) many_blockscan_kernel(u64 *block_sum, u32 nblocks, u64 *total)
{
    __shared__ u64 s_w[32];
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u32 chunk = (nblocks + 1023) / 1024;
    const u32 lo = tid * chunk < nblocks ? tid * chunk : nblocks, hi = lo + chunk < nblocks ? lo + chunk : nblocks;
    u64 sum = 0;
    for (u32 i = lo; i < hi; i++) sum += block_sum[i];
    u64 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= (u32)o) inc += t; }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    u64 run = inc - sum;
    for (u32 ww = 0; ww < w; ww++) run += s_w[ww];
    if (tid == 1023) *total = run + sum;
    for (u32 i = lo; i < hi; i++) { const u64 v = block_sum[i]; block_sum[i] = run; run += v; }
}

// off[i] = output row of the first pair of left element i (exclusive scan inside the block + the block's offset)
__global__ void __launch_bounds__(256) many_offsets_kernel(const uint2 *__restrict__ runs, u32 m1, const u64 *__restrict__ block_off, u64 *off)
{
    __shared__ u64 s_w[8];
    const u32 base = blockIdx.x * EX_BLOCK;
    constexpr int PER = EX_BLOCK / 256;   // 16 consecutive elements per thread
    const u32 t0 = base + threadIdx.x * PER;
    u32 v[PER];
    u64 sum = 0;
#pragma unroll
    for (int q = 0; q < PER; q++) { v[q] = (t0 + q < m1) ? runs[t0 + q].y : 0u; sum += v[q]; }
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    u64 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= (u32)o) inc += t; }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    u64 run = block_off[blockIdx.x] + inc - sum;
    for (u32 ww = 0; ww < w; ww++) run += s_w[ww];
#pragma unroll
    for (int q = 0; q < PER; q++) { if (t0 + q < m1) off[t0 + q] = run; run += v[q]; }
}

// pair o -> (left row id, right row id): left element = the last one whose offset is <= o
__global__ void __launch_bounds__(256)
many_expand_kernel(const u64 *__restrict__ L, const u64 *__restrict__ R, const uint2 *__restrict__ runs, const u64 *__restrict__ off, u32 m1,
                   u64 total, uint2 *__restrict__ dense)
{
    const u64 o = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    u32 lo = 0, hi = m1;                      // upper_bound(off, o) - 1
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (off[mid] <= o) lo = mid + 1; else hi = mid; }
    u32 i = lo - 1;
    // elements with an empty run share their offset with the next one: step back to the one that owns row o
    while (runs[i].y == 0u || off[i] + runs[i].y <= o) i--;
    const uint2 r = runs[i];
    dense[o] = make_uint2(pair_row(L[i]), pair_row(R[r.x + (u32)(o - off[i])]));
}
}  // namespace

// Expands the per-left-element runs written by smj_launch_join_match(mode = SMJ_JOIN_MANY, d_matches = runs) into the dense
// (left row id, right row id) list, order (key, left position, right position).  total = the count the match step returned.
size_t smj_join_many_scratch_bytes(u32 m1) { return ((size_t)(m1 + EX_BLOCK - 1) / EX_BLOCK + 2) * 8 + (size_t)m1 * 8 + 64; }
int smj_launch_join_many_expand(SmjCtx *c, const u64 *d_l, const u64 *d_r, const uint2 *d_runs, u32 m1, u64 total, char *d_scratch,
                                uint2 *d_dense)
{
    if (m1 == 0 || total == 0) return SMJ_OK;
    const u32 nblocks = (m1 + EX_BLOCK - 1) / EX_BLOCK;
    u64 *d_total = (u64 *)d_scratch;
    u64 *d_bsum = (u64 *)(d_scratch + 64);
    u64 *d_off = d_bsum + nblocks + 1;
    many_blocksum_kernel<<<nblocks, 256, 0, c->stream>>>(d_runs, m1, d_bsum);
    KERNEL_CHECK(c);
    many_blockscan_kernel<<<1, 1024, 0, c->stream>>>(d_bsum, nblocks, d_total);
    KERNEL_CHECK(c);
    many_offsets_kernel<<<nblocks, 256, 0, c->stream>>>(d_runs, m1, d_bsum, d_off);
    KERNEL_CHECK(c);
    many_expand_kernel<<<(u32)((total + 255) / 256), 256, 0, c->stream>>>(d_l, d_r, d_runs, d_off, m1, total, d_dense);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

size_t smj_join_num_tiles(u64 total) { return (size_t)((total + JN_TILE - 1) / JN_TILE); }
// tile counts one CTA of the many-CTA scan takes (SMJ_SCAN_CHUNK lets the tests force that path at small sizes)
u32 smj_join_scan_chunk(void)
{
    static const u32 chunk = [] {
        const char *e = getenv("SMJ_SCAN_CHUNK");
        const long v = e ? atol(e) : 0;
        return (u32)((v >= 32 && v <= SCAN1_STAGE) ? v : SCAN1_STAGE);
    }();
    return chunk;
}
// u64 words the scan wants behind the `tiles` tile offsets (one block sum per CTA, whatever the chunk)
size_t smj_join_scan_blocks(size_t tiles) { return tiles / 32 + 2; }
size_t smj_join_tile_size(void) { return JN_TILE; }

static int sm_count(SmjCtx *c)
{
    static int sms[16] = {};
    if (c->device >= 0 && c->device < 16 && sms[c->device]) return sms[c->device];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, c->device);
    if (c->device >= 0 && c->device < 16) sms[c->device] = n;
    return n;
}

// Scratch layout of one join (u32 words): [part tiles+1][runstart tiles+1][tile_count tiles (zeroed by caller)] then
// u64-aligned [tile_off tiles].  smj_join_scratch_* give the sizes; tiles = smj_join_num_tiles(m1_max + m2_max).
int smj_launch_join_match(SmjCtx *c, const u64 *d_l, const u64 *d_r, const u64 *d_counts, u32 m1_max, u32 m2_max,
                          int mode, u32 *d_part, u32 *d_tile_count, u64 *d_tile_off, uint2 *d_matches, uint2 *d_dense,
                          u64 *d_count)
{
    if (m1_max == 0 || m2_max == 0) return SMJ_OK;   // caller zeroed *d_count
    const u32 tiles = (u32)smj_join_num_tiles((u64)m1_max + m2_max);   // upper bound; the kernels use the device counts
    u32 *d_runstart = d_part + (tiles + 1);
    smj_launch(c, join_partition_kernel, (tiles + 1 + 7) / 8, 256, 0, d_l, d_r, d_counts, m1_max, m2_max, d_part, d_runstart);
    KERNEL_CHECK(c);
    const int sms = sm_count(c);
    const u32 grid = tiles < (u32)(sms * SMJ_JN_GRID) ? tiles : (u32)(sms * SMJ_JN_GRID);   // ~60 registers x 256 threads: 4 CTAs per SM
    if (mode == SMJ_JOIN_ZIP) {
        smj_launch(c, join_match_kernel<SMJ_JOIN_ZIP>, grid, JN_THREADS, 0, d_l, d_r, d_counts, m1_max, m2_max, d_part, d_runstart,
                   d_matches, d_tile_count, d_count, (uint2 *)nullptr, c->d_err);
        KERNEL_CHECK(c);
        const u32 chunk = smj_join_scan_chunk();
        if (tiles > chunk) {   // d_tile_off has smj_join_scan_blocks(tiles) spare words behind its `tiles` entries
            const u32 nb = (tiles + chunk - 1) / chunk;
            smj_launch(c, join_blocksum_kernel, nb, JS_THREADS, 0, d_tile_count, tiles, chunk, d_tile_off + tiles, d_counts, m1_max, m2_max);
            KERNEL_CHECK(c);
            smj_launch(c, join_apply_kernel, nb, JS_THREADS, 0, d_tile_count, tiles, chunk, d_tile_off, d_tile_off + tiles, d_count, d_counts,
                       m1_max, m2_max);
        } else
            smj_launch(c, join_scan_kernel, 1, JS_THREADS, 0, d_tile_count, tiles, d_tile_off, d_count, d_counts, m1_max, m2_max);
        if (d_dense) {
            KERNEL_CHECK(c);
            const u32 cgrid = (tiles + 7) / 8 < (u32)(sms * 8) ? (tiles + 7) / 8 : (u32)(sms * 8);
            smj_launch(c, join_compact_kernel, cgrid, 256, 0, d_matches, d_tile_count, d_tile_off, tiles, d_dense, d_counts, m1_max, m2_max,
                       (u64)(m1_max < m2_max ? m1_max : m2_max));
        }
    } else {
        // many-to-many: d_matches (if given) receives one (first right position, run length) entry per LEFT element
        smj_launch(c, join_match_kernel<SMJ_JOIN_MANY>, grid, JN_THREADS, 0, d_l, d_r, d_counts, m1_max, m2_max, d_part, d_runstart,
                   (uint2 *)nullptr, d_tile_count, d_count, d_matches, c->d_err);
    }
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// Joined rows from the dense match list (*d_nj of them, at most nj_max).
int smj_launch_join_materialize(SmjCtx *c, const uint2 *d_dense, const u64 *d_nj, int64_t nj_max, const int32_t *d_t1, int c1,
                                const int32_t *d_t2, int c2, int key2, int32_t *d_out, int32_t *const *d_out_indirect,
                                const u32 *d_use_store, const int32_t *d_store1, const int32_t *d_store2, const int32_t *const *d_tables_indirect)
{
    if (nj_max <= 0) return SMJ_OK;
    const int c_out = c1 + c2 - 1;
    if (c_out > MT_SMEM_CELLS) return smj_set_error(SMJ_EINVAL, "joined rows of %d columns exceed the %d-cell staging tile", c_out, MT_SMEM_CELLS);
    int rpb = MT_SMEM_CELLS / c_out;
    if (rpb > SMJ_MT_RPB) rpb = SMJ_MT_RPB;
    const int64_t nblocks = (nj_max + rpb - 1) / rpb;
    const int sms = sm_count(c);
    const u32 grid = (u32)(nblocks < (int64_t)sms * SMJ_MT_GRID ? nblocks : (int64_t)sms * SMJ_MT_GRID);
    const bool vec = (c1 % 4 == 0) && (c2 % 4 == 0) &&
                     ((((uintptr_t)d_t1) | ((uintptr_t)d_t2) | ((uintptr_t)d_store1) | ((uintptr_t)d_store2)) & 15) == 0;
    if (vec)
        smj_launch(c, join_materialize_kernel<true>, grid, MT_THREADS, 0, d_dense, d_nj, nj_max, d_t1, c1, d_t2, c2, key2, d_out, d_out_indirect, rpb,
                   d_use_store, d_store1, d_store2, d_tables_indirect);
    else
        smj_launch(c, join_materialize_kernel<false>, grid, MT_THREADS, 0, d_dense, d_nj, nj_max, d_t1, c1, d_t2, c2, key2, d_out, d_out_indirect, rpb,
                   d_use_store, d_store1, d_store2, d_tables_indirect);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

// Loads this file's pipeline kernels on the current device.  CUDA loads a kernel lazily at its first launch, and that load can
// wait for other GPUs' running kernels when peer access is enabled; a process that drives several GPUs (smj_dist.cu) must
// not meet such a load while another rank's kernel spins on this rank's flags, so it loads everything up front.
void smj_preload_join(void)
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, join_partition_kernel);
    cudaFuncGetAttributes(&a, join_match_kernel<SMJ_JOIN_ZIP>);
    cudaFuncGetAttributes(&a, join_match_kernel<SMJ_JOIN_MANY>);
    cudaFuncGetAttributes(&a, join_scan_kernel);
    cudaFuncGetAttributes(&a, join_blocksum_kernel);
    cudaFuncGetAttributes(&a, join_apply_kernel);
    cudaFuncGetAttributes(&a, join_compact_kernel);
    cudaFuncGetAttributes(&a, join_materialize_kernel<true>);
    cudaFuncGetAttributes(&a, join_materialize_kernel<false>);
    cudaGetLastError();
}
