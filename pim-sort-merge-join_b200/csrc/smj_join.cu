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
// kernel).  Count -> scan -> write: each tile counts its matches, a decoupled look-back gives the tile's
// output offset, and the (left row id, right row id) matches are written coalesced; a second kernel
// materialises the joined rows (all left columns, then right columns except key2: cpu_app.c:240-251).
#include "smj_internal.h"
#include "smj_dev.cuh"

namespace {

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
                if (pair_key(L[a - 1]) == k) rs = lower_bound_key(L, 0, a, k);
            }
            runstart[t] = rs;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(JN_THREADS)
join_match_kernel(const u64 *__restrict__ L, const u64 *__restrict__ R, const u64 *__restrict__ counts, u32 m1_max,
                  u32 m2_max, const u32 *__restrict__ part, const u32 *__restrict__ runstart, u64 *status,
                  u32 *tile_counter, uint2 *__restrict__ matches, u64 *count, u32 *err)
{
    u32 m1, m2;
    load_counts(counts, m1_max, m2_max, m1, m2);
    const u32 num_tiles = (m1 == 0 || m2 == 0) ? 0u : (u32)(((u64)m1 + m2 + JN_TILE - 1) / JN_TILE);
    __shared__ __align__(16) u64 s[JN_TILE + 2];
    __shared__ u32 s_wsum[JN_WARPS];
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u64 total = (u64)m1 + m2;
    u64 many_total = 0;

    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= num_tiles) break;
        const u64 d0 = (u64)tile * JN_TILE;
        const u64 d1 = (d0 + JN_TILE < total) ? d0 + JN_TILE : total;
        const u32 a0 = part[tile], a1 = part[tile + 1];
        const u32 b0 = (u32)(d0 - a0), b1 = (u32)(d1 - a1);
        const u32 na = a1 - a0, nb = b1 - b0;
        const u32 nbh = nb + (b1 < m2 ? 1u : 0u);   // right segment plus one halo element
        u64 *sA = s, *sB = s + na;
        for (u32 i = tid; i < na; i += JN_THREADS) sA[i] = L[a0 + i];
        for (u32 i = tid; i < nbh; i += JN_THREADS) sB[i] = R[b0 + i];
        const bool has_prev = a0 > 0;
        const u32 prevk = has_prev ? pair_key(L[a0 - 1]) : 0u;
        const u32 tile_rs = runstart[tile];
        __syncthreads();

        const u32 ntile = na + nb;
        const u32 diag = (tid * JN_VT < ntile) ? tid * JN_VT : ntile;
        u32 a = merge_path(sA, na, sB, nb, diag);
        u32 b = diag - a;
        const u32 steps = (ntile - diag < (u32)JN_VT) ? ntile - diag : (u32)JN_VT;
        u32 ka = a < na ? pair_key(sA[a]) : 0u, kb = b < nb ? pair_key(sB[b]) : 0u;
        uint2 mt[JN_VT];
        u32 mmask = 0;
        bool have = false;
        u32 runk = 0, run_val = 0;   // zip: first left position of the current key; many: its right run length
#pragma unroll
        for (int st = 0; st < JN_VT; st++) {
            if (st < (int)steps) {
                const bool takeA = (b >= nb) || (a < na && ka <= kb);
                if (takeA) {
                    const u32 k = ka;
                    if (MODE == SMJ_JOIN_ZIP) {
                        if (!(have && runk == k)) {
                            u32 rs;
                            if (a == 0) rs = (has_prev && prevk == k) ? tile_rs : a0;
                            else if (pair_key(sA[a - 1]) != k) rs = a0 + a;
                            else {
                                const u32 lo = lower_bound_key(sA, 0, a, k);
                                rs = (lo == 0 && has_prev && prevk == k) ? tile_rs : a0 + lo;
                            }
                            run_val = rs; runk = k; have = true;
                        }
                        const u64 j = (u64)(b0 + b) + ((a0 + a) - run_val);
                        if (j < m2) {
                            const u64 rp = (j - b0 < nbh) ? sB[j - b0] : R[j];
                            if (pair_key(rp) == k) {
                                mt[st] = make_uint2(pair_row(sA[a]), pair_row(rp));
                                mmask |= 1u << st;
                            }
                        }
                    } else {
                        if (!(have && runk == k)) {
                            const u32 e = upper_bound_key(sB, b, nbh, k);
                            u32 ub = b0 + e;
                            if (e == nbh && ub < m2 && pair_key(R[ub]) == k) ub = upper_bound_key(R, ub, m2, k);
                            run_val = ub - (b0 + b); runk = k; have = true;
                        }
                        many_total += run_val;
                    }
                    a++;
                    ka = a < na ? pair_key(sA[a]) : 0u;
                } else {
                    b++;
                    kb = b < nb ? pair_key(sB[b]) : 0u;
                }
            }
        }
        if (MODE == SMJ_JOIN_ZIP) {
            const u32 cnt = __popc(mmask);
            const u32 inc = warp_incl_scan(cnt);
            if (lane == 31) s_wsum[w] = inc;
            __syncthreads();   // also: every thread is done reading sA/sB, s[] may be reused below
            u32 wp = 0, tile_total = 0;
#pragma unroll
            for (int ww = 0; ww < JN_WARPS; ww++) {
                const u32 v = s_wsum[ww];
                if (ww < (int)w) wp += v;
                tile_total += v;
            }
            if (w == 0) {
                const u64 e = lookback_warp(status, tile, (u64)tile_total, err, SMJ_ERR_SPIN_JOIN);
                if (lane == 0) {
                    s_base = e;
                    if (tile == num_tiles - 1) *count = e + tile_total;
                }
            }
            uint2 *s_out = reinterpret_cast<uint2 *>(s);
            u32 o = wp + inc - cnt;
#pragma unroll
            for (int st = 0; st < JN_VT; st++)
                if ((mmask >> st) & 1u) s_out[o++] = mt[st];
            __syncthreads();
            const u64 base = s_base;
            for (u32 i = tid; i < tile_total; i += JN_THREADS) matches[base + i] = s_out[i];
        }
    }
    if (MODE != SMJ_JOIN_ZIP) {
        many_total = warp_sum(many_total);
        if (lane == 0 && many_total) atomicAdd(count, many_total);
    }
}

constexpr int MT_THREADS = 256;
constexpr int MT_SMEM_CELLS = 8192;   // 32 KB staging: rows per block = MT_SMEM_CELLS / c_out (<= 1024)

// out[o] = t1[matches[o].x][0..c1) ++ t2[matches[o].y][c != key2]   (cpu_app.c:240-251, join.c:214-229)
// One thread gathers one source row (128-bit loads when the row is a whole number of 16-byte words and aligned,
// i.e. the 4- and 8-column configs), rows are assembled in shared memory and leave as one contiguous,
// fully coalesced block.  The row count comes from device memory so no host round trip precedes the launch.
template <bool VEC>
__global__ void __launch_bounds__(MT_THREADS)
join_materialize_kernel(const uint2 *__restrict__ matches, const u64 *__restrict__ nj_dev, int64_t nj_max,
                        const int32_t *__restrict__ t1, int c1, const int32_t *__restrict__ t2, int c2, int key2,
                        int32_t *__restrict__ out, int rows_per_block)
{
    __shared__ __align__(16) int32_t s_out[MT_SMEM_CELLS];
    int64_t nj = nj_max;
    if (nj_dev) { const u64 v = *nj_dev; nj = v < (u64)nj_max ? (int64_t)v : nj_max; }
    const int c_out = c1 + c2 - 1;
    const int64_t nblocks = (nj + rows_per_block - 1) / rows_per_block;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int64_t row0 = blk * rows_per_block;
        const int nrows = (int)((nj - row0 < rows_per_block) ? (nj - row0) : rows_per_block);
        __syncthreads();   // previous block's copy-out is done with s_out
        // 2 * nrows gather tasks: task i < nrows is the left row of match i, the others the right rows.  Four tasks
        // per thread are in flight at once (match entries first, then the dependent row loads): the random 16..32-byte
        // row reads are latency-bound, so memory-level parallelism per thread is what sets the rate.
        for (int task0 = threadIdx.x; task0 < 2 * nrows; task0 += 4 * MT_THREADS) {
            uint2 m[4];
            const int32_t *src[4];
            int32_t *dst[4];
            int cc[4];
            bool right[4], live[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int task = task0 + u * MT_THREADS;
                live[u] = task < 2 * nrows;
                right[u] = task >= nrows;
                const int r = right[u] ? task - nrows : task;
                m[u] = live[u] ? matches[row0 + r] : make_uint2(0u, 0u);
                dst[u] = s_out + r * c_out + (right[u] ? c1 : 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                cc[u] = right[u] ? c2 : c1;
                src[u] = right[u] ? t2 + (size_t)m[u].y * c2 : t1 + (size_t)m[u].x * c1;
            }
            if (VEC) {
                int4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (live[u]) v[u] = __ldg(reinterpret_cast<const int4 *>(src[u]));
#pragma unroll
                for (int u = 0; u < 4; u++) {
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
                for (int u = 0; u < 4; u++) {
                    if (!live[u]) continue;
                    const int skip = right[u] ? key2 : -1;
                    for (int q = 0; q < cc[u]; q++)
                        if (q != skip) dst[u][q - ((skip >= 0 && q > skip) ? 1 : 0)] = __ldg(src[u] + q);
                }
            }
        }
        __syncthreads();
        const int ncell = nrows * c_out;
        int32_t *o = out + row0 * c_out;
        for (int cell = threadIdx.x; cell < ncell; cell += MT_THREADS) o[cell] = s_out[cell];
    }
}

}  // namespace

size_t smj_join_num_tiles(u64 total) { return (size_t)((total + JN_TILE - 1) / JN_TILE); }

static int sm_count(SmjCtx *c)
{
    static int sms[16] = {};
    if (c->device >= 0 && c->device < 16 && sms[c->device]) return sms[c->device];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, c->device);
    if (c->device >= 0 && c->device < 16) sms[c->device] = n;
    return n;
}

int smj_launch_join_match(SmjCtx *c, const u64 *d_l, const u64 *d_r, const u64 *d_counts, u32 m1_max, u32 m2_max,
                          int mode, u32 *d_part, u64 *d_status, u32 *d_tile_counter, uint2 *d_matches, u64 *d_count)
{
    if (m1_max == 0 || m2_max == 0) return SMJ_OK;   // caller zeroed *d_count
    const u32 tiles = (u32)smj_join_num_tiles((u64)m1_max + m2_max);   // upper bound; the kernels use the device counts
    u32 *d_runstart = d_part + (tiles + 1);
    join_partition_kernel<<<(tiles + 1 + 7) / 8, 256, 0, c->stream>>>(d_l, d_r, d_counts, m1_max, m2_max, d_part, d_runstart);
    KERNEL_CHECK(c);
    const int sms = sm_count(c);
    const u32 grid = tiles < (u32)(sms * 8) ? tiles : (u32)(sms * 8);
    if (mode == SMJ_JOIN_ZIP)
        join_match_kernel<SMJ_JOIN_ZIP><<<grid, JN_THREADS, 0, c->stream>>>(d_l, d_r, d_counts, m1_max, m2_max, d_part, d_runstart,
                                                                           d_status, d_tile_counter, d_matches, d_count, c->d_err);
    else
        join_match_kernel<SMJ_JOIN_MANY><<<grid, JN_THREADS, 0, c->stream>>>(d_l, d_r, d_counts, m1_max, m2_max, d_part, d_runstart,
                                                                            d_status, d_tile_counter, d_matches, d_count, c->d_err);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

int smj_launch_join_materialize(SmjCtx *c, const uint2 *d_matches, const u64 *d_nj, int64_t nj_max, const int32_t *d_t1,
                                int c1, const int32_t *d_t2, int c2, int key2, int32_t *d_out)
{
    if (nj_max <= 0) return SMJ_OK;
    const int c_out = c1 + c2 - 1;
    if (c_out > MT_SMEM_CELLS) return smj_set_error(SMJ_EINVAL, "joined rows of %d columns exceed the %d-cell staging tile", c_out, MT_SMEM_CELLS);
    int rpb = MT_SMEM_CELLS / c_out;
    if (rpb > 1024) rpb = 1024;
    const int64_t nblocks = (nj_max + rpb - 1) / rpb;
    const int sms = sm_count(c);
    const u32 grid = (u32)(nblocks < (int64_t)sms * 8 ? nblocks : (int64_t)sms * 8);
    const bool vec = (c1 % 4 == 0) && (c2 % 4 == 0) && ((((uintptr_t)d_t1) | ((uintptr_t)d_t2)) & 15) == 0;
    if (vec)
        join_materialize_kernel<true><<<grid, MT_THREADS, 0, c->stream>>>(d_matches, d_nj, nj_max, d_t1, c1, d_t2, c2, key2, d_out, rpb);
    else
        join_materialize_kernel<false><<<grid, MT_THREADS, 0, c->stream>>>(d_matches, d_nj, nj_max, d_t1, c1, d_t2, c2, key2, d_out, rpb);
    KERNEL_CHECK(c);
    return SMJ_OK;
}
