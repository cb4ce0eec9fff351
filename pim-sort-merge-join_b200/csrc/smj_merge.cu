// smj_merge.cu -- merge-path partitioned merge of two key-sorted (flipped key << 32 | row id) arrays,
// staged through shared memory.
//
// Replaces the DPU run merge (sort-merge-join/merge_dpu.c:92-103 binary-search split per tasklet, :130-169
// exchange merge, :190-217 shift compaction) and the host tournament that feeds it (app.c:413-547).
// Tie rule: every element of A precedes every equal-keyed element of B, which is what makes
// merge(ssort(A), ssort(B)) == ssort(A ++ B) when A's rows precede B's rows in the original order
// (SURVEY.md appendix A).  Equal tiles of the OUTPUT are found by co-ranking on the merge-path diagonals,
// so skewed or disjoint runs still give every CTA the same amount of work.
#include "smj_internal.h"
#include "smj_dev.cuh"

namespace {

constexpr int MG_THREADS = 256;
constexpr int MG_VT = 8;
constexpr int MG_TILE = MG_THREADS * MG_VT;

__global__ void merge_partition_kernel(const u64 *__restrict__ A, u32 na, const u64 *__restrict__ B, u32 nb,
                                       u32 num_tiles, u32 *part)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > num_tiles) return;
    const u64 total = (u64)na + nb;
    u64 d = (u64)t * MG_TILE;
    if (d > total) d = total;
    part[t] = merge_path(A, na, B, nb, (u32)d);
}

__device__ __forceinline__ void mg_cp_async8(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}

// Persistent CTAs, tiles dealt round-robin, two shared-memory buffers: while tile i is merged, the A and B segments of
// tile i + gridDim.x travel into the other buffer with cp.async (LDGSTS), so a CTA's global loads are always one tile
// ahead of its merge.  Inside a tile every thread finds its diagonal with a merge-path search in shared memory, merges
// MG_VT elements serially into registers, the tile is re-assembled in the buffer it came from and leaves as one
// contiguous, coalesced block.  (The first version was one CTA per tile with plain loads: load, barrier, merge, store,
// nothing overlapped.)
__global__ void __launch_bounds__(MG_THREADS)
merge_pairs_kernel(const u64 *__restrict__ A, u32 na_all, const u64 *__restrict__ B, u32 nb_all,
                   const u32 *__restrict__ part, u64 *__restrict__ out, u32 num_tiles, u32 *__restrict__ err)
{
    // (+ MG_TILE / 16: the merged tile is staged with one pad word per 16 elements -- thread t's eight results sit at 8 t .. 8 t + 7,
    // a 64-byte stride that made the stores 16-way bank conflicts: 70 % of all shared-memory wavefronts in the first capture)
    __shared__ __align__(16) u64 s[2][MG_TILE + MG_TILE / 16];
    const u32 tid = threadIdx.x;
    const u64 total = (u64)na_all + nb_all;
    struct Geo { u32 a0, na, b0, nb, bad_out; u64 d0; };   // bad_out: output entries of a tile that is skipped (0 for a good tile)
    auto geo = [&](u32 tile) {
        Geo g;
        g.d0 = (u64)tile * MG_TILE;
        const u64 d1 = (g.d0 + MG_TILE < total) ? g.d0 + MG_TILE : total;
        const u32 a1 = part[tile + 1];
        g.a0 = part[tile];
        g.b0 = (u32)(g.d0 - g.a0);
        g.na = a1 - g.a0;
        g.nb = (u32)(d1 - a1) - g.b0;
        g.bad_out = 0;
        // Runs that are not sorted give co-ranks that do not grow: the tile is skipped (never fetch past a buffer), its output is
        // filled with pair 0 (row 0: the gather that follows must not meet stale row ids) and the call reports SMJ_EINVAL (code 6).
        if (a1 < g.a0 || g.na > (u32)MG_TILE || g.nb > (u32)MG_TILE || g.na + g.nb != (u32)(d1 - g.d0)) { g.na = 0; g.nb = 0; g.bad_out = (u32)(d1 - g.d0); }
        return g;
    };
    auto fetch = [&](const Geo &g, u64 *buf) {
        for (u32 i = tid; i < g.na; i += MG_THREADS) mg_cp_async8(buf + i, A + g.a0 + i);
        for (u32 i = tid; i < g.nb; i += MG_THREADS) mg_cp_async8(buf + g.na + i, B + g.b0 + i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    u32 tile = blockIdx.x, cur = 0;
    Geo g = {}, gn = {};
    if (tile < num_tiles) {
        g = geo(tile);
        fetch(g, s[0]);
        if (tile + gridDim.x < num_tiles) gn = geo(tile + gridDim.x);
    }
    while (tile < num_tiles) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // this tile's data visible; the other buffer's stores are done
        const u32 next = tile + gridDim.x;
        if (next < num_tiles) fetch(gn, s[cur ^ 1]);
        const Geo gc = g;
        g = gn;
        if (next + gridDim.x < num_tiles) gn = geo(next + gridDim.x);   // partition entries two tiles ahead
        u64 *buf = s[cur];
        const u32 na = gc.na, nb = gc.nb, ntile = na + nb;
        const u64 *sA = buf, *sB = buf + na;
        const u32 diag = (tid * MG_VT < ntile) ? tid * MG_VT : ntile;
        u32 a = merge_path(sA, na, sB, nb, diag);
        u32 b = diag - a;
        u64 va = a < na ? sA[a] : 0ull, vb = b < nb ? sB[b] : 0ull;
        u64 r[MG_VT];
#pragma unroll
        for (int st = 0; st < MG_VT; st++) {
            const bool takeA = (b >= nb) || (a < na && pair_key(va) <= pair_key(vb));
            if (takeA) { r[st] = va; a++; va = a < na ? sA[a] : 0ull; }
            else       { r[st] = vb; b++; vb = b < nb ? sB[b] : 0ull; }
        }
        __syncthreads();                                   // every thread has read its inputs: the buffer becomes the output tile
#pragma unroll
        for (int st = 0; st < MG_VT; st++)
            if (diag + st < ntile) buf[(diag + st) + ((diag + st) >> 4)] = r[st];
        __syncthreads();
        u64 *dst = out + gc.d0;
        for (u32 i = tid; i < ntile; i += MG_THREADS) dst[i] = buf[i + (i >> 4)];
        if (gc.bad_out) {
            for (u32 i = tid; i < gc.bad_out; i += MG_THREADS) dst[i] = 0ull;
            if (tid == 0) atomicExch(err, SMJ_ERR_UNSORTED);
        }
        tile = next;
        cur ^= 1;
        // the next iteration's first barrier orders these reads of buf before the fetch that refills it one round later
    }
}

}  // namespace

size_t smj_merge_num_tiles(u64 total) { return (size_t)((total + MG_TILE - 1) / MG_TILE); }

int smj_launch_merge_pairs(SmjCtx *c, const u64 *d_a, u32 na, const u64 *d_b, u32 nb, u64 *d_out, u32 *d_part)
{
    const u64 total = (u64)na + nb;
    if (total == 0) return SMJ_OK;
    const u32 tiles = (u32)smj_merge_num_tiles(total);
    merge_partition_kernel<<<(tiles + 1 + 127) / 128, 128, 0, c->stream>>>(d_a, na, d_b, nb, tiles, d_part);
    KERNEL_CHECK(c);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const u32 grid = tiles < (u32)(sms * 5) ? tiles : (u32)(sms * 5);   // 34 KB of shared memory, 48 registers x 256 threads: five CTAs per SM
    merge_pairs_kernel<<<grid, MG_THREADS, 0, c->stream>>>(d_a, na, d_b, nb, d_part, d_out, tiles, c->d_err);
    KERNEL_CHECK(c);
    return SMJ_OK;
}
