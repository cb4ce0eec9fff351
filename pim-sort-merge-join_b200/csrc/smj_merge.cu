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

__global__ void __launch_bounds__(MG_THREADS)
merge_pairs_kernel(const u64 *__restrict__ A, u32 na_all, const u64 *__restrict__ B, u32 nb_all,
                   const u32 *__restrict__ part, u64 *__restrict__ out)
{
    __shared__ __align__(16) u64 s[MG_TILE];
    const u32 tid = threadIdx.x, tile = blockIdx.x;
    const u64 total = (u64)na_all + nb_all;
    const u64 d0 = (u64)tile * MG_TILE;
    const u64 d1 = (d0 + MG_TILE < total) ? d0 + MG_TILE : total;
    const u32 a0 = part[tile], a1 = part[tile + 1];
    const u32 b0 = (u32)(d0 - a0), b1 = (u32)(d1 - a1);
    const u32 na = a1 - a0, nb = b1 - b0, ntile = na + nb;
    u64 *sA = s, *sB = s + na;
    for (u32 i = tid; i < na; i += MG_THREADS) sA[i] = A[a0 + i];
    for (u32 i = tid; i < nb; i += MG_THREADS) sB[i] = B[b0 + i];
    __syncthreads();
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
    __syncthreads();
#pragma unroll
    for (int st = 0; st < MG_VT; st++)
        if (diag + st < ntile) s[diag + st] = r[st];
    __syncthreads();
    for (u32 i = tid; i < ntile; i += MG_THREADS) out[d0 + i] = s[i];
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
    merge_pairs_kernel<<<tiles, MG_THREADS, 0, c->stream>>>(d_a, na, d_b, nb, d_part, d_out);
    KERNEL_CHECK(c);
    return SMJ_OK;
}
