// smj_gather.cu -- payload gather: out[i][:] = in[rowid(pairs[i])][:].
//
// The sort moves only 8-byte (key,rowid) pairs; whole rows move exactly once, here.  This is the row
// movement the reference does on every insertion-sort shift (cpu_app.c:186-198, sort_dpu.c:163-184).
// Rows whose byte size is a multiple of 16 (4 / 8 int32 columns) and whose tables are 16-byte aligned go
// through 128-bit loads/stores; everything else (e.g. the 5-column config) through 32-bit cells.
// Both paths write the output fully coalesced: consecutive threads own consecutive output cells.
#include "smj_internal.h"
#include "smj_dev.cuh"

namespace {

constexpr int GA_THREADS = 256;
constexpr int GA_ROWS = 512;   // output rows per CTA

template <typename CELL, bool TWO>
__global__ void __launch_bounds__(GA_THREADS)
gather_rows_kernel(const u64 *__restrict__ pairs, int64_t m, const CELL *__restrict__ a, const CELL *__restrict__ b,
                   u32 split, int cols, CELL *__restrict__ out)
{
    const int64_t row0 = (int64_t)blockIdx.x * GA_ROWS;
    const u32 nrows = (u32)((m - row0 < GA_ROWS) ? (m - row0) : GA_ROWS);
    __shared__ u32 s_row[GA_ROWS];
    for (u32 i = threadIdx.x; i < nrows; i += GA_THREADS) s_row[i] = pair_row(pairs[row0 + i]);
    __syncthreads();
    const u32 ncell = nrows * (u32)cols;
    CELL *o = out + row0 * cols;
    for (u32 cell = threadIdx.x; cell < ncell; cell += GA_THREADS) {
        const u32 r = cell / (u32)cols, col = cell - r * (u32)cols;
        const u32 rid = s_row[r];
        const CELL *src;
        if (TWO && rid >= split) src = b + (size_t)(rid - split) * cols;
        else src = a + (size_t)rid * cols;
        o[cell] = src[col];
    }
}

template <bool TWO>
int launch(SmjCtx *c, const u64 *d_pairs, int64_t m, const int32_t *d_a, const int32_t *d_b, u32 split, int cols,
           int32_t *d_out)
{
    if (m <= 0) return SMJ_OK;
    const u32 grid = (u32)((m + GA_ROWS - 1) / GA_ROWS);
    const bool vec = (cols % 4 == 0) && (((uintptr_t)d_a | (uintptr_t)d_out | (uintptr_t)(TWO ? d_b : d_a)) & 15) == 0;
    if (vec)
        gather_rows_kernel<int4, TWO><<<grid, GA_THREADS, 0, c->stream>>>(
            d_pairs, m, (const int4 *)d_a, (const int4 *)d_b, split, cols / 4, (int4 *)d_out);
    else
        gather_rows_kernel<int32_t, TWO><<<grid, GA_THREADS, 0, c->stream>>>(d_pairs, m, d_a, d_b, split, cols, d_out);
    KERNEL_CHECK(c);
    return SMJ_OK;
}

}  // namespace

int smj_launch_gather_rows(SmjCtx *c, const u64 *d_pairs, int64_t m, const int32_t *d_in, int cols, int32_t *d_out)
{
    return launch<false>(c, d_pairs, m, d_in, d_in, 0, cols, d_out);
}

int smj_launch_gather_rows2(SmjCtx *c, const u64 *d_pairs, int64_t m, const int32_t *d_a, const int32_t *d_b, u32 split,
                            int cols, int32_t *d_out)
{
    return launch<true>(c, d_pairs, m, d_a, d_b, split, cols, d_out);
}
