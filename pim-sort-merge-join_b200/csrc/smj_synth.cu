// smj_synth.cu -- deterministic synthetic tables, generated directly in HBM.
//
// Replaces the unseeded generator sort-merge-join/data/generate_data.py:4-26 (unique key column drawn
// without replacement from [1, 3n], other columns uniform in [1, 3n)).  Every cell is a closed-form function
// of (seed, row, column), so any row range of a table of any size can be produced independently on any GPU
// and re-produced bit for bit by pim-sort-merge-join_b200/datagen.py (numpy) for the CPU checkers.
#include "smj_internal.h"
#include "smj_dev.cuh"

namespace {

__host__ __device__ __forceinline__ u64 mix64(u64 x)   // splitmix64 finaliser
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// A bijection on [0, 2^bits): add, odd multiply and xor-shift-right are each invertible mod 2^bits.
__host__ __device__ __forceinline__ u64 perm_bits(u64 x, int bits, u64 k0, u64 k1, u64 k2)
{
    const u64 mask = (bits >= 64) ? ~0ull : ((1ull << bits) - 1);
    const int s = bits > 1 ? bits / 2 : 1;
    x = (x + k0) & mask;
    x = (x * (k1 | 1ull)) & mask;
    x ^= x >> s;
    x = (x * (k2 | 1ull)) & mask;
    x ^= x >> s;
    x = (x + k1) & mask;
    x = (x * (k0 | 1ull)) & mask;
    x ^= x >> s;
    return x;
}

__global__ void synth_kernel(int32_t *out, int64_t row0, int64_t rows, int cols, int key_col, u64 seed, int kind,
                             u64 key_domain, u64 val_domain, int bits, u64 k0, u64 k1, u64 k2)
{
    const int64_t ncell = rows * cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell; cell += stride) {
        const int64_t r = cell / cols;
        const int col = (int)(cell - r * cols);
        const u64 row = (u64)(row0 + r);
        u64 v;
        if (col == key_col) {
            if (kind == 0) {   // unique: cycle-walk the 2^bits bijection into [0, key_domain)
                u64 x = perm_bits(row, bits, k0, k1, k2);
                while (x >= key_domain) x = perm_bits(x, bits, k0, k1, k2);
                v = 1 + x;
            } else {
                v = 1 + mix64(seed * 0x9E3779B97F4A7C15ull + 0x51ed270b7f4a7c15ull + row) % key_domain;
            }
        } else {
            v = 1 + mix64((seed * 0x9E3779B97F4A7C15ull) ^ (row * (u64)cols + (u64)col + 1)) % val_domain;
        }
        out[cell] = (int32_t)v;
    }
}

}  // namespace

int smj_launch_synth(SmjCtx *c, int32_t *d_out, int64_t row0, int64_t rows, int64_t total_rows, int cols, int key_col,
                     u64 seed, int kind, int64_t key_domain)
{
    if (rows <= 0) return SMJ_OK;
    if (cols < 1 || key_col < 0 || key_col >= cols || total_rows < 1)
        return smj_set_error(SMJ_EINVAL, "smj_synth_table: bad shape");
    u64 dom = key_domain > 0 ? (u64)key_domain : 3ull * (u64)total_rows;
    if (dom > 2147483647ull - 1) dom = 2147483647ull - 1;           // keys 1 + x must stay int32
    if (kind == 0 && dom < (u64)total_rows) return smj_set_error(SMJ_EINVAL, "unique keys need key_domain >= total_rows");
    u64 vdom = 3ull * (u64)total_rows - 1;                           // values in [1, 3n)
    if (vdom > 2147483647ull - 1) vdom = 2147483647ull - 1;
    if (vdom < 1) vdom = 1;
    int bits = 1;
    while ((1ull << bits) < dom) bits++;
    const u64 k0 = mix64(seed + 1), k1 = mix64(seed + 2), k2 = mix64(seed + 3);
    const int64_t ncell = rows * cols;
    int64_t grid = (ncell + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    synth_kernel<<<(u32)grid, 256, 0, c->stream>>>(d_out, row0, rows, cols, key_col, seed, kind, dom, vdom, bits, k0, k1, k2);
    KERNEL_CHECK(c);
    return SMJ_OK;
}
