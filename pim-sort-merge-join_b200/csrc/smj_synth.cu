// smj_synth.cu -- deterministic synthetic tables, generated directly in HBM.
//
// Replaces the unseeded generator sort-merge-join/data/generate_data.py:4-26 (unique key column drawn
// without replacement from [1, 3n], other columns uniform in [1, 3n)).  Every cell is a closed-form function
// of (seed, row, column), so any row range of a table of any size can be produced independently on any GPU
// and re-produced bit for bit by pim-sort-merge-join_b200/datagen.py (numpy) for the CPU checkers.
#include "smj_internal.h"
#include "smj_dev.cuh"
#include <math.h>
#include <stdlib.h>

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
                             u64 key_domain, u64 val_domain, int bits, u64 k0, u64 k1, u64 k2, const u64 *__restrict__ zipf_cdf)
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
            } else if (kind == 2) {   // Zipf: the first rank whose cumulative threshold exceeds a 64-bit uniform draw
                const u64 u = mix64(seed * 0x9E3779B97F4A7C15ull + 0x51ed270b7f4a7c15ull + row);
                u64 lo = 0, hi = key_domain - 1;       // zipf_cdf[key_domain - 1] = 2^64 - 1 >= u
                while (lo < hi) {
                    const u64 mid = (lo + hi) >> 1;
                    if (zipf_cdf[mid] < u) lo = mid + 1; else hi = mid;
                }
                v = 1 + lo;
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

// Zipf(s) over ranks 1..domain as 64-bit cumulative thresholds: cdf[k-1] = floor(2^64 * P(rank <= k)), last = 2^64 - 1.
// Host-only and exported, so that the numpy twin (datagen.py) searches the very same table: no floating point is
// evaluated twice, and the device table and the CPU checker's table agree bit for bit.
extern "C" int smj_synth_zipf_cdf(int64_t key_domain, double s, uint64_t *cdf_out)
{
    if (key_domain < 1 || !cdf_out || !(s > 0)) return smj_set_error(SMJ_EINVAL, "smj_synth_zipf_cdf: bad arguments");
    long double norm = 0;
    for (int64_t k = key_domain; k >= 1; k--) norm += powl((long double)k, -(long double)s);   // small terms first
    long double run = 0;
    for (int64_t k = 1; k <= key_domain; k++) {
        run += powl((long double)k, -(long double)s);
        long double f = run / norm;
        if (f > 1) f = 1;
        const long double scaled = f * 18446744073709551615.0L;
        uint64_t t = scaled >= 18446744073709551615.0L ? UINT64_MAX : (uint64_t)scaled;
        if (k > 1 && t < cdf_out[k - 2]) t = cdf_out[k - 2];
        cdf_out[k - 1] = t;
    }
    cdf_out[key_domain - 1] = UINT64_MAX;
    return SMJ_OK;
}

#define SMJ_ZIPF_S 1.1
#define SMJ_ZIPF_DEFAULT_DOMAIN (1 << 20)

int smj_launch_synth(SmjCtx *c, int32_t *d_out, int64_t row0, int64_t rows, int64_t total_rows, int cols, int key_col,
                     u64 seed, int kind, int64_t key_domain)
{
    if (rows <= 0) return SMJ_OK;
    if (cols < 1 || key_col < 0 || key_col >= cols || total_rows < 1)
        return smj_set_error(SMJ_EINVAL, "smj_synth_table: bad shape");
    u64 dom = key_domain > 0 ? (u64)key_domain : (kind == 2 ? (u64)SMJ_ZIPF_DEFAULT_DOMAIN : 3ull * (u64)total_rows);
    if (dom > 2147483647ull - 1) dom = 2147483647ull - 1;           // keys 1 + x must stay int32
    const u64 *d_cdf = nullptr;
    if (kind == 2) {   // Zipf(1.1) over [1, dom]: the threshold table is built once per domain and kept in a workspace slot
        static u64 cdf_dom[16] = {};
        if (dom > (1ull << 26)) return smj_set_error(SMJ_EINVAL, "Zipf key_domain %llu exceeds 2^26", (unsigned long long)dom);
        u64 *d = (u64 *)smj_ws(c, WS_ZIPF, (size_t)dom * 8);
        if (!d) return SMJ_ENOMEM;
        if (cdf_dom[c->device & 15] != dom) {
            uint64_t *h = (uint64_t *)malloc((size_t)dom * 8);
            if (!h) return smj_set_error(SMJ_ENOMEM, "Zipf table");
            smj_synth_zipf_cdf((int64_t)dom, SMJ_ZIPF_S, h);
            cudaError_t e = cudaMemcpyAsync(d, h, (size_t)dom * 8, cudaMemcpyHostToDevice, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            free(h);
            if (e != cudaSuccess) return smj_cuda_fail(e, "Zipf table upload", __FILE__, __LINE__);
            cdf_dom[c->device & 15] = dom;
        }
        d_cdf = d;
    }
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
    synth_kernel<<<(u32)grid, 256, 0, c->stream>>>(d_out, row0, rows, cols, key_col, seed, kind, dom, vdom, bits, k0, k1, k2, d_cdf);
    KERNEL_CHECK(c);
    return SMJ_OK;
}
