// smj_csv.cu -- CSV text <-> int32 tables on the GPU ("data1.csv in, result.csv out" at
// device speed).
//
// Parse replaces set_csv_size + load_csv (sort-merge-join/cpu_app.c:15-79 == app.c:28-92) for REGULAR files:
//   - the file is a sequence of '\n'-terminated lines, every line shorter than 1023 characters (the reference reads
//     with fgets(line, 1024, f), so longer lines are split there);
//   - cols = number of ','-separated non-empty tokens of the first line, rows = lines - 1;
//   - every other line has exactly cols tokens; cell = atoi(token) as glibc computes it: leading isspace() skipped,
//     optional sign, digits, stop at the first other character (so "12\r\n" is 12), saturating like strtol and then
//     truncated to int.
// Anything else -- a long line, a row with a different token count (the reference then writes cells of the NEXT row or
// leaves cells uninitialised, which only a sequential pass can reproduce), an embedded NUL -- makes smj_csv_parse
// return SMJ_EIRREGULAR and the caller falls back to the sequential host parser (host/csv.c).
//
// Format replaces save_to_csv (cpu_app.c:268-301 == app.c:720-755): header col1..colN, cells printed like %ld,
// ',' between cells, '\n' after every row.
#include "smj_internal.h"
#include "smj_dev.cuh"

#include <stdio.h>
#include <string.h>

namespace {

constexpr int CSV_BLOCK = 1024;      // bytes per newline-counting block
constexpr int CSV_MAX_LINE = 1023;   // fgets(line, 1024, f)

__device__ __forceinline__ bool csv_isspace(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// newline count per block; flag 1 if the text holds a NUL byte
__global__ void __launch_bounds__(256) csv_count_kernel(const char *__restrict__ text, u64 bytes, u32 *block_cnt, u32 *flags)
{
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per CSV_BLOCK bytes
    const u64 lo = b * CSV_BLOCK;
    if (lo >= bytes) return;
    const u64 hi = lo + CSV_BLOCK < bytes ? lo + CSV_BLOCK : bytes;
    u32 n = 0, nul = 0;
    for (u64 i = lo; i < hi; i++) { const char ch = text[i]; n += ch == '\n'; nul |= ch == 0; }
    block_cnt[b] = n;
    if (nul) atomicOr(flags, 1u);
}

// line_start[1 + k] = position after the k-th newline; line_start[0] = 0
__global__ void __launch_bounds__(256) csv_starts_kernel(const char *__restrict__ text, u64 bytes, const u64 *__restrict__ block_off, u64 *line_start)
{
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 lo = b * CSV_BLOCK;
    if (b == 0 && threadIdx.x == 0) line_start[0] = 0;
    if (lo >= bytes) return;
    const u64 hi = lo + CSV_BLOCK < bytes ? lo + CSV_BLOCK : bytes;
    u64 k = block_off[b];
    for (u64 i = lo; i < hi; i++)
        if (text[i] == '\n') line_start[1 + k++] = i + 1;
}

// tokens of [p, e): maximal runs of non-',' characters
__device__ int csv_count_tokens(const char *p, const char *e)
{
    int n = 0;
    while (p < e) {
        while (p < e && *p == ',') p++;
        if (p >= e) break;
        n++;
        while (p < e && *p != ',') p++;
    }
    return n;
}

__device__ int32_t csv_atoi(const char *p, const char *e)
{
    while (p < e && csv_isspace((unsigned char)*p)) p++;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
    const u64 lim = neg ? 9223372036854775808ull : 9223372036854775807ull;   // |LONG_MIN| / LONG_MAX
    u64 v = 0;
    bool sat = false;
    for (; p < e && *p >= '0' && *p <= '9'; p++) {
        const u64 d = (u64)(*p - '0');
        if (sat) continue;
        if (v > (lim - d) / 10) { sat = true; v = lim; }
        else v = v * 10 + d;
    }
    const long long r = neg ? (long long)(0ull - v) : (long long)v;
    return (int32_t)(int)r;
}

__global__ void csv_header_kernel(const char *__restrict__ text, const u64 *__restrict__ line_start, u64 bytes, u64 num_lines, u32 *cols_out,
                                  u32 *flags)
{
    if (num_lines == 0) { *cols_out = 0; return; }
    const u64 lo = line_start[0], hi = num_lines > 1 ? line_start[1] : bytes;
    if (hi - lo > CSV_MAX_LINE) atomicOr(flags, 2u);
    *cols_out = (u32)csv_count_tokens(text + lo, text + hi);
}

// one thread per data line: row r = line r + 1
__global__ void __launch_bounds__(256)
csv_parse_kernel(const char *__restrict__ text, const u64 *__restrict__ line_start, u64 bytes, u64 num_lines, int cols, int32_t *__restrict__ out,
                 u32 *flags)
{
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r + 1 >= num_lines) return;
    const u64 lo = line_start[r + 1], hi = (r + 2 < num_lines) ? line_start[r + 2] : bytes;
    if (hi - lo > CSV_MAX_LINE) { atomicOr(flags, 2u); return; }
    const char *p = text + lo, *e = text + hi;
    int32_t *row = out + r * cols;
    int col = 0;
    while (p < e) {
        while (p < e && *p == ',') p++;
        if (p >= e) break;
        const char *t = p;
        while (p < e && *p != ',') p++;
        if (col < cols) row[col] = csv_atoi(t, p);
        col++;
    }
    if (col != cols) atomicOr(flags, 4u);
}

__device__ __forceinline__ u32 csv_digits(int32_t v)
{
    u32 u = v < 0 ? 0u - (u32)v : (u32)v;
    u32 n = 1;
    while (u >= 10) { u /= 10; n++; }
    return n + (v < 0 ? 1u : 0u);
}

__global__ void __launch_bounds__(256) csv_rowlen_kernel(const int32_t *__restrict__ t, int64_t rows, int cols, u32 *len)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    u32 n = (u32)cols;   // cols - 1 commas + the newline
    const int32_t *row = t + r * cols;
    for (int c = 0; c < cols; c++) n += csv_digits(row[c]);
    len[r] = n;
}

__global__ void __launch_bounds__(256)
csv_write_kernel(const int32_t *__restrict__ t, int64_t rows, int cols, const u64 *__restrict__ off, char *__restrict__ text)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    char *p = text + off[r];
    const int32_t *row = t + r * cols;
    for (int c = 0; c < cols; c++) {
        const int32_t v = row[c];
        u32 u = v < 0 ? 0u - (u32)v : (u32)v;
        const u32 nd = csv_digits(v);
        char *q = p + nd;
        do { *--q = (char)('0' + u % 10); u /= 10; } while (u);
        if (v < 0) *--q = '-';
        p += nd;
        *p++ = (c + 1 < cols) ? ',' : '\n';
    }
}

// exclusive scan of n u32 values into u64 offsets, total in *total (one CTA; each thread owns a contiguous chunk)
constexpr int CS_THREADS = 1024;
__global__ void __launch_bounds__(CS_THREADS) csv_scan_kernel(const u32 *__restrict__ v, u64 n, u64 *off, u64 *total)
{
    __shared__ u64 s_w[CS_THREADS / 32];
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const u64 chunk = (n + CS_THREADS - 1) / CS_THREADS;
    const u64 lo = (u64)tid * chunk < n ? (u64)tid * chunk : n;
    const u64 hi = lo + chunk < n ? lo + chunk : n;
    u64 sum = 0;
    for (u64 i = lo; i < hi; i++) sum += v[i];
    u64 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(FULL_MASK, inc, o);
        if (lane >= (u32)o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    u64 run = inc - sum;
    for (u32 ww = 0; ww < w; ww++) run += s_w[ww];
    if (tid == CS_THREADS - 1) *total = run + sum;
    for (u64 i = lo; i < hi; i++) { off[i] = run; run += v[i]; }
}

}  // namespace

extern SmjCtx *g_ctx[8];
int smj_ensure_init(void);
int smj_alloc_out(SmjCtx *c, smj_table_t *out, int64_t rows, int cols);

// workspace layout for parse: [flags u32 x4][cols u32][pad][total u64][block_cnt u32 x nb][block_off u64 x nb][line_start u64 x (L+1)]
extern "C" int smj_csv_parse(const char *text, size_t bytes, smj_table_t *out)
{
    SMJ_TRY(smj_ensure_init());
    if (!out || (!text && bytes)) return smj_set_error(SMJ_EINVAL, "smj_csv_parse: null argument");
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    out->data = nullptr; out->rows = 0; out->cols = 0; out->on_device = 1;
    if (bytes == 0) { out->rows = -1; return SMJ_OK; }   // empty file: rows = lines - 1 = -1, like the reference
    char *d_text = (char *)smj_ws(c, WS_TMP_ROWS, bytes);
    if (!d_text) return SMJ_ENOMEM;
    CUDA_TRY(cudaMemcpyAsync(d_text, text, bytes, cudaMemcpyHostToDevice, c->stream));
    const u64 nb = (bytes + CSV_BLOCK - 1) / CSV_BLOCK;
    char *ws = (char *)smj_ws(c, WS_TMP_ROWS2, 64 + nb * 4 + 8 + nb * 8);
    if (!ws) return SMJ_ENOMEM;
    u32 *d_flags = (u32 *)ws;
    u32 *d_cols = d_flags + 4;
    u64 *d_total = (u64 *)(ws + 32);
    u32 *d_bcnt = (u32 *)(ws + 64);
    u64 *d_boff = (u64 *)(ws + 64 + align_up(nb * 4, 8));
    CUDA_TRY(cudaMemsetAsync(ws, 0, 64, c->stream));
    csv_count_kernel<<<(u32)((nb + 255) / 256), 256, 0, c->stream>>>(d_text, bytes, d_bcnt, d_flags);
    KERNEL_CHECK(c);
    csv_scan_kernel<<<1, CS_THREADS, 0, c->stream>>>(d_bcnt, nb, d_boff, d_total);
    KERNEL_CHECK(c);
    u64 *h = (u64 *)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(h, d_total, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const u64 newlines = h[0];
    const u64 num_lines = newlines + (text[bytes - 1] != '\n' ? 1 : 0);   // a last line without '\n' is one more fgets() chunk
    u64 *d_ls = (u64 *)smj_ws(c, WS_PART, (newlines + 2) * 8);
    if (!d_ls) return SMJ_ENOMEM;
    csv_starts_kernel<<<(u32)((nb + 255) / 256), 256, 0, c->stream>>>(d_text, bytes, d_boff, d_ls);
    KERNEL_CHECK(c);
    csv_header_kernel<<<1, 1, 0, c->stream>>>(d_text, d_ls, bytes, num_lines, d_cols, d_flags);
    KERNEL_CHECK(c);
    u32 *h32 = (u32 *)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(h32, d_flags, 32, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h32[0]) return smj_set_error(SMJ_EIRREGULAR, "CSV needs the sequential parser (flags %u: 1 NUL byte, 2 line of 1023+ characters)", h32[0]);
    const int cols = (int)h32[4];
    const int64_t rows = (int64_t)num_lines - 1;
    if (cols < 1) return smj_set_error(SMJ_EIRREGULAR, "CSV header has no columns");
    SMJ_TRY(smj_alloc_out(c, out, rows, cols));
    if (rows > 0) {
        csv_parse_kernel<<<(u32)((rows + 255) / 256), 256, 0, c->stream>>>(d_text, d_ls, bytes, num_lines, cols, out->data, d_flags);
        KERNEL_CHECK(c);
    }
    CUDA_TRY(cudaMemcpyAsync(h32, d_flags, 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h32[0]) {
        smj_table_free(out);
        out->rows = 0; out->cols = 0;
        return smj_set_error(SMJ_EIRREGULAR, "CSV needs the sequential parser (flags %u: 2 line of 1023+ characters, 4 row with a token count "
                                             "different from the header's)", h32[0]);
    }
    return SMJ_OK;
}

extern "C" int smj_csv_format(const smj_table_t *t, char **text, size_t *bytes)
{
    SMJ_TRY(smj_ensure_init());
    if (!t || !text || !bytes || t->cols < 1 || t->rows < 0 || (t->rows > 0 && !t->data))
        return smj_set_error(SMJ_EINVAL, "smj_csv_format: bad argument");
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int64_t rows = t->rows;
    const int cols = t->cols;
    char header[32];
    size_t hlen = 0;
    for (int i = 1; i <= cols; i++) hlen += (size_t)snprintf(header, sizeof header, "col%d", i) + 1;   // name + ',' or '\n'
    const int32_t *d_t = t->data;
    if (!t->on_device && rows > 0) {
        int32_t *p = (int32_t *)smj_ws(c, WS_T1, (size_t)rows * cols * 4);
        if (!p) return SMJ_ENOMEM;
        CUDA_TRY(cudaMemcpyAsync(p, t->data, (size_t)rows * cols * 4, cudaMemcpyHostToDevice, c->stream));
        d_t = p;
    }
    u64 body = 0;
    char *d_text = nullptr;
    if (rows > 0) {
        char *ws = (char *)smj_ws(c, WS_TMP_ROWS2, 64 + (size_t)rows * 4 + 8 + (size_t)rows * 8);
        if (!ws) return SMJ_ENOMEM;
        u64 *d_total = (u64 *)ws;
        u32 *d_len = (u32 *)(ws + 64);
        u64 *d_off = (u64 *)(ws + 64 + align_up((size_t)rows * 4, 8));
        const u32 grid = (u32)((rows + 255) / 256);
        csv_rowlen_kernel<<<grid, 256, 0, c->stream>>>(d_t, rows, cols, d_len);
        KERNEL_CHECK(c);
        csv_scan_kernel<<<1, CS_THREADS, 0, c->stream>>>(d_len, (u64)rows, d_off, d_total);
        KERNEL_CHECK(c);
        u64 *h = (u64 *)c->h_pinned;
        CUDA_TRY(cudaMemcpyAsync(h, d_total, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        body = h[0];
        d_text = (char *)smj_ws(c, WS_TMP_ROWS, body);
        if (!d_text) return SMJ_ENOMEM;
        csv_write_kernel<<<grid, 256, 0, c->stream>>>(d_t, rows, cols, d_off, d_text);
        KERNEL_CHECK(c);
    }
    char *host = nullptr;
    CUDA_TRY(cudaMallocHost((void **)&host, hlen + body + 1));
    size_t pos = 0;
    for (int i = 1; i <= cols; i++) {
        pos += (size_t)sprintf(host + pos, "col%d", i);
        host[pos++] = (i < cols) ? ',' : '\n';
    }
    if (body) CUDA_TRY(cudaMemcpyAsync(host + pos, d_text, body, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    host[pos + body] = 0;
    *text = host;
    *bytes = pos + body;
    return SMJ_OK;
}
