// smj_convert.cu -- the reference's in-memory cell type at the boundary.
//
// The reference keeps every table as T[rows * cols] with T = int64_t (sort-merge-join/common.h:1-9; the UINT64 / DOUBLE
// branches are never selected), but every cell is an atoi() result (cpu_app.c:71, app.c:84) and is printed with %ld
// (cpu_app.c:290), so the values are int32 -- which is what the engine's tables hold (include/smj.h).  A host that keeps the
// reference's T* arrays hands them over as they are: smj_table_from_i64 narrows them on the GPU (the H2D copy of the
// 8-byte cells is the cost; a host loop over 80 M cells is 50x slower) and REFUSES a cell that is not an int32 value
// instead of truncating it; smj_table_to_i64 widens a result back into a T* array.
#include "smj_internal.h"
#include "smj_dev.cuh"

extern SmjCtx *g_ctx[8];
int smj_ensure_init(void);
int smj_alloc_out(SmjCtx *c, smj_table_t *out, int64_t rows, int cols);

namespace {

constexpr int CV_THREADS = 256;
constexpr int CV_VEC = 4;        // cells per thread and step: two 16-byte loads, one 16-byte store
constexpr size_t CV_CHUNK_CELLS = (size_t)8 << 20;   // host sources travel in 64 MB chunks through two staging buffers

// dst[i] = (int32) src[i]; *bad = 1 + index of some cell outside [INT32_MIN, INT32_MAX] (0: none)
__global__ void __launch_bounds__(CV_THREADS)
narrow_i64_kernel(const long long *__restrict__ src, size_t n, int32_t *__restrict__ dst, size_t index_base, unsigned long long *bad)
{
    const size_t stride = (size_t)gridDim.x * CV_THREADS * CV_VEC;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    size_t first_bad = ~(size_t)0;
    for (size_t i0 = ((size_t)blockIdx.x * CV_THREADS + threadIdx.x) * CV_VEC; i0 < n; i0 += stride) {
        long long v[CV_VEC];
        if (vec && i0 + CV_VEC <= n) {
            const longlong2 a = *reinterpret_cast<const longlong2 *>(src + i0), b = *reinterpret_cast<const longlong2 *>(src + i0 + 2);
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
            *reinterpret_cast<int4 *>(dst + i0) = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
#pragma unroll
            for (int k = 0; k < CV_VEC; k++)
                if (v[k] != (long long)(int)v[k] && first_bad == ~(size_t)0) first_bad = i0 + k;
        } else {
            for (int k = 0; k < CV_VEC && i0 + k < n; k++) {
                const long long x = src[i0 + k];
                dst[i0 + k] = (int)x;
                if (x != (long long)(int)x && first_bad == ~(size_t)0) first_bad = i0 + k;
            }
        }
    }
    if (first_bad != ~(size_t)0) atomicMax(bad, (unsigned long long)(index_base + first_bad) + 1ull);
}

__global__ void __launch_bounds__(CV_THREADS)
widen_i32_kernel(const int32_t *__restrict__ src, size_t n, long long *__restrict__ dst)
{
    const size_t stride = (size_t)gridDim.x * CV_THREADS * CV_VEC;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    for (size_t i0 = ((size_t)blockIdx.x * CV_THREADS + threadIdx.x) * CV_VEC; i0 < n; i0 += stride) {
        if (vec && i0 + CV_VEC <= n) {
            const int4 v = *reinterpret_cast<const int4 *>(src + i0);
            *reinterpret_cast<longlong2 *>(dst + i0) = make_longlong2(v.x, v.y);
            *reinterpret_cast<longlong2 *>(dst + i0 + 2) = make_longlong2(v.z, v.w);
        } else {
            for (int k = 0; k < CV_VEC && i0 + k < n; k++) dst[i0 + k] = src[i0 + k];
        }
    }
}

u32 cv_grid(SmjCtx *c, size_t cells)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const size_t want = (cells + (size_t)CV_THREADS * CV_VEC - 1) / ((size_t)CV_THREADS * CV_VEC);
    const size_t cap = (size_t)sms * 16;
    return (u32)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

extern "C" int smj_table_from_i64(const int64_t *cells, int64_t rows, int32_t cols, int cells_on_device, smj_table_t *out)
{
    SMJ_TRY(smj_ensure_init());
    if (!out || rows < 0 || cols < 1 || (rows > 0 && !cells)) return smj_set_error(SMJ_EINVAL, "smj_table_from_i64: bad argument");
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int to_device = out->on_device;
    const size_t n = (size_t)rows * (size_t)cols;
    smj_table_t dev = {nullptr, 0, cols, 1};
    SMJ_TRY(smj_alloc_out(c, &dev, rows, cols));
    // every failure below gives the device table back
    struct Guard { smj_table_t *t; ~Guard() { if (t && t->data) smj_table_free(t); } } guard = {&dev};
    unsigned long long *d_bad = (unsigned long long *)((char *)c->d_err + 64);   // a spare cell of the context's flag block
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, 8, c->stream));
    if (n) {
        if (cells_on_device) {
            narrow_i64_kernel<<<cv_grid(c, n), CV_THREADS, 0, c->stream>>>((const long long *)cells, n, dev.data, 0, d_bad);
            KERNEL_CHECK(c);
        } else {
            // host cells: 64 MB chunks alternate between two staging buffers, so chunk i + 1 crosses PCIe on the copy stream while
            // chunk i is narrowed
            const size_t chunk = n < CV_CHUNK_CELLS ? n : CV_CHUNK_CELLS;
            long long *stage = (long long *)smj_ws(c, WS_TMP_ROWS, 2 * chunk * 8);
            if (!stage) return SMJ_ENOMEM;
            cudaEvent_t copied[2] = {c->ev[12], c->ev[13]}, freed[2] = {c->ev[14], c->ev[15]};
            CUDA_TRY(cudaEventRecord(freed[0], c->stream));   // also orders the copy stream behind the staging buffer's last user
            CUDA_TRY(cudaEventRecord(freed[1], c->stream));
            int b = 0;
            for (size_t off = 0; off < n; off += chunk, b ^= 1) {
                const size_t m = n - off < chunk ? n - off : chunk;
                CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, freed[b], 0));
                CUDA_TRY(cudaMemcpyAsync(stage + (size_t)b * chunk, cells + off, m * 8, cudaMemcpyHostToDevice, c->copy_stream));
                CUDA_TRY(cudaEventRecord(copied[b], c->copy_stream));
                CUDA_TRY(cudaStreamWaitEvent(c->stream, copied[b], 0));
                narrow_i64_kernel<<<cv_grid(c, m), CV_THREADS, 0, c->stream>>>(stage + (size_t)b * chunk, m, dev.data + off, off, d_bad);
                KERNEL_CHECK(c);
                CUDA_TRY(cudaEventRecord(freed[b], c->stream));
            }
        }
    }
    unsigned long long *h = (unsigned long long *)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(h, d_bad, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h[0]) {
        const unsigned long long i = h[0] - 1;
        return smj_set_error(SMJ_ERANGE, "smj_table_from_i64: cell [%llu][%llu] is not an int32 value (the engine's cells are int32, like every "
                                         "atoi() result of the reference's load_csv)", i / (unsigned long long)cols, i % (unsigned long long)cols);
    }
    if (to_device) {
        *out = dev;
        guard.t = nullptr;
        return SMJ_OK;
    }
    out->data = nullptr; out->rows = 0; out->cols = cols;
    SMJ_TRY(smj_alloc_out(c, out, rows, cols));   // pinned host memory
    if (n) CUDA_TRY(cudaMemcpyAsync(out->data, dev.data, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJ_OK;
}

extern "C" int smj_table_to_i64(const smj_table_t *t, int64_t *cells, int cells_on_device)
{
    SMJ_TRY(smj_ensure_init());
    if (!t || t->rows < 0 || t->cols < 1 || (t->rows > 0 && (!t->data || !cells))) return smj_set_error(SMJ_EINVAL, "smj_table_to_i64: bad argument");
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)t->rows * (size_t)t->cols;
    if (n == 0) return SMJ_OK;
    const int32_t *d_src = t->data;
    if (!t->on_device) {
        int32_t *p = (int32_t *)smj_ws(c, WS_T1, n * 4);
        if (!p) return SMJ_ENOMEM;
        CUDA_TRY(cudaMemcpyAsync(p, t->data, n * 4, cudaMemcpyHostToDevice, c->stream));
        d_src = p;
    }
    long long *d_dst = (long long *)cells;
    if (!cells_on_device) {
        d_dst = (long long *)smj_ws(c, WS_TMP_ROWS, n * 8);
        if (!d_dst) return SMJ_ENOMEM;
    }
    widen_i32_kernel<<<cv_grid(c, n), CV_THREADS, 0, c->stream>>>(d_src, n, d_dst);
    KERNEL_CHECK(c);
    if (!cells_on_device) CUDA_TRY(cudaMemcpyAsync(cells, d_dst, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJ_OK;
}
