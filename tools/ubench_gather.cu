// tools/ubench_gather.cu -- how should random 16-byte row gathers from a table larger than L2 be issued on B200?
// Times 3.3M random int4 reads out of a 160 MB table (the join_materialize access pattern at the 10M-row config)
// with different load flavours; optional first arg = L2 fetch granularity limit to request (0 = leave default).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32;

template <int MODE>
__device__ __forceinline__ int4 ld16(const int4 *p)
{
    int4 v;
    if (MODE == 0) v = *p;
    if (MODE == 1) v = __ldg(p);
    if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 3) asm volatile("ld.global.nc.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 4) asm volatile("ld.global.nc.L2::128B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 5) asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 6) asm volatile("ld.global.lu.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 7) asm volatile("ld.global.cv.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int MODE, int ILP>
__global__ void gather(const int4 *__restrict__ tab, const u32 *__restrict__ idx, int n, int4 *__restrict__ out)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * ILP) {
        u32 r[ILP];
        int4 v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) r[u] = (i0 + u * stride < n) ? idx[i0 + u * stride] : 0u;
#pragma unroll
        for (int u = 0; u < ILP; u++) v[u] = ld16<MODE>(tab + r[u]);
#pragma unroll
        for (int u = 0; u < ILP; u++)
            if (i0 + u * stride < n) out[i0 + u * stride] = v[u];
    }
}

__global__ void fill_idx(u32 *idx, int n, u32 rows, u32 seed)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull + seed;
        x ^= x >> 31; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 29;
        idx[i] = (u32)(x % rows);
    }
}

template <int MODE, int ILP>
void run(const char *name, const int4 *tab, const u32 *idx, int n, int4 *out, int blocks)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        gather<MODE, ILP><<<blocks, 256>>>(tab, idx, n, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-34s ilp=%d blocks=%5d : %7.1f us  (%5.1f Mgathers/ms, %6.0f GB/s of 128-B lines)%s\n", name, ILP, blocks, best * 1e3,
           n / best / 1e3, n * 128.0 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char **argv)
{
    if (argc > 1 && atoi(argv[1]) > 0) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1]));
        size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
        printf("requested L2 fetch granularity %s -> %s, now %zu\n", argv[1], cudaGetErrorString(e), v);
    } else {
        size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
        printf("default L2 fetch granularity %zu\n", v);
    }
    const u32 rows = 10000000;   // 160 MB table
    const int n = 3333954;
    int4 *tab, *out; u32 *idx;
    cudaMalloc(&tab, (size_t)rows * 16); cudaMalloc(&out, (size_t)n * 16); cudaMalloc(&idx, (size_t)n * 4);
    cudaMemset(tab, 1, (size_t)rows * 16);
    fill_idx<<<1024, 256>>>(idx, n, rows, 12345u);
    cudaDeviceSynchronize();
    run<0, 4>("ld.global", tab, idx, n, out, 148 * 8);
    run<1, 4>("ld.global.nc (__ldg)", tab, idx, n, out, 148 * 8);
    run<1, 1>("ld.global.nc (__ldg)", tab, idx, n, out, 148 * 8);
    run<1, 8>("ld.global.nc (__ldg)", tab, idx, n, out, 148 * 8);
    run<1, 8>("ld.global.nc (__ldg)", tab, idx, n, out, 148 * 2);
    run<2, 4>("ld.global.nc.L1::no_allocate", tab, idx, n, out, 148 * 8);
    run<3, 4>("ld.global.nc.L2::64B", tab, idx, n, out, 148 * 8);
    run<4, 4>("ld.global.nc.L2::128B", tab, idx, n, out, 148 * 8);
    run<5, 4>("ld.global.cs", tab, idx, n, out, 148 * 8);
    run<6, 4>("ld.global.lu", tab, idx, n, out, 148 * 8);
    run<7, 4>("ld.global.cv", tab, idx, n, out, 148 * 8);
    // sequential read of the whole table for reference
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        int4 *dst; cudaMalloc(&dst, (size_t)rows * 16);
        cudaMemcpyAsync(dst, tab, (size_t)rows * 16, cudaMemcpyDeviceToDevice);
        cudaEventRecord(a);
        cudaMemcpyAsync(dst, tab, (size_t)rows * 16, cudaMemcpyDeviceToDevice);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("sequential copy of the 160 MB table: %.1f us (%.0f GB/s read+write)\n", ms * 1e3, 2.0 * rows * 16 / ms / 1e6);
    }
    return 0;
}
