// tools/cub_compare.cu -- same-box comparator for the sort stage (NOT the product path; libsmj.so links no CUB):
// cub::DeviceRadixSort from the CUDA 12.9 toolkit on the pair arrays smj_radix.cu sorts, so that the onesweep pass
// this repository hand-writes can be read against the library state of the art on the same B200.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/cub_compare tools/cub_compare.cu
// Three library formulations of "stable sort of (key, rowid) by the 32-bit key":
//   A  SortPairs<u32 key, u32 value>, bits [0, 32)          (structure-of-arrays, 4 + 4 B per pair per pass side)
//   B  SortPairs<u32 key, u32 value>, bits [0, 24)          (what a key range < 2^24 needs: three 8-bit digits)
//   C  SortKeys<u64>, bits [32, 64)                         (the packed (key << 32 | rowid) words libsmj moves)
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void fill(u32 *k, u32 *v, u64 *p, u32 n, u32 mask)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u64 x = (u64)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
        x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
        const u32 key = (u32)x & mask;
        k[i] = key; v[i] = i; p[i] = ((u64)key << 32) | i;
    }
}

template <class F> static float best_ms(F run, cudaStream_t st, int iters)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < iters + 2; it++) {
        CK(cudaEventRecord(e0, st));
        run();
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 2 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv)
{
    std::vector<u32> sizes = {5000000u, 10000000u, 50000000u, 200000000u};
    if (argc > 1) { sizes.clear(); for (int i = 1; i < argc; i++) sizes.push_back((u32)atol(argv[i])); }
    cudaStream_t st; CK(cudaStreamCreate(&st));
    printf("# cub::DeviceRadixSort (CUDA %d.%d toolkit CUB) on one B200; best of 5 after 2 warm-ups; the input is re-sorted in place\n",
           CUDART_VERSION / 1000, (CUDART_VERSION % 1000) / 10);
    printf("# (double buffers: every run sorts the previous run's output, i.e. sorted input -- radix sort time does not depend on order)\n");
    printf("n, variant, us, Gpairs/s, us_per_8bit_pass, GB/s_per_pass_at_16B_per_pair\n");
    for (u32 n : sizes) {
        u32 *k0, *k1, *v0, *v1; u64 *p0, *p1;
        CK(cudaMalloc(&k0, (size_t)n * 4)); CK(cudaMalloc(&k1, (size_t)n * 4));
        CK(cudaMalloc(&v0, (size_t)n * 4)); CK(cudaMalloc(&v1, (size_t)n * 4));
        CK(cudaMalloc(&p0, (size_t)n * 8)); CK(cudaMalloc(&p1, (size_t)n * 8));
        for (int variant = 0; variant < 3; variant++) {
            const u32 mask = variant == 1 ? 0x00ffffffu : 0xffffffffu;
            fill<<<148 * 8, 256, 0, st>>>(k0, v0, p0, n, mask);
            cub::DoubleBuffer<u32> dk(k0, k1), dv(v0, v1);
            cub::DoubleBuffer<u64> dp(p0, p1);
            size_t tb = 0;
            if (variant == 0) CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)n, 0, 32, st));
            if (variant == 1) CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)n, 0, 24, st));
            if (variant == 2) CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, dp, (int)n, 32, 64, st));
            void *tmp; CK(cudaMalloc(&tmp, tb + 256));
            const float ms = best_ms([&] {
                if (variant == 0) CK(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int)n, 0, 32, st));
                if (variant == 1) CK(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int)n, 0, 24, st));
                if (variant == 2) CK(cub::DeviceRadixSort::SortKeys(tmp, tb, dp, (int)n, 32, 64, st));
            }, st, 5);
            const int passes = variant == 1 ? 3 : 4;
            const char *names[] = {"A SortPairs<u32,u32> bits[0,32)", "B SortPairs<u32,u32> bits[0,24)", "C SortKeys<u64> bits[32,64)"};
            printf("%u, %s, %.1f, %.2f, %.1f, %.0f\n", n, names[variant], ms * 1e3, n / ms / 1e6, ms * 1e3 / passes,
                   16.0 * n / (ms / passes) / 1e6);
            CK(cudaFree(tmp));
        }
        CK(cudaFree(k0)); CK(cudaFree(k1)); CK(cudaFree(v0)); CK(cudaFree(v1)); CK(cudaFree(p0)); CK(cudaFree(p1));
    }
    return 0;
}
