// tools/ubench.cu -- throughput of the warp/shared-memory primitives the ranking kernels choose between (sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench tools/ubench.cu ; run on a B200.
// Reports cycles per warp-instruction per SM with 4/8/16/32 resident warps per SM all issuing the same primitive.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;
#define ITER 2048

template <int OP>
__global__ void k(u32 *out, u32 seed, u64 *cycles)
{
    __shared__ u32 s[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = 0;
    __syncthreads();
    u32 x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 9176u;
    u32 acc = 0;
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u64 t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITER; i++) {
        x = x * 1664525u + 1013904223u;
        const u32 d = (x >> 24);
        if (OP == 0) acc += __match_any_sync(0xffffffffu, d);
        if (OP == 1) {   // 8-ballot match
            u32 peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const u32 bit = (d >> b) & 1u;
                const u32 m = __ballot_sync(0xffffffffu, bit);
                peers &= m ^ (bit - 1u);
            }
            acc += peers;
        }
        if (OP == 2) acc += atomicAdd(&s[(w & 7) * 256 + d], 1u);          // ATOMS with return, random bin
        if (OP == 3) atomicAdd(&s[(w & 7) * 256 + d], 1u);                  // ATOMS no return (RED), random bin
        if (OP == 4) acc += atomicOr(&s[(w & 7) * 256 + d], 1u << lane);    // ATOMS.OR with return
        if (OP == 5) acc += __shfl_sync(0xffffffffu, x, d & 31);
        if (OP == 6) acc += __ballot_sync(0xffffffffu, d & 1);
        if (OP == 7) { s[(w & 7) * 1024 + (d << 2 | (lane & 3))] = x; acc += s[(w & 7) * 1024 + ((d ^ 5) << 2 | (lane & 3))]; }  // STS+LDS random
        if (OP == 8) acc += __reduce_add_sync(0xffffffffu, d);
        if (OP == 9) { u64 *s8 = (u64 *)s; s8[((w & 7) * 512 + (x >> 23)) & 4095] = x; }   // STS.64 random
        if (OP == 10) acc += __popc(x) + __ffs(d);
        if (OP == 11) atomicAdd(&s[d], 1u);                                  // block-shared bins (cross-warp contention)
    }
    u64 t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + s[threadIdx.x];
}

template <int OP>
void run(const char *name)
{
    u32 *out; u64 *cyc;
    cudaMalloc(&out, 148 * 1024 * 4 * 4);
    cudaMalloc(&cyc, 148 * 8 * 8);
    printf("%-28s", name);
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        const int blocks = 148;
        k<OP><<<blocks, threads>>>(out, 1, cyc);
        cudaDeviceSynchronize();
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k<OP><<<blocks, threads>>>(out, 2, cyc);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        u64 h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
        // cycles per warp-instruction per SM = cycles / (ITER * warps)
        printf("  w=%2d: %6.2f cyc/winst/SM (%.3f ms)", warps, avg / ((double)ITER * warps), ms);
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf(" ERR %s", cudaGetErrorString(e));
    }
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<10>("alu popc+ffs (baseline)");
    run<0>("match.any 8-bit");
    run<1>("8x ballot match");
    run<6>("1x ballot");
    run<2>("ATOMS.add ret, warp bins");
    run<3>("ATOMS.add noret, warp bins");
    run<11>("ATOMS.add noret, block bins");
    run<4>("ATOMS.or ret");
    run<5>("shfl idx");
    run<8>("redux.add");
    run<7>("STS+LDS random 32b");
    run<9>("STS.64 random");
    return 0;
}
