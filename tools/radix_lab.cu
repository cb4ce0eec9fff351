// tools/radix_lab.cu -- stand-alone lab for csrc/smj_radix.cu: times the onesweep passes warm (L2-resident, as inside
// the pipeline) on C2-shaped input, checks the result, and with -DSMJ_PHASE_TIMING prints where a CTA's cycles go.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSMJ_PHASE_TIMING -I. -o tools/bin/radix_lab tools/radix_lab.cu
#include "../pim-sort-merge-join_b200/csrc/smj_radix.cu"
#include <algorithm>
#include <vector>
#include <cstdarg>
#include <cstdio>

int smj_set_error(int code, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return code; }
bool smj_pdl_enabled(void) { return false; }
bool smj_stage_events(void) { return true; }
int smj_cuda_fail(cudaError_t e, const char *what, const char *file, int line) { fprintf(stderr, "CUDA %s at %s:%d (%s)\n", cudaGetErrorString(e), file, line, what); return SMJ_ECUDA; }

int main(int argc, char **argv)
{
    const u32 n = argc > 1 ? (u32)atol(argv[1]) : 5000000u;
    const u32 key_bits = argc > 2 ? (u32)atoi(argv[2]) : 25;
    SmjCtx ctx;
    ctx.device = 0;
    cudaSetDevice(0);
    cudaStreamCreate(&ctx.stream);
    cudaMalloc(&ctx.d_err, 256); cudaMemset(ctx.d_err, 0, 256);
    for (auto &e : ctx.pass_ev) cudaEventCreate(&e);
    std::vector<u64> h(n);
    u64 x = 88172645463325252ull;
    for (u32 i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const u32 key = (u32)(x >> 20) & ((key_bits >= 32) ? 0xffffffffu : ((1u << key_bits) - 1));
        h[i] = ((u64)(key ^ 0x80000000u) << 32) | i;
    }
    u64 *a, *b, *src; u32 *hist, *scratch;
    cudaMalloc(&a, (size_t)n * 8); cudaMalloc(&b, (size_t)n * 8); cudaMalloc(&src, (size_t)n * 8);
    cudaMalloc(&hist, 4096);
    const size_t sb = smj_radix_scratch_bytes(n);
    cudaMalloc(&scratch, sb);
    cudaMemcpy(src, h.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 6; it++) {
        cudaMemcpyAsync(a, src, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx.stream);
        cudaMemsetAsync(hist, 0, 4096, ctx.stream);
        cudaMemsetAsync(scratch, 0, sb, ctx.stream);
        smj_launch_radix_hist(&ctx, a, n, hist);
#ifdef SMJ_PHASE_TIMING
        if (it == 5) { unsigned long long z[16] = {}; cudaMemcpyToSymbolAsync(g_phase_cycles, z, sizeof z, 0, cudaMemcpyHostToDevice, ctx.stream); }
#endif
        ctx.pass_count = 0;
        cudaEventRecord(e0, ctx.stream);
        if (smj_radix_sort_pairs(&ctx, a, b, nullptr, n, hist, scratch) != SMJ_OK) return 1;
        cudaEventRecord(e1, ctx.stream);
        cudaStreamSynchronize(ctx.stream);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
        if (it == 5) {
            printf("passes:");
            for (int p = 0; p < ctx.pass_count; p++) { float pm; cudaEventElapsedTime(&pm, ctx.pass_ev[2 * p], ctx.pass_ev[2 * p + 1]); printf(" %.1f us", pm * 1e3); }
            printf("\n");
        }
    }
    printf("n=%u key_bits=%u: sort (scan + 4 passes) best %.1f us -> %.2f Gpairs/s, %.0f GB/s algorithmic per pass\n", n, key_bits, best * 1e3,
           n / best / 1e6, 4 * 16.0 * n / best / 1e6);
#ifdef SMJ_PHASE_TIMING
    {
        unsigned long long c[16]; cudaMemcpyFromSymbol(c, g_phase_cycles, sizeof c);
        const double tiles = 4.0 * ((n + RS_TILE - 1) / RS_TILE);
        const char *names[] = {"zero+sync", "load wait + count", "digit scan/publish", "rank+reorder", "look-back", "barrier wait", "next loads + copy-out", "rank stragglers"};
        double tot = 0; for (int i = 0; i < 8; i++) tot += (double)c[i];
        for (int i = 0; i < 8; i++) printf("  phase %d %-24s %8.0f cycles/tile  %5.1f %%\n", i, names[i], c[i] / tiles, 100.0 * c[i] / tot);
        printf("  total %.0f cycles/tile (thread 0 of each CTA)\n", tot / tiles);
    }
#endif
    if (argc > 3 && atoi(argv[3]) == 0) return 0;   // third argument 0: timing only (the CPU check of 400 M pairs takes minutes)
    std::vector<u64> out(n);
    cudaMemcpy(out.data(), a, (size_t)n * 8, cudaMemcpyDeviceToHost);
    std::stable_sort(h.begin(), h.end(), [](u64 p, u64 q) { return (p >> 32) < (q >> 32); });
    size_t bad = 0; for (u32 i = 0; i < n; i++) bad += out[i] != h[i];
    u32 err = 0; cudaMemcpy(&err, ctx.d_err, 4, cudaMemcpyDeviceToHost);
    printf("check: %zu mismatches, device flag %u\n", bad, err);
    return bad != 0;
}
