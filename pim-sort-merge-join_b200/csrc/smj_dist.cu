// smj_dist.cu -- key-range partitioned sort-merge-join across the GPUs of one box.
//
// Replaces the reference's host-mediated data movement between stages -- the per-DPU dpu_push_xfer gathers and
// scatters (sort-merge-join/app.c:222-288), the log-depth merge tournament that round-trips whole tables through
// host memory (app.c:413-547) and the host-side key-range split for the join (app.c:585-633) -- with ONE exchange.
//
// Default path (the "fabric" path; one process per GPU under torchrun, or all GPUs from one process for the C driver):
//   1. splitters: every rank stores regular row samples of both tables (predicate applied) into every peer's mailbox
//      and sorts the same G x 2S samples itself -- ONE kernel, no collective call, identical splitters everywhere;
//   2. select fused with key-range partitioning of the ROWS (smj_partition.cu): survivors grouped by destination rank
//      inside their tile's slot, original order kept inside each bucket; a many-CTA scan gives every segment its offset;
//   3. counts: every rank stores its G bucket totals into every peer's count matrix and derives, from the same matrix,
//      where its buckets start in the owners' receive buffers -- on the device, never waited for on the host;
//   4. exchange fused into the compaction kernel: every tile's survivors are stored straight into the destination
//      ranks' receive buffers (peer mappings: NVLink stores from the SMs), then an arrival flag;
//   5. the single-GPU pipeline (smj_run_*, select disabled) sorts and joins what arrived, sized from the device-resident
//      row count: runs sit in source-rank order with original order inside each, so the stable sort reproduces the
//      reference's order;
//   6. the result shards, concatenated in rank order, are the single-GPU result (splitters are key values, so all
//      rows of one key meet on one rank and the zip pairing of equal keys is local).
// Receive buffers are sized once for an upper bound; a step whose shares do not fit stores nothing, says so in a verdict
// that every rank computes from the same matrix, and is re-run after a collective re-sizing.
//
// SMJ_DIST_EXCHANGE=nccl: the same partitioning with NCCL collectives, a host wait for the counts and one grouped
// ncclSend/ncclRecv all-to-all.  SMJ_DIST_MODE=merge, smj_run_multi_sorted (sort first; also what tables of more than 32
// columns take): local select + sort + payload gather, splitters from samples of the sorted keys (host), bucket bounds by
// binary search, one grouped ncclSend/ncclRecv of the contiguous sorted slices, a merge-path merge tree over the G
// received runs per table (ties: lower source rank first = original row order), local join.
//
// NCCL is loaded lazily (dlopen of libnccl.so.2) so the single-GPU library has no hard dependency on it and a
// process that already loaded torch's NCCL shares that copy.
#include "smj_internal.h"
#include "smj_dev.cuh"

#include <dlfcn.h>
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

// ------------------------------------------------------------------ minimal NCCL surface, resolved at run time
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclUint32 = 3, ncclUint64 = 5 };   // ncclDataType_t values (nccl.h)

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int nccl_load()
{
    if (g_nccl.lib) return SMJ_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return smj_set_error(SMJ_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return smj_set_error(SMJ_ENCCL, "libnccl lacks %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return SMJ_OK;
}

int nccl_fail(ncclResult_t r, const char *what)
{
    return smj_set_error(SMJ_ENCCL, "NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
}
#define NCCL_TRY(x) do { ncclResult_t r_ = (x); if (r_ != 0) return nccl_fail(r_, #x); } while (0)

constexpr int DIST_SAMPLES = 256;   // regular samples per table per rank

// samples[i] = key of the pair at position floor((2i+1) * m / (2S)) of the sorted pairs (0xffffffff when m == 0)
__global__ void sample_keys_kernel(const u64 *__restrict__ pairs, u32 m, u32 *samples, int S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    if (m == 0) { samples[i] = 0xffffffffu; return; }
    const u64 pos = ((u64)(2 * i + 1) * m) / (u64)(2 * S);
    samples[i] = pair_key(pairs[pos < m ? pos : m - 1]);
}

// bnd[b] = first position of the sorted pairs whose key is >= splitters[b-1]; bnd[0] = 0, bnd[G] = m
__global__ void bucket_bounds_kernel(const u64 *__restrict__ pairs, u32 m, const u32 *__restrict__ splitters, int world, u32 *bnd)
{
    const int b = threadIdx.x;
    if (b > world) return;
    if (b == 0) bnd[0] = 0;
    else if (b == world) bnd[world] = m;
    else bnd[b] = lower_bound_key(pairs, 0, m, splitters[b - 1]);
}

}  // namespace

// ------------------------------------------------------------------ host-side planning (pure C, tested on CPU)
// Splitters from the gathered samples: sorted ascending, splitter b (b = 1..G-1) is the sample at quantile b/G.
// Keys are compared in the engine's order-preserving unsigned form (int32 key ^ 0x80000000).  Buckets are
// [splitter[b-1], splitter[b]) with splitter[-1] = 0 and splitter[G-1] = +inf, so equal keys never straddle ranks.
extern "C" int smj_plan_splitters(const uint32_t *samples, int64_t n_samples, int world, uint32_t *splitters)
{
    if (!samples || !splitters || world < 1 || n_samples < 0) return smj_set_error(SMJ_EINVAL, "smj_plan_splitters: bad arguments");
    std::vector<uint32_t> s;
    s.reserve((size_t)n_samples);
    for (int64_t i = 0; i < n_samples; i++)
        if (samples[i] != 0xffffffffu) s.push_back(samples[i]);   // 0xffffffff marks "no sample" (empty table on that rank)
    std::sort(s.begin(), s.end());
    for (int b = 1; b < world; b++) {
        if (s.empty()) { splitters[b - 1] = 0xffffffffu; continue; }
        size_t pos = (size_t)((unsigned long long)b * s.size() / (unsigned long long)world);
        if (pos >= s.size()) pos = s.size() - 1;
        splitters[b - 1] = s[pos];
    }
    return SMJ_OK;
}

// Exchange plan from the all-gathered bucket sizes: counts[src * world + dst] rows go from src to dst.
// recv_offsets[src] = row offset of src's run inside rank `me`'s receive buffer (runs in source-rank order, which
// keeps equal keys in original row order), *recv_total = rows `me` receives.
extern "C" int smj_plan_exchange(const int64_t *counts, int world, int me, int64_t *recv_offsets, int64_t *recv_total)
{
    if (!counts || !recv_offsets || !recv_total || world < 1 || me < 0 || me >= world)
        return smj_set_error(SMJ_EINVAL, "smj_plan_exchange: bad arguments");
    int64_t run = 0;
    for (int src = 0; src < world; src++) {
        recv_offsets[src] = run;
        const int64_t c = counts[(size_t)src * world + me];
        if (c < 0) return smj_set_error(SMJ_EINVAL, "smj_plan_exchange: negative count");
        run += c;
    }
    *recv_total = run;
    return SMJ_OK;
}

// The fabric path's step 3 on the host (the device twin is dist_counts_kernel): from the gathered matrix
// counts[src * world + dst] every rank derives, identically, where ITS bucket b starts inside rank b's receive buffer
// (row0[b] = rows the lower ranks send to b: runs sit in source-rank order), how many rows it will hold (*rows_mine) and the
// collective verdict: 1 when some rank's share exceeds cap_rows -- then nobody stores anything (*rows_mine = 0) and the
// buffers are re-sized for *need_rows, the largest share, before the step is run again.
extern "C" int smj_plan_fabric(const int64_t *counts, int world, int me, int64_t cap_rows, int64_t *row0, int64_t *rows_mine,
                               int *verdict, int64_t *need_rows)
{
    if (!counts || !row0 || !rows_mine || !verdict || !need_rows || world < 1 || world > SMJ_MAX_G || me < 0 || me >= world)
        return smj_set_error(SMJ_EINVAL, "smj_plan_fabric: bad arguments");
    int over = 0;
    int64_t worst = 0, mine = 0;
    for (int dst = 0; dst < world; dst++) {
        int64_t tot = 0, before = 0;
        for (int src = 0; src < world; src++) {
            const int64_t c = counts[(size_t)src * world + dst];
            if (c < 0) return smj_set_error(SMJ_EINVAL, "smj_plan_fabric: negative count");
            tot += c;
            if (src < me) before += c;
        }
        row0[dst] = before;
        if (tot > cap_rows) over = 1;
        if (tot > worst) worst = tot;
        if (dst == me) mine = tot;
    }
    *verdict = over;
    *rows_mine = over ? 0 : mine;
    *need_rows = worst;
    return SMJ_OK;
}

// ------------------------------------------------------------------ the peer fabric
// Every rank owns a WINDOW in device memory that every other rank maps (CUDA IPC between processes, peer access inside
// one process): flag mailboxes, the sample mailbox and the G x G count matrices.  A rank "all-gathers" by storing its
// piece into every peer's window and then its step sequence number into its flag there (st.release.sys); it waits by
// polling its OWN window's flags (ld.acquire.sys).  Sequence numbers only grow, so nothing is ever reset and a flag left
// over from an earlier step can never satisfy a later wait.  The waits are bounded (DIST_WAIT_NS): a rank that never
// arrives surfaces as SMJ_EINTERNAL on its peers, not as a hung GPU.
extern SmjCtx *g_ctx[8];
extern int g_nctx;
int smj_stage_in(SmjCtx *c, const smj_table_t *t, int slot, const int32_t **d);
int smj_alloc_out(SmjCtx *c, smj_table_t *out, int64_t rows, int cols);
int smj_check_device_flag(SmjCtx *c);
int smj_join_pairs_to_table(SmjCtx *c, const u64 *pl, u32 m1, const u64 *pr, u32 m2, const int32_t *d_t1, int c1, const int32_t *d_t2,
                            int c2, int key2, smj_table_t *out, int64_t *rows_out);
int smj_sorted_pairs_of_table(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int select_all,
                              int key_col, int table_idx, u64 **d_sorted, int64_t *m_out);
int smj_run_single(SmjCtx *c, const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats);
int smj_ensure_init(void);

namespace {

constexpr int DIST_S_MAX = 512;                      // samples per table per rank
constexpr unsigned long long DIST_WAIT_NS = 30ull * 1000 * 1000 * 1000;
enum { PH_SAMPLES = 0, PH_COUNTS0, PH_COUNTS1, PH_XCHG0, PH_XCHG1, PH_N = 8 };
#define SMJ_ERR_DIST_WAIT 4u

struct DistWindow {                                   // mapped by every peer
    u64 flag[PH_N][SMJ_MAX_G];                        // [phase][source rank]: the last step the source signalled
    u64 counts[2][SMJ_MAX_G][SMJ_MAX_G];              // [table][src][dst]: rows src sends to dst (row src written by src)
    u32 samples[SMJ_MAX_G][2 * DIST_S_MAX];           // [src][table * S + i]
};

struct DistLocal {                                    // private to the rank; copied to the host at the end of a step
    u32 split[SMJ_MAX_G];                             // G - 1 key splitters
    u64 rows[2];                                      // rows of table t this rank holds after the exchange (0 after an overflow verdict)
    u64 row0[2][SMJ_MAX_G];                           // first row of this rank's bucket b inside rank b's receive buffer
    u64 matrix[2][SMJ_MAX_G][SMJ_MAX_G];              // the gathered count matrices
    u32 verdict[2];                                   // non-zero: some rank's receive buffer cannot take its share of table t
    u32 pad[2];
    u64 seq_now;                                      // this step's sequence number (written by the step's first kernel)
    u64 arrived[2];                                   // = seq_now once every rank's rows of table t have landed here
};

struct DistPeers { DistWindow *win[SMJ_MAX_G]; };     // every rank's window as seen from this rank's device

__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p)
{
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spins until *p >= seq; false (and the error flag set) after DIST_WAIT_NS
__device__ __forceinline__ bool wait_flag(const u64 *p, u64 seq, u32 *err)
{
    if (ld_acquire_sys(p) >= seq) return true;
    const unsigned long long t0 = global_ns();
    for (;;) {
        for (int i = 0; i < 64; i++)
            if (ld_acquire_sys(p) >= seq) return true;
        if (global_ns() - t0 > DIST_WAIT_NS) { atomicExch(err, SMJ_ERR_DIST_WAIT); return false; }
    }
}

// "I have arrived at `phase` of step `seq`" to every rank, then wait until every rank has.  Everything this rank's
// earlier kernels on the stream stored -- into its own or into peer memory -- happens-before the flag store (stream order,
// then a system-scope release), so a peer that sees the flag sees the data.
__device__ __forceinline__ void signal_and_wait(const DistPeers &P, int me, int G, int phase, u64 seq, u32 *err)
{
    const int tid = threadIdx.x;
    __threadfence_system();
    __syncthreads();
    if (tid < G) st_release_sys(&P.win[tid]->flag[phase][me], seq);
    if (tid < G) wait_flag(&P.win[me]->flag[phase][tid], seq, err);
    __syncthreads();
}

// ... and, for kernels of this rank that run on ANOTHER stream (the local pipeline's select of the table that arrives last
// starts with a device-side wait, SmjWait), the arrival cell of table t
__global__ void __launch_bounds__(32) dist_arrive_kernel(const DistPeers P, DistLocal *loc, int me, int G, int t, u64 seq, u32 *err)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");   // (launched with programmatic serialization; it never triggers its successor early)
    signal_and_wait(P, me, G, PH_XCHG0 + t, seq, err);
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&loc->arrived[t]), "l"(seq) : "memory");
    }
}

// ---- step 1, ONE kernel, one CTA: regular row samples of both tables (predicate applied) -> every peer's sample mailbox
// -> wait for everybody's -> sort all G x 2S samples in shared memory (bitonic) -> G - 1 splitters.  Every rank sorts the
// same samples, so every rank derives the same splitters; no collective library call, no host round trip.
struct SampleJob { const int32_t *in; int64_t n; int cols, sel_col; int32_t sel_val; int select_all, key_col; };
constexpr int SPL2_THREADS = 1024;

__global__ void __launch_bounds__(SPL2_THREADS)
dist_splitters_kernel(const DistPeers P, DistLocal *loc, int me, int G, u64 seq, int S, const SampleJob j0, const SampleJob j1,
                      int n_pow2, u32 *err)
{
    extern __shared__ u32 s_s[];
    __shared__ int s_valid;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    if (tid == 0) { s_valid = 0; loc->verdict[0] = 0; loc->verdict[1] = 0; loc->rows[0] = 0; loc->rows[1] = 0; loc->seq_now = seq; }
    if (tid < 2 * S) {
        const SampleJob &J = tid < S ? j0 : j1;
        const int i = tid < S ? tid : tid - S;
        u32 v = 0xffffffffu;
        if (J.n > 0) {
            int64_t pos = (int64_t)(((unsigned long long)(2 * i + 1) * (unsigned long long)J.n) / (unsigned long long)(2 * S));
            if (pos >= J.n) pos = J.n - 1;
            const int32_t *r = J.in + pos * J.cols;
            if (J.select_all || r[J.sel_col] > J.sel_val) v = (u32)r[J.key_col] ^ 0x80000000u;
        }
        for (int r = 0; r < G; r++) P.win[r]->samples[me][tid] = v;
    }
    signal_and_wait(P, me, G, PH_SAMPLES, seq, err);
    const DistWindow *mine = P.win[me];
    const int n_samples = G * 2 * S;
    // Bitonic sort, four elements per thread (element i = 4 * tid + e): partners closer than 4 are in the thread's own
    // registers, closer than 128 in the same warp (shuffles), only the others go through shared memory -- 15 block-wide
    // stages for 4096 samples instead of 78 (the first version walked every stage through shared memory: 42 us).
    {
        const int nthr = n_pow2 >> 2;                 // threads that hold elements (n_pow2 >= 128, a multiple of 128)
        const bool act = tid < nthr;
        u32 v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int i = 4 * tid + e;
            v[e] = (act && i < n_samples) ? __ldcg(&mine->samples[i / (2 * S)][i % (2 * S)]) : 0xffffffffu;
        }
        for (int k = 2; k <= n_pow2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= 128) {
                    if (act) {
#pragma unroll
                        for (int e = 0; e < 4; e++) s_s[4 * tid + e] = v[e];
                    }
                    __syncthreads();
                    if (act) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int i = 4 * tid + e;
                            const u32 o = s_s[i ^ j];
                            const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                            v[e] = keep_min ? (v[e] < o ? v[e] : o) : (v[e] > o ? v[e] : o);
                        }
                    }
                    __syncthreads();
                } else if (j >= 4) {
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int i = 4 * tid + e;
                        const u32 o = __shfl_xor_sync(FULL_MASK, v[e], j >> 2);
                        const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                        v[e] = keep_min ? (v[e] < o ? v[e] : o) : (v[e] > o ? v[e] : o);
                    }
                } else {   // j = 2: pairs (0,2) (1,3); j = 1: pairs (0,1) (2,3) -- compile-time register indices
                    const bool up = ((4 * tid) & k) == 0;   // k >= 2 > the element bits only when k > 2; for k = 2 see below
#define CE(a, b, asc) { const u32 lo_ = v[a] < v[b] ? v[a] : v[b], hi_ = v[a] < v[b] ? v[b] : v[a]; v[a] = (asc) ? lo_ : hi_; v[b] = (asc) ? hi_ : lo_; }
                    if (j == 2) {
                        // direction of element i: (i & k) == 0; for k >= 8 it is the thread's, for k = 4 elements 0..3 share (i & 4) = tid's bit
                        CE(0, 2, up) CE(1, 3, up)
                    } else {
                        if (k == 2) { CE(0, 1, true) CE(2, 3, false) }   // (i & 2) == 0 for e = 0,1; != 0 for e = 2,3
                        else { CE(0, 1, up) CE(2, 3, up) }
                    }
#undef CE
                }
            }
        if (act) {
#pragma unroll
            for (int e = 0; e < 4; e++) s_s[4 * tid + e] = v[e];
        }
        __syncthreads();
    }
    int local = 0;
    for (int i = tid; i < n_pow2; i += SPL2_THREADS) local += s_s[i] != 0xffffffffu;
    atomicAdd(&s_valid, local);
    __syncthreads();
    const int nv = s_valid;
    if (tid >= 1 && tid < G) {   // the device twin of smj_plan_splitters: the "none" marks sort last, splitter b = quantile b / G
        u32 v = 0xffffffffu;
        if (nv > 0) {
            long long pos = (long long)tid * nv / G;
            if (pos >= nv) pos = nv - 1;
            v = s_s[pos];
        }
        loc->split[tid - 1] = v;
    }
}

// ---- step 3, one CTA per table: this rank's G bucket totals into every peer's count matrix, wait for everybody's, then
// everything the exchange needs, computed identically on every rank from the same matrix: where my bucket b starts inside
// rank b's receive buffer (rows the lower ranks send there), how many rows I will hold, and the VERDICT -- does any rank's
// share exceed its receive capacity?  If so every rank skips the stores of this table and the host re-sizes the buffers
// collectively after the step (the counts are never waited for on the host in a step that fits).
__global__ void __launch_bounds__(64)
dist_counts_kernel(const DistPeers P, DistLocal *loc, int me, int G, u64 seq, int t, const u64 *__restrict__ bucket_total, u64 cap_rows, u32 *err)
{
    __shared__ u64 s_m[SMJ_MAX_G][SMJ_MAX_G];
    __shared__ u32 s_over;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    if (tid == 0) s_over = 0;
    if (tid < G * G) P.win[tid / G]->counts[t][me][tid % G] = bucket_total[tid % G];
    signal_and_wait(P, me, G, PH_COUNTS0 + t, seq, err);
    if (tid < G * G) {
        const u64 v = __ldcg(&P.win[me]->counts[t][tid / G][tid % G]);
        s_m[tid / G][tid % G] = v;
        loc->matrix[t][tid / G][tid % G] = v;
    }
    __syncthreads();
    if (tid < G) {
        u64 tot = 0, before = 0;
        for (int src = 0; src < G; src++) { tot += s_m[src][tid]; if (src < me) before += s_m[src][tid]; }
        loc->row0[t][tid] = before;
        if (tot > cap_rows) atomicOr(&s_over, 1u);
    }
    __syncthreads();
    if (tid == 0) {
        u64 mine = 0;
        for (int src = 0; src < G; src++) mine += s_m[src][me];
        loc->verdict[t] = s_over;
        loc->rows[t] = s_over ? 0ull : mine;
    }
}

// ------------------------------------------------------------------ host side of the fabric
struct DistRank {
    SmjCtx *c = nullptr;
    int me = 0;
    DistWindow *win = nullptr;
    DistLocal *loc = nullptr, *h_loc = nullptr;
    DistPeers peers = {};
    void *ipc_win[SMJ_MAX_G] = {};                    // IPC mappings this process opened (closed at shutdown / re-setup)
    void *ipc_recv[2][SMJ_MAX_G] = {};
    int32_t *recv[2] = {nullptr, nullptr};            // my receive buffers
    int32_t *peer_recv[2][SMJ_MAX_G] = {};            // everybody's, as seen from my device
    int64_t cap_rows[2] = {0, 0};
    int cap_cols[2] = {0, 0};
    cudaStream_t aux = nullptr;                       // table 2's partition / exchange chain
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev[8] = {};
    // one step in flight
    SmjRun run;
    smj_table_t blk[2];
    int64_t launches0 = 0;
    int32_t *slots[2] = {};
    char *pscr[2] = {};
    int none[2] = {};
};
enum { DE_START, DE_H2D, DE_SPLIT, DE_PART, DE_XCHG, DE_XCHG2 };

struct DistState {
    bool active = false;          // smj_init_dist: one process per GPU
    bool local = false;           // one process drives all ranks (smj_run with nr_gpus > 1 after smj_init)
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;    // multi-process: bootstrap collectives + the NCCL exchange paths
    u64 seq = 0;                  // step sequence number, the same on every rank
    int nlocal = 0;
    DistRank rk[SMJ_MAX_G];
} g_dist;

int dist_rank_create(DistRank &K, SmjCtx *c, int me)
{
    K = DistRank();
    K.c = c; K.me = me;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMalloc((void **)&K.win, sizeof(DistWindow)));
    CUDA_TRY(cudaMemset(K.win, 0, sizeof(DistWindow)));
    CUDA_TRY(cudaMalloc((void **)&K.loc, sizeof(DistLocal)));
    CUDA_TRY(cudaMemset(K.loc, 0, sizeof(DistLocal)));
    CUDA_TRY(cudaMallocHost((void **)&K.h_loc, sizeof(DistLocal)));
    {   // every kernel a step launches is loaded now: a lazy load must not happen while a peer spins on this rank (smj_preload_*)
        cudaFuncAttributes a;
        cudaFuncGetAttributes(&a, dist_arrive_kernel);
        cudaFuncGetAttributes(&a, dist_splitters_kernel);
        cudaFuncGetAttributes(&a, dist_counts_kernel);
        cudaGetLastError();
        smj_preload_partition(); smj_preload_select(); smj_preload_radix(); smj_preload_join();
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&K.aux, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&K.ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&K.ev_join, cudaEventDisableTiming));
    for (auto &e : K.ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaDeviceSynchronize());
    return SMJ_OK;
}

void dist_rank_unmap_recv(DistRank &K)
{
    for (int t = 0; t < 2; t++)
        for (int r = 0; r < SMJ_MAX_G; r++) {
            if (K.ipc_recv[t][r]) cudaIpcCloseMemHandle(K.ipc_recv[t][r]);
            K.ipc_recv[t][r] = nullptr;
            K.peer_recv[t][r] = nullptr;
        }
    cudaGetLastError();
}

void dist_rank_destroy(DistRank &K)
{
    if (!K.c) return;
    cudaSetDevice(K.c->device);
    cudaDeviceSynchronize();
    dist_rank_unmap_recv(K);
    for (int r = 0; r < SMJ_MAX_G; r++) if (K.ipc_win[r]) cudaIpcCloseMemHandle(K.ipc_win[r]);
    if (K.win) cudaFree(K.win);
    if (K.loc) cudaFree(K.loc);
    if (K.h_loc) cudaFreeHost(K.h_loc);
    if (K.aux) cudaStreamDestroy(K.aux);
    if (K.ev_fork) cudaEventDestroy(K.ev_fork);
    if (K.ev_join) cudaEventDestroy(K.ev_join);
    for (auto &e : K.ev) if (e) cudaEventDestroy(e);
    cudaGetLastError();
    K = DistRank();
}

// host-side all-gather of `bytes` per rank over the bootstrap communicator (multi-process mode; setup paths only)
int boot_allgather(SmjCtx *c, const void *mine, void *all, size_t bytes)
{
    const int G = g_dist.world;
    char *d = (char *)smj_ws(c, WS_SAMPLES, bytes * (size_t)(G + 1) + 256);
    if (!d) return SMJ_ENOMEM;
    CUDA_TRY(cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllGather(d, d + bytes, bytes, /*ncclUint8*/ 1, g_dist.comm, c->stream));
    CUDA_TRY(cudaMemcpyAsync(all, d + bytes, bytes * (size_t)G, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJ_OK;
}

// maps every rank's window into this process (multi-process mode, once, at smj_init_dist)
int dist_map_windows_ipc(DistRank &K)
{
    const int G = g_dist.world;
    cudaIpcMemHandle_t mine;
    CUDA_TRY(cudaIpcGetMemHandle(&mine, K.win));
    std::vector<cudaIpcMemHandle_t> all((size_t)G);
    SMJ_TRY(boot_allgather(K.c, &mine, all.data(), sizeof mine));
    for (int r = 0; r < G; r++) {
        if (r == K.me) { K.peers.win[r] = K.win; continue; }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return smj_cuda_fail(e, "cudaIpcOpenMemHandle(window): peer access between the GPUs is required", __FILE__, __LINE__);
        K.ipc_win[r] = p;
        K.peers.win[r] = (DistWindow *)p;
    }
    return SMJ_OK;
}

// ---- receive buffers.  Sized ONCE per shape for an upper bound of what a rank can receive (cap_pct % of the largest
// input block plus slack), so that a step needs no host round trip for the counts; (re)sized collectively: at the first
// step, when the column counts change, or after a step whose verdict said some rank's share did not fit (need_rows = what
// the matrix asked for).  All ranks take the same decision from the same facts, so nobody waits in a collective alone.
int dist_cap_pct(void)
{
    static const int pct = [] {
        const char *e = getenv("SMJ_DIST_CAP_PCT");
        const long v = e ? atol(e) : 0;
        return (int)((v >= 1 && v <= 100000) ? v : 150);
    }();
    return pct;
}

int64_t dist_cap_from(const int64_t *n_local, const int64_t *need, int G)
{
    int64_t cap = 0;
    for (int r = 0; r < G; r++) {
        const int64_t a = n_local[r] / 100 * dist_cap_pct() + (n_local[r] % 100) * dist_cap_pct() / 100 + 4096;
        const int64_t b = need[r] + need[r] / 4 + 4096;
        if (a > cap) cap = a;
        if (need[r] > 0 && b > cap) cap = b;
    }
    if (cap > SMJ_MAX_SORT_ROWS) cap = SMJ_MAX_SORT_ROWS;   // the local sort's limit; a share beyond it is refused in the verdict check
    return cap;
}

// multi-process: K = the one local rank; n_local / need: this rank's rows per table / rows it must be able to receive (0: unknown)
int dist_setup_recv_ipc(DistRank &K, const int64_t n_local[2], const int cols[2], const int64_t need[2])
{
    const int G = g_dist.world;
    SmjCtx *c = K.c;
    int64_t mine[4] = {n_local[0], n_local[1], need[0], need[1]};
    std::vector<int64_t> all((size_t)4 * G);
    SMJ_TRY(boot_allgather(c, mine, all.data(), sizeof mine));
    // everybody closes its mappings of the old buffers BEFORE any owner frees one (the gather above and below are the barriers)
    dist_rank_unmap_recv(K);
    int64_t dummy = 0;
    std::vector<int64_t> dummies((size_t)G);
    SMJ_TRY(boot_allgather(c, &dummy, dummies.data(), sizeof dummy));
    cudaIpcMemHandle_t hnd[2];
    for (int t = 0; t < 2; t++) {
        std::vector<int64_t> nl((size_t)G), nd((size_t)G);
        for (int r = 0; r < G; r++) { nl[(size_t)r] = all[(size_t)4 * r + t]; nd[(size_t)r] = all[(size_t)4 * r + 2 + t]; }
        const int64_t cap = dist_cap_from(nl.data(), nd.data(), G);
        K.recv[t] = (int32_t *)smj_ws(c, t ? WS_DIST_RECV2 : WS_DIST_RECV1, (size_t)cap * cols[t] * 4);
        if (!K.recv[t]) return SMJ_ENOMEM;
        K.cap_rows[t] = cap;
        K.cap_cols[t] = cols[t];
        // The owner touches its buffer before any peer can: the very first step on a fresh 8-GPU allocation once returned a
        // result whose checksum differed (row count right, every later step right, never with SMJ_DIST_POISON) -- whatever
        // happens to device memory between cudaMalloc and its first use must not race with a peer's first stores.
        CUDA_TRY(cudaMemsetAsync(K.recv[t], 0, (size_t)cap * cols[t] * 4, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaIpcGetMemHandle(&hnd[t], K.recv[t]));
    }
    std::vector<cudaIpcMemHandle_t> hall((size_t)2 * G);
    SMJ_TRY(boot_allgather(c, hnd, hall.data(), sizeof hnd));
    for (int r = 0; r < G; r++)
        for (int t = 0; t < 2; t++) {
            if (r == K.me) { K.peer_recv[t][r] = K.recv[t]; continue; }
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, hall[(size_t)2 * r + t], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return smj_cuda_fail(e, "cudaIpcOpenMemHandle(receive buffer)", __FILE__, __LINE__);
            K.ipc_recv[t][r] = p;
            K.peer_recv[t][r] = (int32_t *)p;
        }
    return SMJ_OK;
}

// one process, G ranks: the same decision, no collectives; peers address each other's memory directly (peer access)
int dist_setup_recv_local(const int64_t n_local[][2], const int cols[2], const int64_t need[][2])
{
    const int G = g_dist.world;
    for (int t = 0; t < 2; t++) {
        int64_t nl[SMJ_MAX_G], nd[SMJ_MAX_G];
        for (int r = 0; r < G; r++) { nl[r] = n_local[r][t]; nd[r] = need[r][t]; }
        const int64_t cap = dist_cap_from(nl, nd, G);
        for (int r = 0; r < G; r++) {
            DistRank &K = g_dist.rk[r];
            CUDA_TRY(cudaSetDevice(K.c->device));
            K.recv[t] = (int32_t *)smj_ws(K.c, t ? WS_DIST_RECV2 : WS_DIST_RECV1, (size_t)cap * cols[t] * 4);
            if (!K.recv[t]) return SMJ_ENOMEM;
            K.cap_rows[t] = cap;
            K.cap_cols[t] = cols[t];
            CUDA_TRY(cudaMemsetAsync(K.recv[t], 0, (size_t)cap * cols[t] * 4, K.c->stream));   // owner's first touch (see the IPC variant)
            CUDA_TRY(cudaStreamSynchronize(K.c->stream));
        }
        for (int r = 0; r < G; r++)
            for (int q = 0; q < G; q++) g_dist.rk[r].peer_recv[t][q] = g_dist.rk[q].recv[t];
    }
    return SMJ_OK;
}

float dist_ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ms;
}

bool dist_two_streams(void)
{
    static const bool two = !(getenv("SMJ_DIST_STREAMS") && atoi(getenv("SMJ_DIST_STREAMS")) == 1);
    return two;
}

// ---- one step of the partition-first pipeline on one rank, in three phases (see smj_run_prepare in smj_api.cu for why)
// prepare: stage the rank's row blocks, size every buffer (the receive buffers were sized by dist_setup_recv_*)
int dist_step_prepare(DistRank &K, const smj_config_t *cfg, const smj_table_t *b1, const smj_table_t *b2)
{
    SmjCtx *c = K.c;
    CUDA_TRY(cudaSetDevice(c->device));
    K.blk[0] = *b1; K.blk[1] = *b2;
    K.launches0 = c->launches;
    const int64_t sel_val[2] = {cfg->select_val1, cfg->select_val2};
    CUDA_TRY(cudaEventRecord(K.ev[DE_START], c->stream));
    for (int t = 0; t < 2; t++) {
        const int32_t *d = nullptr;
        smj_table_t tt = K.blk[t];
        if (tt.on_device && tt.rows > 0 && ((uintptr_t)tt.data & 15)) {
            // a caller's device view that is not 16-byte aligned: a local copy, so that every rank takes the same path
            int32_t *p = (int32_t *)smj_ws(c, t ? WS_T2 : WS_T1, (size_t)tt.rows * tt.cols * 4);
            if (!p) return SMJ_ENOMEM;
            CUDA_TRY(cudaMemcpyAsync(p, tt.data, (size_t)tt.rows * tt.cols * 4, cudaMemcpyDeviceToDevice, c->stream));
            d = p;
        } else {
            SMJ_TRY(smj_stage_in(c, &tt, t ? WS_T2 : WS_T1, &d));
        }
        K.blk[t].data = const_cast<int32_t *>(d);
        K.blk[t].on_device = 1;
        const size_t cells = (size_t)tt.rows * tt.cols;
        K.slots[t] = (int32_t *)smj_ws(c, t ? WS_TMP_ROWS2 : WS_TMP_ROWS, cells * 4);
        K.pscr[t] = (char *)smj_ws(c, t ? WS_MERGE_B : WS_MERGE_A, smj_partition_scratch_bytes(tt.rows, tt.cols));
        if (!K.slots[t] || !K.pscr[t]) return SMJ_ENOMEM;
        K.none[t] = (sel_val[t] >= (int64_t)INT32_MAX) ? 1 : 0;
    }
    // the local pipeline on what will arrive: select disabled, row counts device-resident (loc->rows)
    smj_config_t local = *cfg;
    local.nr_gpus = 1;
    local.select_col1 = 0; local.select_col2 = 0;
    local.select_val1 = INT64_MIN; local.select_val2 = INT64_MIN;
    const smj_table_t r1 = {K.recv[0], K.cap_rows[0], K.blk[0].cols, 1}, r2 = {K.recv[1], K.cap_rows[1], K.blk[1].cols, 1};
    const u64 *d_rows[2] = {&K.loc->rows[0], &K.loc->rows[1]};
    SMJ_TRY(smj_run_prepare(c, &local, &r1, &r2, d_rows, &K.run));
    // one process, several ranks: no graph capture / instantiation (an allocation) between one rank's spinning kernels and the next rank's launches
    K.run.no_graph = g_dist.local;
    CUDA_TRY(cudaEventRecord(K.ev[DE_H2D], c->stream));   // inputs staged, every buffer sized: the device part of the step starts here
    return SMJ_OK;
}

// enqueue: splitters -> per table: select + partition, offsets, counts, exchange, arrival -> the local pipeline.
// Table 2's chain runs on a second stream and starts when table 1's partition kernel is done, so that its HBM-bound
// partition pass overlaps table 1's NVLink-bound exchange.
int dist_step_enqueue(DistRank &K, const smj_config_t *cfg, u64 seq)
{
    SmjCtx *c = K.c;
    CUDA_TRY(cudaSetDevice(c->device));
    const int G = g_dist.world, me = K.me;
    const int sel_col[2] = {cfg->select_col1, cfg->select_col2};
    const int64_t sel_val[2] = {cfg->select_val1, cfg->select_val2};
    const int key[2] = {cfg->join_key1, cfg->join_key2};
    const int S = G <= 4 ? DIST_S_MAX : DIST_S_MAX / 2;
    int n_pow2 = 128;
    while (n_pow2 < G * 2 * S) n_pow2 <<= 1;
    SampleJob sj[2];
    for (int t = 0; t < 2; t++) {
        int select_all = sel_val[t] < (int64_t)INT32_MIN;
        int64_t n = K.blk[t].rows;
        if (!select_all && sel_val[t] >= (int64_t)INT32_MAX) n = 0;
        sj[t] = {K.blk[t].data, n, K.blk[t].cols, sel_col[t], (int32_t)sel_val[t], select_all, key[t]};
    }
    // SMJ_DIST_POISON=1 (debugging): the receive buffers are filled with 0xff before the step, so a row that arrives late or
    // not at all cannot hide behind the identical row a previous step left at the same place.  Race-free: the fill is ordered
    // before this rank's sample flag, and no peer stores a row here before it has seen that flag.
    static const bool poison = getenv("SMJ_DIST_POISON") && atoi(getenv("SMJ_DIST_POISON")) != 0;
    if (poison)
        for (int t = 0; t < 2; t++)
            if (K.recv[t]) CUDA_TRY(cudaMemsetAsync(K.recv[t], 0xff, (size_t)K.cap_rows[t] * K.cap_cols[t] * 4, c->stream));
    smj_launch_on(c, c->stream, dist_splitters_kernel, 1, SPL2_THREADS, (size_t)n_pow2 * 4, K.peers, K.loc, me, G, seq, S, sj[0], sj[1], n_pow2, c->d_err);
    KERNEL_CHECK(c);
    CUDA_TRY(cudaEventRecord(K.ev[DE_SPLIT], c->stream));
    const bool two = dist_two_streams();
    // the table the local pipeline selects FIRST (the one with the smaller receive capacity, smj_launch_select_plan2's rule)
    // takes the main stream; the other one's chain runs on the second stream and starts when the first one's partition
    // kernel is done, so that its HBM-bound partition pass overlaps the first one's NVLink-bound exchange -- and its own
    // exchange overlaps the local pipeline's select of the first table, which is enqueued without waiting for it
    const int tfirst = K.cap_rows[0] <= K.cap_rows[1] ? 0 : 1;
    const int order[2] = {tfirst, tfirst ^ 1};
    for (int o = 0; o < 2; o++) {
        const int t = order[o];
        cudaStream_t st = (o == 1 && two) ? K.aux : c->stream;
        SMJ_TRY(smj_launch_select_partition(c, st, K.blk[t].data, K.blk[t].rows, K.blk[t].cols, sel_col[t], sel_val[t], key[t], K.loc->split, G,
                                            K.slots[t], K.pscr[t]));
        if (o == 0) {
            CUDA_TRY(cudaEventRecord(K.ev[DE_PART], c->stream));
            if (two) {   // the other table's chain starts here
                CUDA_TRY(cudaEventRecord(K.ev_fork, c->stream));
                CUDA_TRY(cudaStreamWaitEvent(K.aux, K.ev_fork, 0));
            }
        }
        const SmjPartScratch PS = smj_partition_scratch(K.pscr[t], K.none[t] ? 0 : K.blk[t].rows, K.blk[t].cols);
        smj_launch_on(c, st, dist_counts_kernel, 1, 64, 0, K.peers, K.loc, me, G, seq, t, (const u64 *)PS.bucket_total, (u64)K.cap_rows[t], c->d_err);
        KERNEL_CHECK(c);
        SmjPartitionDst D = {};
        for (int b = 0; b < G; b++) D.base[b] = K.peer_recv[t][b];
        D.row0 = K.loc->row0[t];
        D.skip = &K.loc->verdict[t];
        SMJ_TRY(smj_launch_partition_exchange(c, st, K.blk[t].rows, K.blk[t].cols, K.none[t], G, K.slots[t], K.pscr[t], D));
        // every rank's stores must have landed before anybody reads its receive buffer
        smj_launch_on(c, st, dist_arrive_kernel, 1, 32, 0, K.peers, K.loc, me, G, t, seq, c->d_err);
        KERNEL_CHECK(c);
        CUDA_TRY(cudaEventRecord(K.ev[o == 0 ? DE_XCHG : DE_XCHG2], st));
    }
    // the local pipeline: its select of the second table waits ON THE DEVICE for that table's arrival cell
    K.run.wait[order[1]].flag = &K.loc->arrived[order[1]];
    K.run.wait[order[1]].seq = &K.loc->seq_now;
    SMJ_TRY(smj_run_enqueue(c, &K.run));
    if (two) {
        CUDA_TRY(cudaEventRecord(K.ev_join, K.aux));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, K.ev_join, 0));
    }
    CUDA_TRY(cudaMemcpyAsync(K.h_loc, K.loc, sizeof(DistLocal), cudaMemcpyDeviceToHost, c->stream));
    return SMJ_OK;
}

// finish: the one host wait.  *retry = the verdict (the same on every rank): the step stored nothing, `need` says what
// every rank must be able to receive, the caller re-sizes collectively and runs the step again.
int dist_step_finish(DistRank &K, const smj_config_t *cfg, smj_table_t *out, smj_stats_t *stats, bool *retry, int64_t need[2])
{
    SmjCtx *c = K.c;
    const int G = g_dist.world, me = K.me;
    smj_stats_t ls;
    SMJ_TRY(smj_run_finish(c, &K.run, out, &ls));
    const DistLocal *h = K.h_loc;
    *retry = h->verdict[0] || h->verdict[1];
    int64_t sel[2] = {0, 0}, own[2] = {0, 0};
    double sent = 0;
    for (int t = 0; t < 2; t++) {
        int64_t worst = 0;
        for (int dst = 0; dst < G; dst++) {
            int64_t tot = 0;
            for (int src = 0; src < G; src++) tot += (int64_t)h->matrix[t][src][dst];
            if (tot > worst) worst = tot;
            sel[t] += (int64_t)h->matrix[t][me][dst];
            if (dst != me) sent += (double)h->matrix[t][me][dst] * K.blk[t].cols * 4;
        }
        for (int src = 0; src < G; src++) own[t] += (int64_t)h->matrix[t][src][me];
        need[t] = worst;
        if (worst > SMJ_MAX_SORT_ROWS) {
            if (out->data) smj_table_free(out);
            return smj_set_error(SMJ_ETOOBIG, "a rank would receive %lld rows of table %d (limit 2^30 - 1 per GPU)", (long long)worst, t + 1);
        }
    }
    if (*retry) {
        if (out->data) smj_table_free(out);
        return SMJ_OK;
    }
    static const bool trace = getenv("SMJ_DIST_TRACE") != nullptr;
    if (trace && me == 0)
        fprintf(stderr, "[dist] fabric step %llu; dev ms: samples+splitters %.3f | select/partition first table %.3f | first arrival +%.3f | second arrival +%.3f | end of join +%.3f\n",
                (unsigned long long)g_dist.seq, dist_ev_ms(K.ev[DE_H2D], K.ev[DE_SPLIT]), dist_ev_ms(K.ev[DE_SPLIT], K.ev[DE_PART]),
                dist_ev_ms(K.ev[DE_PART], K.ev[DE_XCHG]), dist_ev_ms(K.ev[DE_PART], K.ev[DE_XCHG2]), dist_ev_ms(K.ev[DE_PART], c->ev[4]));
    if (cfg->debug) {
        printf("==================\n#   exchange.cu  #\n==================\n");
        for (int t = 0; t < 2; t++)
            printf("Table %d - GPU %d selected %lld rows, owns %lld rows after the key-range exchange\n", t, me, (long long)sel[t], (long long)own[t]);
        printf("####################\n\n");
    }
    if (stats) {
        *stats = ls;
        stats->h2d_ms = dist_ev_ms(K.ev[DE_START], K.ev[DE_H2D]);
        stats->select_ms = dist_ev_ms(K.ev[DE_H2D], K.ev[DE_PART]);      // samples + splitters + select/partition of table 1
        // from the end of the first table's partition pass to the later of the two arrivals (the other table's partition pass
        // overlaps the first exchange; the local pipeline's first select overlaps the second exchange)
        stats->exchange_ms = dist_ev_ms(K.ev[DE_PART], K.ev[DE_XCHG]);
        { const double x2 = dist_ev_ms(K.ev[DE_PART], K.ev[DE_XCHG2]); if (x2 > stats->exchange_ms) stats->exchange_ms = x2; }
        // pairs of the received rows + the radix passes: from the end of the exchange to the start of the join stage
        stats->sort_ms = dist_ev_ms(K.ev[DE_PART], c->ev[4]) - stats->exchange_ms - ls.join_ms;
        if (stats->sort_ms < 0) stats->sort_ms = 0;
        stats->merge_ms = 0;
        stats->total_device_ms = dist_ev_ms(K.ev[DE_H2D], c->ev[4]);     // through the local pipeline's end-of-join event
        for (int t = 0; t < 2; t++) { stats->rows_in[t] = K.blk[t].rows; stats->rows_selected[t] = sel[t]; }
        stats->bytes_nvlink = sent;
        stats->kernel_launches = c->launches - K.launches0;
    }
    return SMJ_OK;
}

}  // namespace

// ------------------------------------------------------------------ init / shutdown
bool smj_dist_active(void) { return g_dist.active; }

int smj_dist_shutdown(void)
{
    if (g_dist.active && g_dist.comm && g_dist.rk[0].c) {
        // close this process's mappings of the peers' memory, then meet the peers, and only then free what they had mapped
        DistRank &K = g_dist.rk[0];
        cudaSetDevice(K.c->device);
        cudaDeviceSynchronize();
        dist_rank_unmap_recv(K);
        for (int r = 0; r < SMJ_MAX_G; r++) { if (K.ipc_win[r]) cudaIpcCloseMemHandle(K.ipc_win[r]); K.ipc_win[r] = nullptr; }
        int64_t dummy = 0;
        std::vector<int64_t> all((size_t)g_dist.world);
        boot_allgather(K.c, &dummy, all.data(), sizeof dummy);
        cudaGetLastError();
    }
    for (int r = 0; r < SMJ_MAX_G; r++) dist_rank_destroy(g_dist.rk[r]);
    if (g_dist.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_dist.comm);
    g_dist = DistState();
    return SMJ_OK;
}

extern "C" int smj_dist_unique_id(void *nccl_id_128)
{
    if (!nccl_id_128) return smj_set_error(SMJ_EINVAL, "smj_dist_unique_id: null buffer");
    SMJ_TRY(nccl_load());
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(nccl_id_128, &id, sizeof id);
    return SMJ_OK;
}

int smj_init_on_device(const smj_config_t *cfg, int device);   // smj_api.cu

extern "C" int smj_init_dist(const smj_config_t *cfg, int rank, int world, int local_device, const void *nccl_id_128)
{
    if (world < 1 || world > SMJ_MAX_G || rank < 0 || rank >= world || !nccl_id_128)
        return smj_set_error(SMJ_EINVAL, "smj_init_dist: bad rank/world/id (world <= %d)", SMJ_MAX_G);
    SMJ_TRY(nccl_load());
    SMJ_TRY(smj_init_on_device(cfg, local_device));   // (shuts a previous mode down first)
    ncclUniqueId id;
    memcpy(&id, nccl_id_128, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&g_dist.comm, world, id, rank));
    g_dist.rank = rank;
    g_dist.world = world;
    g_dist.active = true;
    g_dist.nlocal = 1;
    SMJ_TRY(dist_rank_create(g_dist.rk[0], g_ctx[0], rank));
    SMJ_TRY(dist_map_windows_ipc(g_dist.rk[0]));
    return SMJ_OK;
}

// One process, G GPUs (smj_init with nr_gpus = G, then smj_run): rank r = device r.
static int dist_init_local(int G)
{
    if (g_dist.local && g_dist.world == G) return SMJ_OK;
    if (g_dist.active) return smj_set_error(SMJ_EINVAL, "smj_init_dist mode is active: it drives one GPU per process");
    if (G > g_nctx || G > SMJ_MAX_G) return smj_set_error(SMJ_EINVAL, "nr_gpus=%d but %d contexts", G, g_nctx);
    smj_dist_shutdown();
    for (int r = 0; r < G; r++)
        for (int q = 0; q < G; q++) {
            if (r == q || g_ctx[r]->device == g_ctx[q]->device) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, g_ctx[r]->device, g_ctx[q]->device));
            if (!can) return smj_set_error(SMJ_EINVAL, "GPU %d cannot access GPU %d's memory: the key-range exchange needs peer access", r, q);
            CUDA_TRY(cudaSetDevice(g_ctx[r]->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(g_ctx[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return smj_cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
    g_dist.world = G;
    g_dist.nlocal = G;
    for (int r = 0; r < G; r++) SMJ_TRY(dist_rank_create(g_dist.rk[r], g_ctx[r], r));
    for (int r = 0; r < G; r++)
        for (int q = 0; q < G; q++) g_dist.rk[r].peers.win[q] = g_dist.rk[q].win;
    g_dist.local = true;
    return SMJ_OK;
}

// ------------------------------------------------------------------ the distributed pipeline (default path)
static int smj_run_multi_sorted(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats);
static int smj_run_multi_nccl(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats);

static int dist_check_knobs(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2)
{
    const smj_table_t *tb[2] = {t1, t2};
    const int sel_col[2] = {cfg->select_col1, cfg->select_col2};
    const int key[2] = {cfg->join_key1, cfg->join_key2};
    if (cfg->join_mode != SMJ_JOIN_ZIP) return smj_set_error(SMJ_EINVAL, "smj_run on several GPUs materialises SMJ_JOIN_ZIP only (the reference semantics)");
    for (int t = 0; t < 2; t++) {
        if (sel_col[t] < 0 || sel_col[t] >= tb[t]->cols) return smj_set_error(SMJ_EINVAL, "SELECT_COL%d=%d out of range", t + 1, sel_col[t]);
        if (key[t] < 0 || key[t] >= tb[t]->cols) return smj_set_error(SMJ_EINVAL, "JOIN_KEY%d=%d out of range", t + 1, key[t]);
    }
    return SMJ_OK;
}

// One process per GPU.  select + key-range partition of the ROWS (one streaming pass, smj_partition.cu) -> exchange fused into
// the compaction kernel (stores into the owners' receive buffers over NVLink) -> the single-GPU pipeline (sort + join, select
// disabled) on what arrived.  All synchronisation between the ranks is flag mailboxes in peer memory; the host waits once.
// Which path a step takes depends only on facts every rank shares (environment, column counts), never on one rank's data.
int smj_run_multi(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats)
{
    if (!g_dist.active) return smj_set_error(SMJ_EINVAL, "smj_run_multi: call smj_init_dist first (one process per GPU)");
    SMJ_TRY(dist_check_knobs(cfg, t1, t2));
    const char *mode = getenv("SMJ_DIST_MODE");
    if ((mode && strcmp(mode, "merge") == 0) || t1->cols > 32 || t2->cols > 32) return smj_run_multi_sorted(cfg, t1, t2, out, stats);
    const char *xmode = getenv("SMJ_DIST_EXCHANGE");
    if (xmode && strcmp(xmode, "nccl") == 0) return smj_run_multi_nccl(cfg, t1, t2, out, stats);
    DistRank &K = g_dist.rk[0];
    const int cols[2] = {t1->cols, t2->cols};
    const int64_t n_local[2] = {t1->rows, t2->rows};
    int64_t need[2] = {0, 0};
    for (int attempt = 0; attempt < 4; attempt++) {
        if (attempt > 0 || K.cap_rows[0] == 0 || K.cap_cols[0] != cols[0] || K.cap_cols[1] != cols[1])
            SMJ_TRY(dist_setup_recv_ipc(K, n_local, cols, need));
        const u64 seq = ++g_dist.seq;
        SMJ_TRY(dist_step_prepare(K, cfg, t1, t2));
        int rc = dist_step_enqueue(K, cfg, seq);
        if (rc != SMJ_OK) { smj_run_abandon(K.c, &K.run); cudaStreamSynchronize(K.c->stream); cudaStreamSynchronize(K.aux); return rc; }
        bool retry = false;
        SMJ_TRY(dist_step_finish(K, cfg, out, stats, &retry, need));
        if (!retry) return SMJ_OK;
    }
    return smj_set_error(SMJ_EINTERNAL, "the key-range exchange did not fit its receive buffers after three re-sizings");
}

// One process, G GPUs (the C driver: NR_GPUS / SMJ_NR_GPUS = G): the host tables are cut into G contiguous row blocks
// (rank order = row order, as app.c:155-218 deals rows to DPUs), every rank runs the step above on its own device and
// stream, and the shards are copied back in rank order = key order into ONE result table (app.c:739-753).
int smj_run_multi_local(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats)
{
    const int G = cfg->nr_gpus;
    SMJ_TRY(dist_check_knobs(cfg, t1, t2));
    if (t1->cols > 32 || t2->cols > 32) return smj_set_error(SMJ_EINVAL, "several GPUs in one process: tables of at most 32 columns");
    SMJ_TRY(dist_init_local(G));
    const smj_table_t *tb[2] = {t1, t2};
    const int cols[2] = {t1->cols, t2->cols};
    smj_table_t blk[SMJ_MAX_G][2];
    int64_t n_local[SMJ_MAX_G][2], need[SMJ_MAX_G][2] = {};
    for (int t = 0; t < 2; t++) {
        int owner = -1;   // a device table (the GPU CSV parser's output) lives on one GPU: the other ranks' blocks travel over NVLink
        if (tb[t]->on_device && tb[t]->rows > 0) {
            cudaPointerAttributes at;
            CUDA_TRY(cudaPointerGetAttributes(&at, tb[t]->data));
            owner = at.device;
        }
        for (int r = 0; r < G; r++) {
            const int64_t lo = tb[t]->rows * r / G, hi = tb[t]->rows * (r + 1) / G;
            int32_t *src = tb[t]->data ? tb[t]->data + (size_t)lo * cols[t] : nullptr;
            blk[r][t] = {src, hi - lo, cols[t], tb[t]->on_device};
            n_local[r][t] = hi - lo;
            SmjCtx *c = g_ctx[r];
            if (owner >= 0 && owner != c->device && hi > lo) {
                const size_t bytes = (size_t)(hi - lo) * cols[t] * 4;
                CUDA_TRY(cudaSetDevice(c->device));
                int32_t *p = (int32_t *)smj_ws(c, t ? WS_T2 : WS_T1, bytes);
                if (!p) return SMJ_ENOMEM;
                CUDA_TRY(cudaMemcpyPeerAsync(p, c->device, src, owner, bytes, c->stream));
                blk[r][t].data = p;
            }
        }
    }
    smj_table_t shard[SMJ_MAX_G];
    smj_stats_t st[SMJ_MAX_G];
    bool done = false;
    for (int attempt = 0; attempt < 4 && !done; attempt++) {
        DistRank &K0 = g_dist.rk[0];
        if (attempt > 0 || K0.cap_rows[0] == 0 || K0.cap_cols[0] != cols[0] || K0.cap_cols[1] != cols[1])
            SMJ_TRY(dist_setup_recv_local(n_local, cols, need));
        const u64 seq = ++g_dist.seq;
        int rc = SMJ_OK;
        int prepared = 0;
        for (int r = 0; r < G && rc == SMJ_OK; r++) { rc = dist_step_prepare(g_dist.rk[r], cfg, &blk[r][0], &blk[r][1]); if (rc == SMJ_OK) prepared = r + 1; }
        for (int r = 0; r < G && rc == SMJ_OK; r++) rc = dist_step_enqueue(g_dist.rk[r], cfg, seq);
        if (rc != SMJ_OK) {
            for (int r = 0; r < prepared; r++) {
                DistRank &K = g_dist.rk[r];
                cudaSetDevice(K.c->device);
                cudaStreamSynchronize(K.c->stream); cudaStreamSynchronize(K.aux);
                smj_run_abandon(K.c, &K.run);
            }
            cudaSetDevice(g_ctx[0]->device);
            return rc;
        }
        bool retry = false;
        for (int r = 0; r < G; r++) {
            shard[r] = {nullptr, 0, 0, 1};   // shards stay on their devices until the total is known
            bool rr = false;
            int frc = dist_step_finish(g_dist.rk[r], cfg, &shard[r], &st[r], &rr, need[r]);
            if (frc != SMJ_OK && rc == SMJ_OK) rc = frc;
            retry = retry || rr;
        }
        if (rc != SMJ_OK || retry) {
            for (int r = 0; r < G; r++) if (shard[r].data) smj_table_free(&shard[r]);
            if (rc != SMJ_OK) { cudaSetDevice(g_ctx[0]->device); return rc; }
            continue;
        }
        done = true;
    }
    if (!done) return smj_set_error(SMJ_EINTERNAL, "the key-range exchange did not fit its receive buffers after three re-sizings");
    // ---- GPU -> CPU: the shards in rank order = key order (app.c:739-753 writes the DPUs' results in order)
    int64_t total = 0;
    for (int r = 0; r < G; r++) total += shard[r].rows;
    const int c_out = cols[0] + cols[1] - 1;
    out->on_device = 0;
    SMJ_TRY(smj_alloc_out(g_ctx[0], out, total, c_out));
    int64_t at = 0;
    for (int r = 0; r < G; r++) {
        DistRank &K = g_dist.rk[r];
        CUDA_TRY(cudaSetDevice(K.c->device));
        CUDA_TRY(cudaEventRecord(K.ev[5], K.c->stream));
        if (shard[r].rows)
            CUDA_TRY(cudaMemcpyAsync(out->data + (size_t)at * c_out, shard[r].data, (size_t)shard[r].rows * c_out * 4, cudaMemcpyDeviceToHost, K.c->stream));
        CUDA_TRY(cudaEventRecord(K.ev[6], K.c->stream));
        at += shard[r].rows;
    }
    double d2h = 0;
    for (int r = 0; r < G; r++) {
        DistRank &K = g_dist.rk[r];
        CUDA_TRY(cudaSetDevice(K.c->device));
        CUDA_TRY(cudaStreamSynchronize(K.c->stream));
        const double ms = dist_ev_ms(K.ev[5], K.ev[6]);
        if (ms > d2h) d2h = ms;
        if (shard[r].data) smj_table_free(&shard[r]);
    }
    CUDA_TRY(cudaSetDevice(g_ctx[0]->device));
    if (stats) {   // times: the slowest rank; counts and bytes: summed
        *stats = st[0];
        for (int r = 1; r < G; r++) {
            const smj_stats_t &x = st[r];
#define MAXF(f) if (x.f > stats->f) stats->f = x.f
            MAXF(h2d_ms); MAXF(select_ms); MAXF(sort_ms); MAXF(exchange_ms); MAXF(join_ms); MAXF(total_device_ms); MAXF(sort_pass_ms_avg);
#undef MAXF
            for (int t = 0; t < 2; t++) { stats->rows_in[t] += x.rows_in[t]; stats->rows_selected[t] += x.rows_selected[t]; }
            stats->rows_joined += x.rows_joined;
            stats->bytes_model += x.bytes_model; stats->bytes_planned += x.bytes_planned; stats->bytes_nvlink += x.bytes_nvlink;
            stats->kernel_launches += x.kernel_launches;
        }
        stats->d2h_ms = d2h;
    }
    return SMJ_OK;
}

// ------------------------------------------------------------------ the same partitioning through NCCL (SMJ_DIST_EXCHANGE=nccl)
// The A/B partner of the fabric path: splitters from an ncclAllGather of the samples, the count matrix all-gathered and
// WAITED FOR on the host (exact receive buffers), a send buffer and one grouped ncclSend/ncclRecv all-to-all.
static int smj_run_multi_nccl(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats)
{
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int G = g_dist.world, me = g_dist.rank;
    const smj_table_t *tb[2] = {t1, t2};
    const int sel_col[2] = {cfg->select_col1, cfg->select_col2};
    const int64_t sel_val[2] = {cfg->select_val1, cfg->select_val2};
    const int key[2] = {cfg->join_key1, cfg->join_key2};
    const int64_t launches0 = c->launches;
    enum { E_START, E_H2D, E_PART, E_XCHG };
    cudaEvent_t *ev = c->ev + 8;   // smj_run_single below uses c->ev[0..5]
    CUDA_TRY(cudaEventRecord(ev[E_START], c->stream));
    const int32_t *d_t[2];
    const int cc[2] = {t1->cols, t2->cols};
    for (int t = 0; t < 2; t++) {
        if (tb[t]->on_device && tb[t]->rows > 0 && ((uintptr_t)tb[t]->data & 15)) {   // unaligned device view: local copy, same path on every rank
            int32_t *p = (int32_t *)smj_ws(c, t ? WS_T2 : WS_T1, (size_t)tb[t]->rows * cc[t] * 4);
            if (!p) return SMJ_ENOMEM;
            CUDA_TRY(cudaMemcpyAsync(p, tb[t]->data, (size_t)tb[t]->rows * cc[t] * 4, cudaMemcpyDeviceToDevice, c->stream));
            d_t[t] = p;
        } else SMJ_TRY(smj_stage_in(c, tb[t], t ? WS_T2 : WS_T1, &d_t[t]));
    }
    CUDA_TRY(cudaEventRecord(ev[E_H2D], c->stream));

    // ---- 1. splitters: regular row samples of both tables on every rank (predicate applied), all-gathered
    const int S = G <= 4 ? 1024 : 4096 / G;          // <= 8192 samples in all: one CTA sorts them in shared memory
    const int MSG = 2 * (G + 1);                       // per-rank count message: bucket starts of both tables
    u32 *d_samp = (u32 *)smj_ws(c, WS_SAMPLES, (size_t)(2 * S) * 4 * (G + 1) + 4096 + (size_t)MSG * 8 * (G + 1) + 1024);
    if (!d_samp) return SMJ_ENOMEM;
    u32 *d_samp_all = d_samp + 2 * S;
    u32 *d_split = d_samp_all + (size_t)G * 2 * S;
    u64 *d_msg = (u64 *)(d_split + 16);                // [MSG]
    u64 *d_msg_all = d_msg + MSG;                      // [G][MSG]
    for (int t = 0; t < 2; t++)
        SMJ_TRY(smj_launch_sample_rows(c, d_t[t], tb[t]->rows, cc[t], sel_col[t], sel_val[t], key[t], S, d_samp + t * S));
    NCCL_TRY(g_nccl.AllGather(d_samp, d_samp_all, (size_t)2 * S, ncclUint32, g_dist.comm, c->stream));
    SMJ_TRY(smj_launch_splitters(c, d_samp_all, G * 2 * S, G, d_split));   // same samples, same kernel, same splitters on every rank

    // ---- 2. select + partition of the rows by destination rank (rows grouped by bucket inside every tile's slot)
    int32_t *slots[2];
    char *pscr[2];
    int none[2];
    for (int t = 0; t < 2; t++) {
        const size_t cells = (size_t)tb[t]->rows * cc[t];
        slots[t] = (int32_t *)smj_ws(c, t ? WS_TMP_ROWS2 : WS_TMP_ROWS, cells * 4);
        pscr[t] = (char *)smj_ws(c, t ? WS_MERGE_B : WS_MERGE_A, smj_partition_scratch_bytes(tb[t]->rows, cc[t]));
        if (!slots[t] || !pscr[t]) return SMJ_ENOMEM;
        none[t] = (sel_val[t] >= (int64_t)INT32_MAX) ? 1 : 0;
        SMJ_TRY(smj_launch_select_partition(c, c->stream, d_t[t], tb[t]->rows, cc[t], sel_col[t], sel_val[t], key[t], d_split, G, slots[t], pscr[t]));
        const SmjPartScratch PS = smj_partition_scratch(pscr[t], none[t] ? 0 : tb[t]->rows, cc[t]);
        CUDA_TRY(cudaMemcpyAsync(d_msg + t * (G + 1), PS.bucket_start, (size_t)(G + 1) * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    CUDA_TRY(cudaEventRecord(ev[E_PART], c->stream));

    // ---- 3. the G x G row-count matrix, waited for on the host
    NCCL_TRY(g_nccl.AllGather(d_msg, d_msg_all, (size_t)MSG, ncclUint64, g_dist.comm, c->stream));
    uint64_t *h_msg = (uint64_t *)((char *)c->h_pinned + 4096);         // [G][MSG] <= 8 * 18 * 8 bytes
    CUDA_TRY(cudaMemcpyAsync(h_msg, d_msg_all, (size_t)G * MSG * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<int64_t> counts[2], recv_off[2];
    int64_t recv_total[2], m[2];
    int too_big = 0;
    for (int t = 0; t < 2; t++) {
        counts[t].assign((size_t)G * G, 0);
        recv_off[t].assign((size_t)G, 0);
        for (int src = 0; src < G; src++)
            for (int dst = 0; dst < G; dst++) {
                const uint64_t *b = h_msg + (size_t)src * MSG + (size_t)t * (G + 1);
                counts[t][(size_t)src * G + dst] = (int64_t)(b[dst + 1] - b[dst]);
            }
        SMJ_TRY(smj_plan_exchange(counts[t].data(), G, me, recv_off[t].data(), &recv_total[t]));
        m[t] = (int64_t)h_msg[(size_t)me * MSG + (size_t)t * (G + 1) + G];
        for (int r = 0; r < G; r++) {   // every rank checks every rank's share: the refusal is collective
            int64_t tot = 0;
            for (int src = 0; src < G; src++) tot += counts[t][(size_t)src * G + r];
            if (tot > SMJ_MAX_SORT_ROWS) too_big = t + 1;
        }
    }
    if (too_big) return smj_set_error(SMJ_ETOOBIG, "a rank would receive more than 2^30 - 1 rows of table %d", too_big);
    int32_t *recv[2], *send[2];
    for (int t = 0; t < 2; t++) {
        recv[t] = (int32_t *)smj_ws(c, t ? WS_XCHG_RECV2 : WS_XCHG_RECV1, (size_t)recv_total[t] * cc[t] * 4);
        send[t] = (int32_t *)smj_ws(c, t ? WS_XCHG_SEND2 : WS_XCHG_SEND1, (size_t)tb[t]->rows * cc[t] * 4);
        if (!recv[t] || !send[t]) return SMJ_ENOMEM;
    }

    // ---- 4. send buffer (buckets contiguous) + one grouped ncclSend/ncclRecv all-to-all
    double sent_bytes = 0;
    for (int t = 0; t < 2; t++) {
        const SmjPartScratch PS = smj_partition_scratch(pscr[t], none[t] ? 0 : tb[t]->rows, cc[t]);
        SmjPartitionDst D = {};
        for (int b = 0; b < G; b++) D.base[b] = send[t];
        D.row0 = PS.bucket_start;
        SMJ_TRY(smj_launch_partition_exchange(c, c->stream, tb[t]->rows, cc[t], none[t], G, slots[t], pscr[t], D));
    }
    NCCL_TRY(g_nccl.GroupStart());
    for (int t = 0; t < 2; t++) {
        const uint64_t *mine = h_msg + (size_t)me * MSG + (size_t)t * (G + 1);
        for (int peer = 0; peer < G; peer++) {
            const int64_t scount = counts[t][(size_t)me * G + peer] * cc[t];
            const int64_t rcount = counts[t][(size_t)peer * G + me] * cc[t];
            const int32_t *sbuf = send[t] + (size_t)mine[peer] * cc[t];
            int32_t *rbuf = recv[t] + (size_t)recv_off[t][peer] * cc[t];
            if (peer == me) {
                if (scount) CUDA_TRY(cudaMemcpyAsync(rbuf, sbuf, (size_t)scount * 4, cudaMemcpyDeviceToDevice, c->stream));
                continue;
            }
            if (scount) { NCCL_TRY(g_nccl.Send(sbuf, (size_t)scount, ncclInt32, peer, g_dist.comm, c->stream)); sent_bytes += (double)scount * 4; }
            if (rcount) NCCL_TRY(g_nccl.Recv(rbuf, (size_t)rcount, ncclInt32, peer, g_dist.comm, c->stream));
        }
    }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaEventRecord(ev[E_XCHG], c->stream));

    // ---- 5. this rank's key range: the single-GPU pipeline on the received rows, select disabled.  Runs arrived in
    // source-rank order with original order inside each, so the stable sort reproduces the reference's order.
    smj_config_t local = *cfg;
    local.nr_gpus = 1;
    local.select_col1 = 0; local.select_col2 = 0;
    local.select_val1 = INT64_MIN; local.select_val2 = INT64_MIN;
    const smj_table_t r1 = {recv[0], recv_total[0], cc[0], 1}, r2 = {recv[1], recv_total[1], cc[1], 1};
    smj_stats_t ls;
    SMJ_TRY(smj_run_single(c, &local, &r1, &r2, out, &ls));
    if (cfg->debug) {
        printf("==================\n#   exchange.cu  #\n==================\n");
        for (int t = 0; t < 2; t++)
            printf("Table %d - GPU %d selected %lld rows, owns %lld rows after the key-range exchange\n", t, me, (long long)m[t], (long long)recv_total[t]);
        printf("####################\n\n");
    }
    if (stats) {
        *stats = ls;
        stats->h2d_ms = dist_ev_ms(ev[E_START], ev[E_H2D]);
        stats->select_ms = dist_ev_ms(ev[E_H2D], ev[E_PART]);          // samples + splitters + select/partition of both tables
        stats->exchange_ms = dist_ev_ms(ev[E_PART], ev[E_XCHG]);       // count all-gather + compaction + all-to-all
        stats->sort_ms = ls.select_ms + ls.sort_ms;                     // pairs of the received rows + the radix passes
        stats->merge_ms = 0;
        stats->total_device_ms = dist_ev_ms(ev[E_H2D], c->ev[4]);      // through smj_run_single's end-of-join event
        for (int t = 0; t < 2; t++) { stats->rows_in[t] = tb[t]->rows; stats->rows_selected[t] = m[t]; }
        stats->bytes_nvlink = sent_bytes;
        stats->kernel_launches = c->launches - launches0;
    }
    return SMJ_OK;
}

// ------------------------------------------------------------------ sort-first pipeline (SMJ_DIST_MODE=merge)
static int smj_run_multi_sorted(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats)
{
    if (!g_dist.active) return smj_set_error(SMJ_EINVAL, "nr_gpus > 1 needs one process per GPU: call smj_init_dist first (see INTEGRATION.md)");
    if (cfg->join_mode != SMJ_JOIN_ZIP) return smj_set_error(SMJ_EINVAL, "smj_run materialises SMJ_JOIN_ZIP only (the reference semantics)");
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int G = g_dist.world, me = g_dist.rank;
    const smj_table_t *tb[2] = {t1, t2};
    const int sel_col[2] = {cfg->select_col1, cfg->select_col2};
    const int64_t sel_val[2] = {cfg->select_val1, cfg->select_val2};
    const int key[2] = {cfg->join_key1, cfg->join_key2};
    for (int t = 0; t < 2; t++) {
        if (sel_col[t] < 0 || sel_col[t] >= tb[t]->cols) return smj_set_error(SMJ_EINVAL, "SELECT_COL%d=%d out of range", t + 1, sel_col[t]);
        if (key[t] < 0 || key[t] >= tb[t]->cols) return smj_set_error(SMJ_EINVAL, "JOIN_KEY%d=%d out of range", t + 1, key[t]);
    }
    const int64_t launches0 = c->launches;
    c->pass_count = 0;
    enum { E_START, E_H2D, E_SORT, E_XCHG, E_MERGE, E_JOIN, E_D2H };
    CUDA_TRY(cudaEventRecord(c->ev[E_START], c->stream));
    const int32_t *d_t[2];
    SMJ_TRY(smj_stage_in(c, t1, WS_T1, &d_t[0]));
    SMJ_TRY(smj_stage_in(c, t2, WS_T2, &d_t[1]));
    const int cc[2] = {t1->cols, t2->cols};
    CUDA_TRY(cudaEventRecord(c->ev[E_H2D], c->stream));

    // ---- 1. local select + sort + payload gather of this rank's row blocks
    u64 *sorted[2];
    int64_t m[2];
    int32_t *rows_sorted[2];
    for (int t = 0; t < 2; t++) {
        SMJ_TRY(smj_sorted_pairs_of_table(c, d_t[t], tb[t]->rows, cc[t], sel_col[t], sel_val[t], 0, key[t], t, &sorted[t], &m[t]));
        int32_t *rs = (int32_t *)smj_ws(c, t ? WS_XCHG_SEND2 : WS_XCHG_SEND1, (size_t)m[t] * cc[t] * 4);
        if (!rs) return SMJ_ENOMEM;
        SMJ_TRY(smj_launch_gather_rows(c, sorted[t], m[t], d_t[t], cc[t], rs));
        rows_sorted[t] = rs;
    }
    CUDA_TRY(cudaEventRecord(c->ev[E_SORT], c->stream));

    // ---- 2. splitters from regular samples of both tables on every rank
    const int S = DIST_SAMPLES;
    u32 *d_samp = (u32 *)smj_ws(c, WS_SAMPLES, (size_t)(2 * S) * 4 * (G + 1) + (size_t)(G + 1) * 4 * 2 * (G + 1) + 4096);
    if (!d_samp) return SMJ_ENOMEM;
    u32 *d_samp_all = d_samp + 2 * S;                    // [G][2S]
    u32 *d_split = d_samp_all + (size_t)G * 2 * S;       // [G-1] (room for G)
    u32 *d_bnd = d_split + G;                            // [2][G+1]
    u32 *d_cnt_all = d_bnd + 2 * (G + 1);                // [G][2][G+1]
    for (int t = 0; t < 2; t++) {
        sample_keys_kernel<<<(S + 255) / 256, 256, 0, c->stream>>>(sorted[t], (u32)m[t], d_samp + t * S, S);
        KERNEL_CHECK(c);
    }
    NCCL_TRY(g_nccl.AllGather(d_samp, d_samp_all, (size_t)2 * S, ncclUint32, g_dist.comm, c->stream));
    std::vector<uint32_t> h_samp((size_t)G * 2 * S), h_split((size_t)std::max(G - 1, 1));
    CUDA_TRY(cudaMemcpyAsync(h_samp.data(), d_samp_all, h_samp.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    SMJ_TRY(smj_plan_splitters(h_samp.data(), (int64_t)h_samp.size(), G, h_split.data()));
    if (G > 1) CUDA_TRY(cudaMemcpyAsync(d_split, h_split.data(), (size_t)(G - 1) * 4, cudaMemcpyHostToDevice, c->stream));

    // ---- 3. bucket boundaries and the G x G row-count matrix
    for (int t = 0; t < 2; t++) {
        bucket_bounds_kernel<<<1, 32, 0, c->stream>>>(sorted[t], (u32)m[t], d_split, G, d_bnd + t * (G + 1));
        KERNEL_CHECK(c);
    }
    NCCL_TRY(g_nccl.AllGather(d_bnd, d_cnt_all, (size_t)2 * (G + 1), ncclUint32, g_dist.comm, c->stream));
    std::vector<uint32_t> h_bnd_all((size_t)G * 2 * (G + 1));
    CUDA_TRY(cudaMemcpyAsync(h_bnd_all.data(), d_cnt_all, h_bnd_all.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<int64_t> counts[2], recv_off[2];
    int64_t recv_total[2];
    for (int t = 0; t < 2; t++) {
        counts[t].assign((size_t)G * G, 0);
        recv_off[t].assign((size_t)G, 0);
        for (int src = 0; src < G; src++)
            for (int dst = 0; dst < G; dst++) {
                const uint32_t *b = &h_bnd_all[((size_t)src * 2 + t) * (G + 1)];
                counts[t][(size_t)src * G + dst] = (int64_t)b[dst + 1] - (int64_t)b[dst];
            }
        SMJ_TRY(smj_plan_exchange(counts[t].data(), G, me, recv_off[t].data(), &recv_total[t]));
        if (recv_total[t] > SMJ_MAX_SORT_ROWS)
            return smj_set_error(SMJ_ETOOBIG, "rank %d would receive %lld rows of table %d (limit 2^30 - 1 per GPU)", me, (long long)recv_total[t], t + 1);
    }
    const uint32_t *my_bnd[2] = {&h_bnd_all[((size_t)me * 2 + 0) * (G + 1)], &h_bnd_all[((size_t)me * 2 + 1) * (G + 1)]};

    // ---- 4. one grouped all-to-all of both tables' rows over NVLink
    int32_t *recv[2];
    for (int t = 0; t < 2; t++) {
        recv[t] = (int32_t *)smj_ws(c, t ? WS_XCHG_RECV2 : WS_XCHG_RECV1, (size_t)recv_total[t] * cc[t] * 4);
        if (!recv[t]) return SMJ_ENOMEM;
    }
    double sent_bytes = 0;
    NCCL_TRY(g_nccl.GroupStart());
    for (int t = 0; t < 2; t++)
        for (int peer = 0; peer < G; peer++) {
            const int64_t scount = counts[t][(size_t)me * G + peer] * cc[t];
            const int64_t rcount = counts[t][(size_t)peer * G + me] * cc[t];
            const int32_t *sbuf = rows_sorted[t] + (size_t)my_bnd[t][peer] * cc[t];
            int32_t *rbuf = recv[t] + (size_t)recv_off[t][peer] * cc[t];
            if (peer == me) {
                if (scount) CUDA_TRY(cudaMemcpyAsync(rbuf, sbuf, (size_t)scount * 4, cudaMemcpyDeviceToDevice, c->stream));
                continue;
            }
            if (scount) { NCCL_TRY(g_nccl.Send(sbuf, (size_t)scount, ncclInt32, peer, g_dist.comm, c->stream)); sent_bytes += (double)scount * 4; }
            if (rcount) NCCL_TRY(g_nccl.Recv(rbuf, (size_t)rcount, ncclInt32, peer, g_dist.comm, c->stream));
        }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaEventRecord(c->ev[E_XCHG], c->stream));

    // ---- 5. pairs of the received rows (row id = position in the receive buffer), then a merge tree over the G runs
    u64 *merged[2];
    for (int t = 0; t < 2; t++) {
        const int64_t n = recv_total[t];
        u64 *pa = (u64 *)smj_ws(c, t ? WS_PAIRS_A2 : WS_PAIRS_A1, (size_t)n * 8);
        u64 *pb = (u64 *)smj_ws(c, t ? WS_PAIRS_B2 : WS_PAIRS_B1, (size_t)n * 8);
        if (!pa || !pb) return SMJ_ENOMEM;
        const size_t sw = smj_select_num_tiles(n);
        char *scr = (char *)smj_ws(c, WS_SCRATCH, 1024 + sw * 8);
        if (!scr) return SMJ_ENOMEM;
        CUDA_TRY(cudaMemsetAsync(scr, 0, 1024 + sw * 8, c->stream));
        SMJ_TRY(smj_launch_select_pairs(c, recv[t], n, cc[t], key[t], 0, /*select_all=*/1, key[t], 0, pa, pb, (u64 *)(scr + 1024),
                                        (u32 *)(scr + 64), nullptr, (u64 *)scr));
        // runs: [recv_off[src], recv_off[src+1]) each sorted; merge neighbours until one run is left
        std::vector<int64_t> bnd(recv_off[t].begin(), recv_off[t].end());
        bnd.push_back(n);
        u64 *src = pa, *dst = pb;
        u32 *part = (u32 *)smj_ws(c, WS_PART, (smj_merge_num_tiles((u64)n) + 2) * 4);
        if (!part) return SMJ_ENOMEM;
        while (bnd.size() > 2) {
            std::vector<int64_t> nb;
            nb.push_back(0);
            const size_t runs = bnd.size() - 1;
            for (size_t r = 0; r < runs; r += 2) {
                const int64_t lo = bnd[r], mid = bnd[r + 1];
                if (r + 1 < runs) {
                    const int64_t hi = bnd[r + 2];
                    SMJ_TRY(smj_launch_merge_pairs(c, src + lo, (u32)(mid - lo), src + mid, (u32)(hi - mid), dst + lo, part));
                    nb.push_back(hi);
                } else {   // odd run out: carried to the next round unchanged (app.c:505-520 does the same with its odd chunk)
                    if (mid > lo) CUDA_TRY(cudaMemcpyAsync(dst + lo, src + lo, (size_t)(mid - lo) * 8, cudaMemcpyDeviceToDevice, c->stream));
                    nb.push_back(mid);
                }
            }
            bnd.swap(nb);
            std::swap(src, dst);
        }
        merged[t] = src;
    }
    CUDA_TRY(cudaEventRecord(c->ev[E_MERGE], c->stream));

    // ---- 6. local join of this rank's key range, payload straight from the receive buffers
    int64_t j = 0;
    smj_table_t dev_out = {nullptr, 0, cc[0] + cc[1] - 1, 1};
    SMJ_TRY(smj_join_pairs_to_table(c, merged[0], (u32)recv_total[0], merged[1], (u32)recv_total[1], recv[0], cc[0], recv[1], cc[1], key[1],
                                    &dev_out, &j));
    CUDA_TRY(cudaEventRecord(c->ev[E_JOIN], c->stream));
    const int c_out = cc[0] + cc[1] - 1;
    if (out->on_device) {
        *out = dev_out;
    } else {
        SMJ_TRY(smj_alloc_out(c, out, j, c_out));
        if (j) CUDA_TRY(cudaMemcpyAsync(out->data, dev_out.data, (size_t)j * c_out * 4, cudaMemcpyDeviceToHost, c->stream));
        smj_table_free(&dev_out);
    }
    CUDA_TRY(cudaEventRecord(c->ev[E_D2H], c->stream));
    SMJ_TRY(smj_check_device_flag(c));
    if (cfg->debug) {
        printf("==================\n#   exchange.cu  #\n==================\n");
        for (int t = 0; t < 2; t++)
            printf("Table %d - GPU %d selected %lld rows, owns %lld rows after the key-range exchange\n", t, me, (long long)m[t], (long long)recv_total[t]);
        printf("GPU %d results: %lld rows\n####################\n\n", me, (long long)j);
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->h2d_ms = dist_ev_ms(c->ev[E_START], c->ev[E_H2D]);
        stats->sort_ms = dist_ev_ms(c->ev[E_H2D], c->ev[E_SORT]);        // select + sort + gather of the local blocks
        stats->exchange_ms = dist_ev_ms(c->ev[E_SORT], c->ev[E_XCHG]);
        stats->merge_ms = dist_ev_ms(c->ev[E_XCHG], c->ev[E_MERGE]);
        stats->join_ms = dist_ev_ms(c->ev[E_MERGE], c->ev[E_JOIN]);
        stats->d2h_ms = dist_ev_ms(c->ev[E_JOIN], c->ev[E_D2H]);
        stats->total_device_ms = dist_ev_ms(c->ev[E_H2D], c->ev[E_JOIN]);
        for (int t = 0; t < 2; t++) { stats->rows_in[t] = tb[t]->rows; stats->rows_selected[t] = m[t]; }
        stats->rows_joined = j;
        stats->bytes_nvlink = sent_bytes;
        stats->kernel_launches = c->launches - launches0;
        double sum = 0;
        for (int p = 0; p < c->pass_count; p++) sum += dist_ev_ms(c->pass_ev[2 * p], c->pass_ev[2 * p + 1]);
        stats->sort_passes = c->pass_count * SMJ_KEY_PASSES;   // each timed group is one table's four passes
        stats->sort_pass_ms_avg = c->pass_count ? sum / (c->pass_count * SMJ_KEY_PASSES) : 0;
    }
    return SMJ_OK;
}
