// smj_dist.cu -- key-range partitioned sort-merge-join across the GPUs of one box, one process per GPU.
//
// Replaces the reference's host-mediated data movement between stages -- the per-DPU dpu_push_xfer gathers and
// scatters (sort-merge-join/app.c:222-288), the log-depth merge tournament that round-trips whole tables through
// host memory (app.c:413-547) and the host-side key-range split for the join (app.c:585-633) -- with ONE exchange.
//
// Default path, smj_run_multi (partition first):
//   1. splitters: every rank contributes regular row samples of both tables with the predicate applied
//      (ncclAllGather); every rank sorts the same gathered samples on the device and so derives the same G-1 key
//      splitters (splitters_kernel, the device twin of smj_plan_splitters) -- no host round trip;
//   2. select fused with key-range partitioning of the ROWS (smj_partition.cu): survivors grouped by destination rank
//      inside their tile's slot, original order kept inside each bucket;
//   3. the per-bucket totals (and every rank's receive capacity) are all-gathered: the G x G row-count matrix, the one
//      host wait of this path (buffer sizing, smj_plan_exchange);
//   4. exchange fused into the compaction kernel: each (tile, bucket) segment is stored straight into the destination
//      rank's receive buffer through a CUDA-IPC mapping (NVLink stores from the SMs); SMJ_DIST_EXCHANGE=nccl uses a
//      send buffer and one grouped ncclSend/ncclRecv all-to-all instead;
//   5. the single-GPU pipeline (smj_run_single, select disabled) sorts and joins what arrived: runs sit in source-rank
//      order with original order inside each, so the stable sort reproduces the reference's order;
//   6. the result shards, concatenated in rank order, are the single-GPU result (splitters are key values, so all
//      rows of one key meet on one rank and the zip pairing of equal keys is local).
//
// SMJ_DIST_MODE=merge, smj_run_multi_sorted (sort first; also the fallback for tables the partition kernel cannot
// take): local select + sort + payload gather, splitters from samples of the sorted keys (host), bucket bounds by
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

struct DistState {
    bool active = false;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
} g_dist;

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

// ------------------------------------------------------------------ init / shutdown
bool smj_dist_active(void) { return g_dist.active; }

static void peer_unmap_all();

int smj_dist_shutdown(void)
{
    peer_unmap_all();
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
    if (world < 1 || rank < 0 || rank >= world || !nccl_id_128) return smj_set_error(SMJ_EINVAL, "smj_init_dist: bad rank/world/id");
    SMJ_TRY(nccl_load());
    SMJ_TRY(smj_init_on_device(cfg, local_device));
    smj_dist_shutdown();
    ncclUniqueId id;
    memcpy(&id, nccl_id_128, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&g_dist.comm, world, id, rank));
    g_dist.rank = rank;
    g_dist.world = world;
    g_dist.active = true;
    return SMJ_OK;
}

// ------------------------------------------------------------------ the distributed pipeline
extern SmjCtx *g_ctx[8];
int smj_stage_in(SmjCtx *c, const smj_table_t *t, int slot, const int32_t **d);
int smj_alloc_out(SmjCtx *c, smj_table_t *out, int64_t rows, int cols);
int smj_check_device_flag(SmjCtx *c);
int smj_join_pairs_to_table(SmjCtx *c, const u64 *pl, u32 m1, const u64 *pr, u32 m2, const int32_t *d_t1, int c1, const int32_t *d_t2,
                            int c2, int key2, smj_table_t *out, int64_t *rows_out);
int smj_sorted_pairs_of_table(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int select_all,
                              int key_col, int table_idx, u64 **d_sorted, int64_t *m_out);

static float dist_ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ms;
}

// ------------------------------------------------------------------ partition-first pipeline (default)
bool smj_partition_supported(const int32_t *d_in, int cols);
size_t smj_partition_scratch_bytes(int64_t n, int cols);
int smj_launch_sample_rows(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int key_col, int S,
                           u32 *d_samples);
int smj_launch_select_partition(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int key_col,
                                const u32 *d_splitters, int G, int32_t *d_slots, char *d_scratch, u64 **d_bucket_start);
int smj_launch_partition_compact(SmjCtx *c, int64_t n, int cols, int sel_val_none, int G, const int32_t *d_slots, char *d_scratch,
                                 int32_t *d_send, int32_t *const *d_dst_by_bucket);
int smj_launch_splitters(SmjCtx *c, const u32 *d_samples, int n_samples, int G, u32 *d_splitters);
int smj_run_single(SmjCtx *c, const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats);
static int smj_run_multi_sorted(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats);

// Peer receive buffers mapped through CUDA IPC (one process per GPU on one box): g_peer[t][r] is rank r's receive buffer
// of table t as seen from this process.  Re-mapped whenever the owner reallocated (its handle changed).
struct PeerMap { void *base = nullptr; cudaIpcMemHandle_t handle; bool have = false; };
static PeerMap g_peer[2][8];

static void peer_unmap_all()
{
    for (int t = 0; t < 2; t++)
        for (int r = 0; r < 8; r++) {
            if (g_peer[t][r].base) cudaIpcCloseMemHandle(g_peer[t][r].base);
            g_peer[t][r] = PeerMap();
        }
    cudaGetLastError();
}

// select + key-range partition of the ROWS (one streaming pass, smj_partition.cu) -> exchange -> the single-GPU pipeline
// (sort + join, select disabled) on what arrived.  Compared with sorting first and merging the received runs
// (smj_run_multi_sorted below, SMJ_DIST_MODE=merge) no row is gathered at random before it travels and there are no merge
// rounds (0.51 ms per step at 8 GPUs).  The exchange itself is FUSED into the compaction kernel: each (tile, bucket)
// segment is stored straight into the destination rank's receive buffer through a CUDA-IPC mapping, i.e. over NVLink
// from the SMs (SMJ_DIST_EXCHANGE=nccl falls back to a send buffer + grouped ncclSend/ncclRecv, which reached 240 GB/s
// per GPU here against ~770 GB/s for peer copies).
int smj_run_multi(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats)
{
    if (!g_dist.active) return smj_set_error(SMJ_EINVAL, "nr_gpus > 1 needs one process per GPU: call smj_init_dist first (see INTEGRATION.md)");
    if (cfg->join_mode != SMJ_JOIN_ZIP) return smj_set_error(SMJ_EINVAL, "smj_run materialises SMJ_JOIN_ZIP only (the reference semantics)");
    const char *mode = getenv("SMJ_DIST_MODE");
    if (mode && strcmp(mode, "merge") == 0) return smj_run_multi_sorted(cfg, t1, t2, out, stats);
    const char *xmode = getenv("SMJ_DIST_EXCHANGE");
    const bool use_peer = !(xmode && strcmp(xmode, "nccl") == 0);
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
    enum { E_START, E_H2D, E_PART, E_XCHG, E_SAMP, E_SPLIT, E_CNT };
    cudaEvent_t *ev = c->ev + 8;   // smj_run_single below uses c->ev[0..5]
    static const bool trace = getenv("SMJ_DIST_TRACE") != nullptr;
    CUDA_TRY(cudaEventRecord(ev[E_START], c->stream));
    const int32_t *d_t[2];
    SMJ_TRY(smj_stage_in(c, t1, WS_T1, &d_t[0]));
    SMJ_TRY(smj_stage_in(c, t2, WS_T2, &d_t[1]));
    const int cc[2] = {t1->cols, t2->cols};
    if ((tb[0]->rows && !smj_partition_supported(d_t[0], cc[0])) || (tb[1]->rows && !smj_partition_supported(d_t[1], cc[1])))
        return smj_run_multi_sorted(cfg, t1, t2, out, stats);   // > 32 columns or a table that is not 16-byte aligned
    CUDA_TRY(cudaEventRecord(ev[E_H2D], c->stream));

    // ---- 1. splitters: regular row samples of both tables on every rank (predicate applied), all-gathered
    const int S = G <= 4 ? 1024 : 4096 / G;          // <= 8192 samples in all: one CTA sorts them in shared memory
    const int MSG = 2 * (G + 1) + 2;                   // per-rank count message: bucket starts of both tables + receive capacities
    u32 *d_samp = (u32 *)smj_ws(c, WS_SAMPLES, (size_t)(2 * S) * 4 * (G + 1) + 4096 + (size_t)MSG * 8 * (G + 1) + 1024);
    if (!d_samp) return SMJ_ENOMEM;
    u32 *d_samp_all = d_samp + 2 * S;
    u32 *d_split = d_samp_all + (size_t)G * 2 * S;
    u64 *d_msg = (u64 *)(d_split + 16);                // [MSG]
    u64 *d_msg_all = d_msg + MSG;                      // [G][MSG]
    int32_t **d_dst = (int32_t **)(d_msg_all + (size_t)G * MSG);   // [2][8] destination pointers per bucket
    unsigned char *d_hnd = (unsigned char *)(d_dst + 16);          // [G+1][2][64] IPC handles
    for (int t = 0; t < 2; t++)
        SMJ_TRY(smj_launch_sample_rows(c, d_t[t], tb[t]->rows, cc[t], sel_col[t], sel_val[t], key[t], S, d_samp + t * S));
    NCCL_TRY(g_nccl.AllGather(d_samp, d_samp_all, (size_t)2 * S, ncclUint32, g_dist.comm, c->stream));
    CUDA_TRY(cudaEventRecord(ev[E_SAMP], c->stream));
    char *hp = (char *)c->h_pinned;                    // small pinned mailbox for the host legs below
    SMJ_TRY(smj_launch_splitters(c, d_samp_all, G * 2 * S, G, d_split));   // same samples, same kernel, same splitters on every rank
    CUDA_TRY(cudaEventRecord(ev[E_SPLIT], c->stream));

    // ---- 2. select + partition of the rows by destination rank (rows grouped by bucket inside every tile's slot)
    int32_t *slots[2];
    char *pscr[2];
    u64 *d_bs[2];
    int none[2];
    for (int t = 0; t < 2; t++) {
        const size_t cells = (size_t)tb[t]->rows * cc[t];
        slots[t] = (int32_t *)smj_ws(c, t ? WS_TMP_ROWS2 : WS_TMP_ROWS, cells * 4);
        pscr[t] = (char *)smj_ws(c, t ? WS_MERGE_B : WS_MERGE_A, smj_partition_scratch_bytes(tb[t]->rows, cc[t]));
        if (!slots[t] || !pscr[t]) return SMJ_ENOMEM;
        none[t] = (sel_val[t] >= (int64_t)INT32_MAX) ? 1 : 0;
        SMJ_TRY(smj_launch_select_partition(c, d_t[t], tb[t]->rows, cc[t], sel_col[t], sel_val[t], key[t], d_split, G, slots[t], pscr[t], &d_bs[t]));
        CUDA_TRY(cudaMemcpyAsync(d_msg + t * (G + 1), d_bs[t], (size_t)(G + 1) * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    // receive capacities (rows) ride along so that every rank knows who must grow its buffer this step
    uint64_t *hp_cap = (uint64_t *)(hp + 2048);
    for (int t = 0; t < 2; t++) hp_cap[t] = (uint64_t)(c->slot_bytes[t ? WS_XCHG_RECV2 : WS_XCHG_RECV1] / ((size_t)cc[t] * 4));
    CUDA_TRY(cudaMemcpyAsync(d_msg + 2 * (G + 1), hp_cap, 16, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(ev[E_PART], c->stream));

    // ---- 3. the G x G row-count matrix
    NCCL_TRY(g_nccl.AllGather(d_msg, d_msg_all, (size_t)MSG, ncclUint64, g_dist.comm, c->stream));
    CUDA_TRY(cudaEventRecord(ev[E_CNT], c->stream));
    uint64_t *h_msg = (uint64_t *)(hp + 4096);         // [G][MSG] <= 8 * 20 * 8 bytes
    CUDA_TRY(cudaMemcpyAsync(h_msg, d_msg_all, (size_t)G * MSG * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<int64_t> counts[2], recv_off[2];
    int64_t recv_total[2], m[2];
    bool someone_grows = false;
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
        if (recv_total[t] > SMJ_MAX_SORT_ROWS)
            return smj_set_error(SMJ_ETOOBIG, "rank %d would receive %lld rows of table %d (limit 2^30 - 1 per GPU)", me, (long long)recv_total[t], t + 1);
        for (int r = 0; r < G; r++) {                  // the same verdict on every rank: does rank r have to grow table t?
            int64_t tot = 0;
            for (int src = 0; src < G; src++) tot += counts[t][(size_t)src * G + r];
            const uint64_t cap = h_msg[(size_t)r * MSG + 2 * (G + 1) + t];
            if ((uint64_t)tot > cap || !g_peer[t][r].have) someone_grows = true;
        }
    }
    int32_t *recv[2];
    for (int t = 0; t < 2; t++) {
        // 25 % head room so that a slightly different split next step does not force a re-map on every rank
        const size_t want = (size_t)recv_total[t] * cc[t] * 4;
        const int slot = t ? WS_XCHG_RECV2 : WS_XCHG_RECV1;
        recv[t] = (int32_t *)smj_ws(c, slot, c->slot_bytes[slot] >= want ? want : want + want / 4);
        if (!recv[t]) return SMJ_ENOMEM;
    }

    double sent_bytes = 0;
    bool peer_ok = use_peer;
    if (use_peer && someone_grows) {
        // every rank publishes the IPC handles of its two receive buffers; peers map what changed
        unsigned char *hp_h = (unsigned char *)(hp + 8192);        // [2][64] mine, then [G][2][64]
        for (int t = 0; t < 2; t++) {
            cudaIpcMemHandle_t hnd;
            if (cudaIpcGetMemHandle(&hnd, recv[t]) != cudaSuccess) { cudaGetLastError(); memset(&hnd, 0, sizeof hnd); }
            memcpy(hp_h + t * 64, &hnd, 64);
        }
        CUDA_TRY(cudaMemcpyAsync(d_hnd, hp_h, 128, cudaMemcpyHostToDevice, c->stream));
        NCCL_TRY(g_nccl.AllGather(d_hnd, d_hnd + 128, 128, /*ncclUint8*/ 1, g_dist.comm, c->stream));
        CUDA_TRY(cudaMemcpyAsync(hp_h + 128, d_hnd + 128, (size_t)G * 128, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < G; r++)
            for (int t = 0; t < 2; t++) {
                PeerMap &pm = g_peer[t][r];
                const unsigned char *hb = hp_h + 128 + (size_t)r * 128 + t * 64;
                if (r == me) { pm.have = true; pm.base = nullptr; continue; }
                if (pm.have && memcmp(&pm.handle, hb, 64) == 0) continue;
                if (pm.base) { cudaIpcCloseMemHandle(pm.base); pm.base = nullptr; }
                memcpy(&pm.handle, hb, 64);
                cudaError_t e = cudaIpcOpenMemHandle(&pm.base, pm.handle, cudaIpcMemLazyEnablePeerAccess);
                pm.have = (e == cudaSuccess);
                if (e != cudaSuccess) { cudaGetLastError(); pm.base = nullptr; peer_ok = false; }
            }
        // all ranks must agree on the path: one failed mapping anywhere sends everybody to NCCL for this step
        uint64_t *hp_ok = (uint64_t *)(hp + 3072);
        *hp_ok = peer_ok ? 1 : 0;
        CUDA_TRY(cudaMemcpyAsync(d_msg, hp_ok, 8, cudaMemcpyHostToDevice, c->stream));
        NCCL_TRY(g_nccl.AllGather(d_msg, d_msg_all, 1, ncclUint64, g_dist.comm, c->stream));
        CUDA_TRY(cudaMemcpyAsync(h_msg, d_msg_all, (size_t)G * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < G; r++) if (!h_msg[r]) peer_ok = false;
        if (!peer_ok) for (int t = 0; t < 2; t++) for (int r = 0; r < G; r++) g_peer[t][r].have = false;   // try again next step
    }

    if (peer_ok) {
        // ---- 4a. fused compaction + exchange: segments are stored straight into the owners' receive buffers (NVLink)
        int32_t **hp_dst = (int32_t **)(hp + 3200);
        for (int t = 0; t < 2; t++)
            for (int b = 0; b < 8; b++) {
                int32_t *p = nullptr;
                if (b < G) {
                    int64_t before = 0;                              // rows the lower ranks put in front of mine at rank b
                    for (int src = 0; src < me; src++) before += counts[t][(size_t)src * G + b];
                    int32_t *base = (b == me) ? recv[t] : (int32_t *)g_peer[t][b].base;
                    p = base + (size_t)before * cc[t];
                    if (b != me) sent_bytes += (double)counts[t][(size_t)me * G + b] * cc[t] * 4;
                }
                hp_dst[t * 8 + b] = p;
            }
        CUDA_TRY(cudaMemcpyAsync(d_dst, hp_dst, 16 * sizeof(int32_t *), cudaMemcpyHostToDevice, c->stream));
        for (int t = 0; t < 2; t++)
            SMJ_TRY(smj_launch_partition_compact(c, tb[t]->rows, cc[t], none[t], G, slots[t], pscr[t], nullptr, d_dst + t * 8));
        // every rank's stores must have landed before anybody sorts: a collective after the kernels is that barrier
        NCCL_TRY(g_nccl.AllGather(d_msg, d_msg_all, 1, ncclUint64, g_dist.comm, c->stream));
    } else {
        // ---- 4b. send buffer + one grouped ncclSend/ncclRecv all-to-all
        int32_t *send[2];
        for (int t = 0; t < 2; t++) {
            send[t] = (int32_t *)smj_ws(c, t ? WS_XCHG_SEND2 : WS_XCHG_SEND1, (size_t)tb[t]->rows * cc[t] * 4);
            if (!send[t]) return SMJ_ENOMEM;
            SMJ_TRY(smj_launch_partition_compact(c, tb[t]->rows, cc[t], none[t], G, slots[t], pscr[t], send[t], nullptr));
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
    }
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
    if (trace && me == 0)
        fprintf(stderr, "[dist] %s exchange; dev ms: samples+allgather %.3f | splitters %.3f | select/partition %.3f | counts %.3f | compaction+exchange %.3f | local %.3f\n",
                peer_ok ? "peer-store" : "nccl", dist_ev_ms(ev[E_H2D], ev[E_SAMP]), dist_ev_ms(ev[E_SAMP], ev[E_SPLIT]),
                dist_ev_ms(ev[E_SPLIT], ev[E_PART]), dist_ev_ms(ev[E_PART], ev[E_CNT]), dist_ev_ms(ev[E_CNT], ev[E_XCHG]), ls.total_device_ms);
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
        stats->exchange_ms = dist_ev_ms(ev[E_PART], ev[E_XCHG]);       // count all-gather + compaction/exchange
        stats->sort_ms = ls.select_ms + ls.sort_ms;                     // pairs of the received rows + the four passes
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
