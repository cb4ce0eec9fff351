// smj_api.cu -- the C-ABI of libsmj.so (include/smj.h): contexts, workspace arena, error reporting and the
// single-GPU stage entry points smj_select / smj_sort / smj_merge / smj_join / smj_run.
//
// Stage order and semantics follow sort-merge-join/app.c:main (select :221-307, sort :315-373,
// merge :413-547, join :585-688) and cpu_app.c:336-344.  There is no CPU fallback: every entry point
// needs a CUDA device and returns SMJ_ENODEVICE without one.
#include "smj_internal.h"
#include "../../include/user.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

// ------------------------------------------------------------------ errors
static thread_local char g_errbuf[512] = "";

int smj_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_errbuf, sizeof g_errbuf, fmt, ap);
    va_end(ap);
    return code;
}

// SMJ_STAGE_EVENTS=0 drops the event records between the stages (and around the radix passes) from the pipeline, so
// that the whole chain of kernels is linked by programmatic dependencies; the per-stage times of smj_stats_t are then 0.
bool smj_stage_events(void)
{
    static const bool on = !(getenv("SMJ_STAGE_EVENTS") && atoi(getenv("SMJ_STAGE_EVENTS")) == 0);
    return on;
}

bool smj_pdl_enabled(void)
{
    static const bool on = !(getenv("SMJ_PDL") && atoi(getenv("SMJ_PDL")) == 0);
    return on;
}

int smj_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    const char *base = strrchr(file, '/');
    snprintf(g_errbuf, sizeof g_errbuf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e),
             base ? base + 1 : file, line, what);
    if (e == cudaErrorMemoryAllocation) return SMJ_ENOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return SMJ_ENODEVICE;
    return SMJ_ECUDA;
}

extern "C" const char *smj_strerror(int code)
{
    switch (code) {
    case SMJ_OK: return "success";
    case SMJ_EINVAL: return "invalid argument";
    case SMJ_ENODEVICE: return "no CUDA device / library not initialised";
    case SMJ_ECUDA: return "CUDA runtime error";
    case SMJ_ENOMEM: return "out of memory";
    case SMJ_ETOOBIG: return "input exceeds the supported row count";
    case SMJ_ENCCL: return "NCCL error";
    case SMJ_EINTERNAL: return "device-side consistency check failed";
    case SMJ_EIRREGULAR: return "CSV text needs the sequential host parser";
    case SMJ_ERANGE: return "cell value outside the int32 range of the engine's tables";
    default: return "unknown error";
    }
}
extern "C" const char *smj_last_error(void) { return g_errbuf; }

// ------------------------------------------------------------------ global state
SmjCtx *g_ctx[8] = {};
int g_nctx = 0;
smj_config_t g_cfg;
static bool g_inited = false;

extern "C" void smj_config_default(smj_config_t *cfg)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->nr_gpus = NR_GPUS;
    cfg->select_col1 = SELECT_COL1; cfg->select_val1 = SELECT_VAL1;
    cfg->select_col2 = SELECT_COL2; cfg->select_val2 = SELECT_VAL2;
    cfg->join_key1 = JOIN_KEY1; cfg->join_key2 = JOIN_KEY2;
    cfg->join_mode = SMJ_JOIN_ZIP;
#ifdef DEBUG
    cfg->debug = 1;
#endif
}

static int ctx_create(int device, SmjCtx **out)
{
    CUDA_TRY(cudaSetDevice(device));
    SmjCtx *c = new SmjCtx();
    c->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc(&c->d_err, 256));
    CUDA_TRY(cudaMemset(c->d_err, 0, 256));
    CUDA_TRY(cudaMalloc((void **)&c->d_out_ptr, 256));
    c->h_pinned_bytes = 1 << 16;
    CUDA_TRY(cudaMallocHost(&c->h_pinned, c->h_pinned_bytes));
    for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    for (auto &e : c->pass_ev) CUDA_TRY(cudaEventCreate(&e));
    // The payload gathers read one 16..32-byte row per random address: ask L2 to fetch single 32-byte sectors
    // instead of 64/128-byte lines (a hint; ncu showed 3.4x DRAM read amplification on join_materialize without it).
    // It is a device-wide limit: the host application's value is put back at shutdown (SMJ_L2_FETCH=0 leaves it alone).
    static const bool l2_fetch = !(getenv("SMJ_L2_FETCH") && atoi(getenv("SMJ_L2_FETCH")) == 0);
    if (l2_fetch && cudaDeviceGetLimit(&c->l2_fetch_saved, cudaLimitMaxL2FetchGranularity) == cudaSuccess) {
        c->l2_fetch_set = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32) == cudaSuccess;
    }
    cudaGetLastError();
    // output buffers come from a PRIVATE stream-ordered pool that keeps freed memory cached (the device's default pool,
    // which the host application may share, is left as it is)
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CUDA_TRY(cudaMemPoolCreate(&c->pool, &props));
    uint64_t thr = UINT64_MAX;
    CUDA_TRY(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    {   // the pool's first allocation sets up its backing memory (17 ms on a B200 box, measured inside the first smj_csv_parse of a
        // one-shot run): take it here, with the other set-up costs, not inside the first call that returns a device table
        void *warm = nullptr;
        if (cudaMallocFromPoolAsync(&warm, (size_t)1 << 20, c->pool, c->stream) == cudaSuccess) cudaFreeAsync(warm, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaGetLastError();
    }
    *out = c;
    return SMJ_OK;
}

static void ctx_destroy(SmjCtx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < SmjCtx::kSlots; i++) if (c->slot[i]) cudaFree(c->slot[i]);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->pass_ev) if (e) cudaEventDestroy(e);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->pool) cudaMemPoolDestroy(c->pool);
    if (c->l2_fetch_set) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, c->l2_fetch_saved);
    if (c->d_out_ptr) cudaFree(c->d_out_ptr);
    if (c->d_err) cudaFree(c->d_err);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

void *smj_ws(SmjCtx *c, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 256;
    if (c->slot_bytes[slot] >= bytes) return c->slot[slot];
    if (c->slot[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->slot[slot]); c->slot[slot] = nullptr; c->slot_bytes[slot] = 0; }
    const size_t want = align_up(bytes + bytes / 16, 1 << 20);   // a little slack so near-equal sizes do not realloc
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { smj_cuda_fail(e, "workspace cudaMalloc", __FILE__, __LINE__); return nullptr; }
    c->slot[slot] = p;
    c->slot_bytes[slot] = want;
    c->ws_gen++;
    return p;
}
#define WS_TRY(var, type, c, slot, bytes) type var = (type)smj_ws(c, slot, bytes); if (!var) return SMJ_ENOMEM

extern "C" int smj_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int smj_version(void) { return SMJ_VERSION; }

extern "C" int smj_init(const smj_config_t *cfg)
{
    if (g_inited) {
        if (cfg) {
            if (cfg->nr_gpus > g_nctx) { smj_shutdown(); }
            else { g_cfg = *cfg; return SMJ_OK; }
        } else return SMJ_OK;
    }
    if (cfg) g_cfg = *cfg; else smj_config_default(&g_cfg);
    if (g_cfg.nr_gpus < 1) g_cfg.nr_gpus = 1;
    // all ranks on one GPU: up to 8 x 3 streams whose kernels wait for each other must not share a hardware queue (the default
    // is 8 connections: rank 7's kernel queued behind rank 0's spinning one never starts).  Read when the context is created.
    if (getenv("SMJ_RANKS_ON_ONE_GPU") && atoi(getenv("SMJ_RANKS_ON_ONE_GPU")) != 0) setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    const int ndev = smj_device_count();
    if (ndev < 1) return smj_set_error(SMJ_ENODEVICE, "no CUDA device visible (libsmj has no CPU fallback)");
    // SMJ_RANKS_ON_ONE_GPU=1: every rank of the one-process multi-GPU mode gets a context (streams, workspace, pool) on
    // device 0.  The whole key-range exchange -- mailboxes, routing to G buckets, receive buffers, verdicts -- then runs
    // on a single GPU, rank against rank on concurrent streams; what the tests and the profiler use where one GPU is all
    // there is.  (Only one-CTA kernels ever spin on a peer, so the ranks cannot starve each other of SMs.)
    const bool one_gpu = getenv("SMJ_RANKS_ON_ONE_GPU") && atoi(getenv("SMJ_RANKS_ON_ONE_GPU")) != 0;
    if ((!one_gpu && g_cfg.nr_gpus > ndev) || g_cfg.nr_gpus > 8)
        return smj_set_error(SMJ_EINVAL, "nr_gpus=%d but %d CUDA devices visible (max 8)", g_cfg.nr_gpus, ndev);
    for (int g = 0; g < g_cfg.nr_gpus; g++) {
        int r = ctx_create(one_gpu ? 0 : g, &g_ctx[g]);
        if (r != SMJ_OK) { smj_shutdown(); return r; }
        g_nctx = g + 1;
    }
    cudaSetDevice(g_ctx[0]->device);
    g_inited = true;
    return SMJ_OK;
}

int smj_dist_shutdown(void);
static void pinned_clear(void);

// One-process-per-GPU mode: this process drives exactly `device` (as g_ctx[0]).
int smj_init_on_device(const smj_config_t *cfg, int device)
{
    if (g_inited) smj_shutdown();
    if (cfg) g_cfg = *cfg; else smj_config_default(&g_cfg);
    const int ndev = smj_device_count();
    if (ndev < 1) return smj_set_error(SMJ_ENODEVICE, "no CUDA device visible (libsmj has no CPU fallback)");
    if (device < 0 || device >= ndev) return smj_set_error(SMJ_EINVAL, "local device %d out of range (%d visible)", device, ndev);
    int r = ctx_create(device, &g_ctx[0]);
    if (r != SMJ_OK) { smj_shutdown(); return r; }
    g_nctx = 1;
    g_inited = true;
    return SMJ_OK;
}

extern "C" void smj_shutdown(void)
{
    smj_dist_shutdown();
    pinned_clear();
    for (int g = 0; g < 8; g++) { ctx_destroy(g_ctx[g]); g_ctx[g] = nullptr; }
    g_nctx = 0;
    g_inited = false;
}

int smj_ensure_init(void)
{
    if (g_inited) return SMJ_OK;
    return smj_init(nullptr);
}

extern "C" int64_t smj_kernel_launches(void)
{
    int64_t n = 0;
    for (int g = 0; g < g_nctx; g++) n += g_ctx[g]->launches;
    return n;
}

// ------------------------------------------------------------------ memory helpers
extern "C" int smj_host_alloc(void **p, size_t bytes)
{
    SMJ_TRY(smj_ensure_init());
    CUDA_TRY(cudaMallocHost(p, bytes ? bytes : 1));
    return SMJ_OK;
}
extern "C" void smj_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int smj_device_alloc(void **p, size_t bytes)
{
    SMJ_TRY(smj_ensure_init());
    CUDA_TRY(cudaSetDevice(g_ctx[0]->device));
    CUDA_TRY(cudaMalloc(p, bytes ? bytes : 1));
    return SMJ_OK;
}
extern "C" void smj_device_free(void *p) { if (p) cudaFree(p); }
extern "C" int smj_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
    SMJ_TRY(smj_ensure_init());
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx[0]->stream));
    CUDA_TRY(cudaStreamSynchronize(g_ctx[0]->stream));
    return SMJ_OK;
}
extern "C" int smj_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
    SMJ_TRY(smj_ensure_init());
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx[0]->stream));
    CUDA_TRY(cudaStreamSynchronize(g_ctx[0]->stream));
    return SMJ_OK;
}
extern "C" int smj_device_sync(void)
{
    SMJ_TRY(smj_ensure_init());
    for (int g = 0; g < g_nctx; g++) {
        CUDA_TRY(cudaSetDevice(g_ctx[g]->device));
        CUDA_TRY(cudaDeviceSynchronize());
    }
    CUDA_TRY(cudaSetDevice(g_ctx[0]->device));
    return SMJ_OK;
}

extern "C" int smj_synth_table(int32_t *dev_out, int64_t row0, int64_t rows, int64_t total_rows, int cols, int key_col,
                               uint64_t seed, int kind, int64_t key_domain)
{
    SMJ_TRY(smj_ensure_init());
    SmjCtx *c = g_ctx[0];
    SMJ_TRY(smj_launch_synth(c, dev_out, row0, rows, total_rows, cols, key_col, seed, kind, key_domain));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJ_OK;
}

// ------------------------------------------------------------------ table staging
static int check_table(const smj_table_t *t, const char *name)
{
    if (!t) return smj_set_error(SMJ_EINVAL, "%s: null table", name);
    if (t->cols < 1 || t->rows < 0) return smj_set_error(SMJ_EINVAL, "%s: bad shape %lld x %d", name, (long long)t->rows, t->cols);
    if (t->rows > 0 && !t->data) return smj_set_error(SMJ_EINVAL, "%s: null data", name);
    if (t->rows > 0xffffffffll) return smj_set_error(SMJ_ETOOBIG, "%s: %lld rows exceed 32-bit row ids", name, (long long)t->rows);
    return SMJ_OK;
}

// Device view of an input table: the caller's pointer when it already lives in HBM, else an async H2D copy
// into a workspace slot (the analogue of dpu_prepare_xfer + dpu_push_xfer, app.c:222-243).
int smj_stage_in(SmjCtx *c, const smj_table_t *t, int slot, const int32_t **d)
{
    if (t->on_device || t->rows == 0) { *d = t->data; return SMJ_OK; }
    const size_t bytes = (size_t)t->rows * t->cols * sizeof(int32_t);
    WS_TRY(p, int32_t *, c, slot, bytes);
    CUDA_TRY(cudaMemcpyAsync(p, t->data, bytes, cudaMemcpyHostToDevice, c->stream));
    *d = p;
    return SMJ_OK;
}

// Pinned host buffers for host-side outputs are recycled: cudaMallocHost costs milliseconds per call,
// which would dominate the GPU->CPU leg (app.c timer 2) of a sub-millisecond pipeline.
struct PinnedEntry { void *p; size_t bytes; bool used; };
static PinnedEntry g_pinned[16];

static int pinned_get(void **out, size_t bytes)
{
    int best = -1, spare = -1;
    for (int i = 0; i < 16; i++) {
        if (g_pinned[i].p && !g_pinned[i].used && g_pinned[i].bytes >= bytes &&
            (best < 0 || g_pinned[i].bytes < g_pinned[best].bytes)) best = i;
        if (!g_pinned[i].p && spare < 0) spare = i;
    }
    if (best >= 0 && g_pinned[best].bytes <= 4 * bytes + (1 << 20)) { g_pinned[best].used = true; *out = g_pinned[best].p; return SMJ_OK; }
    if (spare < 0)
        for (int i = 0; i < 16; i++)
            if (!g_pinned[i].used) { cudaFreeHost(g_pinned[i].p); g_pinned[i].p = nullptr; spare = i; break; }
    void *p = nullptr;
    CUDA_TRY(cudaMallocHost(&p, bytes));
    if (spare >= 0) g_pinned[spare] = {p, bytes, true};
    *out = p;
    return SMJ_OK;
}
static void pinned_put(void *p)
{
    for (int i = 0; i < 16; i++)
        if (g_pinned[i].p == p) { g_pinned[i].used = false; return; }
    cudaFreeHost(p);
}
// frees the pooled buffers that are not handed out; an output table a caller still holds leaves the pool (its
// smj_table_free then releases it with cudaFreeHost) instead of dangling after a shutdown or re-init
static void pinned_clear(void)
{
    for (int i = 0; i < 16; i++) {
        if (g_pinned[i].p && !g_pinned[i].used) cudaFreeHost(g_pinned[i].p);
        g_pinned[i] = {nullptr, 0, false};
    }
}

// Output allocation: device memory from the stream-ordered pool, or (recycled) pinned host memory.
int smj_alloc_out(SmjCtx *c, smj_table_t *out, int64_t rows, int cols)
{
    const int on_device = out->on_device;
    out->rows = rows; out->cols = cols; out->data = nullptr;
    const size_t bytes = (size_t)rows * cols * sizeof(int32_t);
    if (bytes == 0) return SMJ_OK;
    if (on_device) CUDA_TRY(cudaMallocFromPoolAsync((void **)&out->data, bytes, c->pool, c->stream));
    else SMJ_TRY(pinned_get((void **)&out->data, bytes));
    return SMJ_OK;
}

extern "C" void smj_table_free(smj_table_t *t)
{
    if (!t || !t->data) return;
    if (t->on_device) {
        cudaPointerAttributes at;
        SmjCtx *c = g_nctx > 0 ? g_ctx[0] : nullptr;
        if (cudaPointerGetAttributes(&at, t->data) == cudaSuccess)
            for (int g = 0; g < g_nctx; g++) if (g_ctx[g]->device == at.device) c = g_ctx[g];
        if (c) { cudaSetDevice(c->device); cudaFreeAsync(t->data, c->stream); cudaSetDevice(g_ctx[0]->device); }
        else cudaFree(t->data);
    } else pinned_put(t->data);
    t->data = nullptr; t->rows = 0;
}

// Produces the output table from a device buffer of rows (device->device copy is avoided by allocating the
// final buffer up front when the destination is the device).
static int emit_out_from_device(SmjCtx *c, smj_table_t *out, const int32_t *d_rows, int64_t rows, int cols)
{
    SMJ_TRY(smj_alloc_out(c, out, rows, cols));
    const size_t bytes = (size_t)rows * cols * sizeof(int32_t);
    if (bytes)
        CUDA_TRY(cudaMemcpyAsync(out->data, d_rows, bytes, out->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                 c->stream));
    return SMJ_OK;
}

int smj_check_device_flag(SmjCtx *c)
{
    u32 *h = (u32 *)((char *)c->h_pinned + c->h_pinned_bytes - 64);
    CUDA_TRY(cudaMemcpyAsync(h, c->d_err, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (*h) {
        const u32 code = *h;
        cudaMemsetAsync(c->d_err, 0, 4, c->stream);
        if (code == 6)
            return smj_set_error(SMJ_EINVAL, "an input table is not sorted by its key column (smj_join, smj_join_count and smj_merge take "
                                             "tables sorted by key: sort them with smj_sort first)");
        if (code >= 4)
            return smj_set_error(SMJ_EINTERNAL, "device-side check failed (code %u: %s)", code,
                                 code == 4 ? "a rank waited 30 s for a peer's flag: some rank did not reach this step of the exchange"
                                           : "the local select waited 30 s for the exchange of its table to arrive");
        return smj_set_error(SMJ_EINTERNAL, "device-side check failed (code %u: look-back spin limit in %s)", code,
                             code == 1 ? "select" : code == 2 ? "radix sort" : "join");
    }
    return SMJ_OK;
}

// Scratch header, zeroed at the start of every call that uses it.
struct ScratchHeader {
    u32 hist[2][SMJ_KEY_PASSES * SMJ_RADIX];
    u64 count[2];
    u64 jcount;
    u64 pad0;
    u32 counter[16];     // 0,1: select tickets; 2: join ticket
    SmjSortPlan plan[2]; // smj_run: key range of each table's survivors -> radix passes to run
    u64 sel_count[2];    // smj_run: rows that passed the predicate (count[] = those that also passed the semi-join filter)
    u64 kept_count[2];   // smj_run: rows of the table selected second that passed the in-select bitmap probe
    u32 use_store[2];    // smj_run: the table's pairs carry dense row-store indices (decided on the device, plan_scan_kernel)
    u32 pad1[2];
};

// (key,rowid) pairs of every row of a device table (no predicate): used by sort / merge / join entry points.
static int pairs_of_table(SmjCtx *c, const int32_t *d_t, int64_t rows, int cols, int key_col, u32 rowid_base, u64 *d_pairs,
                          char *scratch_status, u32 *d_counter, u32 *d_hist, u64 *d_count)
{
    return smj_launch_select_pairs(c, d_t, rows, cols, key_col, 0, /*select_all=*/1, key_col, rowid_base, d_pairs, nullptr,
                                   (u64 *)scratch_status, d_counter, d_hist, d_count);
}

// ------------------------------------------------------------------ smj_select
extern "C" int smj_select(const smj_table_t *in, int col, int64_t val, smj_table_t *out)
{
    SMJ_TRY(smj_ensure_init());
    SMJ_TRY(check_table(in, "smj_select"));
    if (!out) return smj_set_error(SMJ_EINVAL, "smj_select: null out");
    if (col < 0 || col >= in->cols) return smj_set_error(SMJ_EINVAL, "smj_select: column %d out of range", col);
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int64_t n = in->rows;
    const int32_t *d_in;
    SMJ_TRY(smj_stage_in(c, in, WS_T1, &d_in));
    const size_t tiles = smj_select_num_tiles(n);
    const size_t sbytes = sizeof(ScratchHeader) + tiles * 8;
    WS_TRY(scr, char *, c, WS_SCRATCH, sbytes);
    CUDA_TRY(cudaMemsetAsync(scr, 0, sbytes, c->stream));
    ScratchHeader *h = (ScratchHeader *)scr;
    WS_TRY(pairs, u64 *, c, WS_PAIRS_A1, (size_t)n * 8);
    WS_TRY(tmp_pairs, u64 *, c, WS_PAIRS_B1, (size_t)n * 8);
    SMJ_TRY(smj_launch_select_pairs(c, d_in, n, in->cols, col, val, 0, col, 0, pairs, tmp_pairs, (u64 *)(scr + sizeof(ScratchHeader)),
                                    &h->counter[0], nullptr, &h->count[0]));
    u64 *hm = (u64 *)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(hm, &h->count[0], 8, cudaMemcpyDeviceToHost, c->stream));
    SMJ_TRY(smj_check_device_flag(c));
    const int64_t m = (int64_t)hm[0];
    if (out->on_device) {
        SMJ_TRY(smj_alloc_out(c, out, m, in->cols));
        SMJ_TRY(smj_launch_gather_rows(c, pairs, m, d_in, in->cols, out->data));
    } else {
        WS_TRY(tmp, int32_t *, c, WS_TMP_ROWS, (size_t)m * in->cols * 4);
        SMJ_TRY(smj_launch_gather_rows(c, pairs, m, d_in, in->cols, tmp));
        SMJ_TRY(emit_out_from_device(c, out, tmp, m, in->cols));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJ_OK;
}

// ------------------------------------------------------------------ smj_sort
// Sorts the n pairs in `pairs` (ping buffer) using `pong`; d_hist must hold their digit histogram.  Result in `pairs`.
static int sort_pairs_with_hist(SmjCtx *c, u64 *pairs, u64 *pong, u32 n, const u32 *d_hist, int scratch_slot)
{
    const size_t rbytes = smj_radix_scratch_bytes(n);
    WS_TRY(rs, u32 *, c, scratch_slot, rbytes);
    CUDA_TRY(cudaMemsetAsync(rs, 0, rbytes, c->stream));
    return smj_radix_sort_pairs(c, pairs, pong, nullptr, n, d_hist, rs);
}

extern "C" int smj_sort(smj_table_t *inout, int key_col)
{
    SMJ_TRY(smj_ensure_init());
    SMJ_TRY(check_table(inout, "smj_sort"));
    if (key_col < 0 || key_col >= inout->cols) return smj_set_error(SMJ_EINVAL, "smj_sort: column %d out of range", key_col);
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int64_t n = inout->rows;
    if (n < 2) return SMJ_OK;
    if (n > SMJ_MAX_SORT_ROWS) return smj_set_error(SMJ_ETOOBIG, "smj_sort: %lld rows exceed 2^30 - 1", (long long)n);
    const int32_t *d_in;
    SMJ_TRY(smj_stage_in(c, inout, WS_T1, &d_in));
    const size_t tiles = smj_select_num_tiles(n);
    const size_t sbytes = sizeof(ScratchHeader) + tiles * 8;
    WS_TRY(scr, char *, c, WS_SCRATCH, sbytes);
    CUDA_TRY(cudaMemsetAsync(scr, 0, sbytes, c->stream));
    ScratchHeader *h = (ScratchHeader *)scr;
    WS_TRY(ping, u64 *, c, WS_PAIRS_A1, (size_t)n * 8);
    WS_TRY(pong, u64 *, c, WS_PAIRS_B1, (size_t)n * 8);
    SMJ_TRY(pairs_of_table(c, d_in, n, inout->cols, key_col, 0, ping, scr + sizeof(ScratchHeader), &h->counter[0],
                           h->hist[0], &h->count[0]));
    c->pass_count = 0;
    SMJ_TRY(sort_pairs_with_hist(c, ping, pong, (u32)n, h->hist[0], WS_RADIX));
    const size_t bytes = (size_t)n * inout->cols * 4;
    WS_TRY(tmp, int32_t *, c, WS_TMP_ROWS, bytes);
    SMJ_TRY(smj_launch_gather_rows(c, ping, n, d_in, inout->cols, tmp));
    CUDA_TRY(cudaMemcpyAsync(inout->data, tmp, bytes, inout->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             c->stream));
    SMJ_TRY(smj_check_device_flag(c));
    return SMJ_OK;
}

// ------------------------------------------------------------------ smj_merge
extern "C" int smj_merge(const smj_table_t *a, const smj_table_t *b, int key_col, smj_table_t *out)
{
    SMJ_TRY(smj_ensure_init());
    SMJ_TRY(check_table(a, "smj_merge(a)"));
    SMJ_TRY(check_table(b, "smj_merge(b)"));
    if (!out) return smj_set_error(SMJ_EINVAL, "smj_merge: null out");
    if (a->cols != b->cols) return smj_set_error(SMJ_EINVAL, "smj_merge: runs of one table must have equal column counts");
    if (key_col < 0 || key_col >= a->cols) return smj_set_error(SMJ_EINVAL, "smj_merge: column %d out of range", key_col);
    const int64_t total = a->rows + b->rows;
    if (total > SMJ_MAX_SORT_ROWS) return smj_set_error(SMJ_ETOOBIG, "smj_merge: %lld rows exceed 2^30 - 1", (long long)total);
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int cols = a->cols;
    const int32_t *d_a, *d_b;
    SMJ_TRY(smj_stage_in(c, a, WS_T1, &d_a));
    SMJ_TRY(smj_stage_in(c, b, WS_T2, &d_b));
    const size_t ta = smj_select_num_tiles(a->rows), tb = smj_select_num_tiles(b->rows);
    const size_t sbytes = sizeof(ScratchHeader) + (ta + tb) * 8;
    WS_TRY(scr, char *, c, WS_SCRATCH, sbytes);
    CUDA_TRY(cudaMemsetAsync(scr, 0, sbytes, c->stream));
    ScratchHeader *h = (ScratchHeader *)scr;
    WS_TRY(pa, u64 *, c, WS_PAIRS_A1, (size_t)a->rows * 8);
    WS_TRY(pb, u64 *, c, WS_PAIRS_A2, (size_t)b->rows * 8);
    WS_TRY(pm, u64 *, c, WS_PAIRS_B1, (size_t)total * 8);
    SMJ_TRY(pairs_of_table(c, d_a, a->rows, cols, key_col, 0, pa, scr + sizeof(ScratchHeader), &h->counter[0], nullptr, &h->count[0]));
    SMJ_TRY(pairs_of_table(c, d_b, b->rows, cols, key_col, (u32)a->rows, pb, scr + sizeof(ScratchHeader) + ta * 8,
                           &h->counter[1], nullptr, &h->count[1]));
    WS_TRY(part, u32 *, c, WS_PART, (smj_merge_num_tiles((u64)total) + 2) * 4);
    SMJ_TRY(smj_launch_merge_pairs(c, pa, (u32)a->rows, pb, (u32)b->rows, pm, part));
    if (out->on_device) {
        SMJ_TRY(smj_alloc_out(c, out, total, cols));
        SMJ_TRY(smj_launch_gather_rows2(c, pm, total, d_a, d_b, (u32)a->rows, cols, out->data));
    } else {
        WS_TRY(tmp, int32_t *, c, WS_TMP_ROWS, (size_t)total * cols * 4);
        SMJ_TRY(smj_launch_gather_rows2(c, pm, total, d_a, d_b, (u32)a->rows, cols, tmp));
        SMJ_TRY(emit_out_from_device(c, out, tmp, total, cols));
    }
    const int rc = smj_check_device_flag(c);
    if (rc != SMJ_OK) smj_table_free(out);   // runs that were not sorted: no half-merged table for the caller
    return rc;
}

// ------------------------------------------------------------------ smj_join
// Join scratch of `tiles` tiles inside a byte arena:
// [tile_count u32 x tiles (zeroed)] [tile_off u64 x tiles + the scan's block sums] [part u32 x 2(tiles+1)]
struct JoinScratch { u32 *tile_count; u64 *tile_off; u32 *part; size_t zero_bytes, bytes; };
static JoinScratch join_scratch(char *base, size_t tiles)
{
    JoinScratch j;
    const size_t off_bytes = align_up((tiles + smj_join_scan_blocks(tiles)) * 8, 256);
    j.zero_bytes = align_up(tiles * 4, 256);
    j.tile_count = (u32 *)base;
    j.tile_off = (u64 *)(base + j.zero_bytes);
    j.part = (u32 *)(base + j.zero_bytes + off_bytes);
    j.bytes = j.zero_bytes + off_bytes + align_up((tiles + 1) * 2 * 4, 256);
    return j;
}

static int join_sorted_pairs(SmjCtx *c, const u64 *pl, u32 m1, const u64 *pr, u32 m2, int mode, bool count_only,
                             const int32_t *d_t1, int c1, const int32_t *d_t2, int c2, int key2, smj_table_t *out,
                             int64_t *rows_out)
{
    const size_t tiles = (m1 == 0 || m2 == 0) ? 0 : smj_join_num_tiles((u64)m1 + m2);
    const JoinScratch sz = join_scratch(nullptr, tiles);
    WS_TRY(js, char *, c, WS_PART, 256 + sz.bytes);
    const JoinScratch jsr = join_scratch(js + 256, tiles);
    CUDA_TRY(cudaMemsetAsync(js, 0, 256 + jsr.zero_bytes, c->stream));
    u64 *d_jcount = (u64 *)js;
    uint2 *d_matches = nullptr;
    uint2 *d_dense = nullptr;
    if (mode == SMJ_JOIN_ZIP) {
        WS_TRY(mm, uint2 *, c, WS_MATCH, tiles * smj_join_tile_size() * 8); d_matches = mm;
        if (!count_only) { WS_TRY(dd, uint2 *, c, WS_MATCH_DENSE, (size_t)(m1 < m2 ? m1 : m2) * 8); d_dense = dd; }
    } else if (!count_only) {
        WS_TRY(mm, uint2 *, c, WS_MATCH, (size_t)m1 * 8 + 8); d_matches = mm;   // (first right position, run length) per left element
        CUDA_TRY(cudaMemsetAsync(mm, 0, (size_t)m1 * 8 + 8, c->stream));       // left elements past the right table's end keep an empty run
    }
    SMJ_TRY(smj_launch_join_match(c, pl, pr, nullptr, m1, m2, mode, jsr.part, jsr.tile_count, jsr.tile_off, d_matches, d_dense, d_jcount));
    u64 *hm = (u64 *)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(hm, d_jcount, 8, cudaMemcpyDeviceToHost, c->stream));
    SMJ_TRY(smj_check_device_flag(c));
    const int64_t j = (int64_t)hm[0];
    *rows_out = j;
    if (mode == SMJ_JOIN_ZIP && j > (int64_t)(m1 < m2 ? m1 : m2))   // impossible for sorted inputs (every left row pairs with its own right row)
        return smj_set_error(SMJ_EINVAL, "an input table is not sorted by its key column (smj_join and smj_join_count take tables sorted by key)");
    if (count_only) return SMJ_OK;
    if (mode == SMJ_JOIN_MANY) {
        // the many-to-many result can dwarf its inputs (cL * cR rows per key): refuse what cannot be held, expand the rest
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const double need = (double)j * (8.0 + 4.0 * (c1 + c2 - 1)) + (double)smj_join_many_scratch_bytes(m1);
        if (need > 0.8 * (double)free_b)
            return smj_set_error(SMJ_ETOOBIG, "SMJ_JOIN_MANY result of %lld rows needs %.1f GB; use smj_join_count or SMJ_JOIN_ZIP", (long long)j, need / 1e9);
        WS_TRY(dd, uint2 *, c, WS_MATCH_DENSE, (size_t)j * 8);
        WS_TRY(xs, char *, c, WS_MERGE_A, smj_join_many_scratch_bytes(m1));
        d_dense = dd;
        SMJ_TRY(smj_launch_join_many_expand(c, pl, pr, d_matches, m1, (u64)j, xs, d_dense));
    }
    const int c_out = c1 + c2 - 1;
    if (out->on_device) {
        SMJ_TRY(smj_alloc_out(c, out, j, c_out));
        SMJ_TRY(smj_launch_join_materialize(c, d_dense, nullptr, j, d_t1, c1, d_t2, c2, key2, out->data));
    } else {
        WS_TRY(tmp, int32_t *, c, WS_TMP_ROWS, (size_t)j * c_out * 4);
        SMJ_TRY(smj_launch_join_materialize(c, d_dense, nullptr, j, d_t1, c1, d_t2, c2, key2, tmp));
        SMJ_TRY(emit_out_from_device(c, out, tmp, j, c_out));
    }
    return SMJ_OK;
}

int smj_join_pairs_to_table(SmjCtx *c, const u64 *pl, u32 m1, const u64 *pr, u32 m2, const int32_t *d_t1, int c1, const int32_t *d_t2,
                            int c2, int key2, smj_table_t *out, int64_t *rows_out)
{
    return join_sorted_pairs(c, pl, m1, pr, m2, SMJ_JOIN_ZIP, false, d_t1, c1, d_t2, c2, key2, out, rows_out);
}

// select (or all rows) -> (key,rowid) pairs -> stable sort; one host wait for the survivor count.
// table_idx picks the workspace slots (0: table 1, 1: table 2).  *d_sorted stays valid until those slots are reused.
int smj_sorted_pairs_of_table(SmjCtx *c, const int32_t *d_in, int64_t n, int cols, int sel_col, int64_t sel_val, int select_all,
                              int key_col, int table_idx, u64 **d_sorted, int64_t *m_out)
{
    if (n > SMJ_MAX_SORT_ROWS) return smj_set_error(SMJ_ETOOBIG, "%lld rows exceed 2^30 - 1 per table per GPU", (long long)n);
    const size_t sw = smj_select_num_tiles(n);
    const size_t off_sel = align_up(sizeof(ScratchHeader), 256);
    const size_t off_radix = align_up(off_sel + sw * 8, 256);
    const size_t rb = align_up(smj_radix_scratch_bytes((u32)n), 256);
    WS_TRY(scr, char *, c, WS_RADIX, off_radix + rb);
    WS_TRY(ping, u64 *, c, table_idx ? WS_PAIRS_A2 : WS_PAIRS_A1, (size_t)n * 8);
    WS_TRY(pong, u64 *, c, table_idx ? WS_PAIRS_B2 : WS_PAIRS_B1, (size_t)n * 8);
    CUDA_TRY(cudaMemsetAsync(scr, 0, off_radix + rb, c->stream));
    ScratchHeader *h = (ScratchHeader *)scr;
    SMJ_TRY(smj_launch_select_pairs(c, d_in, n, cols, sel_col, sel_val, select_all, key_col, 0, ping, pong, (u64 *)(scr + off_sel),
                                    &h->counter[0], h->hist[0], &h->count[0]));
    SMJ_TRY(smj_radix_sort_pairs(c, ping, pong, &h->count[0], (u32)n, h->hist[0], (u32 *)(scr + off_radix)));
    u64 *hm = (u64 *)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(hm, &h->count[0], 8, cudaMemcpyDeviceToHost, c->stream));
    SMJ_TRY(smj_check_device_flag(c));
    *m_out = (int64_t)hm[0];
    *d_sorted = ping;
    return SMJ_OK;
}

static int join_tables(const smj_table_t *l, const smj_table_t *r, int key1, int key2, int mode, bool count_only,
                       smj_table_t *out, int64_t *rows_out)
{
    SMJ_TRY(smj_ensure_init());
    SMJ_TRY(check_table(l, "smj_join(l)"));
    SMJ_TRY(check_table(r, "smj_join(r)"));
    if (key1 < 0 || key1 >= l->cols || key2 < 0 || key2 >= r->cols) return smj_set_error(SMJ_EINVAL, "smj_join: key column out of range");
    if (mode != SMJ_JOIN_ZIP && mode != SMJ_JOIN_MANY) return smj_set_error(SMJ_EINVAL, "smj_join: bad mode %d", mode);
    if (l->rows > SMJ_MAX_SORT_ROWS || r->rows > SMJ_MAX_SORT_ROWS) return smj_set_error(SMJ_ETOOBIG, "smj_join: more than 2^30 - 1 rows");
    SmjCtx *c = g_ctx[0];
    CUDA_TRY(cudaSetDevice(c->device));
    const int32_t *d_l, *d_r;
    SMJ_TRY(smj_stage_in(c, l, WS_T1, &d_l));
    SMJ_TRY(smj_stage_in(c, r, WS_T2, &d_r));
    const size_t tl = smj_select_num_tiles(l->rows), tr = smj_select_num_tiles(r->rows);
    const size_t sbytes = sizeof(ScratchHeader) + (tl + tr) * 8;
    WS_TRY(scr, char *, c, WS_SCRATCH, sbytes);
    CUDA_TRY(cudaMemsetAsync(scr, 0, sbytes, c->stream));
    ScratchHeader *h = (ScratchHeader *)scr;
    WS_TRY(pl, u64 *, c, WS_PAIRS_A1, (size_t)l->rows * 8);
    WS_TRY(pr, u64 *, c, WS_PAIRS_A2, (size_t)r->rows * 8);
    SMJ_TRY(pairs_of_table(c, d_l, l->rows, l->cols, key1, 0, pl, scr + sizeof(ScratchHeader), &h->counter[0], nullptr, &h->count[0]));
    SMJ_TRY(pairs_of_table(c, d_r, r->rows, r->cols, key2, 0, pr, scr + sizeof(ScratchHeader) + tl * 8, &h->counter[1], nullptr,
                           &h->count[1]));
    SMJ_TRY(join_sorted_pairs(c, pl, (u32)l->rows, pr, (u32)r->rows, mode, count_only, d_l, l->cols, d_r, r->cols, key2, out, rows_out));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SMJ_OK;
}

extern "C" int smj_join(const smj_table_t *l, const smj_table_t *r, int key1, int key2, int mode, smj_table_t *out)
{
    if (!out) return smj_set_error(SMJ_EINVAL, "smj_join: null out");
    int64_t rows;
    return join_tables(l, r, key1, key2, mode, false, out, &rows);
}

extern "C" int smj_join_count(const smj_table_t *l, const smj_table_t *r, int key1, int key2, int mode, int64_t *rows)
{
    if (!rows) return smj_set_error(SMJ_EINVAL, "smj_join_count: null rows");
    return join_tables(l, r, key1, key2, mode, true, nullptr, rows);
}

// ------------------------------------------------------------------ smj_run (single GPU)
// SURVEY.md section 8d algorithmic-bytes model (w = 4, P = 4 radix passes).
static double bytes_model(const int64_t n[2], const int c[2], const int64_t m[2], int64_t j)
{
    double b = 0;
    for (int t = 0; t < 2; t++)
        b += (double)n[t] * c[t] * 4 + 8.0 * m[t] + 8.0 * m[t] + 16.0 * SMJ_KEY_PASSES * m[t] + 4.0 * m[t] + 2.0 * m[t] * c[t] * 4;
    b += 8.0 * (m[0] + m[1]) + (double)j * (c[0] + c[1] + (c[0] + c[1] - 1)) * 4;
    return b;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ms;
}

enum { E_START, E_H2D, E_SELECT, E_SORT, E_JOIN, E_D2H };   // c->ev[] of the single-GPU pipeline

// Device pipeline of one GPU, no host round trip between the stages: the survivor counts, the radix
// histograms and the match count stay in device memory, every kernel sizes its work from them, and buffers are
// sized by their host-known upper bounds (input rows).  The host waits once, at the end.
int smj_run_single(SmjCtx *c, const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out,
                   smj_stats_t *stats)
{
    CUDA_TRY(cudaSetDevice(c->device));
    const smj_table_t *tb[2] = {t1, t2};
    const int sel_col[2] = {cfg->select_col1, cfg->select_col2};
    const int64_t sel_val[2] = {cfg->select_val1, cfg->select_val2};
    const int key[2] = {cfg->join_key1, cfg->join_key2};
    for (int t = 0; t < 2; t++) {
        if (sel_col[t] < 0 || sel_col[t] >= tb[t]->cols) return smj_set_error(SMJ_EINVAL, "SELECT_COL%d=%d out of range", t + 1, sel_col[t]);
        if (key[t] < 0 || key[t] >= tb[t]->cols) return smj_set_error(SMJ_EINVAL, "JOIN_KEY%d=%d out of range", t + 1, key[t]);
        if (tb[t]->rows > SMJ_MAX_SORT_ROWS)
            return smj_set_error(SMJ_ETOOBIG, "table %d has %lld rows; this build handles at most 2^30 - 1 per table per GPU", t + 1,
                                 (long long)tb[t]->rows);
    }
    if (cfg->join_mode != SMJ_JOIN_ZIP && cfg->join_mode != SMJ_JOIN_MANY) return smj_set_error(SMJ_EINVAL, "smj_run: bad join_mode %d", cfg->join_mode);
    const int64_t launches0 = c->launches;
    c->pass_count = 0;
    if (cfg->join_mode == SMJ_JOIN_MANY) {
        // Extension (not the reference's semantics): every pair of equal keys.  The result size is data dependent and
        // unbounded by the inputs, so this path waits for the host between the stages (stage entry points underneath).
        CUDA_TRY(cudaEventRecord(c->ev[E_START], c->stream));
        const int32_t *dm[2];
        SMJ_TRY(smj_stage_in(c, t1, WS_T1, &dm[0]));
        SMJ_TRY(smj_stage_in(c, t2, WS_T2, &dm[1]));
        CUDA_TRY(cudaEventRecord(c->ev[E_H2D], c->stream));
        u64 *sp[2];
        int64_t mm_[2];
        for (int t = 0; t < 2; t++)
            SMJ_TRY(smj_sorted_pairs_of_table(c, dm[t], tb[t]->rows, tb[t]->cols, sel_col[t], sel_val[t], 0, key[t], t, &sp[t], &mm_[t]));
        CUDA_TRY(cudaEventRecord(c->ev[E_SORT], c->stream));
        int64_t jm = 0;
        SMJ_TRY(join_sorted_pairs(c, sp[0], (u32)mm_[0], sp[1], (u32)mm_[1], SMJ_JOIN_MANY, false, dm[0], tb[0]->cols, dm[1], tb[1]->cols, key[1], out, &jm));
        CUDA_TRY(cudaEventRecord(c->ev[E_JOIN], c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (stats) {
            memset(stats, 0, sizeof *stats);
            stats->h2d_ms = ev_ms(c->ev[E_START], c->ev[E_H2D]);
            stats->sort_ms = ev_ms(c->ev[E_H2D], c->ev[E_SORT]);
            stats->join_ms = ev_ms(c->ev[E_SORT], c->ev[E_JOIN]);
            stats->total_device_ms = ev_ms(c->ev[E_H2D], c->ev[E_JOIN]);
            for (int t = 0; t < 2; t++) { stats->rows_in[t] = tb[t]->rows; stats->rows_selected[t] = mm_[t]; }
            stats->rows_joined = jm;
            stats->kernel_launches = c->launches - launches0;
        }
        return SMJ_OK;
    }

    SmjRun R;
    SMJ_TRY(smj_run_prepare(c, cfg, t1, t2, nullptr, &R));
    int rc = smj_run_enqueue(c, &R);
    if (rc != SMJ_OK) { smj_run_abandon(c, &R); return rc; }
    return smj_run_finish(c, &R, out, stats);
}

// The three phases of the device pipeline of one GPU.  Split so that a process driving several GPUs (smj_dist.cu) can do
// every rank's allocations first, then enqueue every rank's kernels, then wait for each rank -- no allocation and no
// host wait happens while another rank's kernels spin on this rank's flags.
//   prepare: validation, workspace slots, the (upper-bound sized) output buffer, the H2D copies of host tables
//   enqueue: ~23 launches with no host wait in between (or one graph launch)
//   finish : the one host wait: counts, consistency flag, D2H, stats
// d_rows (may be null, or hold nulls): device-resident row counts of the input tables; the descriptors' rows are then
// upper bounds (the tables are receive buffers whose fill the host never learns: smj_dist.cu).

int smj_run_prepare(SmjCtx *c, const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, const u64 *const *d_rows, SmjRun *R)
{
    CUDA_TRY(cudaSetDevice(c->device));
    *R = SmjRun();
    R->cfg = *cfg;
    R->tb[0] = *t1; R->tb[1] = *t2;
    if (d_rows) { R->d_rows[0] = d_rows[0]; R->d_rows[1] = d_rows[1]; }
    const int sel_col[2] = {cfg->select_col1, cfg->select_col2};
    const int key[2] = {cfg->join_key1, cfg->join_key2};
    for (int t = 0; t < 2; t++) {
        if (sel_col[t] < 0 || sel_col[t] >= R->tb[t].cols) return smj_set_error(SMJ_EINVAL, "SELECT_COL%d=%d out of range", t + 1, sel_col[t]);
        if (key[t] < 0 || key[t] >= R->tb[t].cols) return smj_set_error(SMJ_EINVAL, "JOIN_KEY%d=%d out of range", t + 1, key[t]);
        if (R->tb[t].rows > SMJ_MAX_SORT_ROWS)
            return smj_set_error(SMJ_ETOOBIG, "table %d has %lld rows; this build handles at most 2^30 - 1 per table per GPU", t + 1,
                                 (long long)R->tb[t].rows);
    }
    if (cfg->join_mode != SMJ_JOIN_ZIP) return smj_set_error(SMJ_EINVAL, "smj_run_prepare: zip mode only");
    R->launches0 = c->launches;
    c->pass_count = 0;
    // ---- CPU -> GPU (app.c timer 0)
    CUDA_TRY(cudaEventRecord(c->ev[E_START], c->stream));
    SMJ_TRY(smj_stage_in(c, t1, WS_T1, &R->d_t[0]));
    SMJ_TRY(smj_stage_in(c, t2, WS_T2, &R->d_t[1]));
    const int64_t n[2] = {t1->rows, t2->rows};
    const int cc[2] = {t1->cols, t2->cols};
    R->c_out = cc[0] + cc[1] - 1;
    // one zeroed arena: [header][select status x2][radix scratch x2][join header 64 B + join status]; then join partitions
    R->stiles[0] = smj_select_num_tiles(n[0]); R->stiles[1] = smj_select_num_tiles(n[1]);
    R->rb[0] = align_up(smj_radix_scratch_bytes((u32)n[0]), 256); R->rb[1] = align_up(smj_radix_scratch_bytes((u32)n[1]), 256);
    R->jt = (n[0] == 0 || n[1] == 0) ? 0 : smj_join_num_tiles((u64)n[0] + n[1]);
    R->off_sel = align_up(sizeof(ScratchHeader), 256);
    R->off_radix = align_up(R->off_sel + (R->stiles[0] + R->stiles[1]) * 8, 256);
    R->off_join = R->off_radix + R->rb[0] + R->rb[1];
    const JoinScratch jsz = join_scratch(nullptr, R->jt);
    R->zero_bytes = R->off_join + jsz.zero_bytes;
    const size_t sbytes = R->off_join + jsz.bytes;
    WS_TRY(scr, char *, c, WS_SCRATCH, sbytes);
    WS_TRY(ping0, u64 *, c, WS_PAIRS_A1, (size_t)n[0] * 8);
    WS_TRY(ping1, u64 *, c, WS_PAIRS_A2, (size_t)n[1] * 8);
    WS_TRY(pong0, u64 *, c, WS_PAIRS_B1, (size_t)n[0] * 8);
    WS_TRY(pong1, u64 *, c, WS_PAIRS_B2, (size_t)n[1] * 8);
    R->scr = scr;
    R->ping[0] = ping0; R->ping[1] = ping1; R->pong[0] = pong0; R->pong[1] = pong1;
    R->j_max = n[0] < n[1] ? n[0] : n[1];
    // the join's per-tile match slots; before the sort the same bytes hold the select kernels' per-tile pair slots
    const size_t mm_bytes = R->jt * smj_join_tile_size() * 8 > (size_t)(n[0] + n[1]) * 8 ? R->jt * smj_join_tile_size() * 8 : (size_t)(n[0] + n[1]) * 8;
    WS_TRY(mm, uint2 *, c, WS_MATCH, mm_bytes);
    WS_TRY(md, uint2 *, c, WS_MATCH_DENSE, (size_t)R->j_max * 8);
    R->mm = mm; R->md = md;
    WS_TRY(bloom_ws, char *, c, WS_BLOOM, smj_bloom_bytes(n[0], n[1]));   // sized here: no allocation inside a graph capture
    (void)bloom_ws;
    // dense row stores (rowstore_kernel, smj_select.cu): OFF unless SMJ_ROWSTORE_MB is set.  Measured at 10M x 10M x 4 columns
    // (profiles/r02_rowstore_ab.txt): the materialise kernel's DRAM traffic falls from 392 to 110 MB (1.1x its algorithmic
    // bytes) and its time from 76 to 47 us, but copying the 1.67 M surviving 16-byte rows per table costs 56 us and 326 MB --
    // a sparse ascending read still pulls whole DRAM lines -- so the step goes from 0.374 to 0.400 ms.  With the knob set the
    // device uses a store only if the rows that go on to the sort fit it and are at most a quarter of the table.
    static const long store_mb = getenv("SMJ_ROWSTORE_MB") ? atol(getenv("SMJ_ROWSTORE_MB")) : 0;
    for (int t = 0; t < 2; t++) {
        const size_t row_bytes = (size_t)cc[t] * 4, table_bytes = (size_t)n[t] * row_bytes;
        if (store_mb <= 0 || smj_bloom_bytes(n[0], n[1]) == 0 || table_bytes < ((size_t)96 << 20)) continue;
        u64 max_rows = ((u64)store_mb << 20) / row_bytes;
        if (max_rows > (u64)n[t] / 4) max_rows = (u64)n[t] / 4;
        if (max_rows == 0) continue;
        int32_t *sp = (int32_t *)smj_ws(c, t ? WS_ROWSTORE2 : WS_ROWSTORE1, max_rows * row_bytes);
        if (!sp) return SMJ_ENOMEM;
        R->store[t] = sp;
        R->store_max_rows[t] = max_rows;
    }
    R->dev_out = {nullptr, 0, R->c_out, 1};
    SMJ_TRY(smj_alloc_out(c, &R->dev_out, R->j_max, R->c_out));   // upper bound; rows is set once the count is known
    CUDA_TRY(cudaEventRecord(c->ev[E_H2D], c->stream));
    R->prepared = true;
    return SMJ_OK;
}

// gives back what prepare took when the run cannot go on
void smj_run_abandon(SmjCtx *c, SmjRun *R)
{
    (void)c;
    if (R->prepared && R->dev_out.data) smj_table_free(&R->dev_out);
    R->prepared = false;
}

int smj_run_enqueue(SmjCtx *c, SmjRun *R)
{
    CUDA_TRY(cudaSetDevice(c->device));
    const smj_config_t *cfg = &R->cfg;
    const int sel_col[2] = {cfg->select_col1, cfg->select_col2};
    const int64_t sel_val[2] = {cfg->select_val1, cfg->select_val2};
    const int key[2] = {cfg->join_key1, cfg->join_key2};
    const int64_t n[2] = {R->tb[0].rows, R->tb[1].rows};
    const int cc[2] = {R->tb[0].cols, R->tb[1].cols};
    const int32_t *const *d_t = R->d_t;
    u64 *const *ping = R->ping, *const *pong = R->pong;
    char *scr = R->scr;
    const size_t off_sel = R->off_sel, off_radix = R->off_radix, off_join = R->off_join, zero_bytes = R->zero_bytes;
    const size_t *stiles = R->stiles, *rb = R->rb;
    const size_t jt = R->jt;
    uint2 *mm = R->mm, *md = R->md;
    const int64_t j_max = R->j_max;

    // ---- the device pipeline: ~23 launches with no host wait in between.  A call that repeats the previous call's
    // tables, shapes and knobs replays it as ONE CUDA graph (the launch gaps are ~8 % of a 0.6 ms step); the output
    // buffer is new every call, so the materialise kernel reads its pointer from a device cell written just before.
    ScratchHeader *h = (ScratchHeader *)scr;
    // The tables' addresses reach the kernels through device cells (like the output pointer), so a graph captured for one pair
    // of tables serves any other pair of the same shape, alignment and knobs.  Paths that bake the addresses into kernel
    // arguments (a table the TMA select cannot take, the opt-in row store) keep them in the key.
    const bool ind_ok = smj_select_takes_tma(d_t[0], cc[0]) && smj_select_takes_tma(d_t[1], cc[1]) && !R->store[0] && !R->store[1];
    const int32_t *const *d_tab_ind = ind_ok ? (const int32_t *const *)(c->d_out_ptr + 1) : nullptr;
    const u64 tkey[2] = {ind_ok ? ((u64)(uintptr_t)d_t[0] & 15u) | 16u : (u64)(uintptr_t)d_t[0],
                         ind_ok ? ((u64)(uintptr_t)d_t[1] & 15u) | 16u : (u64)(uintptr_t)d_t[1]};
    u64 key_now[16] = {tkey[0], tkey[1], (u64)n[0], (u64)n[1], (u64)cc[0], (u64)cc[1],
                       (u64)sel_col[0], (u64)sel_col[1], (u64)sel_val[0], (u64)sel_val[1], (u64)key[0], (u64)key[1], c->ws_gen,
                       (u64)(uintptr_t)scr, (u64)(uintptr_t)R->d_rows[0] ^ ((u64)(uintptr_t)R->wait[0].flag << 1),
                       (u64)(uintptr_t)R->d_rows[1] ^ ((u64)(uintptr_t)R->wait[1].flag << 1)};
    static const bool graphs_on = !(getenv("SMJ_NO_GRAPH") && atoi(getenv("SMJ_NO_GRAPH")) != 0);
    if (memcmp(key_now, c->graph_key, sizeof key_now) != 0) {
        memcpy(c->graph_key, key_now, sizeof key_now);
        c->graph_seen = 0;
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    }
    c->graph_seen++;
    const bool replay = graphs_on && !R->no_graph && c->graph_exec != nullptr;
    const bool capture = graphs_on && !R->no_graph && !replay && c->graph_seen >= 2;   // the first call warms attributes and slots eagerly
    R->replayed = replay || capture;
    int32_t **h_ptr = (int32_t **)((char *)c->h_pinned + c->h_pinned_bytes - 128);
    h_ptr[0] = R->dev_out.data;
    h_ptr[1] = const_cast<int32_t *>(d_t[0]);
    h_ptr[2] = const_cast<int32_t *>(d_t[1]);
    CUDA_TRY(cudaMemcpyAsync(c->d_out_ptr, h_ptr, 3 * sizeof(int32_t *), cudaMemcpyHostToDevice, c->stream));
    if (!replay) {
        const int64_t l0 = c->launches;
        if (capture) CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        c->capturing = capture;
        int rc = SMJ_OK;
        do {
#define PIPE_TRY(x) if ((rc = (x)) != SMJ_OK) break
#define PIPE_CUDA(x) if (cudaError_t e_ = (x)) { rc = smj_cuda_fail(e_, #x, __FILE__, __LINE__); break; }
            // select (+ digit histograms), both tables
            PIPE_CUDA(cudaMemsetAsync(scr, 0, zero_bytes, c->stream));
            bool planned = true;
            {
                SmjSelectJob job[2];
                for (int t = 0; t < 2; t++)
                    job[t] = {d_t[t], n[t], cc[t], sel_col[t], sel_val[t], 0, key[t], {ping[t], pong[t]}, (u64 *)mm + (t ? n[0] : 0),
                              (u64 *)(scr + off_sel + (t ? stiles[0] * 8 : 0)), h->hist[t], &h->count[t], &h->plan[t], &h->sel_count[t], &h->kept_count[t],
                              R->d_rows[t], R->wait[t], d_tab_ind ? d_tab_ind + t : nullptr, R->store[t], &h->use_store[t], R->store_max_rows[t]};
                rc = smj_launch_select_plan2(c, job);
                if (rc == 1) { planned = false; rc = SMJ_OK; }   // a table the TMA path cannot take: histograms in the select kernel, four passes
            }
            if (rc == SMJ_OK && !planned && (R->d_rows[0] || R->d_rows[1]))
                rc = smj_set_error(SMJ_EINVAL, "device-resident row counts need the TMA select path (16-byte aligned tables of <= 32 columns)");
            for (int t = 0; t < 2 && rc == SMJ_OK && !planned; t++)
                rc = smj_launch_select_pairs(c, d_t[t], n[t], cc[t], sel_col[t], sel_val[t], 0, key[t], 0, ping[t], pong[t],
                                             (u64 *)(scr + off_sel + (t ? stiles[0] * 8 : 0)), &h->counter[t], h->hist[t], &h->count[t]);
            if (rc != SMJ_OK) break;
            // sort: 4 onesweep passes per table over the device-resident survivor counts
            {   // both tables in the same four launches
                const u64 *dn[2] = {&h->count[0], &h->count[1]};
                const u32 nm[2] = {(u32)n[0], (u32)n[1]};
                const u32 *hs[2] = {h->hist[0], h->hist[1]};
                u32 *sc[2] = {(u32 *)(scr + off_radix), (u32 *)(scr + off_radix + rb[0])};
                const SmjSortPlan *pl[2] = {&h->plan[0], &h->plan[1]};
                rc = smj_radix_sort_pairs_n(c, 2, ping, pong, dn, nm, hs, sc, planned ? pl : nullptr);
                c->run_planned = planned;
            }
            if (rc != SMJ_OK) break;
            // the event pair around the radix passes doubles as the select | sort | join boundaries: every event-record
            // node inside the graph costs ~3.5 us of stream time (0.489 vs 0.475 ms per step with four / no records)
            c->stage_from_pass = c->pass_count > 0;
            if (!c->stage_from_pass && smj_stage_events()) {
                PIPE_CUDA(smj_event_record(c->ev[E_SELECT], c->stream));
                PIPE_CUDA(smj_event_record(c->ev[E_SORT], c->stream));
            }
            // join: co-rank, count, scan, compact the matches; then materialise rows straight from the input tables
            const JoinScratch jsr = join_scratch(scr + off_join, jt);
            PIPE_TRY(smj_launch_join_match(c, ping[0], ping[1], &h->count[0], (u32)n[0], (u32)n[1], SMJ_JOIN_ZIP, jsr.part, jsr.tile_count,
                                           jsr.tile_off, mm, md, &h->jcount));
            PIPE_TRY(smj_launch_join_materialize(c, md, &h->jcount, j_max, d_t[0], cc[0], d_t[1], cc[1], key[1], nullptr, c->d_out_ptr,
                                                 planned ? h->use_store : nullptr, R->store[0], R->store[1], d_tab_ind));
#undef PIPE_TRY
#undef PIPE_CUDA
        } while (0);
        c->capturing = false;
        if (capture) {
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            if (rc == SMJ_OK && e != cudaSuccess) rc = smj_cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
            if (rc == SMJ_OK) {
                e = cudaGraphInstantiate(&c->graph_exec, g, 0);
                if (e != cudaSuccess) { c->graph_exec = nullptr; rc = smj_cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
            }
            if (g) cudaGraphDestroy(g);
            c->graph_launches = c->launches - l0;
            c->launches = l0;   // counted again when the graph is launched below
        }
        if (rc != SMJ_OK) { cudaGetLastError(); return rc; }
    }
    if (replay || capture) {
        CUDA_TRY(cudaGraphLaunch(c->graph_exec, c->stream));
        c->launches += c->graph_launches;
        c->pass_count = c->stage_from_pass ? 1 : 0;   // the timed sort group (both tables, four launches) is a pair of event-record nodes of the graph
    }
    CUDA_TRY(cudaEventRecord(c->ev[E_JOIN], c->stream));
    // counts and sort plans for the host, read after the one wait in smj_run_finish
    ScratchHeader *hh = (ScratchHeader *)c->h_pinned;
    CUDA_TRY(cudaMemcpyAsync(hh, h, sizeof(ScratchHeader), cudaMemcpyDeviceToHost, c->stream));
    return SMJ_OK;
}

int smj_run_finish(SmjCtx *c, SmjRun *R, smj_table_t *out, smj_stats_t *stats)
{
    CUDA_TRY(cudaSetDevice(c->device));
    // every early return below gives the buffer back; handing it to the caller (or freeing it) disarms the guard
    struct OutGuard { smj_table_t *t; ~OutGuard() { if (t && t->data) smj_table_free(t); } } out_guard = {&R->dev_out};
    R->prepared = false;
    const smj_config_t *cfg = &R->cfg;
    const int64_t n[2] = {R->tb[0].rows, R->tb[1].rows};
    const int cc[2] = {R->tb[0].cols, R->tb[1].cols};
    const int c_out = R->c_out;
    smj_table_t &dev_out = R->dev_out;

    // ---- the one host wait: counts and the device-side consistency flag
    ScratchHeader *hh = (ScratchHeader *)c->h_pinned;
    SMJ_TRY(smj_check_device_flag(c));
    // m: pairs that were sorted and joined; m_sel: rows that passed the predicate (more, when the semi-join filter ran)
    const int64_t m[2] = {(int64_t)hh->count[0], (int64_t)hh->count[1]};
    const int64_t m_sel[2] = {c->run_planned ? (int64_t)hh->sel_count[0] : m[0], c->run_planned ? (int64_t)hh->sel_count[1] : m[1]};
    const int64_t j = (int64_t)hh->jcount;
    dev_out.rows = j;
    if (j == 0) smj_table_free(&dev_out), dev_out.data = nullptr;

    // ---- GPU -> CPU (app.c timer 2)
    if (out->on_device) {
        *out = dev_out;
        out_guard.t = nullptr;   // the caller owns the buffer now
    } else {
        SMJ_TRY(smj_alloc_out(c, out, j, c_out));
        if (j) CUDA_TRY(cudaMemcpyAsync(out->data, dev_out.data, (size_t)j * c_out * 4, cudaMemcpyDeviceToHost, c->stream));
        if (dev_out.data) smj_table_free(&dev_out);
    }
    CUDA_TRY(cudaEventRecord(c->ev[E_D2H], c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (cfg->debug) {   // app.c:294-305, 379-400, 694-717 print one line per DPU and stage; one GPU here
        printf("==================\n#    select.cu   #\n==================\n");
        for (int t = 0; t < 2; t++) printf("Table %d : select %lld rows\n", t, (long long)m_sel[t]);
        printf("####################\n\n==================\n#     sort.cu    #\n==================\n");
        for (int t = 0; t < 2; t++) printf("Table %d - GPU %d sort %lld rows\n", t, c->device, (long long)m[t]);
        printf("####################\n\n==================\n#     join.cu    #\n==================\n");
        printf("Rows: %lld\nCOL NUM 1: %d / COL NUM 2: %d\n", (long long)j, cc[0], cc[1]);
        printf("GPU %d results: %lld rows\n####################\n\n", c->device, (long long)j);
    }

    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->h2d_ms = ev_ms(c->ev[E_START], c->ev[E_H2D]);
        if (smj_stage_events()) {   // select = through the histogram scan; sort = the radix pass launches
            cudaEvent_t es = c->stage_from_pass ? c->pass_ev[0] : c->ev[E_SELECT], eo = c->stage_from_pass ? c->pass_ev[1] : c->ev[E_SORT];
            stats->select_ms = ev_ms(c->ev[E_H2D], es);
            stats->sort_ms = ev_ms(es, eo);
            stats->join_ms = ev_ms(eo, c->ev[E_JOIN]);
        }
        stats->d2h_ms = ev_ms(c->ev[E_JOIN], c->ev[E_D2H]);
        stats->total_device_ms = ev_ms(c->ev[E_H2D], c->ev[E_JOIN]);
        for (int t = 0; t < 2; t++) { stats->rows_in[t] = n[t]; stats->rows_selected[t] = m_sel[t]; }
        stats->rows_joined = j;
        stats->bytes_model = bytes_model(n, cc, m_sel, j);
        stats->kernel_launches = c->launches - R->launches0;
        stats->graph_replayed = R->replayed ? 1 : 0;
        double sum = 0;
        if (!smj_stage_events()) c->pass_count = 0;
        for (int p = 0; p < c->pass_count; p++) sum += ev_ms(c->pass_ev[2 * p], c->pass_ev[2 * p + 1]);
        // each timed group is four pass launches (both tables in each); with a sort plan only the first
        // max(npass) of them have tiles, and table t takes part in npass[t] of those
        u32 np[2] = {SMJ_KEY_PASSES, SMJ_KEY_PASSES};
        if (c->run_planned) { np[0] = hh->plan[0].npass; np[1] = hh->plan[1].npass; }
        const u32 np_max = np[0] > np[1] ? np[0] : np[1];
        stats->sort_passes = c->pass_count * (int)np_max;
        stats->sort_pass_ms_avg = stats->sort_passes ? sum / stats->sort_passes : 0;
        stats->sort_pass_bytes_avg = np_max ? 16.0 * ((double)m[0] * np[0] + (double)m[1] * np[1]) / np_max : 0;
        // the model's sort term (16 B x 4 passes x selected rows) restated for the passes and pairs this run's plan sorted
        stats->bytes_planned = stats->bytes_model + 16.0 * ((double)m[0] * np[0] + (double)m[1] * np[1]) -
                               16.0 * SMJ_KEY_PASSES * ((double)m_sel[0] + (double)m_sel[1]);
    }
    return SMJ_OK;
}

int smj_run_multi(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats);
int smj_run_multi_local(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out, smj_stats_t *stats);
bool smj_dist_active(void);

extern "C" int smj_run(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2, smj_table_t *out,
                       smj_stats_t *stats)
{
    smj_config_t local;
    if (!cfg) { smj_config_default(&local); cfg = &local; }
    if (!g_inited || (cfg->nr_gpus > g_nctx && !smj_dist_active())) SMJ_TRY(smj_init(cfg));
    SMJ_TRY(check_table(t1, "smj_run(t1)"));
    SMJ_TRY(check_table(t2, "smj_run(t2)"));
    if (!out) return smj_set_error(SMJ_EINVAL, "smj_run: null out");
    if (smj_dist_active()) return smj_run_multi(cfg, t1, t2, out, stats);         // one process per GPU (smj_init_dist)
    if (cfg->nr_gpus > 1) return smj_run_multi_local(cfg, t1, t2, out, stats);     // this process drives nr_gpus devices
    return smj_run_single(g_ctx[0], cfg, t1, t2, out, stats);
}
