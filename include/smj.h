/*
 * include/smj.h -- C-ABI of the B200 sort-merge-join engine (libsmj.so).
 *
 * Drop-in boundary for the select -> sort -> merge -> join path of
 * 5eoyeon/pim-sort-merge-join.  The reference has no FFI; its operator surface
 * is the implicit host<->DPU stage ABI of sort-merge-join/app.c (a 12-byte
 * dpu_block_t descriptor, common.h:13-18, plus a row-major table at
 * DPU_MRAM_HEAP_POINTER) driven stage by stage from app.c:main.  Each entry
 * point below replaces one of those stage launches; the citation says which.
 *
 * Conventions
 *   - C99, plain pointers and sizes; no C++/torch types cross the boundary.
 *   - Tables are row-major int32 cells, rows*cols of them (reference: T[rows*cols],
 *     common.h:1-9, every value produced by atoi(), cpu_app.c:71 / app.c:84, so
 *     int32-valued).  data may be a host pointer (on_device = 0) or a CUDA device
 *     pointer (on_device = 1).
 *   - Outputs are allocated by the library; the caller sets out->on_device before
 *     the call to choose where (0: pinned host memory, 1: device memory) and
 *     releases them with smj_table_free().
 *   - Return 0 on success, a negative SMJ_E* code otherwise; the library never
 *     calls exit() (the reference's DPU_ASSERT does, include/dpu/dpu.h:144 --
 *     the C driver host/app.c reproduces that behaviour on top of these codes).
 *   - Calls are synchronous on return, like DPU_SYNCHRONOUS launches
 *     (app.c:247,362,465,662).  One caller thread.
 *   - There is no CPU fallback: without a CUDA device every compute entry point
 *     returns SMJ_ENODEVICE.
 */
#ifndef SMJ_H
#define SMJ_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMJ_VERSION 100

/* error codes */
#define SMJ_OK          0
#define SMJ_EINVAL     -1   /* bad argument (column out of range, null pointer, ...) */
#define SMJ_ENODEVICE  -2   /* no CUDA device / not initialised */
#define SMJ_ECUDA      -3   /* CUDA runtime error, text via smj_last_error() */
#define SMJ_ENOMEM     -4   /* host or device allocation failed */
#define SMJ_ETOOBIG    -5   /* row count exceeds what this build supports (2^30 selected rows per table per GPU) */
#define SMJ_ENCCL      -6   /* NCCL error */
#define SMJ_EINTERNAL  -7   /* device-side consistency check failed */
#define SMJ_EIRREGULAR -8   /* smj_csv_parse: the text needs the sequential host parser to reproduce the reference exactly */
#define SMJ_ERANGE     -9   /* smj_table_from_i64: a cell of the int64 table is not an int32 value */

/* join modes */
#define SMJ_JOIN_ZIP   0    /* == cpu_app.c:204-266 / join.c:153-248: i-th left duplicate pairs with i-th right duplicate */
#define SMJ_JOIN_MANY  1    /* true many-to-many equi-join (extension, order (key, left row, right row)); results that would
                               not fit in device memory are refused with SMJ_ETOOBIG -- use smj_join_count for those */

/* row-major int32 table; replaces {dpu_block_t bl; T rows[]} (common.h:13-18, select.c:16) */
typedef struct {
    int32_t *data;
    int64_t  rows;
    int32_t  cols;
    int32_t  on_device;
} smj_table_t;

/* the user.h knobs (user.h:1-13) at run time; NR_DPUS/NR_TASKLETS become nr_gpus */
typedef struct {
    int     nr_gpus;
    int     select_col1, select_col2;     /* SELECT_COL1 / SELECT_COL2 */
    int64_t select_val1, select_val2;     /* SELECT_VAL1 / SELECT_VAL2 ; predicate: cell > val (cpu_app.c:88) */
    int     join_key1, join_key2;         /* JOIN_KEY1 / JOIN_KEY2 */
    int     join_mode;                    /* SMJ_JOIN_ZIP (default) | SMJ_JOIN_MANY */
    int     debug;                        /* DEBUG: per-stage row counts on stdout (app.c:294-305,379-400,549-577,694-717) */
} smj_config_t;

/* replaces the Timer slots of app.c (timer 0 = CPU->DPU, 1 = DPU, 2 = DPU->CPU; app.c:221,246,251) */
typedef struct {
    double  h2d_ms, select_ms, sort_ms, exchange_ms, merge_ms, join_ms, d2h_ms;
    double  total_device_ms;              /* select+sort+exchange+merge+join, CUDA events on the library stream */
    int64_t rows_in[2], rows_selected[2], rows_joined;
    double  bytes_model;                  /* algorithmic HBM bytes of this run */
    double  bytes_nvlink;                 /* bytes this process sent over NVLink in the exchange */
    int64_t kernel_launches;              /* kernels this call launched */
    /* average device time of the radix-sort scatter pass (the dominant kernel) and its launch count */
    double  sort_pass_ms_avg; int32_t sort_passes;
    int32_t graph_replayed;               /* 1: the device pipeline of this call was one CUDA-graph launch (a repeat of the previous call) */
    /* algorithmic bytes (16 B per pair the launch sorts) of an average executed pass launch: the sort runs
     * ceil(bits(max key - min key) / 8) passes per table, decided on the device */
    double  sort_pass_bytes_avg;
    /* bytes_model restated for the work this run's device plan did: the radix passes the key range needed, over the
     * pairs that survived the semi-join filter (bytes_model itself is the fixed formula: four passes, every selected row) */
    double  bytes_planned;
} smj_stats_t;

/* fills *cfg from the user.h macros this library was compiled with (include/user.h) */
void smj_config_default(smj_config_t *cfg);

/* replaces dpu_alloc + dpu_load (app.c:175-176,315-316,422-423,638-639): binds cfg->nr_gpus devices
 * (device 0.. in this process), creates streams and the workspace arena.  cfg may be NULL (defaults). */
int  smj_init(const smj_config_t *cfg);
/* replaces dpu_free (app.c:307,...) */
void smj_shutdown(void);

/* one-process-per-GPU mode (torchrun): this process drives `local_device` as rank `rank` of `world`.
 * nccl_id is the 128-byte ncclUniqueId produced by smj_dist_unique_id() on rank 0 and broadcast by the host. */
int  smj_dist_unique_id(void *nccl_id_128);
int  smj_init_dist(const smj_config_t *cfg, int rank, int world, int local_device, const void *nccl_id_128);

/* Host-only planning steps of the key-range exchange (no GPU needed; the CPU tests drive them under gloo):
 * replace the host range split of app.c:589-633.  Keys are in the engine's order-preserving unsigned form
 * (uint32_t)key ^ 0x80000000u; a sample equal to 0xffffffff means "none" (empty table on that rank).
 * smj_plan_splitters: world-1 splitters = the b/world quantiles of the gathered samples; rank b owns keys in
 *   [splitter[b-1], splitter[b]) (splitter[-1] = 0, splitter[world-1] = +inf), so equal keys share a rank.
 * smj_plan_exchange: counts[src*world+dst] = rows src sends to dst; recv_offsets[src] = row offset of src's run in
 *   rank me's receive buffer (runs in source-rank order), *recv_total = rows me receives. */
int  smj_plan_splitters(const uint32_t *samples, int64_t n_samples, int world, uint32_t *splitters);
int  smj_plan_exchange(const int64_t *counts, int world, int me, int64_t *recv_offsets, int64_t *recv_total);
/* The default (peer-memory) exchange's plan, host twin of the device step: row0[b] = first row of rank me's bucket b inside
 * rank b's receive buffer; *rows_mine = rows me will hold; *verdict = 1 when some rank's share exceeds cap_rows (the
 * receive capacity every rank allocated) -- then no rank stores anything, *rows_mine = 0, and the buffers are re-sized
 * for *need_rows (the largest share) before the step is run again.  Identical on every rank by construction. */
int  smj_plan_fabric(const int64_t *counts, int world, int me, int64_t cap_rows, int64_t *row0, int64_t *rows_mine,
                     int *verdict, int64_t *need_rows);

/* == select.c:63-194 (DPU select) / cpu_app.c:81-112: rows with in[row][col] > val, order preserved */
int  smj_select(const smj_table_t *in, int col, int64_t val, smj_table_t *out);
/* == sort_dpu.c:189-328 / cpu_app.c:172-202: stable ascending sort of whole rows by column key_col, in place */
int  smj_sort(smj_table_t *inout, int key_col);
/* == merge_dpu.c:55-223 + app.c:413-547: merge of two runs sorted by key_col; rows of a precede rows of b on ties */
int  smj_merge(const smj_table_t *a, const smj_table_t *b, int key_col, smj_table_t *out);
/* == join.c:58-266 / cpu_app.c:204-266: merge join of two tables sorted by key1/key2;
 * output row = all columns of l, then the columns of r except key2 */
int  smj_join(const smj_table_t *l, const smj_table_t *r, int key1, int key2, int mode, smj_table_t *out);
/* == app.c:main stages select..join (app.c:221-688) / cpu_app.c:336-344: the whole pipeline.
 * In smj_init_dist mode t1/t2 are THIS rank's contiguous row blocks (rank order = row order) and
 * out is this rank's key-range shard of the result (shards concatenated in rank order = full result). */
int  smj_run(const smj_config_t *cfg, const smj_table_t *t1, const smj_table_t *t2,
             smj_table_t *out, smj_stats_t *stats);
/* join-count only (no materialisation): rows smj_join(...) would produce in `mode` */
int  smj_join_count(const smj_table_t *l, const smj_table_t *r, int key1, int key2, int mode, int64_t *rows);

void        smj_table_free(smj_table_t *t);
const char *smj_strerror(int code);
const char *smj_last_error(void);        /* detail text of the last failure on this thread */

/* ---- memory helpers so hosts without a CUDA binding can stage buffers ---- */
int  smj_host_alloc(void **p, size_t bytes);      /* pinned host memory */
void smj_host_free(void *p);
int  smj_device_alloc(void **p, size_t bytes);    /* on the current library device */
void smj_device_free(void *p);
int  smj_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int  smj_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes);
int  smj_device_sync(void);

/* ---- the reference's in-memory cell type at the boundary ----
 * The reference keeps its tables as T[rows*cols] with T = int64_t (common.h:1-9; its UINT64 / DOUBLE branches are never
 * selected) holding atoi() results (cpu_app.c:71, app.c:84), i.e. int32 values in 8-byte cells.  A host that keeps such
 * T* arrays (app.c:158-159 test_array1/2) passes them as they are:
 * smj_table_from_i64: `cells` (host pointer, or device pointer when cells_on_device) -> *out, a library-owned int32 table
 *   (out->on_device chosen by the caller before the call, as for every output; smj_table_free).  Narrowed on the GPU; a
 *   cell outside [INT32_MIN, INT32_MAX] is refused with SMJ_ERANGE (the message names it), never truncated.
 * smj_table_to_i64: the int32 table t (host or device) widened into the caller's rows*cols int64 buffer. */
int  smj_table_from_i64(const int64_t *cells, int64_t rows, int32_t cols, int cells_on_device, smj_table_t *out);
int  smj_table_to_i64(const smj_table_t *t, int64_t *cells, int cells_on_device);

/* ---- CSV text <-> tables on the GPU ----
 * smj_csv_parse == set_csv_size + load_csv (cpu_app.c:15-79 == app.c:28-92) for regular files: `text` is the whole file
 * in host memory; *out becomes a library-owned DEVICE table (smj_table_free).  cols = tokens of the first line,
 * rows = lines - 1 (rows = -1 for an empty file, as the reference computes), cells = atoi(token) with glibc's
 * semantics.  Returns SMJ_EIRREGULAR when only a sequential pass reproduces the reference: a line of 1023+ characters
 * (fgets(line, 1024) splits it), a row whose token count differs from the header's, an embedded NUL -- the caller then
 * uses its sequential parser (host/csv.c does).
 * smj_csv_format == save_to_csv (cpu_app.c:268-301 == app.c:720-755): header col1..colN, "%ld" cells, ',' separators,
 * '\n' line ends; t may be a host or a device table; *text is pinned host memory (smj_host_free), *bytes its length. */
int  smj_csv_parse(const char *text, size_t bytes, smj_table_t *out);
int  smj_csv_format(const smj_table_t *t, char **text, size_t *bytes);

/* ---- deterministic synthetic tables (replaces the unseeded data/generate_data.py:4-26) ----
 * kind 0: column key_col = a seeded bijection of the row index into [1, 3*total_rows] (unique keys, as
 *         generate_data.py:9 draws them), other columns uniform in [1, 3*total_rows).
 * kind 1: key column uniform in [1, key_domain] (duplicates), other columns as kind 0.
 * kind 2: key column Zipf(1.1) over [1, key_domain] (key_domain 0: 2^20; at most 2^26): key k with probability
 *         proportional to k^-1.1, i.e. key 1 holds ~12 % of the rows of a 2^20-key domain (BASELINE config 3's "heavy
 *         duplicates"); drawn by inverse CDF from the 64-bit threshold table smj_synth_zipf_cdf() builds.
 * Rows [row0, row0+rows) of the virtual table of total_rows rows are written to dev_out (device pointer).
 * pim-sort-merge-join_b200/datagen.py computes the same cells with numpy (tests check they agree). */
int  smj_synth_table(int32_t *dev_out, int64_t row0, int64_t rows, int64_t total_rows, int cols,
                     int key_col, uint64_t seed, int kind, int64_t key_domain);
/* host-only: cdf_out[k-1] = floor(2^64 * P(Zipf(s) rank <= k)) for k = 1..key_domain (the table kind 2 searches) */
int  smj_synth_zipf_cdf(int64_t key_domain, double s, uint64_t *cdf_out);

/* number of CUDA kernels launched by this library since smj_init (claim for bench.py "gpu_launches") */
int64_t smj_kernel_launches(void);
int     smj_device_count(void);
int     smj_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SMJ_H */
