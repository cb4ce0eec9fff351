/* host/app.c -- the C host driver of the B200 sort-merge-join engine.
 *
 * Drop-in for the reference's PIM host program sort-merge-join/app.c (main: app.c:123-775):
 *     ./app data1.csv data2.csv            -> ./data/result.csv + timing banner on stdout
 * Same operator surface: two CSV paths on the command line (app.c:130-131), the user.h knobs
 * (include/user.h, same macro names; NR_DPUS/NR_TASKLETS replaced by NR_GPUS), result written to
 * ./data/result.csv relative to the working directory (app.c:720), and the four-line EXEC TIME banner
 * (app.c:763-772) with CPU-DPU / DPU / DPU-CPU renamed CPU-GPU / GPU / GPU-CPU.  Where the reference spends
 * ~650 lines partitioning rows over DPUs, launching four DPU programs and merging through host memory, this
 * driver makes ONE call into libsmj.so (smj_run) -- all stage logic lives behind the C-ABI of include/smj.h.
 *
 * Extras that do not disturb the surface (a maintainer's scripts keep working without them):
 *   - run-time knob overrides from the environment: SMJ_NR_GPUS, SMJ_SELECT_COL1/2, SMJ_SELECT_VAL1/2,
 *     SMJ_JOIN_KEY1/2, SMJ_DEBUG, SMJ_RESULT (output path), SMJ_JSON=1 (one JSON line with the stage times);
 *   - CSV parse / emit are timed separately from the device pipeline (BASELINE.json: "CSV parse/emit timed
 *     separately"), printed after the banner; both run on the GPU (smj_csv_parse / smj_csv_format) unless SMJ_CSV=host
 *     or the text is irregular (host/csv.c then reproduces the reference's sequential parser).
 * Errors behave like the reference: a file that cannot be opened -> perror + exit(EXIT_FAILURE) (app.c:31-35),
 * a library failure -> message + exit(EXIT_FAILURE) (DPU_ASSERT, include/dpu/dpu.h:144).
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/smj.h"
#include "csv.h"

#define SMJ_ASSERT(call)                                                                        \
    do {                                                                                        \
        int rc_ = (call);                                                                       \
        if (rc_ != SMJ_OK) {                                                                    \
            fprintf(stderr, "%s:%d: %s failed: %s (%s)\n", __FILE__, __LINE__, #call,           \
                    smj_strerror(rc_), smj_last_error());                                       \
            exit(EXIT_FAILURE);                                                                 \
        }                                                                                       \
    } while (0)

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void *pinned_alloc(uint64_t bytes)
{
    void *p = NULL;
    if (smj_host_alloc(&p, (size_t)bytes) != SMJ_OK) return NULL;
    return p;
}

static void env_int(const char *name, int *v)
{
    const char *s = getenv(name);
    if (s && *s) *v = atoi(s);
}
static void env_i64(const char *name, int64_t *v)
{
    const char *s = getenv(name);
    if (s && *s) *v = strtoll(s, NULL, 10);
}

/* CSV -> table.  The file is read whole into pinned memory and parsed ON THE GPU (smj_csv_parse); text only a
 * sequential pass can interpret the reference's way (a 1023+ character line, ragged rows) goes through host/csv.c.
 * SMJ_CSV=host forces the host parser and writer. */
static int g_csv_on_gpu = 1;

static void load_or_die(const char *path, smj_table_t *t)
{
    if (g_csv_on_gpu) {
        FILE *f = fopen(path, "rb");
        if (!f) { perror("Failed to open file"); exit(EXIT_FAILURE); }   /* the reference's message (app.c:33) */
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        /* plain memory: the text crosses PCIe once, and pinning a buffer that is used once costs more than it saves */
        char *buf = (char *)malloc((size_t)sz + 1);
        if (!buf) { fprintf(stderr, "out of memory for %s\n", path); exit(EXIT_FAILURE); }
        size_t got = fread(buf, 1, (size_t)sz, f);
        fclose(f);
        int rc = smj_csv_parse(buf, got, t);
        free(buf);
        if (rc == SMJ_OK) { if (t->rows < 0) t->rows = 0; return; }
        if (rc != SMJ_EIRREGULAR) {
            fprintf(stderr, "smj_csv_parse(%s) failed: %s (%s)\n", path, smj_strerror(rc), smj_last_error());
            exit(EXIT_FAILURE);
        }
    }
    int cols = 0;
    int64_t rows = 0;
    int32_t *data = NULL;
    if (csv_load(path, &data, &rows, &cols, pinned_alloc) != 0) {
        perror("Failed to open file");
        exit(EXIT_FAILURE);
    }
    if (rows < 0) rows = 0;
    t->data = data; t->rows = rows; t->cols = cols; t->on_device = 0;
}

static void release_input(smj_table_t *t)
{
    if (t->on_device) smj_table_free(t);
    else smj_host_free(t->data);
}

int main(int argc, char *argv[])
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s data1.csv data2.csv\n", argv[0]);
        return EXIT_FAILURE;
    }
    smj_config_t cfg;
    smj_config_default(&cfg);
    env_int("SMJ_NR_GPUS", &cfg.nr_gpus);
    env_int("SMJ_SELECT_COL1", &cfg.select_col1); env_i64("SMJ_SELECT_VAL1", &cfg.select_val1);
    env_int("SMJ_SELECT_COL2", &cfg.select_col2); env_i64("SMJ_SELECT_VAL2", &cfg.select_val2);
    env_int("SMJ_JOIN_KEY1", &cfg.join_key1);     env_int("SMJ_JOIN_KEY2", &cfg.join_key2);
    env_int("SMJ_DEBUG", &cfg.debug);
    const char *result_path = getenv("SMJ_RESULT");
    if (!result_path || !*result_path) result_path = "./data/result.csv";
    const char *csv_mode = getenv("SMJ_CSV");
    if (csv_mode && strcmp(csv_mode, "host") == 0) g_csv_on_gpu = 0;

    /* replaces dpu_alloc + dpu_load; done before the CSV load so the tables land in pinned memory */
    SMJ_ASSERT(smj_init(&cfg));

    double t0 = now_ms();
    smj_table_t t1, t2, out = {NULL, 0, 0, 0};
    out.on_device = g_csv_on_gpu;   /* the result stays in HBM when the GPU also formats it */
    load_or_die(argv[1], &t1);
    load_or_die(argv[2], &t2);
    double parse_ms = now_ms() - t0;

    smj_stats_t st;
    SMJ_ASSERT(smj_run(&cfg, &t1, &t2, &out, &st));

    t0 = now_ms();
    if (!out.cols) out.cols = t1.cols + t2.cols - 1;
    if (g_csv_on_gpu) {
        char *text = NULL;
        size_t bytes = 0;
        SMJ_ASSERT(smj_csv_format(&out, &text, &bytes));
        FILE *f = fopen(result_path, "wb");
        if (!f || fwrite(text, 1, bytes, f) != bytes || fclose(f) != 0) {
            perror("Failed to open file");
            exit(EXIT_FAILURE);
        }
        smj_host_free(text);
    } else if (csv_save(result_path, out.data, out.rows, out.cols) != 0) {
        perror("Failed to open file");
        exit(EXIT_FAILURE);
    }
    double emit_ms = now_ms() - t0;

    printf("\n");
    printf("######### GPU #########\n");
    printf("### SORT-MERGE-JOIN ###\n");
    printf("         EXEC TIME     \n");
    printf("CPU-GPU  %f\n", st.h2d_ms);
    printf("GPU      %f\n", st.total_device_ms);
    printf("GPU-CPU  %f\n", st.d2h_ms);
    printf("-----------------------\n");
    printf("TOTAL %f\n", st.h2d_ms + st.total_device_ms + st.d2h_ms);
    printf("#######################\n\n");
    printf("CSV-PARSE %f\nCSV-EMIT  %f\n", parse_ms, emit_ms);
    printf("ROWS %lld x %lld -> selected %lld / %lld -> joined %lld\n", (long long)t1.rows, (long long)t2.rows,
           (long long)st.rows_selected[0], (long long)st.rows_selected[1], (long long)st.rows_joined);

    const char *js = getenv("SMJ_JSON");
    if (js && *js == '1')
        printf("{\"rows\": [%lld, %lld], \"selected\": [%lld, %lld], \"joined\": %lld, \"nr_gpus\": %d, "
               "\"parse_ms\": %.3f, \"h2d_ms\": %.3f, \"select_ms\": %.3f, \"sort_ms\": %.3f, \"exchange_ms\": %.3f, "
               "\"merge_ms\": %.3f, \"join_ms\": %.3f, \"device_ms\": %.3f, \"d2h_ms\": %.3f, \"emit_ms\": %.3f, "
               "\"bytes_model\": %.0f, \"model_gbs\": %.1f, \"kernel_launches\": %lld}\n",
               (long long)t1.rows, (long long)t2.rows, (long long)st.rows_selected[0], (long long)st.rows_selected[1],
               (long long)st.rows_joined, cfg.nr_gpus, parse_ms, st.h2d_ms, st.select_ms, st.sort_ms, st.exchange_ms,
               st.merge_ms, st.join_ms, st.total_device_ms, st.d2h_ms, emit_ms, st.bytes_model,
               st.total_device_ms > 0 ? st.bytes_model / (st.total_device_ms * 1e-3) / 1e9 : 0.0,
               (long long)st.kernel_launches);

    smj_table_free(&out);
    release_input(&t1);
    release_input(&t2);
    smj_shutdown();   /* replaces dpu_free (app.c:756) */
    return 0;
}
