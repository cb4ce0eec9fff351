/*
 * oracle/smj_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference CPU pipeline (sort-merge-join/cpu_app.c)
 * with the same observable results but
 *   - int32 cells (every reference cell is atoi()'d, cpu_app.c:71, so it is
 *     int32-valued even though T = int64_t, common.h:1-9),
 *   - 64-bit sizes (the reference's `int` byte counts overflow, cpu_app.c:49),
 *   - an O(n log n) stable merge sort in place of the O(n^2) stable insertion
 *     sort (cpu_app.c:172-202); any stable ascending sort yields the same table.
 *
 * PARITY PINNED: tests/test_oracle.py checks every function below against
 * oracle/_ref (the reference's own cpu_app.c compiled by oracle/Makefile) on
 * the bundled data, the KATs of SURVEY.md section 8c and seeded random inputs,
 * and against the committed fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libsmj.so) never does.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ---- select: cpu_app.c:81-112 (strict >, signed, order preserved) ---- */
int64_t oracle_select(const int32_t *in, int64_t rows, int cols, int col, int64_t val, int32_t *out)
{
    int64_t m = 0;
    for (int64_t i = 0; i < rows; i++) {
        if ((int64_t)in[i * cols + col] > val) {
            if (out) memcpy(out + m * cols, in + i * cols, (size_t)cols * sizeof(int32_t));
            m++;
        }
    }
    return m;
}

/* ---- sort: ordering contract of cpu_app.c:172-202 (stable ascending) ---- */
typedef struct { int32_t key; uint32_t pad; int64_t idx; } keyidx_t;

static void msort(keyidx_t *a, keyidx_t *tmp, int64_t n)
{
    /* bottom-up stable merge sort; ties keep the left run's element first */
    for (int64_t i = 0; i + 1 < n; i += 2)
        if (a[i].key > a[i + 1].key) { keyidx_t t = a[i]; a[i] = a[i + 1]; a[i + 1] = t; }
    keyidx_t *src = a, *dst = tmp;
    for (int64_t w = 2; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) dst[k++] = (src[j].key < src[i].key) ? src[j++] : src[i++];
            while (i < mid) dst[k++] = src[i++];
            while (j < hi) dst[k++] = src[j++];
        }
        keyidx_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, (size_t)n * sizeof(keyidx_t));
}

void oracle_sort(int32_t *inout, int64_t rows, int cols, int key)
{
    if (rows < 2) return;
    keyidx_t *ki = (keyidx_t *)malloc((size_t)rows * sizeof(keyidx_t));
    keyidx_t *tmp = (keyidx_t *)malloc((size_t)rows * sizeof(keyidx_t));
    for (int64_t i = 0; i < rows; i++) { ki[i].key = inout[i * cols + key]; ki[i].pad = 0; ki[i].idx = i; }
    msort(ki, tmp, rows);
    int32_t *sorted = (int32_t *)malloc((size_t)rows * cols * sizeof(int32_t));
    for (int64_t i = 0; i < rows; i++)
        memcpy(sorted + i * cols, inout + ki[i].idx * cols, (size_t)cols * sizeof(int32_t));
    memcpy(inout, sorted, (size_t)rows * cols * sizeof(int32_t));
    free(sorted); free(tmp); free(ki);
}

/* ---- merge of two sorted runs of one table, a before b on ties.
 * No CPU counterpart in cpu_app.c (it sorts whole tables); this is the
 * contract merge_dpu.c:55-223 + app.c:413-547 must satisfy for the merged
 * table to equal ssort(a ++ b). ---- */
void oracle_merge(const int32_t *a, int64_t ra, const int32_t *b, int64_t rb, int cols, int key, int32_t *out)
{
    int64_t i = 0, j = 0, k = 0;
    size_t rb_ = (size_t)cols * sizeof(int32_t);
    while (i < ra && j < rb) {
        if (b[j * cols + key] < a[i * cols + key]) memcpy(out + (k++) * cols, b + (j++) * cols, rb_);
        else memcpy(out + (k++) * cols, a + (i++) * cols, rb_);
    }
    while (i < ra) memcpy(out + (k++) * cols, a + (i++) * cols, rb_);
    while (j < rb) memcpy(out + (k++) * cols, b + (j++) * cols, rb_);
}

/* ---- join, zip mode: cpu_app.c:204-266.  out==NULL -> count pass only
 * (cpu_app.c:211-227); otherwise emit pass (cpu_app.c:236-265): all left
 * columns, then right columns except key2. ---- */
int64_t oracle_join_zip(const int32_t *l, int64_t r1, int c1, const int32_t *r, int64_t r2, int c2,
                        int key1, int key2, int32_t *out)
{
    int tc = c1 + c2 - 1;
    int64_t i = 0, j = 0, n = 0;
    while (i < r1 && j < r2) {
        int32_t a = l[i * c1 + key1], b = r[j * c2 + key2];
        if (a == b) {
            if (out) {
                int32_t *o = out + n * tc;
                for (int k = 0; k < c1; k++) o[k] = l[i * c1 + k];
                for (int k = 0, q = 0; k < c2; k++)
                    if (k != key2) o[c1 + q++] = r[j * c2 + k];
            }
            n++; i++; j++;
        } else if (a < b) i++;
        else j++;
    }
    return n;
}

/* ---- join, many-to-many mode (extension; no reference counterpart).
 * Order (key, left row, right row); same column layout.  out==NULL counts. ---- */
int64_t oracle_join_many(const int32_t *l, int64_t r1, int c1, const int32_t *r, int64_t r2, int c2,
                         int key1, int key2, int32_t *out, int64_t out_cap_rows)
{
    int tc = c1 + c2 - 1;
    int64_t i = 0, j = 0, n = 0;
    while (i < r1 && j < r2) {
        int32_t a = l[i * c1 + key1], b = r[j * c2 + key2];
        if (a < b) { i++; continue; }
        if (a > b) { j++; continue; }
        int64_t je = j;
        while (je < r2 && r[je * c2 + key2] == a) je++;
        for (; i < r1 && l[i * c1 + key1] == a; i++) {
            for (int64_t jj = j; jj < je; jj++) {
                if (out && n < out_cap_rows) {
                    int32_t *o = out + n * tc;
                    for (int k = 0; k < c1; k++) o[k] = l[i * c1 + k];
                    for (int k = 0, q = 0; k < c2; k++)
                        if (k != key2) o[c1 + q++] = r[jj * c2 + k];
                }
                n++;
            }
        }
        j = je;
    }
    return n;
}

/* ---- CSV in: set_csv_size + load_csv, cpu_app.c:15-79.
 * cols = number of ","-tokens of the header; rows = fgets chunks - 1;
 * cells = atoi(token) with 1024-byte line chunks, exactly as the reference
 * (same libc calls, so the same quirks: CRLF, empty fields collapse). ---- */
int oracle_csv_size(const char *path, int *cols, int64_t *rows)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    int first = 1;
    *cols = 0; *rows = 0;
    while (fgets(line, sizeof line, f)) {
        if (first) {
            first = 0;
            for (char *t = strtok(line, ","); t; t = strtok(NULL, ",")) (*cols)++;
        }
        (*rows)++;
    }
    (*rows)--;
    fclose(f);
    return 0;
}

int oracle_load_csv(const char *path, int cols, int64_t rows, int32_t *out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    int64_t row = 0;
    if (!fgets(line, sizeof line, f)) { fclose(f); return 0; }
    while (fgets(line, sizeof line, f) && row < rows) {
        int col = 0;
        for (char *t = strtok(line, ","); t; t = strtok(NULL, ",")) {
            if (col < cols) out[row * cols + col] = (int32_t)atoi(t);
            col++;
        }
        row++;
    }
    fclose(f);
    return 0;
}

/* ---- CSV out: save_to_csv, cpu_app.c:268-301 (header col1..colN, %ld, LF). ---- */
int oracle_save_csv(const char *path, int cols, int64_t rows, const int32_t *a)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    for (int i = 1; i <= cols; i++) fprintf(f, i < cols ? "col%d," : "col%d", i);
    fprintf(f, "\n");
    for (int64_t i = 0; i < rows; i++) {
        for (int j = 0; j < cols; j++) fprintf(f, j < cols - 1 ? "%ld," : "%ld", (long)a[i * cols + j]);
        fprintf(f, "\n");
    }
    fclose(f);
    return 0;
}

/* ---- whole pipeline in memory, order of cpu_app.c:336-344.
 * t1/t2 are not modified.  *out is malloc'd (free with oracle_free).
 * stage_ms[3] = select, sort, join.  Returns joined rows. ---- */
int64_t oracle_run(const int32_t *t1, int64_t r1, int c1, const int32_t *t2, int64_t r2, int c2,
                   int sel_col1, int64_t sel_val1, int sel_col2, int64_t sel_val2,
                   int key1, int key2, int mode, int32_t **out, int64_t *selected, double *stage_ms)
{
    double t0 = now_ms();
    int64_t m1 = oracle_select(t1, r1, c1, sel_col1, sel_val1, NULL);
    int64_t m2 = oracle_select(t2, r2, c2, sel_col2, sel_val2, NULL);
    int32_t *a = (int32_t *)malloc((size_t)(m1 ? m1 : 1) * c1 * sizeof(int32_t));
    int32_t *b = (int32_t *)malloc((size_t)(m2 ? m2 : 1) * c2 * sizeof(int32_t));
    oracle_select(t1, r1, c1, sel_col1, sel_val1, a);
    oracle_select(t2, r2, c2, sel_col2, sel_val2, b);
    double t1_ = now_ms();
    oracle_sort(a, m1, c1, key1);
    oracle_sort(b, m2, c2, key2);
    double t2_ = now_ms();
    int64_t j;
    int tc = c1 + c2 - 1;
    if (mode == 0) {
        j = oracle_join_zip(a, m1, c1, b, m2, c2, key1, key2, NULL);
        *out = (int32_t *)malloc((size_t)(j ? j : 1) * tc * sizeof(int32_t));
        oracle_join_zip(a, m1, c1, b, m2, c2, key1, key2, *out);
    } else {
        j = oracle_join_many(a, m1, c1, b, m2, c2, key1, key2, NULL, 0);
        *out = (int32_t *)malloc((size_t)(j ? j : 1) * tc * sizeof(int32_t));
        oracle_join_many(a, m1, c1, b, m2, c2, key1, key2, *out, j);
    }
    double t3_ = now_ms();
    if (selected) { selected[0] = m1; selected[1] = m2; }
    if (stage_ms) { stage_ms[0] = t1_ - t0; stage_ms[1] = t2_ - t1_; stage_ms[2] = t3_ - t2_; }
    free(a); free(b);
    return j;
}

void oracle_free(void *p) { free(p); }

#ifdef SMJ_ORACLE_MAIN
/* smj_oracle data1.csv data2.csv out.csv [sc1 sv1 sc2 sv2 k1 k2] */
int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s d1.csv d2.csv out.csv [sc1 sv1 sc2 sv2 k1 k2]\n", argv[0]); return 2; }
    int sc1 = 0, sc2 = 0, k1 = 0, k2 = 0; int64_t sv1 = 5000, sv2 = 5000; /* user.h:6-13 defaults */
    if (argc >= 10) { sc1 = atoi(argv[4]); sv1 = atoll(argv[5]); sc2 = atoi(argv[6]); sv2 = atoll(argv[7]); k1 = atoi(argv[8]); k2 = atoi(argv[9]); }
    int c1, c2; int64_t r1, r2;
    if (oracle_csv_size(argv[1], &c1, &r1) || oracle_csv_size(argv[2], &c2, &r2)) { perror("Failed to open file"); return 1; }
    int32_t *a = (int32_t *)calloc((size_t)(r1 > 0 ? r1 : 1) * c1, 4), *b = (int32_t *)calloc((size_t)(r2 > 0 ? r2 : 1) * c2, 4);
    oracle_load_csv(argv[1], c1, r1, a); oracle_load_csv(argv[2], c2, r2, b);
    int32_t *out; int64_t sel[2]; double ms[3];
    int64_t j = oracle_run(a, r1, c1, b, r2, c2, sc1, sv1, sc2, sv2, k1, k2, 0, &out, sel, ms);
    oracle_save_csv(argv[3], c1 + c2 - 1, j, out);
    printf("{\"selected\": [%ld, %ld], \"joined\": %ld, \"select_ms\": %.3f, \"sort_ms\": %.3f, \"join_ms\": %.3f}\n",
           (long)sel[0], (long)sel[1], (long)j, ms[0], ms[1], ms[2]);
    return 0;
}
#endif
