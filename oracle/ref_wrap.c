/*
 * oracle/ref_wrap.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin wrapper that compiles the UNMODIFIED reference CPU pipeline
 * (sort-merge-join/cpu_app.c) from where it lies under $SMJ_REF_DIR
 * (default /root/reference/sort-merge-join) by #including it through the
 * -I path.  No reference source is copied into this repository.
 *
 * Why a wrapper is needed (reference file:line):
 *   - cpu_app.c:350  the save_to_csv() call is commented out, so the stock
 *     binary never writes a result file;
 *   - cpu_app.c:336-344  the user.h knobs are only ever used as ARGUMENTS in
 *     main(), so calling the (non-static) stage functions directly lets the
 *     tests vary the knobs at run time without touching user.h.
 *
 * Built by oracle/Makefile into oracle/_ref/ (git-ignored, travels with
 * gpurun):  ref_oracle (CLI)  and  libref_oracle.so (ctypes).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use this.
 */
#define main ref_cpu_app_main
#include "cpu_app.c" /* found via -I$SMJ_REF_DIR; the reference file itself */
#undef main

#include <time.h>

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ---- stage-level entry points over caller-owned int64 buffers (ctypes) ---- */

/* select_in_cpu (cpu_app.c:81-112) frees its input, so hand it a malloc'd copy. */
int ref_select(const T *in, int rows, int cols, int col, T val, T *out)
{
    T *buf = (T *)malloc((size_t)(rows > 0 ? rows : 1) * cols * sizeof(T));
    memcpy(buf, in, (size_t)rows * cols * sizeof(T));
    int r = rows;
    select_in_cpu(cols, &r, &buf, col, val);
    memcpy(out, buf, (size_t)r * cols * sizeof(T));
    free(buf);
    return r;
}

/* insertion_sort_in_cpu (cpu_app.c:172-202), in place. */
void ref_sort(T *inout, int rows, int cols, int key)
{
    insertion_sort_in_cpu(cols, rows, key, &inout);
}

/* join_in_cpu (cpu_app.c:204-266) leaves its output in the globals
 * result/result_row_num/result_col_num; returns the row count. */
int ref_join(T *l, int r1, int c1, T *r, int r2, int c2, int key1, int key2)
{
    if (result) { free(result); result = NULL; }
    join_in_cpu(c1, r1, l, c2, r2, r, key1, key2);
    return result_row_num;
}

int ref_join_cols(void) { return result_col_num; }

void ref_join_copy(T *out)
{
    memcpy(out, result, (size_t)result_row_num * result_col_num * sizeof(T));
}

/* set_csv_size + load_csv (cpu_app.c:15-79). Caller frees with ref_free. */
T *ref_load_csv(const char *path, int *cols, int *rows)
{
    T *a = NULL;
    *cols = 0; *rows = 0;
    set_csv_size(path, cols, rows);
    load_csv(path, *cols, *rows, &a);
    return a;
}

void ref_free(void *p) { free(p); }

void ref_save_csv(const char *path, int cols, int rows, T *a)
{
    save_to_csv(path, cols, rows, a);
}

/* Whole pipeline in the order of cpu_app.c:324-344, knobs as arguments.
 * stage_ms[5] = load, select, sort, join, save.  out_path may be NULL.
 * dump_prefix (may be NULL): writes <prefix>_select{1,2}.csv and
 * <prefix>_sort{1,2}.csv for stage-level parity tests. */
int ref_pipeline_csv(const char *f1, const char *f2, const char *out_path,
                     int sel_col1, long sel_val1, int sel_col2, long sel_val2,
                     int key1, int key2, double *stage_ms, int *selected,
                     const char *dump_prefix)
{
    int c1 = 0, r1 = 0, c2 = 0, r2 = 0;
    T *a = NULL, *b = NULL;
    char path[4096];
    double t0 = now_ms();
    set_csv_size(f1, &c1, &r1);
    set_csv_size(f2, &c2, &r2);
    load_csv(f1, c1, r1, &a);
    load_csv(f2, c2, r2, &b);
    double t1 = now_ms();
    select_in_cpu(c1, &r1, &a, sel_col1, sel_val1);
    select_in_cpu(c2, &r2, &b, sel_col2, sel_val2);
    double t2 = now_ms();
    if (dump_prefix) {
        snprintf(path, sizeof path, "%s_select1.csv", dump_prefix); save_to_csv(path, c1, r1, a);
        snprintf(path, sizeof path, "%s_select2.csv", dump_prefix); save_to_csv(path, c2, r2, b);
    }
    double t2b = now_ms();
    insertion_sort_in_cpu(c1, r1, key1, &a);
    insertion_sort_in_cpu(c2, r2, key2, &b);
    double t3 = now_ms();
    if (dump_prefix) {
        snprintf(path, sizeof path, "%s_sort1.csv", dump_prefix); save_to_csv(path, c1, r1, a);
        snprintf(path, sizeof path, "%s_sort2.csv", dump_prefix); save_to_csv(path, c2, r2, b);
    }
    double t3b = now_ms();
    if (result) { free(result); result = NULL; }
    join_in_cpu(c1, r1, a, c2, r2, b, key1, key2);
    double t4 = now_ms();
    if (out_path) save_to_csv(out_path, result_col_num, result_row_num, result);
    double t5 = now_ms();
    if (stage_ms) {
        stage_ms[0] = t1 - t0; stage_ms[1] = t2 - t1; stage_ms[2] = t3 - t2b;
        stage_ms[3] = t4 - t3b; stage_ms[4] = t5 - t4;
    }
    if (selected) { selected[0] = r1; selected[1] = r2; }
    free(a); free(b);
    return result_row_num;
}

#ifdef REF_ORACLE_MAIN
/* ref_oracle data1.csv data2.csv out.csv [sel_col1 sel_val1 sel_col2 sel_val2 key1 key2 [dump_prefix]]
 * Defaults are the reference's user.h macros (user.h:6-13). */
int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s data1.csv data2.csv out.csv|- [sc1 sv1 sc2 sv2 k1 k2 [dump_prefix]]\n", argv[0]);
        return 2;
    }
    int sc1 = SELECT_COL1, sc2 = SELECT_COL2, k1 = JOIN_KEY1, k2 = JOIN_KEY2;
    long sv1 = SELECT_VAL1, sv2 = SELECT_VAL2;
    if (argc >= 10) {
        sc1 = atoi(argv[4]); sv1 = atol(argv[5]); sc2 = atoi(argv[6]);
        sv2 = atol(argv[7]); k1 = atoi(argv[8]); k2 = atoi(argv[9]);
    }
    const char *dump = argc >= 11 ? argv[10] : NULL;
    const char *out = strcmp(argv[3], "-") ? argv[3] : NULL;
    double ms[5]; int sel[2];
    int j = ref_pipeline_csv(argv[1], argv[2], out, sc1, sv1, sc2, sv2, k1, k2, ms, sel, dump);
    printf("{\"selected\": [%d, %d], \"joined\": %d, \"cols\": %d, \"load_ms\": %.3f, \"select_ms\": %.3f, "
           "\"sort_ms\": %.3f, \"join_ms\": %.3f, \"save_ms\": %.3f}\n",
           sel[0], sel[1], j, result_col_num, ms[0], ms[1], ms[2], ms[3], ms[4]);
    return 0;
}
#endif
