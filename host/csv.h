/* host/csv.h -- CSV in / CSV out with the reference's exact text semantics.
 * In:  set_csv_size + load_csv  (sort-merge-join/cpu_app.c:15-79 == app.c:28-92)
 * Out: save_to_csv              (cpu_app.c:268-301 == app.c:720-755) */
#ifndef SMJ_CSV_H
#define SMJ_CSV_H
#include <stdint.h>

/* Parses `path` into a row-major int32 table allocated with alloc(bytes) (e.g. pinned memory).
 * cols = number of ","-separated tokens of the header line, rows = lines - 1, cells = atoi(token).
 * Returns 0, or -1 when the file cannot be opened (errno set, like the reference's fopen check). */
int csv_load(const char *path, int32_t **data, int64_t *rows, int *cols, void *(*alloc)(uint64_t bytes));

/* Writes header col1..colN and the rows as decimal integers, "," separated, LF line ends. */
int csv_save(const char *path, const int32_t *data, int64_t rows, int cols);

#endif
