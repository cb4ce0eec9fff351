/* host/csv.c -- see csv.h.  Semantics reproduced from the reference's parser (cpu_app.c:15-79):
 *   - the file is consumed in fgets() chunks of at most 1023 characters; each chunk is one "line";
 *   - a line is split with strtok(line, ","): tokens are maximal runs of non-',' characters (empty fields
 *     collapse; the trailing "\n" / "\r\n" belongs to the last token, and a bare "\n" is a token too);
 *   - each token is atoi()'d: leading isspace() skipped, optional sign, digits, stop at anything else;
 *     glibc's atoi is (int)strtol(), i.e. saturate to LONG_MIN/LONG_MAX first, then truncate to int;
 *   - token `col` of line `row` is stored at cell row*cols+col (unchecked in the reference; bounds-checked here);
 *   - cells a short line does not reach stay uninitialised in the reference; they are 0 here.
 * The fast path works on the whole file in memory; any line of 1023+ characters makes it fall back to a
 * chunked walk that splits exactly where fgets() would. */
#include "csv.h"
#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHUNK 1023   /* fgets(line, 1024, f) returns at most 1023 characters */

static inline int is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* atoi() of the token [p, e) as glibc computes it. */
static inline int32_t tok_atoi(const char *p, const char *e)
{
    while (p < e && is_space((unsigned char)*p)) p++;
    int neg = 0;
    if (p < e && (*p == '-' || *p == '+')) { neg = (*p == '-'); p++; }
    unsigned long v = 0;
    int sat = 0;
    const unsigned long lim = neg ? (unsigned long)LONG_MAX + 1ul : (unsigned long)LONG_MAX;
    for (; p < e && *p >= '0' && *p <= '9'; p++) {
        unsigned d = (unsigned)(*p - '0');
        if (sat) continue;
        if (v > (lim - d) / 10) { sat = 1; v = lim; }
        else v = v * 10 + d;
    }
    long r = neg ? (long)(0ul - v) : (long)v;
    return (int32_t)(int)r;
}

/* end of the fgets() chunk starting at p: through the first '\n', at most CHUNK characters */
static inline const char *chunk_end(const char *p, const char *end)
{
    const char *lim = (end - p > CHUNK) ? p + CHUNK : end;
    const char *nl = (const char *)memchr(p, '\n', (size_t)(lim - p));
    return nl ? nl + 1 : lim;
}

static int count_tokens(const char *p, const char *e)
{
    int n = 0;
    while (p < e) {
        while (p < e && *p == ',') p++;
        if (p >= e) break;
        n++;
        while (p < e && *p != ',') p++;
    }
    return n;
}

int csv_load(const char *path, int32_t **data, int64_t *rows, int *cols, void *(*alloc)(uint64_t))
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); errno = ENOMEM; return -1; }
    size_t got = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    /* fgets() is a C-string API: an embedded NUL would end every token early; treat it as end of data */
    const char *nul = (const char *)memchr(buf, 0, got);
    const char *end = nul ? nul : buf + got;

    /* pass 1: chunk count and header tokens (set_csv_size) */
    int64_t lines = 0;
    int c = 0;
    for (const char *p = buf; p < end;) {
        const char *e = chunk_end(p, end);
        if (lines == 0) c = count_tokens(p, e);
        lines++;
        p = e;
    }
    int64_t r = lines - 1;
    *cols = c;
    *rows = r;
    if (r < 0) { *rows = r; *data = NULL; free(buf); return 0; }   /* empty file: rows = -1, like the reference */
    uint64_t cells = (uint64_t)r * (uint64_t)c;
    int32_t *out = (int32_t *)alloc(cells ? cells * sizeof(int32_t) : sizeof(int32_t));
    if (!out) { free(buf); errno = ENOMEM; return -1; }
    memset(out, 0, cells ? cells * sizeof(int32_t) : sizeof(int32_t));

    /* pass 2: cells (load_csv) */
    const char *p = buf;
    if (p < end) p = chunk_end(p, end);   /* skip header */
    int64_t row = 0;
    while (p < end) {
        const char *e = chunk_end(p, end);
        int col = 0;
        const char *q = p;
        while (q < e) {
            while (q < e && *q == ',') q++;
            if (q >= e) break;
            const char *t = q;
            while (q < e && *q != ',') q++;
            uint64_t idx = (uint64_t)row * (uint64_t)c + (uint64_t)col;
            if (idx < cells) out[idx] = tok_atoi(t, q);
            col++;
        }
        row++;
        p = e;
    }
    free(buf);
    *data = out;
    return 0;
}

static inline char *put_i32(char *p, int32_t v)
{
    char tmp[12];
    int n = 0;
    uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

int csv_save(const char *path, const int32_t *data, int64_t rows, int cols)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    const size_t cap = 1 << 20;
    char *buf = (char *)malloc(cap + (size_t)cols * 12 + 64);
    if (!buf) { fclose(f); errno = ENOMEM; return -1; }
    char *p = buf;
    for (int i = 1; i <= cols; i++) {
        p += sprintf(p, "col%d", i);
        if (i < cols) *p++ = ',';
        if ((size_t)(p - buf) >= cap) { fwrite(buf, 1, (size_t)(p - buf), f); p = buf; }
    }
    *p++ = '\n';
    for (int64_t i = 0; i < rows; i++) {
        const int32_t *r = data + i * cols;
        for (int j = 0; j < cols; j++) {
            p = put_i32(p, r[j]);
            if (j < cols - 1) *p++ = ',';
        }
        *p++ = '\n';
        if ((size_t)(p - buf) >= cap) { fwrite(buf, 1, (size_t)(p - buf), f); p = buf; }
    }
    fwrite(buf, 1, (size_t)(p - buf), f);
    free(buf);
    return fclose(f) ? -1 : 0;
}
