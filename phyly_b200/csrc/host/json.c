#include "json.h"
#include "par.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

typedef struct {
    const char *p;
    const char *start;
    const char *end;    /* the terminating NUL */
    char *err;
    size_t errlen;
    int depth;
    const char *key;    /* key of the object member whose value is about to be read, else NULL */
    size_t hint;        /* length of the array parsed last: rows of a matrix have one length, so the next one is sized right at once */
} parser;

static void fail(parser *ps, const char *msg)
{
    if (ps->err && ps->errlen && !ps->err[0]) {
        int line = 1;
        for (const char *q = ps->start; q < ps->p; q++) if (*q == '\n') line++;
        snprintf(ps->err, ps->errlen, "error on json line %d: %s", line, msg);
    }
}

static void skip_ws(parser *ps)
{
    while (*ps->p == ' ' || *ps->p == '\t' || *ps->p == '\n' || *ps->p == '\r') ps->p++;
}

static void free_contents(jv *v)
{
    if (v->type == JV_STRING) free(v->u.s);
    else if (v->type == JV_ARRAY && (v->flags & JV_F_COMPACT)) {
        json_codes_release((jv_codes *)v->u.items);
    } else if (v->type == JV_ARRAY) {
        if (v->flags & JV_F_BLOCKS) {           /* rows of integers in shared blocks (parse_int_matrix) */
            const jv *tail = v->u.items + v->len;
            for (uint32_t t = 0; t < tail->len; t++) free(tail[1 + t].u.items);
        } else if (!(v->flags & JV_F_BORROWED)) {
            for (uint32_t i = 0; i < v->len; i++) if (v->u.items[i].type >= JV_STRING) free_contents(&v->u.items[i]);
        }
        if (!(v->flags & JV_F_BORROWED)) free(v->u.items);
    } else if (v->type == JV_OBJECT) {
        for (uint32_t i = 0; i < 2 * v->len; i++) free_contents(&v->u.items[i]);
        free(v->u.items);
    }
    v->type = JV_NULL;
}

static int parse_value(parser *ps, jv *out);

static int parse_string(parser *ps, jv *out)
{
    /* ps->p points at the opening quote */
    const char *q = ps->p + 1;
    size_t cap = 16, len = 0;
    char *s = malloc(cap);
    if (!s) { fail(ps, "out of memory"); return -1; }
    for (;;) {
        unsigned char ch = (unsigned char)*q;
        if (ch == 0) { free(s); ps->p = q; fail(ps, "premature end of input in string"); return -1; }
        if (ch == '"') { q++; break; }
        if (ch < 0x20) { free(s); ps->p = q; fail(ps, "control character in string"); return -1; }
        unsigned out_cp = ch;
        if (ch == '\\') {
            q++;
            switch (*q) {
            case '"': out_cp = '"'; break;
            case '\\': out_cp = '\\'; break;
            case '/': out_cp = '/'; break;
            case 'b': out_cp = '\b'; break;
            case 'f': out_cp = '\f'; break;
            case 'n': out_cp = '\n'; break;
            case 'r': out_cp = '\r'; break;
            case 't': out_cp = '\t'; break;
            case 'u': {
                unsigned cp = 0;
                for (int k = 1; k <= 4; k++) {
                    char h = q[k];
                    cp <<= 4;
                    if (h >= '0' && h <= '9') cp |= (unsigned)(h - '0');
                    else if (h >= 'a' && h <= 'f') cp |= (unsigned)(h - 'a' + 10);
                    else if (h >= 'A' && h <= 'F') cp |= (unsigned)(h - 'A' + 10);
                    else { free(s); ps->p = q; fail(ps, "invalid \\u escape"); return -1; }
                }
                q += 4;
                out_cp = cp;
                break;
            }
            default: free(s); ps->p = q; fail(ps, "invalid escape"); return -1;
            }
        }
        if (len + 5 > cap) { cap *= 2; char *t = realloc(s, cap); if (!t) { free(s); fail(ps, "out of memory"); return -1; } s = t; }
        if (out_cp < 0x80) s[len++] = (char)out_cp;
        else if (out_cp < 0x800) { s[len++] = (char)(0xC0 | (out_cp >> 6)); s[len++] = (char)(0x80 | (out_cp & 0x3F)); }
        else { s[len++] = (char)(0xE0 | (out_cp >> 12)); s[len++] = (char)(0x80 | ((out_cp >> 6) & 0x3F)); s[len++] = (char)(0x80 | (out_cp & 0x3F)); }
        q++;
    }
    s[len] = 0;
    out->type = JV_STRING; out->flags = 0; out->len = (uint32_t)len; out->u.s = s;
    ps->p = q;
    return 0;
}

static int parse_number(parser *ps, jv *out)
{
    const char *q = ps->p;
    int is_real = 0;
    if (*q == '-') q++;
    if (*q == '0') q++;
    else if (*q >= '1' && *q <= '9') { while (*q >= '0' && *q <= '9') q++; }
    else { fail(ps, "invalid token"); return -1; }
    if (*q == '.') {
        is_real = 1; q++;
        if (!(*q >= '0' && *q <= '9')) { ps->p = q; fail(ps, "invalid number"); return -1; }
        while (*q >= '0' && *q <= '9') q++;
    }
    if (*q == 'e' || *q == 'E') {
        is_real = 1; q++;
        if (*q == '+' || *q == '-') q++;
        if (!(*q >= '0' && *q <= '9')) { ps->p = q; fail(ps, "invalid number"); return -1; }
        while (*q >= '0' && *q <= '9') q++;
    }
    if (!is_real) {
        /* fast path for short integers (character codes, indices) */
        size_t nd = (size_t)(q - ps->p);
        if (nd <= 17) {
            const char *r = ps->p;
            int neg = 0;
            if (*r == '-') { neg = 1; r++; }
            int64_t v = 0;
            for (; r < q; r++) v = v * 10 + (*r - '0');
            out->type = JV_INT; out->u.i = neg ? -v : v; out->len = 0;
            ps->p = q;
            return 0;
        }
        errno = 0;
        long long v = strtoll(ps->p, NULL, 10);
        if (errno == ERANGE) { fail(ps, "too big integer"); return -1; }
        out->type = JV_INT; out->u.i = v; out->len = 0;
    } else {
        errno = 0;
        double d = strtod(ps->p, NULL);
        if (errno == ERANGE && (d == HUGE_VAL || d == -HUGE_VAL)) { fail(ps, "real number overflow"); return -1; }
        out->type = JV_REAL; out->u.d = d; out->len = 0;
    }
    ps->p = q;
    return 0;
}

/* ---- the 'character_data' cache (json.h) ---- */

#include <pthread.h>

static pthread_mutex_t g_codes_lock = PTHREAD_MUTEX_INITIALIZER;
static jv_codes *g_codes_entry = NULL;

void json_codes_retain(jv_codes *d)
{
    if (!d) return;
    pthread_mutex_lock(&g_codes_lock);
    d->refs++;
    pthread_mutex_unlock(&g_codes_lock);
}

void json_codes_release(jv_codes *d)
{
    if (!d) return;
    pthread_mutex_lock(&g_codes_lock);
    const int left = --d->refs;
    pthread_mutex_unlock(&g_codes_lock);
    if (left == 0) { free(d->codes); free(d); }
}

jv_codes *json_codes_publish(uint64_t h1, uint64_t h2, size_t text_len, void *codes, int code_bytes, int64_t rows, int row_len, int max_code)
{
    jv_codes *d = malloc(sizeof(jv_codes));
    if (!d) return NULL;
    d->codes = codes; d->code_bytes = code_bytes; d->rows = rows; d->row_len = row_len; d->max_code = max_code;
    d->h1 = h1; d->h2 = h2; d->text_len = text_len;
    d->refs = 2;                /* the cache's and the caller's */
    pthread_mutex_lock(&g_codes_lock);
    jv_codes *old = g_codes_entry;
    g_codes_entry = d;
    pthread_mutex_unlock(&g_codes_lock);
    json_codes_release(old);
    return d;
}

int jv_matrix_text_hash(const jv *v, uint64_t *h1, uint64_t *h2, size_t *text_len)
{
    if (!v || v->type != JV_ARRAY || !(v->flags & JV_F_HASHED) || !(v->flags & JV_F_BLOCKS)) return 0;
    const jv *tail = v->u.items + v->len;
    const jv *slot = tail + 1 + tail->len;
    *h1 = (uint64_t)slot[0].u.i; *h2 = (uint64_t)slot[1].u.i; *text_len = (size_t)slot[2].u.i;
    return 1;
}

/* 128 bits over a text: two multiply-xor chains per 1 MB chunk (the chunks go to the host threads; their size does not
 * depend on the number of threads), the chunk values folded in order.  Not cryptographic: it guards against reading a
 * different alignment as the cached one by accident, not against an adversary. */
#define HASH_CHUNK ((size_t)1 << 20)
typedef struct { const char *p; size_t len, nchunk; uint64_t *h; } hash_job;

static inline uint64_t mix64(uint64_t x) { x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; return x; }

static void hash_worker(int tid, int nthreads, void *ctx)
{
    hash_job *job = ctx;
    for (size_t c = (size_t)tid; c < job->nchunk; c += (size_t)nthreads) {
        const char *p = job->p + c * HASH_CHUNK;
        const size_t n = (c + 1 == job->nchunk) ? job->len - c * HASH_CHUNK : HASH_CHUNK;
        uint64_t a = 0x9e3779b97f4a7c15ull ^ n, b = 0xc2b2ae3d27d4eb4full ^ (n << 1);
        size_t i = 0;
        for (; i + 8 <= n; i += 8) {
            uint64_t w;
            memcpy(&w, p + i, 8);
            a = (a ^ w) * 0x100000001b3ull; a ^= a >> 29;
            b = (b + w) * 0xff51afd7ed558ccdull; b ^= b >> 31;
        }
        uint64_t w = 0;
        memcpy(&w, p + i, n - i);
        a = (a ^ w) * 0x100000001b3ull; b = (b + w) * 0xff51afd7ed558ccdull;
        job->h[2 * c] = mix64(a); job->h[2 * c + 1] = mix64(b);
    }
}

static void text_hash(const char *p, size_t len, uint64_t *h1, uint64_t *h2)
{
    const size_t nchunk = (len + HASH_CHUNK - 1) / HASH_CHUNK;
    uint64_t *h = malloc(sizeof(uint64_t) * 2 * (nchunk ? nchunk : 1));
    *h1 = 0x243f6a8885a308d3ull ^ len; *h2 = 0x13198a2e03707344ull ^ len;
    if (!h) { *h1 = *h2 = 0; return; }
    hash_job job = {p, len, nchunk, h};
    int nthreads = par_threads();
    if ((size_t)nthreads > nchunk) nthreads = nchunk ? (int)nchunk : 1;
    par_run(nthreads, hash_worker, &job);
    for (size_t c = 0; c < nchunk; c++) {
        *h1 = mix64(*h1 ^ h[2 * c]) + c;
        *h2 = mix64(*h2 + h[2 * c + 1]) ^ (c << 7);
    }
    free(h);
}

/* the cached codes if the text at p (the matrix's opening bracket) is the text they were read from */
static jv_codes *cache_probe(const char *p, const char *end)
{
    pthread_mutex_lock(&g_codes_lock);
    jv_codes *d = g_codes_entry;
    if (d) d->refs++;
    pthread_mutex_unlock(&g_codes_lock);
    if (!d) return NULL;
    if ((size_t)(end - p) >= d->text_len && d->text_len >= 2 && p[d->text_len - 1] == ']') {
        uint64_t h1, h2;
        text_hash(p, d->text_len, &h1, &h2);
        if (h1 == d->h1 && h2 == d->h2 && (h1 | h2) != 0) return d;
    }
    json_codes_release(d);
    return NULL;
}

/*
 * A large array whose items are flat arrays of short non-negative integers -- the 'character_data' of an alignment,
 * 2 bytes per code and 10^8 codes at the target size -- is read by the host threads in two passes over disjoint ranges:
 * first every ']' of the remaining text is located (memchr), then each thread reads a contiguous range of rows (the text
 * between one ']' and the next) into one block of its own.  The row after which the array's own ']' follows ends the
 * matrix; threads that were handed text beyond it have read garbage, which is dropped.  A row holding anything but such
 * integers (a sign, a fraction, a nested array, a string, a stray comma ...) before that point makes the whole attempt
 * step aside: the array is then read again by the general route, which owns every check and every error message.
 * Returns 0 when the array has been read, 1 when it is left to the general route.
 */
#define MATRIX_MIN_BYTES (1u << 20)
#define MATRIX_MIN_ROWS 256

static inline const char *skip_ws_to(const char *q)
{
    while (*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r') q++;
    return q;
}

/* a block of many megabytes that is written once, front to back: huge pages (where the kernel grants them on request)
 * cut the page faults of first touch, which otherwise cost as much as reading the digits */
static void *big_block(size_t bytes)
{
    const size_t huge = (size_t)2 << 20;
    if (bytes >= 2 * huge) {
        void *p = NULL;
        if (posix_memalign(&p, huge, (bytes + huge - 1) / huge * huge) == 0 && p) {
#ifdef MADV_HUGEPAGE
            madvise(p, (bytes + huge - 1) / huge * huge, MADV_HUGEPAGE);
#endif
            return p;
        }
    }
    return malloc(bytes);
}

/* pass 1: the positions of every ']' in [lo, hi), one list per thread */
typedef struct { const char *lo, *hi; const char **pos; size_t n; int oom; } scan_part;
typedef struct { const char *lo, *hi; scan_part *parts; } scan_job;

static void scan_worker(int tid, int nthreads, void *ctx)
{
    scan_job *job = ctx;
    const size_t span = (size_t)(job->hi - job->lo);
    scan_part *sp = &job->parts[tid];
    sp->lo = job->lo + span * (size_t)tid / (size_t)nthreads;
    sp->hi = job->lo + span * (size_t)(tid + 1) / (size_t)nthreads;
    size_t cap = (size_t)(sp->hi - sp->lo) / 128 + 64;
    sp->pos = malloc(cap * sizeof(const char *));
    sp->n = 0;
    if (!sp->pos) { sp->oom = 1; return; }
    for (const char *p = sp->lo; p < sp->hi; p++) {
        p = memchr(p, ']', (size_t)(sp->hi - p));
        if (!p) break;
        if (sp->n == cap) {
            cap *= 2;
            const char **t = realloc(sp->pos, cap * sizeof(const char *));
            if (!t) { sp->oom = 1; return; }
            sp->pos = t;
        }
        sp->pos[sp->n++] = p;
    }
}

/* pass 2: rows [r0, r1) of the matrix, row r being the text that ends at closes[r] */
enum { EV_NONE = 0, EV_END, EV_INVALID };
typedef struct { int kind; size_t row; } row_event;

typedef struct {
    const char *first_open;
    const char *const *closes;
    size_t nclose;
    jv *row_out;
    jv **blocks;
    row_event *events;
} matrix_job;

static void matrix_worker(int tid, int nthreads, void *ctx)
{
    matrix_job *job = ctx;
    const size_t r0 = job->nclose * (size_t)tid / (size_t)nthreads, r1 = job->nclose * (size_t)(tid + 1) / (size_t)nthreads;
    row_event *ev = &job->events[tid];
    ev->kind = EV_NONE;
    if (r0 == r1) return;
    /* an item takes at least two bytes of text (digit + separator) */
    const char *from = r0 ? job->closes[r0 - 1] : job->first_open;
    const size_t bound = (size_t)(job->closes[r1 - 1] - from) / 2 + (r1 - r0) + 1;
    jv *block = big_block(bound * sizeof(jv));
    job->blocks[tid] = block;
    if (!block) { ev->kind = EV_INVALID; ev->row = r0; return; }
    jv *cur = block;
    for (size_t r = r0; r < r1; r++) {
        const char *q;
        if (r == 0) q = job->first_open;
        else {
            q = skip_ws_to(job->closes[r - 1] + 1);
            if (*q == ']') { ev->kind = EV_END; ev->row = r; return; }       /* the array's own bracket: rows 0 .. r-1 */
            if (*q != ',') { ev->kind = EV_INVALID; ev->row = r; return; }
            q = skip_ws_to(q + 1);
        }
        if (*q != '[') { ev->kind = EV_INVALID; ev->row = r; return; }
        const char *close = job->closes[r];
        q = skip_ws_to(q + 1);
        jv *first = cur;
        if (q != close) {
            for (;;) {
                int64_t v;
                if (*q >= '1' && *q <= '9') {
                    int nd = 1;
                    v = *q++ - '0';
                    while (*q >= '0' && *q <= '9' && nd < 9) { v = v * 10 + (*q++ - '0'); nd++; }
                } else if (*q == '0') {
                    v = 0; q++;
                } else { ev->kind = EV_INVALID; ev->row = r; return; }
                cur->type = JV_INT; cur->flags = 0; cur->len = 0; cur->u.i = v;
                cur++;
                q = skip_ws_to(q);
                if (q == close) break;
                if (*q != ',') { ev->kind = EV_INVALID; ev->row = r; return; }
                q = skip_ws_to(q + 1);
            }
        }
        jv *row = &job->row_out[r];
        row->type = JV_ARRAY; row->flags = JV_F_BORROWED; row->len = (uint32_t)(cur - first); row->u.items = first;
    }
}

static int parse_int_matrix(parser *ps, jv *out)
{
    /* ps->p is at the '[' of the first row.  A matrix that ends within MATRIX_MIN_ROWS rows (a rate matrix, a small
     * tree's edges) is not worth a pass over the rest of the document */
    {
        const char *p = ps->p;
        for (int r = 0; r < MATRIX_MIN_ROWS; r++) {
            if (*p != '[') return 1;
            const char *close = memchr(p + 1, ']', (size_t)(ps->end - (p + 1)));
            if (!close) return 1;
            p = skip_ws_to(close + 1);
            if (*p != ',') return 1;
            p = skip_ws_to(p + 1);
        }
    }
    int nthreads = par_threads();
    const int scan_threads = nthreads;
    scan_part *parts = calloc((size_t)nthreads, sizeof(scan_part));
    if (!parts) return 1;
    scan_job sj = {ps->p, ps->end, parts};
    par_run(nthreads, scan_worker, &sj);
    size_t nclose = 0;
    int bad = 0;
    for (int t = 0; t < nthreads; t++) { nclose += parts[t].n; bad |= parts[t].oom; }
    const char **closes = bad ? NULL : malloc((nclose + 1) * sizeof(const char *));
    jv *items = NULL;
    jv **blocks = NULL;
    row_event *events = NULL;
    int rc = 1;
    if (!closes || nclose <= MATRIX_MIN_ROWS || nclose >= 0xffffffffu) goto done;
    {
        size_t k = 0;
        for (int t = 0; t < nthreads; t++) { memcpy(closes + k, parts[t].pos, parts[t].n * sizeof(const char *)); k += parts[t].n; }
    }
    if ((size_t)nthreads > nclose / 64) nthreads = (int)(nclose / 64);
    if (nthreads < 1) nthreads = 1;
    items = malloc((nclose + 4 + (size_t)nthreads) * sizeof(jv));
    blocks = calloc((size_t)nthreads, sizeof(jv *));
    events = calloc((size_t)nthreads, sizeof(row_event));
    if (!items || !blocks || !events) goto done;
    {
        matrix_job job = {ps->p, closes, nclose, items, blocks, events};
        par_run(nthreads, matrix_worker, &job);
    }
    {
        /* ranges ascend with the thread index: the first event is the earliest row that is not a row */
        size_t n = 0;
        int kind = EV_NONE;
        for (int t = 0; t < nthreads && kind == EV_NONE; t++) if (events[t].kind != EV_NONE) { kind = events[t].kind; n = events[t].row; }
        if (kind != EV_END || n < MATRIX_MIN_ROWS) goto done;
        jv *tail = items + n;
        tail->type = JV_NULL; tail->flags = 0; tail->len = (uint32_t)nthreads; tail->u.i = 0;
        for (int t = 0; t < nthreads; t++) {
            if (nclose * (size_t)t / (size_t)nthreads >= n) { free(blocks[t]); blocks[t] = NULL; }     /* text beyond the matrix */
            tail[1 + t].type = JV_NULL; tail[1 + t].flags = 0; tail[1 + t].len = 0; tail[1 + t].u.items = blocks[t];
            blocks[t] = NULL;
        }
        if (getenv("ARBPLF_JSON_TRACE")) fprintf(stderr, "json: matrix of %zu rows read by %d threads\n", n, nthreads);
        ps->hint = items[n - 1].len;
        out->type = JV_ARRAY; out->flags = JV_F_BLOCKS; out->len = (uint32_t)n; out->u.items = items;
        items = NULL;
        ps->p = closes[n] + 1;          /* the ']' after the last row is the array's own */
        rc = 0;
    }
done:
    if (blocks) for (int t = 0; t < nthreads; t++) free(blocks[t]);
    free(blocks); free(events); free(items); free(closes);
    for (int t = 0; t < scan_threads; t++) free(parts[t].pos);
    free(parts);
    return rc;
}

static int parse_array(parser *ps, jv *out)
{
    const char *key = ps->key;
    ps->key = NULL;             /* the items of this array are nobody's member */
    if (ps->depth < 2000 && (size_t)(ps->end - ps->p) >= MATRIX_MIN_BYTES) {
        const char *q = skip_ws_to(ps->p + 1);
        if (*q == '[') {
            const char *keep = ps->p;
            const int is_data = key && !strcmp(key, "character_data") && !getenv("ARBPLF_NO_DATA_CACHE");
            if (is_data) {
                jv_codes *d = cache_probe(keep, ps->end);
                if (d) {        /* the bytes of the matrix read last time: nothing to read */
                    if (getenv("ARBPLF_JSON_TRACE")) fprintf(stderr, "json: character_data of %lld rows taken from the cache\n", (long long)d->rows);
                    out->type = JV_ARRAY; out->flags = JV_F_COMPACT;
                    out->len = d->rows > 0xffffffffLL ? 0xffffffffu : (uint32_t)d->rows;
                    out->u.items = (jv *)d;
                    ps->p = keep + d->text_len;
                    return 0;
                }
            }
            ps->p = q;
            if (parse_int_matrix(ps, out) == 0) {
                if (is_data) {
                    /* remember what text this was: three hidden items behind the block list */
                    jv *tail = out->u.items + out->len;
                    jv *slot = tail + 1 + tail->len;
                    uint64_t h1, h2;
                    text_hash(keep, (size_t)(ps->p - keep), &h1, &h2);
                    slot[0].type = slot[1].type = slot[2].type = JV_NULL;
                    slot[0].flags = slot[1].flags = slot[2].flags = 0;
                    slot[0].u.i = (int64_t)h1; slot[1].u.i = (int64_t)h2; slot[2].u.i = (int64_t)(ps->p - keep);
                    out->flags |= JV_F_HASHED;
                }
                return 0;
            }
            ps->p = keep;
        }
    }
    size_t cap = (ps->hint >= 8 && ps->hint <= (1u << 20)) ? ps->hint : 8, len = 0;
    jv *items = malloc(cap * sizeof(jv));
    if (!items) { fail(ps, "out of memory"); return -1; }
    ps->p++;
    skip_ws(ps);
    if (*ps->p == ']') { ps->p++; goto done; }
    for (;;) {
        if (len == cap) {
            cap *= 2;
            jv *t = realloc(items, cap * sizeof(jv));
            if (!t) { fail(ps, "out of memory"); goto bad; }
            items = t;
        }
        skip_ws(ps);
        {
            /* fast path for what a large document consists of: short non-negative integers (character codes) directly
             * followed by ',' or ']'.  Anything else -- signs, fractions, exponents, leading zeros -- takes the general
             * route and its checks. */
            const char *q = ps->p;
            if (*q >= '1' && *q <= '9') {
                int64_t v = *q++ - '0';
                int nd = 1;
                while (*q >= '0' && *q <= '9' && nd < 9) { v = v * 10 + (*q++ - '0'); nd++; }
                if (*q == ',' || *q == ']') {
                    items[len].type = JV_INT; items[len].flags = 0; items[len].len = 0; items[len].u.i = v;
                    len++;
                    ps->p = q + 1;
                    if (*q == ',') continue;
                    break;
                }
            } else if (*q == '0' && (q[1] == ',' || q[1] == ']')) {
                items[len].type = JV_INT; items[len].flags = 0; items[len].len = 0; items[len].u.i = 0;
                len++;
                ps->p = q + 2;
                if (q[1] == ',') continue;
                break;
            }
        }
        if (parse_value(ps, &items[len])) goto bad;
        len++;
        skip_ws(ps);
        if (*ps->p == ',') { ps->p++; continue; }
        if (*ps->p == ']') { ps->p++; break; }
        fail(ps, "']' expected");
        goto bad;
    }
done:
    if (len < cap && len > 0) { jv *t = realloc(items, len * sizeof(jv)); if (t) items = t; }
    ps->hint = len;
    out->type = JV_ARRAY; out->len = (uint32_t)len; out->u.items = items;
    return 0;
bad:
    for (size_t i = 0; i < len; i++) free_contents(&items[i]);
    free(items);
    return -1;
}

static int parse_object(parser *ps, jv *out)
{
    size_t cap = 8, len = 0;   /* pairs */
    jv *items = malloc(2 * cap * sizeof(jv));
    if (!items) { fail(ps, "out of memory"); return -1; }
    ps->p++;
    skip_ws(ps);
    if (*ps->p == '}') { ps->p++; goto done; }
    for (;;) {
        if (len == cap) {
            cap *= 2;
            jv *t = realloc(items, 2 * cap * sizeof(jv));
            if (!t) { fail(ps, "out of memory"); goto bad; }
            items = t;
        }
        skip_ws(ps);
        if (*ps->p != '"') { fail(ps, "string or '}' expected"); goto bad; }
        if (parse_string(ps, &items[2 * len])) goto bad;
        skip_ws(ps);
        if (*ps->p != ':') { free_contents(&items[2 * len]); fail(ps, "':' expected"); goto bad; }
        ps->p++;
        skip_ws(ps);
        ps->key = items[2 * len].u.s;
        if (parse_value(ps, &items[2 * len + 1])) { ps->key = NULL; free_contents(&items[2 * len]); goto bad; }
        ps->key = NULL;
        /* jansson: a repeated key replaces the earlier value */
        {
            int dup = 0;
            for (size_t i = 0; i < len; i++) {
                if (!strcmp(items[2 * i].u.s, items[2 * len].u.s)) {
                    free_contents(&items[2 * i + 1]);
                    items[2 * i + 1] = items[2 * len + 1];
                    free_contents(&items[2 * len]);
                    dup = 1;
                    break;
                }
            }
            if (!dup) len++;
        }
        skip_ws(ps);
        if (*ps->p == ',') { ps->p++; continue; }
        if (*ps->p == '}') { ps->p++; break; }
        fail(ps, "'}' expected");
        goto bad;
    }
done:
    out->type = JV_OBJECT; out->len = (uint32_t)len; out->u.items = items;
    return 0;
bad:
    for (size_t i = 0; i < 2 * len; i++) free_contents(&items[i]);
    free(items);
    return -1;
}

static int parse_value(parser *ps, jv *out)
{
    out->type = JV_NULL; out->flags = 0; out->len = 0; out->u.i = 0;
    if (++ps->depth > 2048) { fail(ps, "maximum parsing depth reached"); return -1; }
    int rc;
    skip_ws(ps);
    char ch = *ps->p;
    if (ch == '{') rc = parse_object(ps, out);
    else if (ch == '[') rc = parse_array(ps, out);
    else if (ch == '"') rc = parse_string(ps, out);
    else if (ch == '-' || (ch >= '0' && ch <= '9')) rc = parse_number(ps, out);
    else if (!strncmp(ps->p, "true", 4)) { out->type = JV_TRUE; ps->p += 4; rc = 0; }
    else if (!strncmp(ps->p, "false", 5)) { out->type = JV_FALSE; ps->p += 5; rc = 0; }
    else if (!strncmp(ps->p, "null", 4)) { out->type = JV_NULL; ps->p += 4; rc = 0; }
    else if (ch == 0) { fail(ps, "unexpected end of input"); rc = -1; }
    else { fail(ps, "invalid token"); rc = -1; }
    ps->depth--;
    return rc;
}

jv *json_parse(const char *text, char *err, size_t errlen)
{
    parser ps = {text, text, text + strlen(text), err, errlen, 0, NULL, 0};
    if (err && errlen) err[0] = 0;
    jv *root = malloc(sizeof(jv));
    if (!root) { fail(&ps, "out of memory"); return NULL; }
    skip_ws(&ps);
    /* jansson's json_loads without JSON_DECODE_ANY accepts only arrays and objects */
    if (*ps.p != '{' && *ps.p != '[') { fail(&ps, "'[' or '{' expected"); free(root); return NULL; }
    if (parse_value(&ps, root)) { free(root); return NULL; }
    skip_ws(&ps);
    if (*ps.p != 0) { fail(&ps, "end of file expected"); free_contents(root); free(root); return NULL; }
    return root;
}

void json_free(jv *v)
{
    if (!v) return;
    free_contents(v);
    free(v);
}

const jv *jv_get(const jv *obj, const char *key)
{
    if (!obj || obj->type != JV_OBJECT) return NULL;
    for (uint32_t i = 0; i < obj->len; i++)
        if (!strcmp(obj->u.items[2 * i].u.s, key)) return &obj->u.items[2 * i + 1];
    return NULL;
}

int jv_unpack_strict(const jv *obj, const char *const *keys, const jv **out)
{
    if (!obj || obj->type != JV_OBJECT) {
        fprintf(stderr, "error: on line -1: Expected object, got %s\n", obj ? "another type" : "NULL");
        return -1;
    }
    size_t nk = 0;
    while (keys[nk]) nk++;
    size_t found = 0;
    for (size_t k = 0; k < nk; k++) {
        int optional = keys[k][0] == '?';
        const char *name = keys[k] + (optional ? 1 : 0);
        out[k] = jv_get(obj, name);
        if (out[k]) found++;
        else if (!optional) {
            fprintf(stderr, "error: on line -1: Object item not found: %s\n", name);
            return -1;
        }
    }
    if (found != obj->len) {
        /* JSON_STRICT: every key of the object must have been unpacked */
        for (uint32_t i = 0; i < obj->len; i++) {
            const char *have = obj->u.items[2 * i].u.s;
            int known = 0;
            for (size_t k = 0; k < nk; k++) if (!strcmp(have, keys[k] + (keys[k][0] == '?' ? 1 : 0))) known = 1;
            if (!known) {
                fprintf(stderr, "error: on line -1: %u object item(s) left unpacked: %s\n",
                        (unsigned)(obj->len - found), have);
                return -1;
            }
        }
    }
    return 0;
}

void jbuf_init(jbuf *b) { b->p = NULL; b->len = 0; b->cap = 0; }

static void jbuf_reserve(jbuf *b, size_t extra)
{
    if (b->len + extra + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 256;
        while (nc < b->len + extra + 1) nc *= 2;
        char *t = realloc(b->p, nc);
        if (!t) { fprintf(stderr, "out of memory building JSON output\n"); abort(); }
        b->p = t; b->cap = nc;
    }
}

void jbuf_puts(jbuf *b, const char *s)
{
    size_t n = strlen(s);
    jbuf_reserve(b, n);
    memcpy(b->p + b->len, s, n);
    b->len += n;
    b->p[b->len] = 0;
}

void jbuf_append(jbuf *b, const char *s, size_t n)
{
    jbuf_reserve(b, n);
    memcpy(b->p + b->len, s, n);
    b->len += n;
    b->p[b->len] = 0;
}

void jbuf_int(jbuf *b, long long v)
{
    char tmp[32];
    snprintf(tmp, sizeof tmp, "%lld", v);
    jbuf_puts(b, tmp);
}

void jbuf_real(jbuf *b, double d)
{
    char tmp[64];
    if (d == 0.0) d = 0.0;                 /* scrub -0.0, util.c:44-48 */
    if (d == 0.0 && signbit(d)) d = fabs(d);
    snprintf(tmp, sizeof tmp, "%.17g", d);
    /* jansson: make sure the token reads back as a real */
    if (!strchr(tmp, '.') && !strchr(tmp, 'e') && !strchr(tmp, 'n') && !strchr(tmp, 'i')) strcat(tmp, ".0");
    /* jansson strips the '+' and leading zeros of the exponent: 1e-05 -> 1e-5 */
    char *e = strchr(tmp, 'e');
    if (e) {
        char *r = e + 1, *w = e + 1;
        if (*r == '-') { w++; r++; }
        else if (*r == '+') r++;
        while (*r == '0' && r[1]) r++;
        memmove(w, r, strlen(r) + 1);
    }
    jbuf_puts(b, tmp);
}

char *jbuf_take(jbuf *b)
{
    if (!b->p) { jbuf_reserve(b, 0); b->p[0] = 0; }
    char *r = b->p;
    b->p = NULL; b->len = b->cap = 0;
    return r;
}
