#include "reduce.h"
#include "par.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void reduction_init(reduction *r) { memset(r, 0, sizeof(*r)); }

void reduction_clear(reduction *r)
{
    free(r->selection); free(r->weights); free(r->first_idx); free(r->second_idx);
    memset(r, 0, sizeof(*r));
}

static int exists(const jv *v) { return v && !jv_is_null(v); }

/* parsereduction.c:19-74; a present-but-null selection is an error there too */
static int parse_selection(reduction *r, int k, const char *name, const jv *root)
{
    if (!root) {
        r->selection_len = k;
        r->selection = malloc(sizeof(int) * (k > 0 ? k : 1));
        for (int i = 0; i < k; i++) r->selection[i] = i;
        return 0;
    }
    if (!jv_is_array(root)) { fprintf(stderr, "error: %s selection: the selection should be an array\n", name); return -1; }
    r->selection_len = (int)root->len;
    r->selection = malloc(sizeof(int) * (root->len > 0 ? root->len : 1));
    for (int i = 0; i < r->selection_len; i++) {
        const jv *x = &root->u.items[i];
        if (!jv_is_int(x)) { fprintf(stderr, "error: %s selection: each index in the selection must be an integer\n", name); return -1; }
        if (x->u.i < 0) { fprintf(stderr, "error: %s selection: each index in the selection must be non-negative\n", name); return -1; }
        if (x->u.i >= k) {
            fprintf(stderr, "error: %s selection: each index in the selection must be less than the total number of "
                            "available %s indices\n", name, name);
            return -1;
        }
        r->selection[i] = (int)x->u.i;
    }
    return 0;
}

/* parsereduction.c:77-159 */
static int parse_aggregation(reduction *r, const char *name, const jv *root)
{
    const char *valid = "{\"sum\", \"avg\", \"only\"}";
    if (!root) { r->agg_mode = AGG_NONE; return 0; }
    if (jv_is_string(root)) {
        if (!strcmp(root->u.s, "sum")) r->agg_mode = AGG_SUM;
        else if (!strcmp(root->u.s, "avg")) r->agg_mode = AGG_AVG;
        else if (!strcmp(root->u.s, "only")) {
            if (r->selection_len != 1) {
                fprintf(stderr, "error: %s aggregation (\"only\"): when using this aggregation mode, the selection length "
                                "must be exactly 1 (selection_len = %d)\n", name, r->selection_len);
                return -1;
            }
            r->agg_mode = AGG_ONLY;
        } else {
            fprintf(stderr, "error: %s aggregation (string): the only valid aggregation strings are %s\n", name, valid);
            return -1;
        }
        return 0;
    }
    if (jv_is_array(root)) {
        r->agg_mode = AGG_WEIGHTED_SUM;
        if ((int)root->len != r->selection_len) {
            fprintf(stderr, "error: %s aggregation (weighted sum): the number of weights must be equal to the number of "
                            "selected %s indices, or to the total number of %s indices if no selection was provided\n",
                    name, name, name);
            return -1;
        }
        r->weights = malloc(sizeof(double) * (root->len > 0 ? root->len : 1));
        for (uint32_t i = 0; i < root->len; i++) {
            const jv *x = &root->u.items[i];
            if (!jv_is_number(x)) { fprintf(stderr, "error: %s aggregation (weighted sum): weights should be numeric\n", name); return -1; }
            r->weights[i] = jv_number(x);
        }
        return 0;
    }
    fprintf(stderr, "error: %s aggregation: if provided, the aggregation should be either one of the strings %s or an "
                    "array of numeric weights for a weighted sum\n", name, valid);
    return -1;
}

int reduction_parse(reduction *r, int k, const char *name, const jv *root)
{
    const jv *v[2] = {NULL, NULL};
    if (root) {
        static const char *const keys[] = {"?selection", "?aggregation", NULL};
        if (jv_unpack_strict(root, keys, v)) return -1;
    }
    if (parse_selection(r, k, name, v[0])) return -1;
    return parse_aggregation(r, name, v[1]);
}

int reduction_parse_pairs(reduction *r, int k, const char *name, const jv *root)
{
    const jv *v[2] = {NULL, NULL};
    if (root) {
        static const char *const keys[] = {"?selection", "?aggregation", NULL};
        if (jv_unpack_strict(root, keys, v)) return -1;
    }
    if (exists(v[0])) {
        const jv *sel = v[0];
        if (!jv_is_array(sel)) { fprintf(stderr, "error: %s selection: the selection should be an array\n", name); return -1; }
        const int len = (int)sel->len;
        r->selection_len = len;
        r->selection = malloc(sizeof(int) * (len > 0 ? len : 1));
        r->first_idx = malloc(sizeof(int) * (len > 0 ? len : 1));
        r->second_idx = malloc(sizeof(int) * (len > 0 ? len : 1));
        for (int i = 0; i < len; i++) {
            const jv *p = &sel->u.items[i];
            if (!jv_is_array(p) || p->len != 2 || !jv_is_int(&p->u.items[0]) || !jv_is_int(&p->u.items[1])) {
                fprintf(stderr, "error: on line -1: expected a pair of integers\n");
                return -1;
            }
            for (int j = 0; j < 2; j++) {
                if (p->u.items[j].u.i < 0 || p->u.items[j].u.i >= k) {
                    fprintf(stderr, "error: %s selection: indices should be nonnegative integers less than the number "
                                    "of column elements\n", name);
                    return -1;
                }
            }
            r->first_idx[i] = (int)p->u.items[0].u.i;
            r->second_idx[i] = (int)p->u.items[1].u.i;
            r->selection[i] = i;
        }
        return parse_aggregation(r, name, v[1]);
    }
    int mode = -1;
    if (exists(v[1])) {
        if (jv_is_string(v[1])) {
            if (!strcmp(v[1]->u.s, "sum")) mode = AGG_SUM;
            else if (!strcmp(v[1]->u.s, "avg")) mode = AGG_AVG;
        }
    } else mode = AGG_NONE;
    if (mode == -1) {
        fprintf(stderr, "error: %s reduction (no selection): if no selection is specified, the only allowed aggregations "
                        "are \"sum\" or \"avg\"\n", name);
        return -1;
    }
    r->agg_mode = mode;
    const int len = k * (k - 1);
    r->selection_len = len;
    r->selection = malloc(sizeof(int) * (len > 0 ? len : 1));
    r->first_idx = malloc(sizeof(int) * (len > 0 ? len : 1));
    r->second_idx = malloc(sizeof(int) * (len > 0 ? len : 1));
    int i = 0;
    for (int a = 0; a < k; a++) for (int b = 0; b < k; b++) if (a != b) {
        r->selection[i] = i; r->first_idx[i] = a; r->second_idx[i] = b; i++;
    }
    return 0;
}

int axis_init(axis *a, const char *name, int n, const reduction *r)
{
    memset(a, 0, sizeof(*a));
    a->name = name; a->n = n; a->r = r;
    a->aggregated = r->agg_mode != AGG_NONE;
    a->requested = calloc(n > 0 ? n : 1, 1);
    for (int i = 0; i < r->selection_len; i++) a->requested[r->selection[i]] = 1;
    if (!a->aggregated) return 0;
    a->w = calloc(n > 0 ? n : 1, sizeof(double));
    long double *acc = calloc(n > 0 ? n : 1, sizeof(long double));
    long double div = 1;
    if (r->agg_mode == AGG_WEIGHTED_SUM) {
        for (int i = 0; i < r->selection_len; i++) acc[r->selection[i]] += r->weights[i];
    } else if (r->agg_mode == AGG_SUM || r->agg_mode == AGG_AVG) {
        for (int i = 0; i < r->selection_len; i++) acc[r->selection[i]] += 1;
        if (r->agg_mode == AGG_AVG) div = r->selection_len;
    } else if (r->agg_mode == AGG_ONLY) {
        if (r->selection_len != 1) {
            fprintf(stderr, "error: when using {\"aggregation\" : \"only\"}, the selection length must be exactly 1\n");
            free(acc);
            return -1;
        }
        acc[r->selection[0]] = 1;
    }
    for (int i = 0; i < n; i++) a->w[i] = (double)(acc[i] / div);
    free(acc);
    return 0;
}

void axis_clear(axis *a) { free(a->w); free(a->requested); memset(a, 0, sizeof(*a)); }

/* rows of the output table in the reference's order (axes outermost first); at axis `split` only the entries
 * [lo, hi) of its selection are visited, so that several threads can each write a contiguous run of rows */
typedef struct { int split, lo, hi; } row_range;

static void emit_rows(jbuf *b, const axis *axes, int ndim, const double *values, int d, size_t offset,
                      const size_t *strides, long long *prefix, int *nprefix, int *first, const row_range *rr)
{
    if (d == ndim) {
        jbuf_puts(b, *first ? "[" : ", [");
        *first = 0;
        for (int i = 0; i < *nprefix; i++) { jbuf_int(b, prefix[i]); jbuf_puts(b, ", "); }
        jbuf_real(b, values[offset]);
        jbuf_puts(b, "]");
        return;
    }
    const axis *a = &axes[d];
    if (a->aggregated) { emit_rows(b, axes, ndim, values, d + 1, offset, strides, prefix, nprefix, first, rr); return; }
    const int i0 = d == rr->split ? rr->lo : 0, i1 = d == rr->split ? rr->hi : a->r->selection_len;
    for (int i = i0; i < i1; i++) {
        int idx = a->r->selection[i];
        int saved = *nprefix;
        if (a->ncomp) for (int c = 0; c < a->ncomp; c++) prefix[(*nprefix)++] = a->comp_idx[c][i];
        else prefix[(*nprefix)++] = idx;
        emit_rows(b, axes, ndim, values, d + 1, offset + (size_t)idx * strides[d], strides, prefix, nprefix, first, rr);
        *nprefix = saved;
    }
}

typedef struct { const axis *axes; int ndim; const double *values; const size_t *strides; int split; jbuf *parts; } emit_job;

static void emit_worker(int tid, int nthreads, void *ctx)
{
    emit_job *job = ctx;
    const int len = job->axes[job->split].r->selection_len;
    row_range rr = {job->split, (int)((long long)len * tid / nthreads), (int)((long long)len * (tid + 1) / nthreads)};
    long long prefix[16];
    int nprefix = 0, first = 1;
    jbuf_init(&job->parts[tid]);
    emit_rows(&job->parts[tid], job->axes, job->ndim, job->values, 0, 0, job->strides, prefix, &nprefix, &first, &rr);
}

char *table_to_json(const axis *axes, int ndim, const double *values)
{
    jbuf b; jbuf_init(&b);
    size_t strides[8];
    size_t st = 1;
    for (int d = ndim - 1; d >= 0; d--) { strides[d] = st; st *= axes[d].aggregated ? 1 : (size_t)axes[d].n; }
    jbuf_puts(&b, "{\"columns\": [");
    for (int d = 0; d < ndim; d++) {
        if (axes[d].aggregated) continue;
        if (axes[d].ncomp) for (int c = 0; c < axes[d].ncomp; c++) { jbuf_puts(&b, "\""); jbuf_puts(&b, axes[d].comp_name[c]); jbuf_puts(&b, "\", "); }
        else { jbuf_puts(&b, "\""); jbuf_puts(&b, axes[d].name); jbuf_puts(&b, "\", "); }
    }
    jbuf_puts(&b, "\"value\"], \"data\": [");
    /* a table with the site axis kept has one row per site pattern (times edges, nodes, states): printing the reals
     * is then the longest phase of the call, so the outermost kept axis is cut into one run of rows per host thread */
    int split = -1;
    size_t rows = 1;
    for (int d = 0; d < ndim; d++) {
        if (axes[d].aggregated) continue;
        if (split < 0) split = d;
        rows *= (size_t)axes[d].r->selection_len;
    }
    int nthreads = (split >= 0 && rows >= 65536) ? par_threads() : 1;
    if (split >= 0 && nthreads > axes[split].r->selection_len) nthreads = axes[split].r->selection_len;
    if (nthreads <= 1) {
        long long prefix[16];
        int nprefix = 0, first = 1;
        row_range all = {-1, 0, 0};
        emit_rows(&b, axes, ndim, values, 0, 0, strides, prefix, &nprefix, &first, &all);
    } else {
        emit_job job = {axes, ndim, values, strides, split, calloc((size_t)nthreads, sizeof(jbuf))};
        if (!job.parts) { fprintf(stderr, "out of memory building JSON output\n"); abort(); }
        par_run(nthreads, emit_worker, &job);
        int first = 1;
        for (int t = 0; t < nthreads; t++) {
            if (!job.parts[t].len) { free(job.parts[t].p); continue; }
            if (!first) jbuf_puts(&b, ", ");
            first = 0;
            jbuf_append(&b, job.parts[t].p, job.parts[t].len);
            free(job.parts[t].p);
        }
        free(job.parts);
    }
    jbuf_puts(&b, "]}");
    return jbuf_take(&b);
}
