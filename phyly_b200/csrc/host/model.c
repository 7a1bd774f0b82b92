#include "model.h"
#include "par.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "plf.h"
#include "dd.h"

void plf_model_init(plf_model *m)
{
    memset(m, 0, sizeof(*m));
    m->root = -1;
    m->rate_divisor = 1.0;
    m->gamma_shape = 1.0;
    m->gamma_categories = 1;
    m->code_bytes = 1;
}

void plf_model_clear(plf_model *m)
{
    free(m->indptr); free(m->indices); free(m->order); free(m->idx_to_user); free(m->preorder);
    free(m->rate_matrix); free(m->edge_rate_user); free(m->root_custom);
    free(m->mix_rates); free(m->mix_prior); free(m->defs);
    if (m->codes_shared) json_codes_release(m->codes_shared); else free(m->codes);
    plf_model_init(m);
}

static int exists(const jv *v) { return v && !jv_is_null(v); }

/* _validate_nonnegative_array, parsemodel.c:26-75 */
static int nonneg_array(double *dest, int want, const jv *root)
{
    if (!jv_is_array(root)) { fprintf(stderr, "_validate_nonnegative_array: not an array\n"); return -1; }
    if ((int)root->len != want) {
        fprintf(stderr, "_validate_nonnegative_array: unexpected array length (actual: %d desired: %d)\n",
                (int)root->len, want);
        return -1;
    }
    for (int i = 0; i < want; i++) {
        const jv *x = &root->u.items[i];
        if (!jv_is_number(x)) { fprintf(stderr, "_validate_nonnegative_array: not a number\n"); return -1; }
        double d = jv_number(x);
        if (d < 0) { fprintf(stderr, "_validate_nonnegative_array: array entries must be nonnegative\n"); return -1; }
        dest[i] = d;
    }
    return 0;
}

/* "[i, i]" with JSON_STRICT, parsemodel.c:191-207 */
static int int_pair(int pair[2], const jv *v)
{
    if (!jv_is_array(v)) { fprintf(stderr, "error: on line -1: Expected array, got something else\n"); return -1; }
    if (v->len != 2) { fprintf(stderr, "error: on line -1: expected an array of exactly two integers\n"); return -1; }
    for (int j = 0; j < 2; j++) {
        if (!jv_is_int(&v->u.items[j])) { fprintf(stderr, "error: on line -1: Expected integer\n"); return -1; }
        int64_t x = v->u.items[j].u.i;
        if (x < -2147483647 || x > 2147483647) { fprintf(stderr, "error: on line -1: integer out of range\n"); return -1; }
        pair[j] = (int)x;
    }
    return 0;
}

/* _validate_edges, parsemodel.c:210-368, with csr_graph.c and model.c:23-45 */
static int parse_edges(plf_model *m, const jv *root)
{
    int rc = -1;
    int *in_deg = NULL, *out_deg = NULL, *fill = NULL, *pairs = NULL, *visited = NULL, *u = NULL, *v = NULL;
    if (!jv_is_array(root)) { fprintf(stderr, "_validate_edges: not an array\n"); return -1; }
    const int E = (int)root->len, N = E + 1;
    in_deg = calloc(N, sizeof(int)); out_deg = calloc(N, sizeof(int)); fill = calloc(N, sizeof(int));
    pairs = malloc(sizeof(int) * 2 * (E > 0 ? E : 1));
    for (int i = 0; i < E; i++) {
        int pr[2];
        if (int_pair(pr, &root->u.items[i])) goto finish;
        for (int j = 0; j < 2; j++) {
            if (pr[j] < 0 || pr[j] >= N) {
                fprintf(stderr, "_validate_edges: node indices must be integers no less than 0 and no greater "
                                "than the number of edges\n");
                goto finish;
            }
        }
        if (pr[0] == pr[1]) { fprintf(stderr, "_validate_edges: edges cannot be loops\n"); goto finish; }
        out_deg[pr[0]]++; in_deg[pr[1]]++;
        pairs[2 * i] = pr[0]; pairs[2 * i + 1] = pr[1];
    }
    {
        int root_count = 0;
        for (int i = 0; i < N; i++) if (!in_deg[i]) { m->root = i; root_count++; }
        if (root_count != 1) { fprintf(stderr, "_validate_edges: exactly one node should have in-degree 0\n"); goto finish; }
    }
    for (int i = 0; i < N; i++) if (in_deg[i] > 1) {
        fprintf(stderr, "_validate_edges: the in-degree of each node must be 0 or 1\n"); goto finish;
    }
    for (int i = 0; i < N; i++) if (in_deg[i] + out_deg[i] < 1) {
        fprintf(stderr, "_validate_edges: node index %d is not an endpoint of any edge\n", i); goto finish;
    }
    m->N = N; m->E = E;
    m->indptr = malloc(sizeof(int) * (N + 1));
    m->indices = malloc(sizeof(int) * (E > 0 ? E : 1));
    m->order = malloc(sizeof(int) * (E > 0 ? E : 1));
    m->idx_to_user = malloc(sizeof(int) * (E > 0 ? E : 1));
    m->preorder = malloc(sizeof(int) * N);
    m->indptr[0] = 0;
    for (int i = 0; i < N; i++) m->indptr[i + 1] = m->indptr[i] + out_deg[i];
    /* children in user edge order within each parent; user edge -> csr idx (csr_graph.c:28-46) */
    for (int i = 0; i < E; i++) {
        int a = pairs[2 * i], b = pairs[2 * i + 1];
        int idx = m->indptr[a] + fill[a]++;
        m->indices[idx] = b;
        m->order[i] = idx;
        m->idx_to_user[idx] = i;
    }
    /* BFS level order from the root (csr_graph.c:102-177) */
    visited = calloc(N, sizeof(int)); u = malloc(sizeof(int) * N); v = malloc(sizeof(int) * N);
    {
        int npre = 0, nv = 1, nu;
        v[0] = m->root; visited[m->root] = 1;
        while (nv) {
            int *tmp = u; u = v; v = tmp;
            nu = nv; nv = 0;
            for (int i = 0; i < nu; i++) {
                int a = u[i];
                m->preorder[npre++] = a;
                for (int j = m->indptr[a]; j < m->indptr[a + 1]; j++) {
                    int b = m->indices[j];
                    if (visited[b]) {
                        fprintf(stderr, "validate_edges: topo sort failed: node index %d already visited\n", b);
                        goto finish;
                    }
                    v[nv++] = b; visited[b] = 1;
                }
            }
        }
        if (npre != N) {
            fprintf(stderr, "validate_edges: the topo sort contains %d of the %d nodes\n", npre, N);
            goto finish;
        }
    }
    rc = 0;
finish:
    free(in_deg); free(out_deg); free(fill); free(pairs); free(visited); free(u); free(v);
    return rc;
}

static int parse_rate_matrix(plf_model *m, const jv *root)
{
    if (!jv_is_array(root)) { fprintf(stderr, "_validate_rate_matrix: not an array\n"); return -1; }
    const int n = (int)root->len;
    m->n = n;
    m->rate_matrix = malloc(sizeof(double) * (n > 0 ? n * n : 1));
    for (int i = 0; i < n; i++) {
        const jv *x = &root->u.items[i];
        if (!jv_is_array(x)) { fprintf(stderr, "_validate_rate_matrix: this row is not an array\n"); return -1; }
        if ((int)x->len != n) {
            fprintf(stderr, "_validate_rate_matrix: this row length does not match the number of rows: "
                            "(actual: %d desired: %d)\n", (int)x->len, n);
            return -1;
        }
        for (int j = 0; j < n; j++) {
            const jv *y = &x->u.items[j];
            if (!jv_is_number(y)) { fprintf(stderr, "_validate_rate_matrix: not a number\n"); return -1; }
            double d = jv_number(y);
            if (d < 0) { fprintf(stderr, "_validate_rate_matrix: rate matrix entries must be nonnegative\n"); return -1; }
            m->rate_matrix[i * n + j] = d;
        }
    }
    return 0;
}

/* open-addressing table of distinct n-vectors (bitwise identity) */
typedef struct { int *slots; size_t cap; double *rows; int count, rowcap, n; } rowtab;

static uint64_t row_hash(const double *r, int n)
{
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; i++) {
        uint64_t b; memcpy(&b, &r[i], 8);
        h ^= b; h *= 1099511628211ull; h ^= h >> 29;
    }
    return h;
}

static int rowtab_intern(rowtab *t, const double *r)
{
    if ((size_t)(t->count + 1) * 2 > t->cap) {
        size_t nc = t->cap ? t->cap * 2 : 64;
        int *ns = malloc(sizeof(int) * nc);
        for (size_t i = 0; i < nc; i++) ns[i] = -1;
        for (int k = 0; k < t->count; k++) {
            size_t pos = row_hash(t->rows + (size_t)k * t->n, t->n) & (nc - 1);
            while (ns[pos] != -1) pos = (pos + 1) & (nc - 1);
            ns[pos] = k;
        }
        free(t->slots); t->slots = ns; t->cap = nc;
    }
    size_t pos = row_hash(r, t->n) & (t->cap - 1);
    while (t->slots[pos] != -1) {
        if (!memcmp(t->rows + (size_t)t->slots[pos] * t->n, r, sizeof(double) * t->n)) return t->slots[pos];
        pos = (pos + 1) & (t->cap - 1);
    }
    if (t->count == t->rowcap) {
        t->rowcap = t->rowcap ? t->rowcap * 2 : 16;
        t->rows = realloc(t->rows, sizeof(double) * (size_t)t->rowcap * t->n);
    }
    memcpy(t->rows + (size_t)t->count * t->n, r, sizeof(double) * t->n);
    t->slots[pos] = t->count;
    return t->count++;
}

static void compress_codes(plf_model *m, int *codes32)
{
    const size_t total = (size_t)m->S * m->N;
    if (m->K <= 256) {
        unsigned char *c8 = malloc(total ? total : 1);
        for (size_t i = 0; i < total; i++) c8[i] = (unsigned char)codes32[i];
        free(codes32);
        m->codes = c8; m->code_bytes = 1;
    } else {
        m->codes = codes32; m->code_bytes = 4;
    }
}

/* parsemodel.c:458-510; the dense array is stored as codes into its distinct rows */
static int parse_probability_array(plf_model *m, const jv *root)
{
    const char *name = "_validate_probability_array";
    if (!jv_is_array(root)) { fprintf(stderr, "%s: expected an array\n", name); return -1; }
    const int n = m->n, N = m->N;
    m->S = root->len;
    int *codes = malloc(sizeof(int) * ((size_t)m->S * N + 1));
    double *row = malloc(sizeof(double) * (n > 0 ? n : 1));
    rowtab tab; memset(&tab, 0, sizeof tab); tab.n = n;
    int rc = -1;
    for (int64_t i = 0; i < m->S; i++) {
        const jv *x = &root->u.items[i];
        if (!jv_is_array(x)) { fprintf(stderr, "%s: expected an array\n", name); goto finish; }
        if ((int)x->len != N) {
            fprintf(stderr, "%s: failed to match the number of nodes: (actual: %d desired: %d)\n", name, (int)x->len, N);
            goto finish;
        }
        for (int j = 0; j < N; j++) {
            if (nonneg_array(row, n, &x->u.items[j])) { fprintf(stderr, "%s: array validation has failed\n", name); goto finish; }
            codes[(size_t)i * N + j] = rowtab_intern(&tab, row);
        }
    }
    m->K = tab.count;
    if (m->K == 0) { /* zero sites: keep a dummy all-ones row so that the tables are never empty */
        for (int k = 0; k < n; k++) row[k] = 1.0;
        rowtab_intern(&tab, row); m->K = 1;
    }
    m->defs = tab.rows; tab.rows = NULL;
    compress_codes(m, codes); codes = NULL;
    rc = 0;
finish:
    free(codes); free(row); free(tab.slots); free(tab.rows);
    return rc;
}

/* one thread's share of parsemodel.c:590-615: rows [S tid / T, S (tid+1) / T) */
typedef struct { int kind, actual, max_code; } cd_fault;
typedef struct { const jv *rows; int64_t S; int N, K, wide; void *codes; cd_fault *faults; } cd_job;

static void cd_worker(int tid, int nthreads, void *ctx)
{
    cd_job *job = ctx;
    const int N = job->N;
    const int64_t r0 = job->S * tid / nthreads, r1 = job->S * (tid + 1) / nthreads;
    cd_fault *f = &job->faults[tid];
    for (int64_t i = r0; i < r1; i++) {
        const jv *x = &job->rows[i];
        if (!jv_is_array(x)) { f->kind = 1; return; }
        if ((int)x->len != N) { f->kind = 2; f->actual = (int)x->len; return; }
        const jv *y = x->u.items;
        unsigned char *c8 = (unsigned char *)job->codes + (size_t)i * N;
        int *c32 = (int *)job->codes + (size_t)i * N;
        for (int j = 0; j < N; j++) {
            if (!jv_is_int(&y[j])) { f->kind = 3; return; }
            if (y[j].u.i < 0) { f->kind = 4; return; }
            if (y[j].u.i >= job->K) { f->kind = 5; return; }
            if (job->wide) c32[j] = (int)y[j].u.i; else c8[j] = (unsigned char)y[j].u.i;
            if ((int)y[j].u.i > f->max_code) f->max_code = (int)y[j].u.i;
        }
    }
}

/* parsemodel.c:514-628 */
static int parse_character_data(plf_model *m, const jv *data, const jv *defs)
{
    const char *name = "_validate_character_data_and_definitions";
    if (!exists(defs))
        fprintf(stderr, "%s: 'character_data' has been provided without 'character_definitions'\n", name);
    if (!jv_is_array(data)) { fprintf(stderr, "%s: expected 'character_data' to be an array\n", name); return -1; }
    if (!jv_is_array(defs)) { fprintf(stderr, "%s: expected 'character_definitions' to be an array\n", name); return -1; }
    const int n = m->n, N = m->N;
    m->S = data->len;
    m->K = (int)defs->len;
    m->defs = malloc(sizeof(double) * ((size_t)m->K * n + 1));
    for (int j = 0; j < m->K; j++) {
        if (nonneg_array(m->defs + (size_t)j * n, n, &defs->u.items[j])) {
            fprintf(stderr, "%s: validation of definition of character %d of %d has failed\n", name, j, m->K);
            return -1;
        }
    }
    /* rows -> codes, several threads for an alignment-sized matrix; the first offence in row-major order is the one
     * reported, as in the reference's single loop */
    const int wide = m->K > 256;
    jv_codes *cached = jv_compact_codes(data);
    if (cached) {
        /* the reader recognised the text of the matrix read last time (json.h): every row is an array of row_len
         * non-negative integers, the largest being max_code */
        m->S = cached->rows;
        if (m->S > 0 && cached->row_len != N) {
            fprintf(stderr, "%s: failed to match the number of nodes: (actual: %d desired: %d)\n", name, cached->row_len, N);
            return -1;
        }
        if (m->S > 0 && cached->max_code >= m->K) {
            fprintf(stderr, "%s: character indices must each be less than the character count (%d)\n", name, m->K);
            return -1;
        }
        if (cached->code_bytes == (wide ? 4 : 1)) {
            json_codes_retain(cached);
            m->codes = cached->codes; m->code_bytes = cached->code_bytes; m->codes_shared = cached;
        } else {
            /* the definitions crossed 256 rows since: same codes, other width */
            const size_t cnt = (size_t)m->S * N;
            void *conv = malloc((wide ? sizeof(int) : 1) * (cnt + 1));
            if (!conv) { fprintf(stderr, "%s: out of memory\n", name); return -1; }
            for (size_t i = 0; i < cnt; i++) {
                const int v = cached->code_bytes == 1 ? ((const unsigned char *)cached->codes)[i] : ((const int *)cached->codes)[i];
                if (wide) ((int *)conv)[i] = v; else ((unsigned char *)conv)[i] = (unsigned char)v;
            }
            m->codes = conv; m->code_bytes = wide ? 4 : 1;
        }
        return 0;
    }
    const size_t total = (size_t)m->S * N;
    void *codes = malloc((wide ? sizeof(int) : 1) * (total + 1));
    if (!codes) { fprintf(stderr, "%s: out of memory\n", name); return -1; }
    int nthreads = total >= ((size_t)1 << 20) ? par_threads() : 1;
    if ((int64_t)nthreads > m->S) nthreads = m->S > 0 ? (int)m->S : 1;
    cd_fault *faults = calloc((size_t)nthreads, sizeof(cd_fault));
    if (!faults) { fprintf(stderr, "%s: out of memory\n", name); free(codes); return -1; }
    cd_job job = {data->u.items, m->S, N, m->K, wide, codes, faults};
    par_run(nthreads, cd_worker, &job);
    const cd_fault *f = NULL;
    for (int t = 0; t < nthreads && !f; t++) if (faults[t].kind) f = &faults[t];     /* ranges ascend with t */
    if (f) {
        switch (f->kind) {
        case 1: fprintf(stderr, "%s: expected an array\n", name); break;
        case 2: fprintf(stderr, "%s: failed to match the number of nodes: (actual: %d desired: %d)\n", name, f->actual, N); break;
        case 3: fprintf(stderr, "%s: character indices must be integers\n", name); break;
        case 4: fprintf(stderr, "%s: character indices must be non-negative\n", name); break;
        default: fprintf(stderr, "%s: character indices must each be less than the character count (%d)\n", name, m->K);
        }
        free(faults); free(codes);
        return -1;
    }
    int *faults_max = malloc(sizeof(int) * (size_t)nthreads);
    for (int t = 0; faults_max && t < nthreads; t++) faults_max[t] = faults[t].max_code;
    free(faults);
    if (!faults_max) { fprintf(stderr, "%s: out of memory\n", name); free(codes); return -1; }
    if (m->K == 0) {   /* only reachable with zero sites */
        free(m->defs);
        m->defs = malloc(sizeof(double) * (n > 0 ? n : 1));
        for (int k = 0; k < n; k++) m->defs[k] = 1.0;
        m->K = 1;
    }
    m->codes = codes; m->code_bytes = wide ? 4 : 1;
    {
        /* a large matrix whose text the reader has hashed: share the codes with the reader's cache, so that the same text
         * in the next call costs neither reading nor conversion (nor, in drivers.c, an upload) */
        uint64_t h1, h2;
        size_t text_len;
        if (m->S > 0 && jv_matrix_text_hash(data, &h1, &h2, &text_len)) {
            int max_code = 0;
            for (int t = 0; t < nthreads; t++) if (faults_max[t] > max_code) max_code = faults_max[t];
            jv_codes *d = json_codes_publish(h1, h2, text_len, codes, m->code_bytes, m->S, N, max_code);
            if (d) m->codes_shared = d;
        }
        free(faults_max);
    }
    return 0;
}

static int parse_gamma_mixture(plf_model *m, const jv *root)
{
    static const char *const keys[] = {"gamma_shape", "gamma_categories", "?invariable_prior", NULL};
    const jv *v[3];
    if (jv_unpack_strict(root, keys, v)) return -1;
    if (!v[2] || jv_is_null(v[2])) m->invariable_prior = 0;
    else if (jv_is_number(v[2])) m->invariable_prior = jv_number(v[2]);
    else { fprintf(stderr, "invariable_prior: not a number\n"); return -1; }
    if (jv_is_number(v[0])) m->gamma_shape = jv_number(v[0]);
    else { fprintf(stderr, "gamma_shape: not a number\n"); return -1; }
    if (jv_is_int(v[1])) m->gamma_categories = (int)v[1]->u.i;
    else { fprintf(stderr, "gamma_categories: not an integer\n"); return -1; }
    /* the reference would loop or abort on these; reject them cleanly */
    if (m->gamma_categories < 1 || !(m->gamma_shape > 0) || m->invariable_prior < 0 || m->invariable_prior >= 1) {
        fprintf(stderr, "gamma rate mixture: need gamma_categories >= 1, gamma_shape > 0 and 0 <= invariable_prior < 1\n");
        return -1;
    }
    return 0;
}

static int parse_rate_mixture(plf_model *m, const jv *root)
{
    static const char *const keys[] = {"rates", "prior", NULL};
    const jv *v[2];
    if (jv_unpack_strict(root, keys, v)) return -1;
    if (!jv_is_array(v[0])) { fprintf(stderr, "_validate_rate_mixture: 'rates' is not an array\n"); return -1; }
    const int k = (int)v[0]->len;
    m->mix_n = k;
    m->mix_rates = malloc(sizeof(double) * (k > 0 ? k : 1));
    m->mix_prior = malloc(sizeof(double) * (k > 0 ? k : 1));
    if (nonneg_array(m->mix_rates, k, v[0])) { fprintf(stderr, "_validate_rate_mixture: invalid 'rates' array\n"); return -1; }
    const char *msg = "_validate_rate_mixture: the 'prior' argument must be either a nonnegative array or the string "
                      "\"uniform_distribution\"\n";
    if (jv_is_string(v[1])) {
        if (!strcmp(v[1]->u.s, "uniform_distribution")) m->mix_mode = MIX_UNIFORM;
        else { fputs(msg, stderr); return -1; }
    } else if (jv_is_array(v[1])) {
        if (nonneg_array(m->mix_prior, k, v[1])) { fprintf(stderr, "_validate_rate_mixture: invalid 'prior' array\n"); return -1; }
        m->mix_mode = MIX_CUSTOM;
    } else {
        /* the reference leaves the mode undefined and aborts later (rate_mixture.c:262-266) */
        fputs(msg, stderr);
        return -1;
    }
    if (k < 1) { fprintf(stderr, "_validate_rate_mixture: at least one rate category is required\n"); return -1; }
    return 0;
}

int plf_model_parse(plf_model *m, const jv *md)
{
    static const char *const keys[] = {
        "edges", "edge_rate_coefficients", "rate_matrix", "?probability_array", "?character_definitions",
        "?character_data", "?rate_divisor", "?root_prior", "?rate_mixture", "?gamma_rate_mixture",
        "?normalized_median_gamma_rate_mixture", NULL};
    const jv *v[11];
    if (jv_unpack_strict(md, keys, v)) return -1;
    const jv *edges = v[0], *erc = v[1], *rmat = v[2], *parr = v[3], *cdefs = v[4], *cdata = v[5];
    const jv *rdiv = v[6], *rprior = v[7], *rmix = v[8], *gmix = v[9], *gmed = v[10];
    int mixtures = exists(rmix) + exists(gmix) + exists(gmed);
    if (mixtures > 1) { fprintf(stderr, "error: conflicting rate mixture options\n"); return -1; }
    if (exists(parr) && exists(cdata)) {
        fprintf(stderr, "error: the mutually exclusive options 'probability_array' and 'character_data' have both been specified\n");
        return -1;
    }
    if (exists(parr) && exists(cdefs)) {
        fprintf(stderr, "error: the mutually exclusive options 'probability_array' and 'character_definitions' have both been specified\n");
        return -1;
    }
    if (parse_edges(m, edges)) return -1;
    m->edge_rate_user = malloc(sizeof(double) * (m->E > 0 ? m->E : 1));
    if (nonneg_array(m->edge_rate_user, m->E, erc)) {
        fprintf(stderr, "_validate_edge_rate_coefficients: array validation has failed\n");
        return -1;
    }
    if (parse_rate_matrix(m, rmat)) return -1;
    if (exists(parr)) { if (parse_probability_array(m, parr)) return -1; }
    else if (exists(cdata)) { if (parse_character_data(m, cdata, cdefs)) return -1; }
    else { fprintf(stderr, "error: either 'probability_array' or 'character_data' must be specified\n"); return -1; }
    /* rate divisor, parsemodel.c:82-126 */
    if (exists(rdiv)) {
        const char *msg = "_validate_rate_divisor: the optional rate_divisor argument must be either a positive number "
                          "or the string \"equilibrium_exit_rate\"\n";
        if (jv_is_string(rdiv)) {
            if (!strcmp(rdiv->u.s, "equilibrium_exit_rate")) m->use_eq_divisor = 1;
            else { fputs(msg, stderr); return -1; }
        } else if (jv_is_number(rdiv)) {
            double d = jv_number(rdiv);
            if (d <= 0) { fputs(msg, stderr); return -1; }
            m->rate_divisor = d;
        } else { fputs(msg, stderr); return -1; }
    }
    /* root prior, parsemodel.c:129-188 */
    if (!exists(rprior)) m->root_mode = PLF_ROOT_NONE;
    else if (jv_is_string(rprior)) {
        if (!strcmp(rprior->u.s, "equilibrium_distribution")) m->root_mode = PLF_ROOT_EQUILIBRIUM;
        else if (!strcmp(rprior->u.s, "uniform_distribution")) m->root_mode = PLF_ROOT_UNIFORM;
        else { fprintf(stderr, "_validate_root_prior: unrecognised string\n"); return -1; }
    } else {
        m->root_mode = PLF_ROOT_CUSTOM;
        m->root_custom = malloc(sizeof(double) * (m->n > 0 ? m->n : 1));
        if (nonneg_array(m->root_custom, m->n, rprior)) {
            fprintf(stderr, "_validate_root_prior: the optional \"root_prior\" must be either a list of probabilities "
                            "or one of the strings {\"equilibrium_distribution\", \"uniform_distribution\"}\n");
            return -1;
        }
    }
    if (exists(gmix)) { m->mix_mode = MIX_GAMMA; if (parse_gamma_mixture(m, gmix)) return -1; }
    else if (exists(gmed)) { m->mix_mode = MIX_GAMMA_MEDIAN; if (parse_gamma_mixture(m, gmed)) return -1; }
    else if (exists(rmix)) { if (parse_rate_mixture(m, rmix)) return -1; }
    else m->mix_mode = MIX_NONE;
    if (m->n < 1) { fprintf(stderr, "_validate_rate_matrix: the rate matrix is empty\n"); return -1; }
    return 0;
}

int plf_model_category_count(const plf_model *m)
{
    if (m->mix_mode == MIX_NONE) return 1;
    if (m->mix_mode == MIX_UNIFORM || m->mix_mode == MIX_CUSTOM) return m->mix_n;
    return m->gamma_categories + (m->invariable_prior ? 1 : 0);
}

/* ------------------------------------------------------------------ */
/* gamma discretisation in extended precision                          */
/* ------------------------------------------------------------------ */

/* regularised lower incomplete gamma P(s, x), s > 0, x >= 0 */
static long double gamma_p(long double s, long double x)
{
    if (x <= 0) return 0;
    if (isinf(x)) return 1;
    const long double lg = lgammal(s + 1);          /* log Gamma(s+1) */
    if (x < s + 1) {
        /* series: x^s e^-x / Gamma(s+1) * sum x^k / ((s+1)...(s+k)) */
        long double term = 1, sum = 1;
        for (int k = 1; k < 100000; k++) {
            term *= x / (s + k);
            sum += term;
            if (term < sum * 1e-22L) break;
        }
        return expl(s * logl(x) - x - lg) * sum;
    }
    /* continued fraction for Q(s, x) (modified Lentz) */
    const long double tiny = 1e-4000L;
    long double b = x + 1 - s, c = 1 / tiny, d = 1 / b, h = d;
    for (int i = 1; i < 100000; i++) {
        long double an = -(long double)i * ((long double)i - s);
        b += 2;
        d = an * d + b; if (fabsl(d) < tiny) d = tiny;
        c = b + an / c; if (fabsl(c) < tiny) c = tiny;
        d = 1 / d;
        long double del = d * c;
        h *= del;
        if (fabsl(del - 1) < 1e-21L) break;
    }
    long double q = expl(s * logl(x) - x - lgammal(s)) * h;
    return 1 - q;
}

/* quantile: P(s, q) = k/nq; gamma_discretization.c:208-282 (bracket, then refine) */
static long double gamma_quantile(int k, int nq, long double s)
{
    const long double c = (long double)k / nq;
    /* work in u = log x; the small-x asymptotic P ~ x^s/Gamma(s+1) gives the start */
    long double u0 = (logl(c) + lgammal(s + 1)) / s;
    if (u0 < -11000) return 0;       /* below the smallest long double: the quantile is numerically zero */
    long double lo = u0 - 1, hi = (u0 > 0 ? u0 : 0) + 1;
    for (int it = 0; it < 200 && gamma_p(s, expl(lo)) > c; it++) { lo -= fabsl(lo) > 1 ? fabsl(lo) : 1; if (lo < -11300) return 0; }
    for (int it = 0; it < 200 && gamma_p(s, expl(hi)) < c; it++) hi += fabsl(hi) > 1 ? fabsl(hi) : 1;
    long double u = u0 < lo ? lo : (u0 > hi ? hi : u0);
    for (int it = 0; it < 400; it++) {
        long double x = expl(u);
        long double f = gamma_p(s, x) - c;
        if (f > 0) hi = u; else lo = u;
        /* Newton step in u: dP/du = x^s e^-x / Gamma(s) */
        long double dpdu = expl(s * u - x - lgammal(s));
        long double un = (dpdu > 0) ? u - f / dpdu : (lo + hi) / 2;
        if (!(un > lo && un < hi)) un = (lo + hi) / 2;
        if (fabsl(un - u) <= 4e-19L * (fabsl(u) > 1 ? fabsl(u) : 1)) { u = un; break; }
        u = un;
        if (hi - lo <= 4e-19L * (fabsl(u) > 1 ? fabsl(u) : 1)) break;
    }
    return expl(u);
}

int plf_gamma_rates(double *rates, int n, double shape, int median)
{
    const long double s = shape;
    if (n < 1 || !(shape > 0)) return -1;
    if (!median) {
        /* gamma_rates, gamma_discretization.c:298-332 */
        long double prev = 0;
        for (int k = 0; k < n; k++) {
            long double q = (k + 1 < n) ? gamma_quantile(k + 1, n, s) : INFINITY;
            long double ex = (k + 1 < n) ? gamma_p(s + 1, q) : 1;
            rates[k] = (double)((ex - prev) * n);
            prev = ex;
        }
    } else {
        /* normalized_median_gamma_rates, gamma_discretization.c:334-370 */
        long double *q = malloc(sizeof(long double) * n), tot = 0;
        for (int k = 0; k < n; k++) { q[k] = gamma_quantile(2 * k + 1, 2 * n, s); tot += q[k]; }
        for (int k = 0; k < n; k++) rates[k] = (double)(q[k] / tot * n);
        free(q);
    }
    for (int k = 0; k < n; k++) if (!isfinite(rates[k])) return -1;
    return 0;
}

/* ------------------------------------------------------------------ */
/* derived, site independent quantities                                */
/* ------------------------------------------------------------------ */

void plf_derived_clear(plf_derived *d)
{
    free(d->cat_rates); free(d->cat_prior); free(d->equilibrium); free(d->q_hi); free(d->q_lo);
    free(d->edge_rates_csr); free(d->root_vec);
    memset(d, 0, sizeof(*d));
}

/* equilibrium.c:20-88: solve [Q^T e; e^T 0] x = 1 ignoring the diagonal of Q */
static int equilibrium(long double *pi, const double *raw, int n)
{
    const int m = n + 1;
    long double *R = calloc((size_t)m * (m + 1), sizeof(long double));
    for (int i = 0; i < n; i++) {
        long double ex = 0;
        for (int j = 0; j < n; j++) if (i != j) { R[i * (m + 1) + j] = raw[j * n + i]; ex += raw[i * n + j]; }
        R[i * (m + 1) + i] = -ex;
        R[n * (m + 1) + i] = 1;
        R[i * (m + 1) + n] = 1;
    }
    for (int i = 0; i < m; i++) R[i * (m + 1) + m] = 1;   /* right-hand side */
    for (int col = 0; col < m; col++) {
        int piv = col;
        for (int r = col + 1; r < m; r++) if (fabsl(R[r * (m + 1) + col]) > fabsl(R[piv * (m + 1) + col])) piv = r;
        if (R[piv * (m + 1) + col] == 0) { free(R); return -1; }
        if (piv != col) for (int j = 0; j <= m; j++) {
            long double t = R[col * (m + 1) + j]; R[col * (m + 1) + j] = R[piv * (m + 1) + j]; R[piv * (m + 1) + j] = t;
        }
        for (int r = col + 1; r < m; r++) {
            long double f = R[r * (m + 1) + col] / R[col * (m + 1) + col];
            if (f != 0) for (int j = col; j <= m; j++) R[r * (m + 1) + j] -= f * R[col * (m + 1) + j];
        }
    }
    long double *x = malloc(sizeof(long double) * m);
    for (int i = m - 1; i >= 0; i--) {
        long double sacc = R[i * (m + 1) + m];
        for (int j = i + 1; j < m; j++) sacc -= R[i * (m + 1) + j] * x[j];
        x[i] = sacc / R[i * (m + 1) + i];
    }
    for (int i = 0; i < n; i++) pi[i] = x[i];
    free(x); free(R);
    return 0;
}

static dd_t dd_from_ld(long double v)
{
    double hi = (double)v;
    double lo = (double)(v - (long double)hi);
    return dd_make(hi, lo);
}

int plf_derive(plf_derived *d, const plf_model *m)
{
    memset(d, 0, sizeof(*d));
    const int n = m->n, C = plf_model_category_count(m), E = m->E;
    d->n = n; d->C = C; d->E = E;
    d->cat_rates = calloc((size_t)(C > 0 ? C : 1), sizeof(double));
    d->cat_prior = calloc((size_t)(C > 0 ? C : 1), sizeof(double));
    long double expect = 1;
    /* rate_mixture_summarize, rate_mixture.c:287-339 */
    if (m->mix_mode == MIX_NONE) { d->cat_rates[0] = 1; d->cat_prior[0] = 1; }
    else if (m->mix_mode == MIX_UNIFORM) {
        long double acc = 0;
        for (int i = 0; i < C; i++) { d->cat_rates[i] = m->mix_rates[i]; d->cat_prior[i] = (double)(1.0L / C); acc += m->mix_rates[i]; }
        expect = acc / C;
    } else if (m->mix_mode == MIX_CUSTOM) {
        long double acc = 0;
        for (int i = 0; i < C; i++) {
            d->cat_rates[i] = m->mix_rates[i]; d->cat_prior[i] = m->mix_prior[i];
            acc += (long double)m->mix_rates[i] * m->mix_prior[i];
        }
        expect = acc;
    } else {
        /* gamma_rate_mixture_summarize, rate_mixture.c:166-229 */
        const int K = m->gamma_categories;
        const long double p = m->invariable_prior, q = 1 - p;
        if (plf_gamma_rates(d->cat_rates, K, m->gamma_shape, m->mix_mode == MIX_GAMMA_MEDIAN)) {
            fprintf(stderr, "gamma rate discretisation failed\n");
            return -1;
        }
        for (int i = 0; i < K; i++) {
            d->cat_rates[i] = (double)((long double)d->cat_rates[i] / q);
            d->cat_prior[i] = (double)(q / K);
        }
        if (m->invariable_prior) { d->cat_prior[K] = m->invariable_prior; d->cat_rates[K] = 0; }
        expect = 1;
    }
    d->expect = (double)expect;
    /* equilibrium, if needed */
    long double *pi = NULL;
    if (m->root_mode == PLF_ROOT_EQUILIBRIUM || m->use_eq_divisor) {
        pi = malloc(sizeof(long double) * (size_t)(n > 0 ? n : 1));
        if (equilibrium(pi, m->rate_matrix, n)) {
            fprintf(stderr, "error: the equilibrium distribution of the rate matrix is not uniquely defined\n");
            free(pi);
            return -1;
        }
        d->equilibrium = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        for (int i = 0; i < n; i++) d->equilibrium[i] = (double)pi[i];
    }
    /* rate divisor, cross_site_ws.c:174-191 */
    long double div;
    if (m->use_eq_divisor) {
        long double acc = 0;
        for (int i = 0; i < n; i++) {
            long double rs = 0;
            for (int j = 0; j < n; j++) if (i != j) rs += m->rate_matrix[i * n + j];
            acc += rs * pi[i];
        }
        div = acc * expect;
    } else div = m->rate_divisor;
    free(pi);
    if (!(div > 0) || !isfinite((double)div)) { fprintf(stderr, "error: the rate divisor is not a positive number\n"); return -1; }
    /* Q = offdiag(raw)/div in double-double, diagonal = -row sum (util.c:214-239) */
    d->q_hi = calloc((size_t)n * n, sizeof(double));
    d->q_lo = calloc((size_t)n * n, sizeof(double));
    const dd_t ddiv = dd_from_ld(div);
    for (int i = 0; i < n; i++) {
        dd_t rs = dd_make(0, 0);
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            dd_t qv = dd_div(dd_from_d(m->rate_matrix[i * n + j]), ddiv);
            d->q_hi[i * n + j] = qv.hi; d->q_lo[i * n + j] = qv.lo;
            rs = dd_add(rs, qv);
        }
        d->q_hi[i * n + i] = -rs.hi; d->q_lo[i * n + i] = -rs.lo;
    }
    d->edge_rates_csr = malloc(sizeof(double) * (E > 0 ? E : 1));
    for (int i = 0; i < E; i++) d->edge_rates_csr[m->order[i]] = m->edge_rate_user[i];
    if (m->root_mode == PLF_ROOT_EQUILIBRIUM) {
        d->root_vec = malloc(sizeof(double) * n);
        memcpy(d->root_vec, d->equilibrium, sizeof(double) * n);
    } else if (m->root_mode == PLF_ROOT_CUSTOM) {
        d->root_vec = malloc(sizeof(double) * n);
        memcpy(d->root_vec, m->root_custom, sizeof(double) * n);
    }
    return 0;
}
