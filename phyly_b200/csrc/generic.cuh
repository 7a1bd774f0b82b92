/*
 * Generic (any state count, any data) per-site kernels.
 *
 * These restate evaluate_site_lhood (evaluate_site_lhood.c:6-63),
 * evaluate_site_forward (evaluate_site_forward.c:31-105),
 * evaluate_site_marginal_unnormalized (evaluate_site_marginal.c:6-21) and
 * evaluate_site_frechet (evaluate_site_frechet.c:4-42) with one GPU thread per
 * site (x rate category for the inside pass).  All per-node vectors live in HBM
 * with the site index fastest, so every load and store of a warp is one
 * contiguous 256-byte run.  Per-site power-of-two rescaling (exponents counted
 * in units of 2^256) replaces the reference's growing precision.
 *
 * Layouts (Sc = sites in the current chunk, s = site within the chunk):
 *   Lg [C][N][n][Sc]  inside ("lhood node") vectors, mantissas
 *   Kg [C][N][Sc]     their exponents (int32, <= 0), true = mant * 2^(256 K)
 *   Cg [C][N][Sc]     1 if the vector is an exact constant column
 *                     (arb_mat_extras.c:35-51), tracked symbolically
 *   Eg [C][E][n][Sc]  edge vectors P_e L_b (mantissas; exponent = K of the child)
 *   Fg [C][N][n][Sc]  outside ("forward node") vectors, FK [C][N][Sc] exponents
 */
#pragma once
#include <stdint.h>

#define PLF_SCALE_BITS 256
#define PLF_TS 128   /* threads (sites) per CTA in the generic kernels */

struct TreeDev {
    int N, E, root;
    const int *indptr;     /* [N+1] */
    const int *indices;    /* [E]   */
    const int *preorder;   /* [N]   */
    const unsigned char *node_has_data;  /* [N] */
};

struct GenericArgs {
    TreeDev t;
    int n, C, K;
    int64_t S;            /* total sites */
    int64_t s0;           /* first site of the chunk */
    int Sc;               /* sites in the chunk */
    const void *codes;    /* [N][S] node-major, uint8 or int32 */
    int code_bytes;
    const double *defs;   /* [K][n] */
    const unsigned char *def_const;   /* [K] row is constant */
    const double *P;      /* [C][E][n][n] */
    const double *Fm;     /* [C][E][n][n] matrices of the edge bilinear form (or NULL) */
    int f_zero_rowsum;    /* Fm has zero row sums: constant columns map to exact 0 (util.c:338-344) */
    const double *edge_coef; /* [C][E] extra factor per (category, edge) on the edge form, or NULL */
    const unsigned char *edge_mask;  /* [E] or NULL */
    const double *cat_prior; /* [C] */
    int root_mode;
    const double *root_vec;  /* [n] */
    double *Lg; int *Kg; unsigned char *Cg; double *Eg; double *Fg; int *FK;
    double *cat_lh;       /* [C][Sc] mantissa of root-prior expectation */
    int *cat_k;           /* [C][Sc] */
    double *site_m;       /* [Sc] mantissa of site likelihood */
    int *site_k;          /* [Sc] */
    double *site_ll;      /* [S]  (global site index) */
    double *edge_out;     /* [E][Sc] per-site edge forms (already divided by site lhood) or NULL */
    double *marg_out;     /* [N][n][Sc] or NULL */
    int want_edge, want_marg;
    /* tip tables for the tile kernel: TP[c][tip edge][k][i] = (P_e def_k)_i, tip_of_edge[csr idx] = tip edge or -1 */
    const double *TP;
    const double *TF;             /* the same with the edge-form matrices (F_e def_k), or NULL */
    const int *tip_of_edge;
    int Et;
    /* csr indices of the edges whose child goes through the GEMM, in the order tile_inside_kernel meets them */
    const int *int_seq;
    int n_int_seq;
    /* tile_inside_kernel stages the tile's tip codes in shared memory (1-byte codes, Et * 64 bytes) */
    int tip_stage;
    const int *tip_edge_csr;      /* [Et] csr index of each tip edge */
};

__device__ __forceinline__ int plf_code_at(const void *codes, int code_bytes, int64_t S, int node, int64_t site)
{
    if (code_bytes == 1) return ((const unsigned char *)codes)[(size_t)node * S + site];
    return ((const int *)codes)[(size_t)node * S + site];
}

#define PLF_TWO_P256 1.157920892373162e+77      /* 2^256  */
#define PLF_TWO_M256 8.636168555094445e-78      /* 2^-256 */

/* inside pass: thread = (site, category) ; blockIdx.y = category */
__global__ void __launch_bounds__(PLF_TS) generic_inside_kernel(GenericArgs a)
{
    extern __shared__ double gsm[];
    const int tid = threadIdx.x;
    const int s = blockIdx.x * PLF_TS + tid;
    const int c = blockIdx.y;
    if (s >= a.Sc) return;
    const int n = a.n, Sc = a.Sc;
    const int64_t gs = a.s0 + s;
    double *acc = gsm + tid;                  /* acc[i*PLF_TS] */
    double *v = gsm + (size_t)n * PLF_TS + tid;
    const size_t cN = (size_t)c * a.t.N;
    const size_t cE = (size_t)c * a.t.E;

    for (int u = a.t.N - 1; u >= 0; u--) {
        const int nd = a.t.preorder[u];
        const int start = a.t.indptr[nd], stop = a.t.indptr[nd + 1];
        int accK = 0;
        int cst = 1;
        if (a.t.node_has_data[nd]) {
            int code = plf_code_at(a.codes, a.code_bytes, a.S, nd, gs);
            for (int i = 0; i < n; i++) acc[i * PLF_TS] = a.defs[(size_t)code * n + i];
            cst = a.def_const[code];
        } else {
            for (int i = 0; i < n; i++) acc[i * PLF_TS] = 1.0;
        }
        for (int idx = start; idx < stop; idx++) {
            const int b = a.t.indices[idx];
            const double *Lb = a.Lg + ((cN + b) * n) * Sc + s;
            for (int k = 0; k < n; k++) v[k * PLF_TS] = Lb[(size_t)k * Sc];
            const int kb = a.Kg[(cN + b) * Sc + s];
            const int bc = a.Cg[(cN + b) * Sc + s];
            const double *Pm = a.P + (cE + idx) * n * n;
            double *Eo = a.Eg + ((cE + idx) * n) * Sc + s;
            double mx = 0.0;
            for (int i = 0; i < n; i++) {
                double em;
                if (bc) {
                    em = v[0];
                } else {
                    em = 0.0;
                    for (int k = 0; k < n; k++) em = fma(Pm[i * n + k], v[k * PLF_TS], em);
                }
                Eo[(size_t)i * Sc] = em;
                double x = acc[i * PLF_TS] * em;
                acc[i * PLF_TS] = x;
                mx = fmax(mx, x);
            }
            accK += kb;
            cst &= bc;
            while (mx > 0.0 && mx < PLF_TWO_M256) {
                for (int i = 0; i < n; i++) acc[i * PLF_TS] *= PLF_TWO_P256;
                mx *= PLF_TWO_P256;
                accK -= 1;
            }
        }
        double *La = a.Lg + ((cN + nd) * n) * Sc + s;
        for (int i = 0; i < n; i++) La[(size_t)i * Sc] = acc[i * PLF_TS];
        a.Kg[(cN + nd) * Sc + s] = accK;
        a.Cg[(cN + nd) * Sc + s] = (unsigned char)cst;
        if (nd == a.t.root) {
            /* root_prior_expectation, model.c:282-350 */
            double lh = 0.0;
            if (a.root_mode == PLF_ROOT_NONE) {
                for (int i = 0; i < n; i++) lh += acc[i * PLF_TS];
            } else if (a.root_mode == PLF_ROOT_UNIFORM) {
                if (cst) lh = acc[0];
                else { for (int i = 0; i < n; i++) lh += acc[i * PLF_TS]; lh /= (double)n; }
            } else if (a.root_mode == PLF_ROOT_EQUILIBRIUM && cst) {
                lh = acc[0];
            } else {
                for (int i = 0; i < n; i++) lh = fma(a.root_vec[i], acc[i * PLF_TS], lh);
            }
            a.cat_lh[(size_t)c * Sc + s] = lh;
            a.cat_k[(size_t)c * Sc + s] = accK;
        }
    }
}

/* combine categories: site likelihood and log-likelihood (arbplfll.c:149-169) */
__global__ void generic_site_kernel(GenericArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.Sc) return;
    const int Sc = a.Sc;
    int k0 = INT_MIN;
    for (int c = 0; c < a.C; c++) {
        double v = a.cat_prior[c] * a.cat_lh[(size_t)c * Sc + s];
        if (v > 0.0) k0 = max(k0, a.cat_k[(size_t)c * Sc + s]);
    }
    double m = 0.0;
    if (k0 != INT_MIN) {
        for (int c = 0; c < a.C; c++) {
            double v = a.cat_prior[c] * a.cat_lh[(size_t)c * Sc + s];
            if (v > 0.0) m += scalbn(v, PLF_SCALE_BITS * (a.cat_k[(size_t)c * Sc + s] - k0));
        }
    } else {
        k0 = 0;
    }
    a.site_m[s] = m;
    a.site_k[s] = k0;
    /* log(m * 2^(256 k0)) ; 256*ln2 split in two parts */
    const double c_hi = 177.445678223346, c_lo = 5.936759843446527e-15;  /* 256*ln(2) = hi + lo */
    double ll = log(m);
    ll = fma((double)k0, c_hi, ll);
    ll = fma((double)k0, c_lo, ll);
    a.site_ll[a.s0 + s] = ll;
}

/*
 * outside pass: thread = site, categories looped inside (their contributions to
 * the same output cell are summed without atomics).
 */
__global__ void __launch_bounds__(PLF_TS) generic_outside_kernel(GenericArgs a)
{
    extern __shared__ double gsm[];
    const int tid = threadIdx.x;
    const int s = blockIdx.x * PLF_TS + tid;
    if (s >= a.Sc) return;
    const int n = a.n, Sc = a.Sc;
    const int64_t gs = a.s0 + s;
    double *tmp = gsm + tid;
    double *fe = gsm + (size_t)n * PLF_TS + tid;
    double *y = gsm + (size_t)2 * n * PLF_TS + tid;
    const double site_m = a.site_m[s];
    const int site_k = a.site_k[s];

    for (int c = 0; c < a.C; c++) {
        const size_t cN = (size_t)c * a.t.N;
        const size_t cE = (size_t)c * a.t.E;
        const double prior = a.cat_prior[c];
        /* arbplfmarginal.c:184-191: skip categories whose posterior is exactly zero */
        if (!(prior * a.cat_lh[(size_t)c * Sc + s] > 0.0)) continue;
        const double coef = prior / site_m;
        /* forward vector at the root: root_prior_mul_col_vec, model.c:225-280 */
        {
            double *Fr = a.Fg + ((cN + a.t.root) * n) * Sc + s;
            for (int i = 0; i < n; i++) {
                double r = 1.0;
                if (a.root_mode == PLF_ROOT_UNIFORM) r = 1.0 / (double)n;
                else if (a.root_mode == PLF_ROOT_EQUILIBRIUM || a.root_mode == PLF_ROOT_CUSTOM) r = a.root_vec[i];
                Fr[(size_t)i * Sc] = r;
            }
            a.FK[(cN + a.t.root) * Sc + s] = 0;
        }
        for (int u = 0; u < a.t.N; u++) {
            const int nd = a.t.preorder[u];
            const int start = a.t.indptr[nd], stop = a.t.indptr[nd + 1];
            const double *Fa = a.Fg + ((cN + nd) * n) * Sc + s;
            const int ka = a.FK[(cN + nd) * Sc + s];
            if (a.want_marg) {
                /* marg_a += prior * fn_a .* L_a / site_L */
                const double *La = a.Lg + ((cN + nd) * n) * Sc + s;
                const int kl = a.Kg[(cN + nd) * Sc + s];
                const int sh = PLF_SCALE_BITS * (ka + kl - site_k);
                double *Mo = a.marg_out + ((size_t)nd * n) * Sc + s;
                for (int i = 0; i < n; i++)
                    Mo[(size_t)i * Sc] += scalbn(coef * Fa[(size_t)i * Sc] * La[(size_t)i * Sc], sh);
            }
            if (start == stop) continue;
            if (a.t.node_has_data[nd]) {
                int code = plf_code_at(a.codes, a.code_bytes, a.S, nd, gs);
                for (int i = 0; i < n; i++) tmp[i * PLF_TS] = Fa[(size_t)i * Sc] * a.defs[(size_t)code * n + i];
            } else {
                for (int i = 0; i < n; i++) tmp[i * PLF_TS] = Fa[(size_t)i * Sc];
            }
            for (int idx = start; idx < stop; idx++) {
                const int b = a.t.indices[idx];
                int kfe = ka;
                for (int i = 0; i < n; i++) fe[i * PLF_TS] = tmp[i * PLF_TS];
                for (int idx2 = start; idx2 < stop; idx2++) {
                    if (idx2 == idx) continue;
                    const double *Es = a.Eg + ((cE + idx2) * n) * Sc + s;
                    double mx = 0.0;
                    for (int i = 0; i < n; i++) {
                        double x = fe[i * PLF_TS] * Es[(size_t)i * Sc];
                        fe[i * PLF_TS] = x;
                        mx = fmax(mx, x);
                    }
                    kfe += a.Kg[(cN + a.t.indices[idx2]) * Sc + s];
                    while (mx > 0.0 && mx < PLF_TWO_M256) {
                        for (int i = 0; i < n; i++) fe[i * PLF_TS] *= PLF_TWO_P256;
                        mx *= PLF_TWO_P256; kfe -= 1;
                    }
                    while (mx > PLF_TWO_P256) {
                        for (int i = 0; i < n; i++) fe[i * PLF_TS] *= PLF_TWO_M256;
                        mx *= PLF_TWO_M256; kfe += 1;
                    }
                }
                const double *Lb = a.Lg + ((cN + b) * n) * Sc + s;
                const int kb = a.Kg[(cN + b) * Sc + s];
                const int bc = a.Cg[(cN + b) * Sc + s];
                if (a.want_edge && (!a.edge_mask || a.edge_mask[idx])) {
                    /* evaluate_site_frechet.c:18-39: fe^T F L_b */
                    double x = 0.0;
                    if (!(a.f_zero_rowsum && bc)) {
                        const double *Fm = a.Fm + (cE + idx) * n * n;
                        for (int k = 0; k < n; k++) y[k * PLF_TS] = Lb[(size_t)k * Sc];
                        for (int i = 0; i < n; i++) {
                            double acc = 0.0;
                            for (int k = 0; k < n; k++) acc = fma(Fm[i * n + k], y[k * PLF_TS], acc);
                            x = fma(fe[i * PLF_TS], acc, x);
                        }
                    }
                    double ec = a.edge_coef ? a.edge_coef[cE + idx] : 1.0;
                    a.edge_out[(size_t)idx * Sc + s] += scalbn(x * coef * ec, PLF_SCALE_BITS * (kfe + kb - site_k));
                }
                /* forward node vector of b: P^T fe  (util.c:464-498) */
                {
                    const double *Pm = a.P + (cE + idx) * n * n;
                    double *Fb = a.Fg + ((cN + b) * n) * Sc + s;
                    double mx = 0.0;
                    for (int j = 0; j < n; j++) {
                        double acc = 0.0;
                        for (int i = 0; i < n; i++) acc = fma(Pm[i * n + j], fe[i * PLF_TS], acc);
                        y[j * PLF_TS] = acc;
                        mx = fmax(mx, acc);
                    }
                    int kfb = kfe;
                    double sc = 1.0;
                    while (mx * sc > PLF_TWO_P256) { sc *= PLF_TWO_M256; kfb += 1; }
                    while (mx > 0.0 && mx * sc < PLF_TWO_M256) { sc *= PLF_TWO_P256; kfb -= 1; }
                    for (int j = 0; j < n; j++) Fb[(size_t)j * Sc] = y[j * PLF_TS] * sc;
                    a.FK[(cN + b) * Sc + s] = kfb;
                }
            }
        }
    }
}

/* leaves: write their base vectors as "inside" vectors once per chunk (so that
 * parents read every child the same way). blockIdx.y = category.  write_vectors = 0 writes the
 * exponents and constant flags only (the tile kernels take the vectors of tips from tip tables). */
__global__ void generic_leaf_kernel(GenericArgs a, int write_vectors)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (s >= a.Sc) return;
    const int n = a.n, Sc = a.Sc;
    const size_t cN = (size_t)c * a.t.N;
    for (int nd = 0; nd < a.t.N; nd++) {
        if (a.t.indptr[nd] != a.t.indptr[nd + 1]) continue;
        double *La = a.Lg + ((cN + nd) * n) * Sc + s;
        int cst = 1;
        if (a.t.node_has_data[nd]) {
            int code = plf_code_at(a.codes, a.code_bytes, a.S, nd, a.s0 + s);
            if (write_vectors) for (int i = 0; i < n; i++) La[(size_t)i * Sc] = a.defs[(size_t)code * n + i];
            cst = a.def_const[code];
        } else if (write_vectors) {
            for (int i = 0; i < n; i++) La[(size_t)i * Sc] = 1.0;
        }
        a.Kg[(cN + nd) * Sc + s] = 0;
        a.Cg[(cN + nd) * Sc + s] = (unsigned char)cst;
    }
}
