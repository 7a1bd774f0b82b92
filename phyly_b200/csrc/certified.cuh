/*
 * Certified mode: enclosures [lower, upper] of the site log-likelihoods with directed rounding.
 *
 * The reference computes in Arb's midpoint-radius arithmetic and raises the precision until every output is a ball
 * that rounds to a unique double (util.c:12-50, arbplfll.c:206-224).  Here the same pruning recursion
 * (evaluate_site_lhood.c:21-57, arbplfll.c:139-170) is run in fp64 INTERVAL arithmetic: every quantity on the path is
 * non-negative, so a lower bound is obtained by rounding every operation down (__dmul_rd, __fma_rd, __dadd_rd) and an
 * upper bound by rounding every operation up; scaling by powers of two is exact.  The enclosure is rigorous with
 * respect to the inputs the device receives, given as intervals themselves:
 *   - rate_c * t_e  within a relative delta_rate (the host's gamma quantiles are accurate to 1e-14, not certified);
 *   - the scaled rate matrix Q and the derived priors within a relative delta_q (long-double host arithmetic);
 *   - exp / log of the CUDA math library within their documented 1 ulp (3 ulps are taken).
 *
 * cert_expm_kernel encloses exp(x Q), x in [x_lo, x_hi]:  with mu >= max_i |x Q_ii| and B = (x Q + mu I) / 2^s >= 0,
 * exp(x Q) = (e^{-beta} exp(B))^(2^s), beta = mu / 2^s <= 1.  The Taylor polynomial of degree K of exp(B) with all
 * operations rounded down is a lower bound; rounded up and increased by the bound beta^(K+1)/(K+1)!/(1 - beta/(K+2)) of
 * the remainder (B has row sums beta) it is an upper bound; squaring keeps both (non-negative matrices).
 */
#pragma once
#include <stdint.h>
#include "plf_consts.h"

#define CERT_TS 64          /* sites per CTA in cert_inside_kernel */
#define CERT_TAYLOR 26

__device__ __forceinline__ double cert_down(double x, int ulps)
{
    for (int i = 0; i < ulps; i++) x = nextafter(x, -__longlong_as_double(0x7ff0000000000000LL));
    return x;
}
__device__ __forceinline__ double cert_up(double x, int ulps)
{
    for (int i = 0; i < ulps; i++) x = nextafter(x, __longlong_as_double(0x7ff0000000000000LL));
    return x;
}

/* one CTA per (edge, category); ws: 8 n^2 doubles per CTA */
__global__ void __launch_bounds__(128) cert_expm_kernel(int n, int E, int C, const double *q_hi, const double *q_lo,
                                                        const double *edge_rate, const double *cat_rate, double delta_rate,
                                                        double delta_q, double *ws, double *Plo, double *Phi)
{
    const int e = blockIdx.x, c = blockIdx.y, nn = n * n;
    double *base = ws + ((size_t)c * E + e) * 8 * nn;
    double *Blo = base, *Bhi = base + nn, *Tlo = base + 2 * nn, *Thi = base + 3 * nn;
    double *Slo = base + 4 * nn, *Shi = base + 5 * nn, *Ulo = base + 6 * nn, *Uhi = base + 7 * nn;
    double *outlo = Plo + ((size_t)c * E + e) * nn, *outhi = Phi + ((size_t)c * E + e) * nn;
    __shared__ double sh_mu, sh_xlo, sh_xhi;
    __shared__ int sh_s;
    const double r = cat_rate[c], t = edge_rate[e];
    if (threadIdx.x == 0) {
        const double xlo = __dmul_rd(__dmul_rd(r, __dsub_rd(1.0, delta_rate)), t);
        const double xhi = __dmul_ru(__dmul_ru(r, __dadd_ru(1.0, delta_rate)), t);
        double mu = 0.0;
        for (int i = 0; i < n; i++) {
            double rs = 0.0;
            for (int j = 0; j < n; j++) {
                if (j == i) continue;
                const double q = __dmul_ru(__dadd_ru(q_hi[i * n + j], q_lo ? q_lo[i * n + j] : 0.0), __dadd_ru(1.0, delta_q));
                rs = __dadd_ru(rs, q);
            }
            mu = fmax(mu, __dmul_ru(xhi, rs));
        }
        int s = 0;
        if (mu > 1.0) { int ex; frexp(mu, &ex); s = ex; }
        sh_mu = mu; sh_s = s; sh_xlo = xlo; sh_xhi = xhi;
    }
    __syncthreads();
    const double mu = sh_mu, xlo = sh_xlo, xhi = sh_xhi;
    const int s = sh_s;
    if (!(mu > 0.0)) {          /* zero rate or zero length: the identity, exactly */
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) { const double v = (idx / n == idx % n) ? 1.0 : 0.0; outlo[idx] = v; outhi[idx] = v; }
        return;
    }
    const double sc = scalbn(1.0, -s);
    /* B, term_0 = sum_0 = I */
    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        double lo, hi;
        if (i != j) {
            const double qd = __dadd_rd(q_hi[idx], q_lo ? q_lo[idx] : 0.0), qu = __dadd_ru(q_hi[idx], q_lo ? q_lo[idx] : 0.0);
            lo = __dmul_rd(xlo, fmax(0.0, __dmul_rd(qd, __dsub_rd(1.0, delta_q))));
            hi = __dmul_ru(xhi, __dmul_ru(fmax(0.0, qu), __dadd_ru(1.0, delta_q)));
        } else {
            double rlo = 0.0, rhi = 0.0;
            for (int k = 0; k < n; k++) {
                if (k == i) continue;
                const double qd = __dadd_rd(q_hi[i * n + k], q_lo ? q_lo[i * n + k] : 0.0), qu = __dadd_ru(q_hi[i * n + k], q_lo ? q_lo[i * n + k] : 0.0);
                rlo = __dadd_rd(rlo, fmax(0.0, __dmul_rd(qd, __dsub_rd(1.0, delta_q))));
                rhi = __dadd_ru(rhi, __dmul_ru(fmax(0.0, qu), __dadd_ru(1.0, delta_q)));
            }
            lo = fmax(0.0, __dsub_rd(mu, __dmul_ru(xhi, rhi)));
            hi = __dsub_ru(mu, __dmul_rd(xlo, rlo));
        }
        Blo[idx] = lo * sc; Bhi[idx] = hi * sc;               /* exact: a power of two */
        const double id = (i == j) ? 1.0 : 0.0;
        Tlo[idx] = id; Thi[idx] = id; Slo[idx] = id; Shi[idx] = id;
    }
    __syncthreads();
    for (int k = 1; k <= CERT_TAYLOR; k++) {
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            const int i = idx / n, j = idx - i * n;
            double lo = 0.0, hi = 0.0;
            for (int m = 0; m < n; m++) {
                lo = __fma_rd(Tlo[i * n + m], Blo[m * n + j], lo);
                hi = __fma_ru(Thi[i * n + m], Bhi[m * n + j], hi);
            }
            Ulo[idx] = __ddiv_rd(lo, (double)k); Uhi[idx] = __ddiv_ru(hi, (double)k);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            Tlo[idx] = Ulo[idx]; Thi[idx] = Uhi[idx];
            Slo[idx] = __dadd_rd(Slo[idx], Ulo[idx]); Shi[idx] = __dadd_ru(Shi[idx], Uhi[idx]);
        }
        __syncthreads();
    }
    /* remainder of the series and the scalar e^{-beta} */
    const double beta = mu * sc;                               /* exact, <= 1 */
    double rem = 1.0;
    for (int k = 1; k <= CERT_TAYLOR + 1; k++) rem = __ddiv_ru(__dmul_ru(rem, beta), (double)k);
    rem = __ddiv_ru(rem, __dsub_rd(1.0, __ddiv_ru(beta, (double)(CERT_TAYLOR + 2))));
    const double ev = exp(-beta);
    const double elo = cert_down(ev, 3), ehi = cert_up(ev, 3);
    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
        Slo[idx] = __dmul_rd(Slo[idx], elo);
        Shi[idx] = __dmul_ru(__dadd_ru(Shi[idx], rem), ehi);
    }
    __syncthreads();
    for (int q = 0; q < s; q++) {
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            const int i = idx / n, j = idx - i * n;
            double lo = 0.0, hi = 0.0;
            for (int m = 0; m < n; m++) {
                lo = __fma_rd(Slo[i * n + m], Slo[m * n + j], lo);
                hi = __fma_ru(Shi[i * n + m], Shi[m * n + j], hi);
            }
            Ulo[idx] = lo; Uhi[idx] = hi;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) { Slo[idx] = Ulo[idx]; Shi[idx] = Uhi[idx]; }
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) { outlo[idx] = fmax(0.0, Slo[idx]); outhi[idx] = fmin(1.0, Shi[idx]); }
}

struct CertArgs {
    TreeDev t;
    int n, C, K;
    int64_t S, s0;
    int Sc;
    const void *codes; int code_bytes;
    const double *defs; const unsigned char *def_const;
    const double *Plo, *Phi;           /* [C][E][n][n] */
    int root_mode;
    const double *root_lo, *root_hi;   /* [n] enclosure of the root prior weights */
    const double *prior_lo, *prior_hi; /* [C] enclosure of the category priors */
    double *Llo, *Lhi;                 /* [C][N][n][Sc] node vectors (mantissas), one exponent for both bounds */
    int *Kg; unsigned char *Cg;        /* [C][N][Sc] */
    double *cat_lo, *cat_hi; int *cat_k;   /* [C][Sc] */
    double *site_lo, *site_hi;         /* [S] enclosures of the site log-likelihoods */
};

/* interval restatement of generic_inside_kernel: thread = (site, category), blockIdx.y = category */
__global__ void __launch_bounds__(CERT_TS) cert_inside_kernel(CertArgs a)
{
    extern __shared__ double csm[];
    const int tid = threadIdx.x;
    const int s = blockIdx.x * CERT_TS + tid;
    const int c = blockIdx.y;
    if (s >= a.Sc) return;
    const int n = a.n, Sc = a.Sc;
    const int64_t gs = a.s0 + s;
    double *alo = csm + tid, *ahi = csm + (size_t)n * CERT_TS + tid;
    double *vlo = csm + (size_t)2 * n * CERT_TS + tid, *vhi = csm + (size_t)3 * n * CERT_TS + tid;
    const size_t cN = (size_t)c * a.t.N, cE = (size_t)c * a.t.E;

    for (int u = a.t.N - 1; u >= 0; u--) {
        const int nd = a.t.preorder[u];
        const int start = a.t.indptr[nd], stop = a.t.indptr[nd + 1];
        int accK = 0, cst = 1;
        if (a.t.node_has_data[nd]) {
            const int code = plf_code_at(a.codes, a.code_bytes, a.S, nd, gs);
            for (int i = 0; i < n; i++) { const double d = a.defs[(size_t)code * n + i]; alo[i * CERT_TS] = d; ahi[i * CERT_TS] = d; }
            cst = a.def_const[code];
        } else {
            for (int i = 0; i < n; i++) { alo[i * CERT_TS] = 1.0; ahi[i * CERT_TS] = 1.0; }
        }
        for (int idx = start; idx < stop; idx++) {
            const int b = a.t.indices[idx];
            const double *Lbl = a.Llo + ((cN + b) * n) * Sc + s, *Lbh = a.Lhi + ((cN + b) * n) * Sc + s;
            for (int k = 0; k < n; k++) { vlo[k * CERT_TS] = Lbl[(size_t)k * Sc]; vhi[k * CERT_TS] = Lbh[(size_t)k * Sc]; }
            const int kb = a.Kg[(cN + b) * Sc + s];
            const int bc = a.Cg[(cN + b) * Sc + s];
            const double *Pl = a.Plo + (cE + idx) * n * n, *Ph = a.Phi + (cE + idx) * n * n;
            double mx = 0.0;
            for (int i = 0; i < n; i++) {
                double elo, ehi;
                if (bc) {                    /* exact constant column under a stochastic matrix (arb_mat_extras.c:84-91) */
                    elo = vlo[0]; ehi = vhi[0];
                } else {
                    elo = 0.0; ehi = 0.0;
                    for (int k = 0; k < n; k++) {
                        elo = __fma_rd(Pl[i * n + k], vlo[k * CERT_TS], elo);
                        ehi = __fma_ru(Ph[i * n + k], vhi[k * CERT_TS], ehi);
                    }
                }
                const double xl = __dmul_rd(alo[i * CERT_TS], elo), xh = __dmul_ru(ahi[i * CERT_TS], ehi);
                alo[i * CERT_TS] = xl; ahi[i * CERT_TS] = xh;
                mx = fmax(mx, xh);
            }
            accK += kb;
            cst &= bc;
            while (mx > 0.0 && mx < PLF_TWO_M256) {      /* exact rescaling of both bounds */
                for (int i = 0; i < n; i++) { alo[i * CERT_TS] *= PLF_TWO_P256; ahi[i * CERT_TS] *= PLF_TWO_P256; }
                mx *= PLF_TWO_P256;
                accK -= 1;
            }
        }
        double *Ll = a.Llo + ((cN + nd) * n) * Sc + s, *Lh = a.Lhi + ((cN + nd) * n) * Sc + s;
        for (int i = 0; i < n; i++) { Ll[(size_t)i * Sc] = alo[i * CERT_TS]; Lh[(size_t)i * Sc] = ahi[i * CERT_TS]; }
        a.Kg[(cN + nd) * Sc + s] = accK;
        a.Cg[(cN + nd) * Sc + s] = (unsigned char)cst;
        if (nd == a.t.root) {
            /* root_prior_expectation, model.c:282-350 */
            double lo = 0.0, hi = 0.0;
            if (cst && (a.root_mode == PLF_ROOT_UNIFORM || a.root_mode == PLF_ROOT_EQUILIBRIUM)) {
                lo = alo[0]; hi = ahi[0];
            } else {
                for (int i = 0; i < n; i++) {
                    lo = __fma_rd(a.root_lo[i], alo[i * CERT_TS], lo);
                    hi = __fma_ru(a.root_hi[i], ahi[i * CERT_TS], hi);
                }
            }
            a.cat_lo[(size_t)c * Sc + s] = lo;
            a.cat_hi[(size_t)c * Sc + s] = hi;
            a.cat_k[(size_t)c * Sc + s] = accK;
        }
    }
}

/* x 2^(256 k), k <= 0, rounded towards zero (down = 1) or away from it, for x >= 0 */
__device__ __forceinline__ double cert_scale(double x, int k, int down)
{
    if (x == 0.0 || k == 0) return x;
    const double y = scalbn(x, PLF_SCALE_BITS * max(k, -16));
    if (scalbn(y, -PLF_SCALE_BITS * max(k, -16)) == x && k >= -16) return y;       /* exact */
    return down ? (y > 0.0 ? cert_down(y, 1) : 0.0) : cert_up(y, 1);
}

/* combine the categories (arbplfll.c:149-169): enclosure of log(sum_c prior_c L_c) */
__global__ void cert_site_kernel(CertArgs a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.Sc) return;
    const int Sc = a.Sc;
    int k0 = INT_MIN;
    for (int c = 0; c < a.C; c++)
        if (a.prior_hi[c] * a.cat_hi[(size_t)c * Sc + s] > 0.0) k0 = max(k0, a.cat_k[(size_t)c * Sc + s]);
    double lo = 0.0, hi = 0.0;
    if (k0 != INT_MIN) {
        for (int c = 0; c < a.C; c++) {
            const int dk = a.cat_k[(size_t)c * Sc + s] - k0;
            if (a.prior_hi[c] * a.cat_hi[(size_t)c * Sc + s] > 0.0) {
                lo = __dadd_rd(lo, cert_scale(__dmul_rd(a.prior_lo[c], a.cat_lo[(size_t)c * Sc + s]), dk, 1));
                hi = __dadd_ru(hi, cert_scale(__dmul_ru(a.prior_hi[c], a.cat_hi[(size_t)c * Sc + s]), dk, 0));
            }
        }
    } else {
        k0 = 0;
    }
    /* log(m 2^(256 k0)): 256 ln 2 lies in [c_lo, c_hi]; k0 <= 0 */
    const double c_lo = 177.445678223346, c_hi = 177.44567822334602;
    const double kd = (double)k0;
    double llo = (lo > 0.0) ? cert_down(log(lo), 3) : -INFINITY;
    double lhi = (hi > 0.0) ? cert_up(log(hi), 3) : -INFINITY;
    llo = __dadd_rd(llo, __dmul_rd(kd, k0 <= 0 ? c_hi : c_lo));
    lhi = __dadd_ru(lhi, __dmul_ru(kd, k0 <= 0 ? c_lo : c_hi));
    a.site_lo[a.s0 + s] = llo;
    a.site_hi[a.s0 + s] = lhi;
}
