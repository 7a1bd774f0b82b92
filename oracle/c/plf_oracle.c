/*
 * CPU restatement (plain C, fp64) of the reference's per-site hot path --
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Nothing under phyly_b200/ links or
 * loads this file; it is the checker and the "port" CPU baseline of bench.py.
 *
 * Parity status: pinned through tests/test_oracle_c.py, which checks it
 * against the 320-bit Python oracle (oracle/arbplf_oracle.py), itself pinned
 * bit-for-bit on the reference's golden vectors.
 *
 * It follows the reference's loop structure literally:
 *   for site: for category:                         arbplfll.c:139-170, arbplfderiv.c:274-357
 *     evaluate_site_lhood (reverse BFS order,       evaluate_site_lhood.c:21-57
 *       emat = P L_b, L_a .*= emat)                 util.c:241-301
 *     root_prior_expectation                        model.c:282-350
 *     evaluate_site_forward (BFS order)             evaluate_site_forward.c:52-102
 *     per edge fe^T (rate Q P) L_b                  evaluate_site_frechet.c:18-39 with F := rate Q P
 * with one simplification the reference's arbitrary precision does not need:
 * per-node power-of-two rescaling.  Sites are independent, so the site loop is
 * split over OpenMP threads (the reference itself is single threaded).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TWO_P256 1.157920892373162e+77
#define TWO_M256 8.636168555094445e-78

int plf_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* returns 0; site_ll [S] (may be NULL), sum_ll [1], sum_deriv [E] (csr order; may be NULL => ll only) */
int plf_oracle_ll_deriv(int N, const int *indptr, const int *indices, const int *preorder,
                        int n, int C, const double *P, const double *D, const double *prior,
                        int root_mode, const double *root_vec,
                        int64_t S, const unsigned char *codes, int K, const double *defs,
                        const double *w, int nthreads,
                        double *site_ll, double *sum_ll, double *sum_deriv)
{
    const int E = N - 1;
    const int root = preorder[0];
    const int want_d = sum_deriv != NULL;
    (void)K;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    double *tsum = calloc((size_t)nthreads * (E + 1), sizeof(double));
#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double *acc = tsum + (size_t)tid * (E + 1);
        double *L = malloc(sizeof(double) * (size_t)C * N * n);     /* inside vectors */
        int *LK = malloc(sizeof(int) * (size_t)C * N);
        double *EM = malloc(sizeof(double) * (size_t)C * E * n);    /* edge vectors */
        double *F = malloc(sizeof(double) * (size_t)N * n);         /* outside vectors */
        int *FK = malloc(sizeof(int) * N);
        double *tmp = malloc(sizeof(double) * 3 * n);
        double *lh = malloc(sizeof(double) * C);
        int *lk = malloc(sizeof(int) * C);
#pragma omp for schedule(static)
        for (int64_t s = 0; s < S; s++) {
            const unsigned char *code = codes + (size_t)s * N;
            /* ---- inside pass per category ---- */
            for (int c = 0; c < C; c++) {
                double *Lc = L + (size_t)c * N * n;
                int *Kc = LK + (size_t)c * N;
                for (int u = N - 1; u >= 0; u--) {
                    const int a = preorder[u];
                    double *La = Lc + (size_t)a * n;
                    const double *base = defs + (size_t)code[a] * n;
                    for (int i = 0; i < n; i++) La[i] = base[i];
                    int k = 0;
                    for (int idx = indptr[a]; idx < indptr[a + 1]; idx++) {
                        const int b = indices[idx];
                        const double *Pm = P + ((size_t)c * E + idx) * n * n;
                        const double *Lb = Lc + (size_t)b * n;
                        double *em = EM + ((size_t)c * E + idx) * n;
                        double mx = 0;
                        for (int i = 0; i < n; i++) {
                            double x = 0;
                            for (int j = 0; j < n; j++) x += Pm[i * n + j] * Lb[j];
                            em[i] = x;
                            La[i] *= x;
                            if (La[i] > mx) mx = La[i];
                        }
                        k += Kc[b];
                        while (mx > 0 && mx < TWO_M256) { for (int i = 0; i < n; i++) La[i] *= TWO_P256; mx *= TWO_P256; k--; }
                    }
                    Kc[a] = k;
                }
                const double *Lr = Lc + (size_t)root * n;
                double x = 0;
                if (root_mode == 0) for (int i = 0; i < n; i++) x += Lr[i];
                else if (root_mode == 1) { for (int i = 0; i < n; i++) x += Lr[i]; x /= n; }
                else for (int i = 0; i < n; i++) x += root_vec[i] * Lr[i];
                lh[c] = x; lk[c] = Kc[root];
            }
            int k0 = -2147483647;
            for (int c = 0; c < C; c++) if (prior[c] * lh[c] > 0 && lk[c] > k0) k0 = lk[c];
            double m = 0;
            for (int c = 0; c < C; c++) if (prior[c] * lh[c] > 0) m += ldexp(prior[c] * lh[c], 256 * (lk[c] - k0));
            const double ll = log(m) + (double)k0 * 177.445678223346 + (double)k0 * 5.936759843446527e-15;
            if (site_ll) site_ll[s] = ll;
            const double ws = w ? w[s] : 1.0;
            if (ws != 0) acc[0] += ws * ll;
            if (!want_d || ws == 0) continue;
            /* ---- outside pass per category ---- */
            for (int c = 0; c < C; c++) {
                if (!(prior[c] * lh[c] > 0)) continue;
                const double *Lc = L + (size_t)c * N * n;
                const int *Kc = LK + (size_t)c * N;
                const double coef = ws * prior[c] / m;
                for (int i = 0; i < n; i++)
                    F[(size_t)root * n + i] = root_mode == 0 ? 1.0 : (root_mode == 1 ? 1.0 / n : root_vec[i]);
                FK[root] = 0;
                for (int u = 0; u < N; u++) {
                    const int a = preorder[u];
                    const int start = indptr[a], stop = indptr[a + 1];
                    if (start == stop) continue;
                    const double *base = defs + (size_t)code[a] * n;
                    const double *Fa = F + (size_t)a * n;
                    for (int idx = start; idx < stop; idx++) {
                        const int b = indices[idx];
                        double *fe = tmp, *y = tmp + n;
                        int kfe = FK[a];
                        for (int i = 0; i < n; i++) fe[i] = Fa[i] * base[i];
                        for (int idx2 = start; idx2 < stop; idx2++) {
                            if (idx2 == idx) continue;
                            const double *em = EM + ((size_t)c * E + idx2) * n;
                            double mx = 0;
                            for (int i = 0; i < n; i++) { fe[i] *= em[i]; if (fe[i] > mx) mx = fe[i]; }
                            kfe += Kc[indices[idx2]];
                            while (mx > 0 && mx < TWO_M256) { for (int i = 0; i < n; i++) fe[i] *= TWO_P256; mx *= TWO_P256; kfe--; }
                            while (mx > TWO_P256) { for (int i = 0; i < n; i++) fe[i] *= TWO_M256; mx *= TWO_M256; kfe++; }
                        }
                        const double *Dm = D + ((size_t)c * E + idx) * n * n;
                        const double *Pm = P + ((size_t)c * E + idx) * n * n;
                        const double *Lb = Lc + (size_t)b * n;
                        double x = 0;
                        for (int i = 0; i < n; i++) {
                            double t = 0;
                            for (int j = 0; j < n; j++) t += Dm[i * n + j] * Lb[j];
                            x += fe[i] * t;
                        }
                        acc[1 + idx] += ldexp(x * coef, 256 * (kfe + Kc[b] - k0));
                        if (indptr[b] != indptr[b + 1]) {
                            double *Fb = F + (size_t)b * n;
                            double mx = 0;
                            for (int j = 0; j < n; j++) {
                                double t = 0;
                                for (int i = 0; i < n; i++) t += Pm[i * n + j] * fe[i];
                                y[j] = t; if (t > mx) mx = t;
                            }
                            int kfb = kfe;
                            double sc = 1;
                            while (mx * sc > TWO_P256) { sc *= TWO_M256; kfb++; }
                            while (mx > 0 && mx * sc < TWO_M256) { sc *= TWO_P256; kfb--; }
                            for (int j = 0; j < n; j++) Fb[j] = y[j] * sc;
                            FK[b] = kfb;
                        }
                    }
                }
            }
        }
        free(L); free(LK); free(EM); free(F); free(FK); free(tmp); free(lh); free(lk);
    }
    double tot = 0;
    for (int t = 0; t < nthreads; t++) tot += tsum[(size_t)t * (E + 1)];
    if (sum_ll) *sum_ll = tot;
    if (sum_deriv) for (int e = 0; e < E; e++) {
        double x = 0;
        for (int t = 0; t < nthreads; t++) x += tsum[(size_t)t * (E + 1) + 1 + e];
        sum_deriv[e] = x;
    }
    free(tsum);
    return 0;
}
