/*
 * Inside (pruning) pass for large state spaces (16 < n <= 64: amino-acid, codon) on the FP64
 * tensor pipe.
 *
 * For these models the per-node contraction P_e . L_b of evaluate_site_lhood
 * (evaluate_site_lhood.c:32-56, _arb_mat_mul_stochastic arb_mat_extras.c:54-113) is a dense
 * GEMM over sites: [n x n] . [n x TS] for a tile of TS sites.  One CTA owns a tile of 64
 * sites of one rate category and walks the tree in post order; for every child edge it stages
 * P_e (padded to 64 x 64) and the child's partials (64 x 64 sites) in shared memory and
 * multiplies them with mma.sync.m8n8k4.f64 (DMMA): warp w owns output rows [8w, 8w+8) for
 * all 64 sites, i.e. 8 accumulator fragments, and keeps the node's running product in the
 * same fragment layout.  Partials live in HBM as [category][node][state][site] with the site
 * index fastest (the layout of the generic kernels, which consume the results in the outside
 * pass).  Row strides of the shared tiles are padded (68 and 72 doubles) so that both
 * fragment loads are bank-conflict free.
 *
 * Algorithmic work per (site, child edge): 2 n^2 flops (here counted on the padded 64 x 64);
 * HBM traffic per (site, internal node): n*8 B written + n*8 B per internal child read
 * (+ the same again for the edge vectors kept for the outside pass).
 */
#pragma once
#include <stdint.h>

#define TL_NP 64          /* padded state count */
#define TL_TS 64          /* sites per tile */
#define TL_PS 68          /* row stride of the staged P (doubles) */
#define TL_LS 68          /* row stride of the staged child partials (doubles): = 4 mod 16, so that the 16 lanes (4 rows x 4 columns)
                             of a half-warp B-fragment load fall on 16 different 8-byte banks (72 put rows q and q + 2 on the same ones) */

__device__ __forceinline__ void tl_dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* asynchronous global -> shared copies (LDGSTS); src_bytes < size zero-fills the remainder */
__device__ __forceinline__ void tl_cp16(void *dst, const void *src, int src_bytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void tl_cp8(void *dst, const void *src, int src_bytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void tl_cp_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void tl_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/*
 * grid = (site tiles, categories), 256 threads.
 *
 * The child tiles of the GEMM edges are double-buffered: while the tensor pipe works on one child's
 * [64 states x 64 sites] tile, cp.async brings in the tile of the next GEMM edge of the traversal
 * (a.int_seq, compiled by the host: the kernel walks the tree in reverse BFS order, so a tile needed next
 * was finished long ago and nothing in flight depends on the current node).  Tip children never enter a
 * GEMM: each thread gathers its 16 entries of P_e def_k from the tip table.  Per-site bookkeeping (scale
 * exponent, constant-column flag) lives in the registers of thread `site`; the column maxima for the
 * rescale are taken after every second child and after the last one.
 */
template <int NP>
__global__ void __launch_bounds__(NP * 4, (NP == 64 ? 2 : 4)) tile_inside_kernel(GenericArgs a, int keep_edges)
{
    constexpr int NT = NP * 4;                        /* threads: one warp per 8 output rows */
    constexpr int NW = NP / 8;
    extern __shared__ __align__(16) double tl_sm[];
    double *Lbuf = tl_sm;                             /* [2][64][TL_LS] */
    double *colmax = Lbuf + 2 * NP * TL_LS;        /* [8][64] */
    double *scale = colmax + NW * TL_TS;               /* [64] */
    int *colmax_i = reinterpret_cast<int *>(colmax);  /* the same storage, for the integer maxima of the rescale */
    int *bcs_s = reinterpret_cast<int *>(scale + TL_TS);     /* [2][64] constant-column flag of the staged child */
    unsigned char *tipc = reinterpret_cast<unsigned char *>(bcs_s + 2 * TL_TS);   /* [Et][64] codes of the tips (if staged) */

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;            /* fragment coordinates */
    const int c = blockIdx.y;
    const int n = a.n, Sc = a.Sc;
    const int s0 = blockIdx.x * TL_TS;
    const int row = warp * 8 + g;                     /* output row owned by this thread */
    const size_t cN = (size_t)c * a.t.N, cE = (size_t)c * a.t.E;
    const bool site_thread = tid < TL_TS;
    const bool site_in = site_thread && s0 + tid < Sc;
    const bool al16 = (Sc & 1) == 0;                  /* rows of the [state][site] arrays are 16-byte aligned */

    /* rows n..63 of both buffers stay zero (k padding of the GEMM) */
    for (int i = tid; i < 2 * NP * TL_LS; i += NT) Lbuf[i] = 0.0;
    /* the tile's tip codes: one coalesced pass instead of dependent code -> table loads per tip edge */
    if (a.tip_stage) {
        const unsigned char *cd = (const unsigned char *)a.codes;
        for (int i = tid; i < a.Et * TL_TS; i += NT) {
            const int te = i >> 6, s = i & 63;
            const int b = a.t.indices[a.tip_edge_csr[te]];
            tipc[i] = (s0 + s < Sc) ? cd[(size_t)b * a.S + a.s0 + s0 + s] : 0;
        }
    }
    __syncthreads();

    /* exponent / flag of the staged child, one tile ahead, in the registers of the site threads */
    int kb_nxt = 0, bc_nxt = 0;
    auto prefetch = [&](int p) {
        const int idx = a.int_seq[p];
        const int b = a.t.indices[idx];
        double *dst = Lbuf + (size_t)(p & 1) * NP * TL_LS;
        const double *Lb = a.Lg + ((cN + b) * n) * Sc;
        if (al16) {
            for (int i = tid; i < n * (TL_TS / 2); i += NT) {
                const int k = i >> 5, s = (i & 31) * 2;
                const int left = Sc - (s0 + s);
                const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
                tl_cp16(dst + k * TL_LS + s, bytes ? (const void *)(Lb + (size_t)k * Sc + s0 + s) : (const void *)Lb, bytes);
            }
        } else {
            for (int i = tid; i < n * TL_TS; i += NT) {
                const int k = i >> 6, s = i & 63;
                const int bytes = (s0 + s < Sc) ? 8 : 0;
                tl_cp8(dst + k * TL_LS + s, bytes ? (const void *)(Lb + (size_t)k * Sc + s0 + s) : (const void *)Lb, bytes);
            }
        }
        tl_cp_commit();
        {
            /* this thread's A-fragment row of the edge's matrix: 61 doubles = 4 lines */
            const char *pr = reinterpret_cast<const char *>(a.P + (cE + idx) * n * n + (size_t)(row < n ? row : 0) * n);
            if (q == 0) {
                for (int o = 0; o < n * 8; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + o));
            }
        }
        if (site_thread) {
            kb_nxt = site_in ? a.Kg[(cN + b) * Sc + s0 + tid] : 0;
            bc_nxt = site_in ? a.Cg[(cN + b) * Sc + s0 + tid] : 0;
            bcs_s[(p & 1) * TL_TS + tid] = bc_nxt;
        }
    };
    /* p = next GEMM edge to consume, issued = GEMM tiles requested so far (p <= issued <= p + 1).  A tile may be
     * requested only when its node has been stored: seq_u[t] is the walk position at which that happens. */
    int p = 0, issued = 0;
    const int *seq_u = a.int_seq + a.n_int_seq;

    for (int u = a.t.N - 1; u >= 0; u--) {
        const int nd = a.t.preorder[u];
        const int start = a.t.indptr[nd], stop = a.t.indptr[nd + 1];
        if (start == stop) continue;                  /* leaves are written by generic_leaf_kernel */
        double acc[8][2];
        int kacc = 0, cst = 1, own_code = -1;         /* meaningful in the site threads */
        if (site_in && a.t.node_has_data[nd]) {
            own_code = plf_code_at(a.codes, a.code_bytes, a.S, nd, a.s0 + s0 + tid);
            cst = a.def_const[own_code];
        }
        for (int idx = start; idx < stop; idx++) {
            const int b = a.t.indices[idx];
            const int te = a.tip_of_edge ? a.tip_of_edge[idx] : -1;
            double em[8][2];
            if (te >= 0) {
                /* tip child: no GEMM, gather P_e def_k for this thread's 16 sites through L1 */
                const double *Tt = a.TP + (((size_t)c * a.Et + te) * a.K) * n + (row < n ? row : 0);
                if (a.tip_stage) {
                    const unsigned char *tc = tipc + te * TL_TS + q * 2;
#pragma unroll
                    for (int nb = 0; nb < 8; nb++) {
                        const uchar2 cc = *reinterpret_cast<const uchar2 *>(tc + nb * 8);
                        em[nb][0] = __ldg(Tt + (int)cc.x * n);
                        em[nb][1] = __ldg(Tt + (int)cc.y * n);
                    }
                    if (row >= n) {
#pragma unroll
                        for (int nb = 0; nb < 8; nb++) { em[nb][0] = 0.0; em[nb][1] = 0.0; }
                    }
                    if (site_in) cst &= a.def_const[tipc[te * TL_TS + tid]];
                } else {
#pragma unroll
                    for (int nb = 0; nb < 8; nb++)
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int s = nb * 8 + q * 2 + h;
                            const int code = (s0 + s < Sc) ? plf_code_at(a.codes, a.code_bytes, a.S, b, a.s0 + s0 + s) : 0;
                            em[nb][h] = (row < n) ? __ldg(Tt + (size_t)code * n) : 0.0;
                        }
                    if (site_in) cst &= a.def_const[plf_code_at(a.codes, a.code_bytes, a.S, b, a.s0 + s0 + tid)];
                }
            } else {
                const int buf = p & 1;
                const double *Ls = Lbuf + (size_t)buf * NP * TL_LS;
                /* A fragments of P_e straight from global memory (hot in L1/L2, shared by all CTAs) */
                const double *Pm = a.P + (cE + idx) * n * n;
                double af[NP / 4];
#pragma unroll
                for (int kk = 0; kk < NP / 4; kk++) {
                    const int k = kk * 4 + q;
                    af[kk] = (row < n && k < n) ? __ldg(Pm + row * n + k) : 0.0;
                }
                if (issued == p) { prefetch(p); issued++; }      /* not requested ahead of time: fetch it now */
                tl_cp_wait();
                __syncthreads();          /* tile p has landed; every warp is done with tile p-1 (other buffer) */
                const int kb_cur = kb_nxt, bc_cur = bc_nxt;
                if (p + 1 < a.n_int_seq && seq_u[p + 1] > u) { prefetch(p + 1); issued++; }
                /* em = P_e . L_b on the FP64 tensor pipe */
#pragma unroll
                for (int nb = 0; nb < 8; nb++) { em[nb][0] = 0.0; em[nb][1] = 0.0; }
#pragma unroll
                for (int kk = 0; kk < NP / 4; kk++) {
#pragma unroll
                    for (int nb = 0; nb < 8; nb++) {
                        const double bf = Ls[(kk * 4 + q) * TL_LS + nb * 8 + g];
                        tl_dmma(em[nb][0], em[nb][1], af[kk], bf);
                    }
                }
                /* a constant column maps to itself (arb_mat_extras.c:84-91): row 0 of the child tile */
#pragma unroll
                for (int nb = 0; nb < 8; nb++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int s = nb * 8 + q * 2 + h;
                        if (bcs_s[buf * TL_TS + s] && row < n) em[nb][h] = Ls[s];
                    }
                kacc += kb_cur;
                cst &= bc_cur;
                p++;
            }
            /* keep the edge vectors for the outside pass; multiply in */
#pragma unroll
            for (int nb = 0; nb < 8; nb++) {
                if (keep_edges && row < n) {
                    const int s = nb * 8 + q * 2;
                    double *Eo = a.Eg + ((cE + idx) * n + row) * Sc + s0 + s;
                    if (s0 + s + 1 < Sc && al16) *reinterpret_cast<double2 *>(Eo) = make_double2(em[nb][0], em[nb][1]);
                    else { if (s0 + s < Sc) Eo[0] = em[nb][0]; if (s0 + s + 1 < Sc) Eo[1] = em[nb][1]; }
                }
                if (idx == start) { acc[nb][0] = em[nb][0]; acc[nb][1] = em[nb][1]; }
                else { acc[nb][0] *= em[nb][0]; acc[nb][1] *= em[nb][1]; }
            }
            /* per-site max over all rows -> rescale (exponents in units of 2^256), every second child */
            if (((idx - start) & 1) || idx == stop - 1) {
#pragma unroll
                for (int nb = 0; nb < 8; nb++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        /* non-negative doubles order like their high words: the column maximum only decides
                         * whether (and how often) to multiply by 2^256, so 32 bits of it are enough */
                        int m = __double2hiint(acc[nb][h]);
                        m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
                        m = max(m, __shfl_xor_sync(0xffffffffu, m, 8));
                        m = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
                        if (g == 0) colmax_i[warp * TL_TS + nb * 8 + q * 2 + h] = m;
                    }
                __syncthreads();
                if (site_thread) {
                    int m = 0;
#pragma unroll
                    for (int w = 0; w < NW; w++) m = max(m, colmax_i[w * TL_TS + tid]);
                    /* high word of 2^-256 is 0x2FF00000; one step of 2^256 adds 0x10000000 to it.  A column
                     * whose high word is 0 (zero, or below 2^-1042) is left alone. */
                    double sc = 1.0;
                    while (m > 0 && m < 0x2FF00000) { m += 0x10000000; sc *= PLF_TWO_P256; kacc -= 1; }
                    scale[tid] = sc;
                }
                __syncthreads();
#pragma unroll
                for (int nb = 0; nb < 8; nb++) {
                    acc[nb][0] *= scale[nb * 8 + q * 2];
                    acc[nb][1] *= scale[nb * 8 + q * 2 + 1];
                }
            }
        }
        /* base vector of a node that carries data (rare) */
        if (a.t.node_has_data[nd]) {
#pragma unroll
            for (int nb = 0; nb < 8; nb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int s = nb * 8 + q * 2 + h;
                    if (s0 + s < Sc && row < n) {
                        const int code = plf_code_at(a.codes, a.code_bytes, a.S, nd, a.s0 + s0 + s);
                        acc[nb][h] *= a.defs[(size_t)code * n + row];
                    }
                }
        }
        /* store the node's partials, exponents and flags */
        if (row < n) {
            double *La = a.Lg + ((cN + nd) * n + row) * Sc + s0;
#pragma unroll
            for (int nb = 0; nb < 8; nb++) {
                const int s = nb * 8 + q * 2;
                if (s0 + s + 1 < Sc && al16) *reinterpret_cast<double2 *>(La + s) = make_double2(acc[nb][0], acc[nb][1]);
                else { if (s0 + s < Sc) La[s] = acc[nb][0]; if (s0 + s + 1 < Sc) La[s + 1] = acc[nb][1]; }
            }
        }
        if (site_in) {
            a.Kg[(cN + nd) * Sc + s0 + tid] = kacc;
            a.Cg[(cN + nd) * Sc + s0 + tid] = (unsigned char)cst;
        }
        if (issued == p && p < a.n_int_seq && seq_u[p] >= u) {
            /* the next GEMM tile is this node or an earlier one (a node without GEMM children, e.g. a cherry,
             * is not followed by a request of its own): request it now, behind this node's stores */
            __syncthreads();
            prefetch(p);
            issued++;
        }
        if (nd == a.t.root) {
            /* root_prior_expectation (model.c:282-350): weighted column sums */
            __syncthreads();
#pragma unroll
            for (int nb = 0; nb < 8; nb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    double r = 0.0;
                    if (row < n) {
                        if (a.root_mode == PLF_ROOT_NONE) r = 1.0;
                        else if (a.root_mode == PLF_ROOT_UNIFORM) r = 1.0 / (double)n;
                        else r = a.root_vec[row];
                    }
                    double v = r * acc[nb][h];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == 0) colmax[warp * TL_TS + nb * 8 + q * 2 + h] = v;
                }
            /* acc row 0 for the constant-column shortcut of the uniform / equilibrium priors */
            if (warp == 0 && g == 0) {
#pragma unroll
                for (int nb = 0; nb < 8; nb++) { scale[nb * 8 + q * 2] = acc[nb][0]; scale[nb * 8 + q * 2 + 1] = acc[nb][1]; }
            }
            __syncthreads();
            if (site_in) {
                double lh = 0.0;
                for (int w = 0; w < NW; w++) lh += colmax[w * TL_TS + tid];
                if (cst && (a.root_mode == PLF_ROOT_UNIFORM || a.root_mode == PLF_ROOT_EQUILIBRIUM)) lh = scale[tid];
                a.cat_lh[(size_t)c * Sc + s0 + tid] = lh;
                a.cat_k[(size_t)c * Sc + s0 + tid] = kacc;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* outside pass on the FP64 tensor pipe                                                         */
/* ------------------------------------------------------------------------------------------ */

/* fragment helpers: a thread owns row `row` and the 16 sites {8 nb + 2 q + h} of the tile */
#define TL_FOR_FRAG(nb, h) _Pragma("unroll") for (int nb = 0; nb < 8; nb++) _Pragma("unroll") for (int h = 0; h < 2; h++)

__device__ __forceinline__ void tl_load_frag(double (&f)[8][2], const double *rows0, int row, int n, int Sc, int s0, int q)
{
    const double *p = rows0 + (size_t)row * Sc + s0;
    TL_FOR_FRAG(nb, h) {
        const int s = nb * 8 + q * 2 + h;
        f[nb][h] = (row < n && s0 + s < Sc) ? p[s] : 0.0;
    }
}

/* per-site (column) reduction over all 64 rows: result[s] for s = tid < 64, valid after the call */
template <bool IS_MAX, int NW>
__device__ __forceinline__ void tl_col_reduce(const double (&f)[8][2], double *colbuf, double *out, int warp, int g, int q, int tid)
{
    double r[8][2];
    TL_FOR_FRAG(nb, h) {
        double m = f[nb][h];
        if (IS_MAX) {
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 16));
        } else {
            m += __shfl_xor_sync(0xffffffffu, m, 4);
            m += __shfl_xor_sync(0xffffffffu, m, 8);
            m += __shfl_xor_sync(0xffffffffu, m, 16);
        }
        r[nb][h] = m;
    }
    __syncthreads();
    if (g == 0) {
        TL_FOR_FRAG(nb, h) colbuf[warp * TL_TS + nb * 8 + q * 2 + h] = r[nb][h];
    }
    __syncthreads();
    if (tid < TL_TS) {
        double m = 0.0;
        for (int w = 0; w < NW; w++) m = IS_MAX ? fmax(m, colbuf[w * TL_TS + tid]) : m + colbuf[w * TL_TS + tid];
        out[tid] = m;
    }
    __syncthreads();
}

/*
 * Outside pass for 16 < n <= 64 (evaluate_site_forward.c:52-102, evaluate_site_frechet.c:18-39,
 * evaluate_site_marginal.c:14-20).  One CTA owns 64 sites and loops over the categories (so that
 * every output cell has a single writer).  Per child edge: fe = fn_a .* base_a .* prod of the
 * siblings' edge vectors (kept by the inside pass), y = F_e L_b and fn_b = P_e^T fe on DMMA.
 * grid = (site tiles), 256 threads.
 */
template <int NP>
__global__ void __launch_bounds__(NP * 4, (NP == 64 ? 2 : 4)) tile_outside_kernel(GenericArgs a)
{
    constexpr int NW = NP / 8;                        /* NP * 4 threads: one warp per 8 rows */
    extern __shared__ __align__(16) double tl_sm[];
    double *Bsm = tl_sm;                              /* [64][TL_LS] B operand: child partials or fe */
    double *colbuf = Bsm + NP * TL_LS;             /* [8][64] */
    double *colres = colbuf + NW * TL_TS;              /* [64] */
    double *coef = colres + TL_TS;                    /* [64] prior_c / site likelihood (0 if the category is dead) */
    double *scl = coef + TL_TS;                       /* [64] */
    int *ka = reinterpret_cast<int *>(scl + TL_TS);   /* [64] exponent of fn_a */
    int *kfe = ka + TL_TS;                            /* [64] */
    int *kbv = kfe + TL_TS;                           /* [64] exponent of the child's partial */
    int *bcv = kbv + TL_TS;                           /* [64] constant-column flag of the child */
    int *codev = bcv + TL_TS;                         /* [64] */
    int *sitek = codev + TL_TS;                       /* [64] */

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int n = a.n, Sc = a.Sc;
    const int s0 = blockIdx.x * TL_TS;
    const int row = warp * 8 + g;
    const bool site_in = (tid < TL_TS) && (s0 + tid < Sc);

    for (int c = 0; c < a.C; c++) {
        const size_t cN = (size_t)c * a.t.N, cE = (size_t)c * a.t.E;
        __syncthreads();
        if (tid < TL_TS) {
            double cf = 0.0;
            int sk = 0;
            if (site_in) {
                const double pl = a.cat_prior[c] * a.cat_lh[(size_t)c * Sc + s0 + tid];
                const double m = a.site_m[s0 + tid];
                if (pl > 0.0 && m > 0.0) cf = a.cat_prior[c] / m;       /* arbplfmarginal.c:184-191 */
                sk = a.site_k[s0 + tid];
            }
            coef[tid] = cf;
            sitek[tid] = sk;
        }
        /* forward vector at the root: root_prior_mul_col_vec (model.c:225-280) */
        if (row < n) {
            double r = 1.0;
            if (a.root_mode == PLF_ROOT_UNIFORM) r = 1.0 / (double)n;
            else if (a.root_mode == PLF_ROOT_EQUILIBRIUM || a.root_mode == PLF_ROOT_CUSTOM) r = a.root_vec[row];
            double *Fr = a.Fg + ((cN + a.t.root) * n + row) * Sc + s0;
            TL_FOR_FRAG(nb, h) { const int s = nb * 8 + q * 2 + h; if (s0 + s < Sc) Fr[s] = r; }
        }
        if (site_in) a.FK[(cN + a.t.root) * Sc + s0 + tid] = 0;
        __syncthreads();

        for (int u = 0; u < a.t.N; u++) {
            const int nd = a.t.preorder[u];
            const int start = a.t.indptr[nd], stop = a.t.indptr[nd + 1];
            const bool leaf = start == stop;
            if (leaf && !a.want_marg) continue;
            double fa[8][2];
            tl_load_frag(fa, a.Fg + (cN + nd) * n * Sc, row, n, Sc, s0, q);
            __syncthreads();
            if (tid < TL_TS) ka[tid] = site_in ? a.FK[(cN + nd) * Sc + s0 + tid] : 0;
            __syncthreads();
            if (a.want_marg) {
                /* marg_a += prior_c fn_a .* L_a / site_L */
                double la[8][2];
                tl_load_frag(la, a.Lg + (cN + nd) * n * Sc, row, n, Sc, s0, q);
                if (tid < TL_TS) kbv[tid] = site_in ? a.Kg[(cN + nd) * Sc + s0 + tid] : 0;
                __syncthreads();
                if (row < n) {
                    double *Mo = a.marg_out + ((size_t)nd * n + row) * Sc + s0;
                    TL_FOR_FRAG(nb, h) {
                        const int s = nb * 8 + q * 2 + h;
                        if (s0 + s < Sc && coef[s] != 0.0)
                            Mo[s] += scalbn(coef[s] * fa[nb][h] * la[nb][h], PLF_SCALE_BITS * (ka[s] + kbv[s] - sitek[s]));
                    }
                }
                __syncthreads();
            }
            if (leaf) continue;
            if (a.t.node_has_data[nd]) {
                TL_FOR_FRAG(nb, h) {
                    const int s = nb * 8 + q * 2 + h;
                    if (row < n && s0 + s < Sc) {
                        const int code = plf_code_at(a.codes, a.code_bytes, a.S, nd, a.s0 + s0 + s);
                        fa[nb][h] *= a.defs[(size_t)code * n + row];
                    }
                }
            }
            for (int idx = start; idx < stop; idx++) {
                const int b = a.t.indices[idx];
                const bool b_leaf = a.t.indptr[b] == a.t.indptr[b + 1];
                const bool want_x = a.want_edge && (!a.edge_mask || a.edge_mask[idx]);
                const bool want_fn = !b_leaf || a.want_marg;
                if (!want_x && !want_fn) continue;
                /* fe = fa .* prod_{sib != idx} Em_sib ; exponent kfe = ka + sum K_sib.  Every factor has a column
                 * maximum of at least 2^-256 p_min, so two of them cannot underflow: the mantissas are pulled
                 * back into range after every second sibling only (the GEMM result is rescaled anyway). */
                double fe[8][2];
                TL_FOR_FRAG(nb, h) fe[nb][h] = fa[nb][h];
                int kfe_r = 0, nsib = 0;                 /* site threads: exponent of fe (on top of ka) */
                for (int idx2 = start; idx2 < stop; idx2++) {
                    if (idx2 == idx) continue;
                    double es[8][2];
                    tl_load_frag(es, a.Eg + (cE + idx2) * n * Sc, row, n, Sc, s0, q);
                    TL_FOR_FRAG(nb, h) fe[nb][h] *= es[nb][h];
                    if (site_in) kfe_r += a.Kg[(cN + a.t.indices[idx2]) * Sc + s0 + tid];
                    if ((++nsib & 1) == 0) {
                        tl_col_reduce<true, NW>(fe, colbuf, colres, warp, g, q, tid);
                        if (tid < TL_TS) {
                            double m = colres[tid], sc = 1.0;
                            while (m > 0.0 && m < PLF_TWO_M256) { m *= PLF_TWO_P256; sc *= PLF_TWO_P256; kfe_r -= 1; }
                            while (m > PLF_TWO_P256) { m *= PLF_TWO_M256; sc *= PLF_TWO_M256; kfe_r += 1; }
                            scl[tid] = sc;
                        }
                        __syncthreads();
                        TL_FOR_FRAG(nb, h) fe[nb][h] *= scl[nb * 8 + q * 2 + h];
                    }
                }
                /* site threads: exponent and flags of the child */
                const int te = (b_leaf && a.TF && a.tip_of_edge) ? a.tip_of_edge[idx] : -1;
                int kb_r = 0, bc_r = 0;
                if (site_in) {
                    kfe_r += ka[tid];
                    if (te < 0) {       /* a tip handled through the tip table has no stored vector (exponent 0) */
                        kb_r = a.Kg[(cN + b) * Sc + s0 + tid];
                        bc_r = a.Cg[(cN + b) * Sc + s0 + tid];
                    }
                }
                if (want_x && te >= 0) {
                    /* tip child: x_e = fe . (F_e def_k) from the tip table, no GEMM */
                    const double *Tt = a.TF + (((size_t)c * a.Et + te) * a.K) * n + (row < n ? row : 0);
                    double y[8][2];
                    TL_FOR_FRAG(nb, h) {
                        const int s = nb * 8 + q * 2 + h;
                        const int code = (s0 + s < Sc) ? plf_code_at(a.codes, a.code_bytes, a.S, b, a.s0 + s0 + s) : 0;
                        y[nb][h] = (row < n) ? __ldg(Tt + (size_t)code * n) * fe[nb][h] : 0.0;
                    }
                    tl_col_reduce<false, NW>(y, colbuf, colres, warp, g, q, tid);
                    if (site_in && coef[tid] != 0.0)
                        a.edge_out[(size_t)idx * Sc + s0 + tid] += scalbn(colres[tid] * coef[tid], PLF_SCALE_BITS * (kfe_r + kb_r - sitek[tid]));
                    if (!want_fn) continue;
                }
                /* fe is the B operand of both products: z = F_e^T fe (then x_e = z . L_b) and fn_b = P_e^T fe */
                __syncthreads();
                TL_FOR_FRAG(nb, h) Bsm[row * TL_LS + nb * 8 + q * 2 + h] = fe[nb][h];
                __syncthreads();
                if (want_x && te < 0) {
                    /* x_e = fe^T (F_e L_b) = (F_e^T fe)^T L_b  (evaluate_site_frechet.c:18-39) */
                    const double *Fm = a.Fm + (cE + idx) * n * n;
                    double af[NP / 4];
#pragma unroll
                    for (int kk = 0; kk < NP / 4; kk++) {
                        const int k = kk * 4 + q;
                        af[kk] = (row < n && k < n) ? __ldg(Fm + k * n + row) : 0.0;     /* A = F^T */
                    }
                    double lb[8][2];
                    tl_load_frag(lb, a.Lg + (cN + b) * n * Sc, row, n, Sc, s0, q);
                    double z[8][2];
                    TL_FOR_FRAG(nb, h) z[nb][h] = 0.0;
#pragma unroll
                    for (int kk = 0; kk < NP / 4; kk++) {
#pragma unroll
                        for (int nb = 0; nb < 8; nb++) tl_dmma(z[nb][0], z[nb][1], af[kk], Bsm[(kk * 4 + q) * TL_LS + nb * 8 + g]);
                    }
                    TL_FOR_FRAG(nb, h) z[nb][h] *= lb[nb][h];
                    tl_col_reduce<false, NW>(z, colbuf, colres, warp, g, q, tid);
                    if (site_in && coef[tid] != 0.0 && !(a.f_zero_rowsum && bc_r))
                        a.edge_out[(size_t)idx * Sc + s0 + tid] += scalbn(colres[tid] * coef[tid], PLF_SCALE_BITS * (kfe_r + kb_r - sitek[tid]));
                }
                if (want_fn) {
                    /* fn_b = P_e^T fe  (util.c:464-498) */
                    const double *Pm = a.P + (cE + idx) * n * n;
                    double af[NP / 4];
#pragma unroll
                    for (int kk = 0; kk < NP / 4; kk++) {
                        const int k = kk * 4 + q;
                        af[kk] = (row < n && k < n) ? __ldg(Pm + k * n + row) : 0.0;     /* A = P^T */
                    }
                    double fb[8][2];
                    TL_FOR_FRAG(nb, h) fb[nb][h] = 0.0;
#pragma unroll
                    for (int kk = 0; kk < NP / 4; kk++) {
#pragma unroll
                        for (int nb = 0; nb < 8; nb++) tl_dmma(fb[nb][0], fb[nb][1], af[kk], Bsm[(kk * 4 + q) * TL_LS + nb * 8 + g]);
                    }
                    tl_col_reduce<true, NW>(fb, colbuf, colres, warp, g, q, tid);
                    if (tid < TL_TS) {
                        double m = colres[tid], sc = 1.0;
                        int k = kfe_r;
                        while (m > PLF_TWO_P256) { m *= PLF_TWO_M256; sc *= PLF_TWO_M256; k += 1; }
                        while (m > 0.0 && m < PLF_TWO_M256) { m *= PLF_TWO_P256; sc *= PLF_TWO_P256; k -= 1; }
                        scl[tid] = sc;
                        if (site_in) a.FK[(cN + b) * Sc + s0 + tid] = k;
                    }
                    __syncthreads();
                    if (row < n) {
                        double *Fb = a.Fg + ((cN + b) * n + row) * Sc + s0;
                        TL_FOR_FRAG(nb, h) {
                            const int s = nb * 8 + q * 2 + h;
                            if (s0 + s < Sc) Fb[s] = fb[nb][h] * scl[s];
                        }
                    }
                }
            }
        }
    }
}
