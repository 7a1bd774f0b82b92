/*
 * Fused 4-state (DNA) pruning kernel: the hot path of arbplf-ll / arbplf-deriv
 * (and of dwell / trans, which only change the per-edge matrices).
 *
 * Reference loop nest being replaced (arbplfll.c:139-170, arbplfderiv.c:274-357):
 *   for site: for category: evaluate_site_lhood (evaluate_site_lhood.c:21-57,
 *   _prune_update_prob util.c:241-301), then per requested edge a root-path
 *   recomputation (arbplfderiv.c:144-205).
 *
 * B200 design: one thread owns one site pattern and walks the whole tree
 * on chip.  The tree is compiled on the host into a post-order "program" of
 * node operations with Sethi-Ullman child ordering; the most recent partial
 * stays in registers, older pending partials sit on a small per-thread stack
 * in shared memory.  Tips never touch fp64 matrix-vector work: P_e.def_k and
 * F_e.def_k are precomputed per (category, edge, character) ("tip tables").
 * The derivative / edge-expectation outputs come from an outside (pre-order)
 * pass that runs the same program backwards, using the identity
 *     d L_c / d t_e = rate_c * fe_e^T (Q P_e) L_b
 * instead of the reference's O(depth) walk per edge.  Per-edge values are
 * reduced over the 32 sites of a warp with shuffles and accumulated per warp in
 * shared memory; one deterministic second-stage kernel adds the per-CTA rows.
 * HBM traffic: 1 byte per (site, tip) of codes, plus the inside partials of the
 * internal nodes (32 B per node, category, site) when the outside pass runs.
 */
#pragma once
#include <stdint.h>

#define F4_MAXD 3          /* max out-degree handled by the fused kernel */
#define F4_KIND_TIP 0
#define F4_KIND_CUR 1
#define F4_KIND_STACK 2

struct F4Op {
    int node;
    int first_child;
    int nchild;
    int slot;            /* scratch slot of this node's inside vector */
    int has_data;
    int spill_before;    /* push the register-resident partial before this op */
};

struct F4Child {
    int edge;            /* csr idx */
    int node;
    int slot;            /* scratch slot if internal, else -1 */
    int kind;
};

struct F4Args {
    int nops;
    const F4Op *ops;
    const F4Child *children;
    int C, E, K;
    int64_t S;
    const unsigned char *codes;      /* [N][S] */
    const unsigned char *node_has_data; /* [N] */
    const double *defs;              /* [K][4] */
    const unsigned char *def_const;  /* [K] */
    const double *P;                 /* [C][E][16] */
    const double *TP;                /* [C][E][K][4] */
    const double *Fm;                /* [C][E][16] or NULL */
    const double *TF;                /* [C][E][K][4] or NULL */
    int f_zero_rowsum;
    const double *cat_prior;
    int root_mode;
    double root_vec[4];
    const double *site_w;            /* [S] or NULL */
    const unsigned char *edge_mask;  /* [E] or NULL */
    int stack_depth;
    int nslots;
    double4 *scratch;                /* [C][nslots][T] */
    signed char *scratchS;           /* [C][nslots][T] */
    double *site_ll;                 /* [S] or NULL */
    double *edge_site_out;           /* [E][S] or NULL */
    double *block_ll;                /* [grid] */
    double *block_edge;              /* [grid][E] */
    int *error_flag;
};

__device__ __forceinline__ void f4_matvec(const double *__restrict__ M, const double v[4], double out[4])
{
    /* M row-major 4x4, uniform address across the warp: broadcast loads */
    const double2 *M2 = reinterpret_cast<const double2 *>(M);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double2 a = __ldg(M2 + 2 * i), b = __ldg(M2 + 2 * i + 1);
        double r = a.x * v[0];
        r = fma(a.y, v[1], r);
        r = fma(b.x, v[2], r);
        r = fma(b.y, v[3], r);
        out[i] = r;
    }
}

/* out = M^T v */
__device__ __forceinline__ void f4_matvec_t(const double *__restrict__ M, const double v[4], double out[4])
{
    const double2 *M2 = reinterpret_cast<const double2 *>(M);
    double2 a0 = __ldg(M2 + 0), b0 = __ldg(M2 + 1);
    out[0] = a0.x * v[0]; out[1] = a0.y * v[0]; out[2] = b0.x * v[0]; out[3] = b0.y * v[0];
#pragma unroll
    for (int i = 1; i < 4; i++) {
        double2 a = __ldg(M2 + 2 * i), b = __ldg(M2 + 2 * i + 1);
        out[0] = fma(a.x, v[i], out[0]);
        out[1] = fma(a.y, v[i], out[1]);
        out[2] = fma(b.x, v[i], out[2]);
        out[3] = fma(b.y, v[i], out[3]);
    }
}

__device__ __forceinline__ void f4_ld4(const double *__restrict__ p, double out[4])
{
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(p2), b = __ldg(p2 + 1);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}

__device__ __forceinline__ int f4_rescale_up(double v[4])
{
    double m = fmax(fmax(v[0], v[1]), fmax(v[2], v[3]));
    int s = 0;
    while (m > 0.0 && m < PLF_TWO_M256) {
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] *= PLF_TWO_P256;
        m *= PLF_TWO_P256;
        s++;
    }
    return s;
}

__device__ __forceinline__ double f4_warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

/*
 * EDGE = false : log-likelihood only (no scratch, no outside pass)
 * EDGE = true  : log-likelihood + per-edge bilinear forms with matrices Fm
 * Launch: persistent grid; dynamic shared memory =
 *   blockDim*(stack_depth*36 + 4*4) + (blockDim/32)*E*8 bytes
 */
template <bool EDGE>
__global__ void __launch_bounds__(256) fused4_kernel(F4Args a)
{
    extern __shared__ __align__(16) unsigned char f4_smem[];
    const int tid = threadIdx.x, bd = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = bd >> 5;
    double *stack = reinterpret_cast<double *>(f4_smem);                    /* [depth][4][bd] */
    int *stackf = reinterpret_cast<int *>(stack + (size_t)a.stack_depth * 4 * bd);   /* [depth][bd] */
    int *kcat = stackf + (size_t)a.stack_depth * bd;                        /* [4][bd] */
    double *accE = reinterpret_cast<double *>(kcat + 4 * bd);               /* [nwarp][E] */
    const int64_t T = (int64_t)gridDim.x * bd;
    const int64_t gtid = (int64_t)blockIdx.x * bd + tid;

    if (EDGE) {
        for (int i = tid; i < nwarp * a.E; i += bd) accE[i] = 0.0;
        __syncthreads();
    }
    double ll_acc = 0.0;
    const int64_t ntiles = (a.S + bd - 1) / bd;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t site_raw = tile * bd + tid;
        const bool valid = site_raw < a.S;
        const int64_t site = valid ? site_raw : a.S - 1;
        const double w = valid ? (a.site_w ? a.site_w[site] : 1.0) : 0.0;

        /* ---------------- inside pass, one category at a time ---------------- */
        double site_m = 0.0;
        int site_k = 0;
        bool have = false;
#pragma unroll 1
        for (int c = 0; c < a.C; c++) {
            double cur[4] = {1.0, 1.0, 1.0, 1.0};
            int curf = 1;
            int sp = 0;
            int ktot = 0;
            const double *Pc = a.P + (size_t)c * a.E * 16;
            const double *TPc = a.TP + (size_t)c * a.E * a.K * 4;
#pragma unroll 1
            for (int o = 0; o < a.nops; o++) {
                const F4Op op = a.ops[o];
                if (op.spill_before) {
#pragma unroll
                    for (int i = 0; i < 4; i++) stack[((size_t)sp * 4 + i) * bd + tid] = cur[i];
                    stackf[(size_t)sp * bd + tid] = curf;
                    sp++;
                }
                double acc[4];
                int cst = 1;
                if (a.node_has_data[op.node]) {
                    int code = a.codes[(size_t)op.node * a.S + site];
                    f4_ld4(a.defs + code * 4, acc);
                    cst = a.def_const[code];
                } else {
                    acc[0] = acc[1] = acc[2] = acc[3] = 1.0;
                }
                int sloc = 0;
#pragma unroll 1
                for (int j = 0; j < op.nchild; j++) {
                    const F4Child ch = a.children[op.first_child + j];
                    double em[4];
                    int bc;
                    if (ch.kind == F4_KIND_TIP) {
                        int code = a.codes[(size_t)ch.node * a.S + site];
                        f4_ld4(TPc + ((size_t)ch.edge * a.K + code) * 4, em);
                        bc = a.def_const[code];
                    } else {
                        double v[4];
                        if (ch.kind == F4_KIND_CUR) {
#pragma unroll
                            for (int i = 0; i < 4; i++) v[i] = cur[i];
                            bc = curf;
                        } else {
                            sp--;
#pragma unroll
                            for (int i = 0; i < 4; i++) v[i] = stack[((size_t)sp * 4 + i) * bd + tid];
                            bc = stackf[(size_t)sp * bd + tid];
                        }
                        if (bc) {
#pragma unroll
                            for (int i = 0; i < 4; i++) em[i] = v[i];
                        } else {
                            f4_matvec(Pc + (size_t)ch.edge * 16, v, em);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) acc[i] *= em[i];
                    sloc += f4_rescale_up(acc);
                    cst &= bc;
                }
                ktot -= sloc;
#pragma unroll
                for (int i = 0; i < 4; i++) cur[i] = acc[i];
                curf = cst;
                if (EDGE) {
                    const size_t off = ((size_t)c * a.nslots + op.slot) * T + gtid;
                    a.scratch[off] = make_double4(acc[0], acc[1], acc[2], acc[3]);
                    a.scratchS[off] = (signed char)(sloc | (cst << 6));
                }
            }
            /* root prior expectation (model.c:282-350) */
            double lh;
            if (a.root_mode == PLF_ROOT_NONE) {
                lh = (cur[0] + cur[1]) + (cur[2] + cur[3]);
            } else if (a.root_mode == PLF_ROOT_UNIFORM) {
                lh = curf ? cur[0] : ((cur[0] + cur[1]) + (cur[2] + cur[3])) * 0.25;
            } else if (a.root_mode == PLF_ROOT_EQUILIBRIUM && curf) {
                lh = cur[0];
            } else {
                lh = a.root_vec[0] * cur[0];
                lh = fma(a.root_vec[1], cur[1], lh);
                lh = fma(a.root_vec[2], cur[2], lh);
                lh = fma(a.root_vec[3], cur[3], lh);
            }
            const double v = a.cat_prior[c] * lh;
            kcat[c * bd + tid] = (v > 0.0) ? ktot : INT_MIN;
            if (v > 0.0) {
                if (!have) { site_m = v; site_k = ktot; have = true; }
                else if (ktot > site_k) { site_m = scalbn(site_m, PLF_SCALE_BITS * (site_k - ktot)) + v; site_k = ktot; }
                else site_m += scalbn(v, PLF_SCALE_BITS * (ktot - site_k));
            }
        }
        /* ---------------- site log-likelihood ---------------- */
        {
            const double c_hi = 177.445678223346, c_lo = 5.936759843446527e-15;
            double ll = log(site_m);
            ll = fma((double)site_k, c_hi, ll);
            ll = fma((double)site_k, c_lo, ll);
            if (valid) {
                if (a.site_ll) a.site_ll[site] = ll;
                if (w != 0.0) {
                    if (!have) atomicOr(a.error_flag, 1);
                    else ll_acc = fma(w, ll, ll_acc);
                }
            }
        }
        if (!EDGE) continue;

        /* ---------------- outside pass ---------------- */
        const double inv_site = (have && w != 0.0) ? (a.edge_site_out ? 1.0 : w) / site_m : 0.0;
#pragma unroll 1
        for (int c = 0; c < a.C; c++) {
            const int kc = kcat[c * bd + tid];
            const bool alive = (kc != INT_MIN);
            /* fn_root = root prior vector * prior_c * w / site_L, scaled so that fn .* L is O(1) */
            const double sc0 = alive ? scalbn(a.cat_prior[c] * inv_site, PLF_SCALE_BITS * (kc - site_k)) : 0.0;
            double curF[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double r = 1.0;
                if (a.root_mode == PLF_ROOT_UNIFORM) r = 0.25;
                else if (a.root_mode == PLF_ROOT_EQUILIBRIUM || a.root_mode == PLF_ROOT_CUSTOM) r = a.root_vec[i];
                curF[i] = r * sc0;
            }
            int sp = 0;
            const double *Pc = a.P + (size_t)c * a.E * 16;
            const double *Fc = a.Fm + (size_t)c * a.E * 16;
            const double *TPc = a.TP + (size_t)c * a.E * a.K * 4;
            const double *TFc = a.TF + (size_t)c * a.E * a.K * 4;
#pragma unroll 1
            for (int o = a.nops - 1; o >= 0; o--) {
                const F4Op op = a.ops[o];
                double fa[4];
                if (o != a.nops - 1 && a.ops[o + 1].spill_before) {
                    sp--;
#pragma unroll
                    for (int i = 0; i < 4; i++) fa[i] = stack[((size_t)sp * 4 + i) * bd + tid];
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) fa[i] = curF[i];
                }
                /* tmp = fn_a .* base_a * 2^(256 s_a) */
                {
                    const size_t off = ((size_t)c * a.nslots + op.slot) * T + gtid;
                    int sa = a.scratchS[off] & 63;
                    if (a.node_has_data[op.node]) {
                        int code = a.codes[(size_t)op.node * a.S + site];
                        double b[4];
                        f4_ld4(a.defs + code * 4, b);
#pragma unroll
                        for (int i = 0; i < 4; i++) fa[i] *= b[i];
                    }
                    while (sa > 0) {
#pragma unroll
                        for (int i = 0; i < 4; i++) fa[i] *= PLF_TWO_P256;
                        sa--;
                    }
                }
                double em[F4_MAXD][4], y[F4_MAXD][4];
                int kinds[F4_MAXD], edges[F4_MAXD];
#pragma unroll
                for (int j = 0; j < F4_MAXD; j++) {
                    kinds[j] = -1; edges[j] = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) { em[j][i] = 1.0; y[j][i] = 0.0; }
                    if (j < op.nchild) {
                        const F4Child ch = a.children[op.first_child + j];
                        kinds[j] = ch.kind; edges[j] = ch.edge;
                        if (ch.kind == F4_KIND_TIP) {
                            int code = a.codes[(size_t)ch.node * a.S + site];
                            f4_ld4(TPc + ((size_t)ch.edge * a.K + code) * 4, em[j]);
                            f4_ld4(TFc + ((size_t)ch.edge * a.K + code) * 4, y[j]);
                        } else {
                            const size_t off = ((size_t)c * a.nslots + ch.slot) * T + gtid;
                            double4 l4 = a.scratch[off];
                            int bc = (a.scratchS[off] >> 6) & 1;
                            double lv[4] = {l4.x, l4.y, l4.z, l4.w};
                            if (bc) {
#pragma unroll
                                for (int i = 0; i < 4; i++) em[j][i] = lv[i];
                            } else {
                                f4_matvec(Pc + (size_t)ch.edge * 16, lv, em[j]);
                            }
                            if (!(bc && a.f_zero_rowsum)) f4_matvec(Fc + (size_t)ch.edge * 16, lv, y[j]);
                        }
                    }
                }
                /* per child: fe_j = tmp .* prod_{i != j} em_i ; x_j = fe_j . y_j ; fn_j = P_j^T fe_j */
                double newcur[4] = {0.0, 0.0, 0.0, 0.0};
                double push[F4_MAXD][4];
#pragma unroll
                for (int j = 0; j < F4_MAXD; j++) {
                    if (j < op.nchild) {
                        double fe[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            double f = fa[i];
#pragma unroll
                            for (int j2 = 0; j2 < F4_MAXD; j2++) if (j2 != j) f *= em[j2][i];
                            fe[i] = f;
                        }
                        double x = fe[0] * y[j][0];
                        x = fma(fe[1], y[j][1], x);
                        x = fma(fe[2], y[j][2], x);
                        x = fma(fe[3], y[j][3], x);
                        const int e = edges[j];
                        if (!a.edge_mask || a.edge_mask[e]) {
                            if (a.edge_site_out) {
                                if (valid) a.edge_site_out[(size_t)e * a.S + site] += x;
                            } else {
                                double xs = f4_warp_sum(x);
                                if (lane == 0) accE[warp * a.E + e] += xs;
                            }
                        }
                        if (kinds[j] != F4_KIND_TIP) {
                            double fb[4];
                            f4_matvec_t(Pc + (size_t)e * 16, fe, fb);
                            if (kinds[j] == F4_KIND_CUR) {
#pragma unroll
                                for (int i = 0; i < 4; i++) newcur[i] = fb[i];
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; i++) push[j][i] = fb[i];
                            }
                        }
                    }
                }
                /* push stack children in reverse list order (mirror of the inside pops) */
#pragma unroll
                for (int j = F4_MAXD - 1; j >= 0; j--) {
                    if (j < op.nchild && kinds[j] == F4_KIND_STACK) {
#pragma unroll
                        for (int i = 0; i < 4; i++) stack[((size_t)sp * 4 + i) * bd + tid] = push[j][i];
                        sp++;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; i++) curF[i] = newcur[i];
            }
        }
    }

    /* ---------------- CTA-level reductions ---------------- */
    {
        __shared__ double red[32];
        double x = f4_warp_sum(ll_acc);
        if (lane == 0) red[warp] = x;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < nwarp; i++) s += red[i];
            a.block_ll[blockIdx.x] = s;
        }
        if (EDGE && !a.edge_site_out) {
            for (int e = tid; e < a.E; e += bd) {
                double s = 0.0;
                for (int wv = 0; wv < nwarp; wv++) s += accE[wv * a.E + e];
                a.block_edge[(size_t)blockIdx.x * a.E + e] = s;
            }
        }
    }
}

/* second stage: out[j] = sum over rows of part[row][j], fixed order */
__global__ void sum_rows_kernel(const double *part, int rows, int cols, double *out)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    double s = 0.0, comp = 0.0;
    for (int r = 0; r < rows; r++) {
        double yv = part[(size_t)r * cols + j] - comp;
        double tsum = s + yv;
        comp = (tsum - s) - yv;
        s = tsum;
    }
    out[j] = s;
}
