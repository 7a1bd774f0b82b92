/*
 * Fused 4-state (DNA) pruning kernels: the hot path of arbplf-ll / arbplf-deriv
 * (and of dwell / trans, which only change the per-edge matrices).
 *
 * Reference loop nest being replaced (arbplfll.c:139-170, arbplfderiv.c:274-357):
 *   for site: for category: evaluate_site_lhood (evaluate_site_lhood.c:21-57,
 *   _prune_update_prob util.c:241-301), then per requested edge a root-path
 *   recomputation (arbplfderiv.c:144-205).
 *
 * B200 design
 *   - one thread owns one (site pattern, rate category) pair and walks the whole
 *     tree with its 4-vector in registers; the categories of a site sit in
 *     adjacent lanes of one warp.  The working set of a thread is a few 4-vectors,
 *     so 16-32 warps are resident per SM.
 *   - the tree is compiled on the host into a post-order program with
 *     Sethi-Ullman child ordering; the program, the transition / derivative
 *     matrices of the internal edges and the tip tables (P_e.def_k, F_e.def_k
 *     per character k) are staged in shared memory once per CTA; the tile's tip
 *     codes are staged in shared memory once per tile (coalesced 1-byte loads).
 *   - three kernels per evaluation: inside pass (writes each internal partial
 *     once to a slab in HBM, [slot][category][site] double4: every warp access is
 *     contiguous), site combine (mixes categories, log, weighted sum), outside
 *     pass (reads each partial once, prefetching one node ahead).
 *   - derivative / dwell / trans outputs come from the outside (pre-order) pass
 *     that runs the same program backwards with the identity
 *         d L_c / d t_e = rate_c * fe_e^T (Q P_e) L_b ,
 *     replacing the reference's O(depth) walk per edge.  Per-edge values are
 *     reduced over the warp (sites x categories) with shuffles, two edges per
 *     butterfly, accumulated per warp in shared memory, and summed over CTAs
 *     by a deterministic second-stage kernel.
 * Algorithmic HBM traffic: 1 byte per (site, tip) of codes plus, when the
 * outside pass runs, 2 x 33 bytes per (internal node, category, site).
 */
#pragma once
#include <stdint.h>

#define F4_MAXD 3          /* max out-degree handled by the fused kernels */
#define F4_KIND_TIP 0
#define F4_KIND_CUR 1
#define F4_KIND_STACK 2

struct F4Op {
    int node;
    int first_child;
    int nchild;
    int slot;            /* slab slot of this node's inside vector */
    int code_row;        /* row of the codes tile if the node carries data, else -1 */
    int spill_before;    /* the register-resident partial is not consumed by this op */
    int shape;           /* has_cur * 16 + n_stack * 4 + n_tip */
    int pad;
};

struct F4Child {
    int kind;
    int slot;            /* slab slot if internal, else -1 */
    int mat;             /* internal child: index into the compact internal-edge matrices;
                            tip child: index into the compact tip tables */
    int code_row;        /* tip child: row of the codes tile */
    int edge;            /* csr idx (output position) */
    int pad[3];
};

struct F4Args {
    int nops, nchildren;
    const F4Op *ops;
    const F4Child *children;
    int C, E, K, Ei, Et;             /* categories, edges, characters, internal-child edges, tip edges */
    int64_t S;                       /* sites in this chunk */
    int64_t S_total, s0;             /* codes / weights / per-site outputs are indexed by s0 + site, stride S_total */
    int ncode_rows;
    const int *code_row_node;        /* [ncode_rows] node whose codes fill the row */
    const unsigned char *codes;      /* [N][S] */
    const double *defs;              /* [K][4] */
    const unsigned char *def_const;  /* [K] */
    const double *Pint;              /* [C][Ei][16] */
    const double *TP;                /* [C][Et][K][4] */
    const double *Fint;              /* [C][Ei][16] or NULL */
    const double *TF;                /* [C][Et][K][4] or NULL */
    int f_zero_rowsum;
    const double *cat_prior;
    int root_mode;
    double root_vec[4];
    const double *site_w;            /* [S] or NULL */
    const unsigned char *edge_mask;  /* [E] or NULL */
    int stack_depth;                 /* ll-only mode: shared-memory stack entries */
    int nslots;
    int64_t slab_stride;             /* elements per slot = C * S */
    double4 *slab;                   /* [nslots][C][S]  (index = slot*stride + c*S + site) */
    unsigned char *slabF;            /* same indexing: rescale count | const flag << 6 */
    double *cat_lh;                  /* [C][S] root-prior expectation (mantissa) */
    int *cat_k;                      /* [C][S] exponent (units of 2^256) */
    double *site_m;                  /* [S] site likelihood mantissa */
    int *site_k;                     /* [S] */
    double *site_ll;                 /* [S] or NULL */
    double *edge_site_out;           /* [E][S] or NULL */
    double *block_ll;                /* [grid of the site kernel] */
    double *block_edge;              /* [grid of the outside kernel][E] */
    int *error_flag;
};

__device__ __forceinline__ void f4_matvec(const double *__restrict__ M, const double v[4], double out[4])
{
    const double2 *M2 = reinterpret_cast<const double2 *>(M);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double2 a = M2[2 * i], b = M2[2 * i + 1];
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
    double2 a0 = M2[0], b0 = M2[1];
    out[0] = a0.x * v[0]; out[1] = a0.y * v[0]; out[2] = b0.x * v[0]; out[3] = b0.y * v[0];
#pragma unroll
    for (int i = 1; i < 4; i++) {
        double2 a = M2[2 * i], b = M2[2 * i + 1];
        out[0] = fma(a.x, v[i], out[0]);
        out[1] = fma(a.y, v[i], out[1]);
        out[2] = fma(b.x, v[i], out[2]);
        out[3] = fma(b.y, v[i], out[3]);
    }
}

__device__ __forceinline__ void f4_ld4(const double *__restrict__ p, double out[4])
{
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    double2 a = p2[0], b = p2[1];
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}

/* number of 2^256 up-scalings applied so that max(v) >= 2^-256 (v >= 0) */
__device__ __forceinline__ int f4_rescale_up(double v[4])
{
    /* non-negative doubles order like their high words */
    int h = max(max(__double2hiint(v[0]), __double2hiint(v[1])), max(__double2hiint(v[2]), __double2hiint(v[3])));
    int s = 0;
    if (h < 0x2FF00000) {          /* high word of 2^-256 */
        double m = fmax(fmax(v[0], v[1]), fmax(v[2], v[3]));
        while (m > 0.0 && m < PLF_TWO_M256 && s < 3) {
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] *= PLF_TWO_P256;
            m *= PLF_TWO_P256;
            s++;
        }
    }
    return s;
}

__device__ __forceinline__ double f4_warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

/* sums of two values over the warp in one butterfly: returns sum(a) in lanes < 16, sum(b) in lanes >= 16 */
__device__ __forceinline__ double f4_warp_sum2(double a, double b, int lane)
{
    const bool hi = lane >= 16;
    double keep = hi ? b : a, send = hi ? a : b;
    keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    return keep;
}

__host__ __device__ inline size_t f4_align16(size_t x) { return (x + 15) & ~(size_t)15; }

__device__ __forceinline__ void f4_prefetch(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

/* shared-memory carve-up common to the inside and outside kernels (mirrored by f4_smem_bytes on the host) */
struct F4Smem {
    F4Op *ops; F4Child *chs;
    unsigned char *tile, *dconst;
    double *defs_s;
    const double *Pint, *TP, *Fint, *TF;
    double *accE; double *stack; int *stackf;
};

template <int BD, bool OUTSIDE, bool STAGED>
__device__ __forceinline__ void f4_setup_smem(const F4Args &a, unsigned char *smem, F4Smem &s, int spc)
{
    const int tid = threadIdx.x;
    size_t off = 0;
    s.ops = reinterpret_cast<F4Op *>(smem + off); off = f4_align16(off + sizeof(F4Op) * a.nops);
    s.chs = reinterpret_cast<F4Child *>(smem + off); off = f4_align16(off + sizeof(F4Child) * a.nchildren);
    s.tile = smem + off; off = f4_align16(off + (size_t)a.ncode_rows * spc);
    s.dconst = smem + off; off = f4_align16(off + a.K);
    s.defs_s = reinterpret_cast<double *>(smem + off); off = f4_align16(off + sizeof(double) * 4 * a.K);
    s.accE = reinterpret_cast<double *>(smem + off); off = f4_align16(off + (OUTSIDE ? sizeof(double) * (BD / 32) * a.E : 0));
    s.stack = reinterpret_cast<double *>(smem + off); off = f4_align16(off + (OUTSIDE ? 0 : sizeof(double) * 4 * BD * a.stack_depth));
    s.stackf = reinterpret_cast<int *>(smem + off); off = f4_align16(off + (OUTSIDE ? 0 : sizeof(int) * BD * a.stack_depth));
    const size_t nP = (size_t)a.C * a.Ei * 16, nT = (size_t)a.C * a.Et * a.K * 4;
    s.Pint = a.Pint; s.TP = a.TP; s.Fint = a.Fint; s.TF = a.TF;
    if (STAGED) {
        double *sP = reinterpret_cast<double *>(smem + off); off += sizeof(double) * nP;
        double *sT = reinterpret_cast<double *>(smem + off); off += sizeof(double) * nT;
        for (size_t i = tid; i < nP; i += BD) sP[i] = a.Pint[i];
        for (size_t i = tid; i < nT; i += BD) sT[i] = a.TP[i];
        s.Pint = sP; s.TP = sT;
        if (OUTSIDE) {
            double *sF = reinterpret_cast<double *>(smem + off); off += sizeof(double) * nP;
            double *sTF = reinterpret_cast<double *>(smem + off); off += sizeof(double) * nT;
            for (size_t i = tid; i < nP; i += BD) sF[i] = a.Fint[i];
            for (size_t i = tid; i < nT; i += BD) sTF[i] = a.TF[i];
            s.Fint = sF; s.TF = sTF;
        }
    }
    for (int i = tid; i < a.nops; i += BD) s.ops[i] = a.ops[i];
    for (int i = tid; i < a.nchildren; i += BD) s.chs[i] = a.children[i];
    for (int i = tid; i < a.K; i += BD) s.dconst[i] = a.def_const[i];
    for (int i = tid; i < 4 * a.K; i += BD) s.defs_s[i] = a.defs[i];
    if (OUTSIDE) for (int i = tid; i < (BD / 32) * a.E; i += BD) s.accE[i] = 0.0;
}

/* cooperative, coalesced load of the tile's character codes: tile[row][site in tile] */
template <int BD>
__device__ __forceinline__ void f4_load_tile(const F4Args &a, unsigned char *tile, int spc, int64_t site0)
{
    const int total = a.ncode_rows * spc;
    for (int idx = threadIdx.x; idx < total; idx += BD) {
        const int row = idx / spc, sl = idx - row * spc;
        int64_t site = site0 + sl;
        if (site >= a.S) site = a.S - 1;
        tile[idx] = a.codes[(size_t)a.code_row_node[row] * a.S_total + a.s0 + site];
    }
}

/*
 * The fused kernel: per tile of sites, inside pass -> category mix -> (EDGE) outside pass.
 * Thread -> (site, category): lane = (site within warp) * C + category; 32 / C sites per warp.
 * EDGE: every internal partial goes to the CTA's private slab ([slot][category][site in tile]
 * double4, reused tile after tile, so it tends to stay in L2); otherwise only a small
 * shared-memory stack is used.
 */
template <int BD, bool EDGE, bool STAGED>
__global__ void __launch_bounds__(BD) fused4_kernel(F4Args a)
{
    extern __shared__ __align__(16) unsigned char f4_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.C;
    const int spw = 32 / C;                 /* sites per warp */
    const int spc = spw * (BD / 32);        /* sites per CTA tile */
    const int sl_w = lane / C, c = lane - sl_w * C;
    const bool lane_on = sl_w < spw;
    const int cc = lane_on ? c : 0;
    const int sl = warp * spw + (lane_on ? sl_w : 0);
    const int base_lane = (lane_on ? sl_w : 0) * C;
    F4Smem s;
    f4_setup_smem<BD, EDGE, STAGED>(a, f4_smem, s, spc);
    __syncthreads();
    const double *Pc = s.Pint + (size_t)cc * a.Ei * 16;
    const double *TPc = s.TP + (size_t)cc * a.Et * a.K * 4;
    const double *Fc = EDGE ? s.Fint + (size_t)cc * a.Ei * 16 : nullptr;
    const double *TFc = EDGE ? s.TF + (size_t)cc * a.Et * a.K * 4 : nullptr;
    double *accW = s.accE + warp * a.E;
    const double prior = a.cat_prior[cc];
    const int64_t ntiles = (a.S + spc - 1) / spc;
    /* this CTA's slab */
    const size_t slot_stride = (size_t)C * spc;
    double4 *slab = EDGE ? a.slab + (size_t)blockIdx.x * a.nslots * slot_stride + (size_t)cc * spc + sl : nullptr;
    unsigned char *slabF = EDGE ? a.slabF + (size_t)blockIdx.x * a.nslots * slot_stride + (size_t)cc * spc + sl : nullptr;
    double ll_acc = 0.0;

    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t site0 = t * spc;
        const int64_t site_raw = site0 + sl;
        const bool valid = lane_on && site_raw < a.S;
        const int64_t site = site_raw < a.S ? site_raw : a.S - 1;
        __syncthreads();                    /* previous tile fully consumed */
        f4_load_tile<BD>(a, s.tile, spc, site0);
        __syncthreads();

        /* ---------------- inside pass ---------------- */
        double cur[4] = {1.0, 1.0, 1.0, 1.0};
        int curf = 1, sp = 0, ktot = 0;
#pragma unroll 1
        for (int o = 0; o < a.nops; o++) {
            const F4Op op = s.ops[o];
            if (!EDGE && op.spill_before) {
#pragma unroll
                for (int i = 0; i < 4; i++) s.stack[(sp * 4 + i) * BD + tid] = cur[i];
                s.stackf[sp * BD + tid] = curf;
                sp++;
            }
            double acc[4];
            int cst = 1;
            bool first = true;
            if (op.code_row >= 0) {
                const int code = s.tile[op.code_row * spc + sl];
                f4_ld4(s.defs_s + code * 4, acc);
                cst = s.dconst[code];
                first = false;
            }
#pragma unroll 1
            for (int j = 0; j < op.nchild; j++) {
                const F4Child ch = s.chs[op.first_child + j];
                double em[4];
                int bc;
                if (ch.kind == F4_KIND_TIP) {
                    const int code = s.tile[ch.code_row * spc + sl];
                    bc = s.dconst[code];
                    f4_ld4(TPc + (ch.mat * a.K + code) * 4, em);
                } else {
                    double v[4];
                    if (ch.kind == F4_KIND_CUR) {
                        bc = curf;
#pragma unroll
                        for (int i = 0; i < 4; i++) v[i] = cur[i];
                    } else if (EDGE) {
                        const size_t so = (size_t)ch.slot * slot_stride;
                        double4 l4 = slab[so];
                        bc = (slabF[so] >> 6) & 1;
                        v[0] = l4.x; v[1] = l4.y; v[2] = l4.z; v[3] = l4.w;
                    } else {
                        sp--;
                        bc = s.stackf[sp * BD + tid];
#pragma unroll
                        for (int i = 0; i < 4; i++) v[i] = s.stack[(sp * 4 + i) * BD + tid];
                    }
                    f4_matvec(Pc + ch.mat * 16, v, em);
                    if (bc) {          /* constant column maps to itself (arb_mat_extras.c:84-91) */
#pragma unroll
                        for (int i = 0; i < 4; i++) em[i] = v[i];
                    }
                }
                if (first) {
#pragma unroll
                    for (int i = 0; i < 4; i++) acc[i] = em[i];
                    first = false;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) acc[i] *= em[i];
                }
                cst &= bc;
            }
            /* one rescale check per node is enough for out-degree <= 3 */
            const int sloc = f4_rescale_up(acc);
            ktot -= sloc;
#pragma unroll
            for (int i = 0; i < 4; i++) cur[i] = acc[i];
            curf = cst;
            if (EDGE) {
                const size_t so = (size_t)op.slot * slot_stride;
                slab[so] = make_double4(acc[0], acc[1], acc[2], acc[3]);
                slabF[so] = (unsigned char)(sloc | (cst << 6));
            }
        }
        /* ---------------- root prior expectation (model.c:282-350), category mix (arbplfll.c:149-169) ---------------- */
        double lh;
        if (a.root_mode == PLF_ROOT_NONE) lh = (cur[0] + cur[1]) + (cur[2] + cur[3]);
        else if (a.root_mode == PLF_ROOT_UNIFORM) lh = curf ? cur[0] : ((cur[0] + cur[1]) + (cur[2] + cur[3])) * 0.25;
        else if (a.root_mode == PLF_ROOT_EQUILIBRIUM && curf) lh = cur[0];
        else {
            lh = a.root_vec[0] * cur[0];
            lh = fma(a.root_vec[1], cur[1], lh);
            lh = fma(a.root_vec[2], cur[2], lh);
            lh = fma(a.root_vec[3], cur[3], lh);
        }
        const double vmine = prior * lh;
        const int kmine = (vmine > 0.0) ? ktot : INT_MIN;
        int k0 = INT_MIN;
        for (int k = 0; k < C; k++) k0 = max(k0, __shfl_sync(0xffffffffu, kmine, (base_lane + k) & 31));
        double site_m = 0.0;
        for (int k = 0; k < C; k++) {
            const double vk = __shfl_sync(0xffffffffu, vmine, (base_lane + k) & 31);
            const int kk = __shfl_sync(0xffffffffu, kmine, (base_lane + k) & 31);
            if (kk != INT_MIN) site_m += (kk == k0) ? vk : scalbn(vk, PLF_SCALE_BITS * (kk - k0));
        }
        const bool have = k0 != INT_MIN;
        const double w = valid ? (a.site_w ? a.site_w[a.s0 + site] : 1.0) : 0.0;
        if (c == 0) {
            const double c_hi = 177.445678223346, c_lo = 5.936759843446527e-15;   /* 256 ln 2 */
            double ll = log(site_m);
            if (have && k0 != 0) { ll = fma((double)k0, c_hi, ll); ll = fma((double)k0, c_lo, ll); }
            if (valid) {
                if (a.site_ll) a.site_ll[a.s0 + site] = ll;
                if (w != 0.0) {
                    if (!have) atomicOr(a.error_flag, 1);
                    else ll_acc = fma(w, ll, ll_acc);
                }
            }
        }
        if (!EDGE) continue;

        /* ---------------- outside pass ---------------- */
        /* fn_root = root prior vector * prior_c * w / site_L, scaled so that fn .* L is O(1) */
        double curF[4];
        {
            const double wq = a.edge_site_out ? (valid ? 1.0 : 0.0) : w;
            double sc0 = 0.0;
            if (wq != 0.0 && have && kmine != INT_MIN) sc0 = scalbn(prior * wq / site_m, PLF_SCALE_BITS * (kmine - k0));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double r = 1.0;
                if (a.root_mode == PLF_ROOT_UNIFORM) r = 0.25;
                else if (a.root_mode == PLF_ROOT_EQUILIBRIUM || a.root_mode == PLF_ROOT_CUSTOM) r = a.root_vec[i];
                curF[i] = r * sc0;
            }
        }
#pragma unroll 1
        for (int o = a.nops - 1; o >= 0; o--) {
            const F4Op op = s.ops[o];
            const bool from_slot = (o != a.nops - 1) && s.ops[o + 1].spill_before;
            const size_t so_a = (size_t)op.slot * slot_stride;
            /* 1. loads: children first (longest latency), then own words */
            int kinds[F4_MAXD], mats[F4_MAXD], edges[F4_MAXD];
            size_t cso[F4_MAXD];
            double lv[F4_MAXD][4];
            int bcs[F4_MAXD];
#pragma unroll
            for (int j = 0; j < F4_MAXD; j++) {
                kinds[j] = -1; mats[j] = 0; edges[j] = 0; cso[j] = 0; bcs[j] = 0;
                if (j < op.nchild) {
                    const F4Child ch = s.chs[op.first_child + j];
                    kinds[j] = ch.kind; mats[j] = ch.mat; edges[j] = ch.edge;
                    if (ch.kind == F4_KIND_TIP) {
                        bcs[j] = s.tile[ch.code_row * spc + sl];        /* the character code */
                    } else {
                        cso[j] = (size_t)ch.slot * slot_stride;
                        double4 l4 = slab[cso[j]];
                        lv[j][0] = l4.x; lv[j][1] = l4.y; lv[j][2] = l4.z; lv[j][3] = l4.w;
                        bcs[j] = (slabF[cso[j]] >> 6) & 1;
                    }
                }
            }
            double fa[4];
            if (from_slot) {
                double4 f4v = slab[so_a];
                fa[0] = f4v.x; fa[1] = f4v.y; fa[2] = f4v.z; fa[3] = f4v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) fa[i] = curF[i];
            }
            const int sa = slabF[so_a] & 63;
            if (op.code_row >= 0) {
                double b[4];
                f4_ld4(s.defs_s + s.tile[op.code_row * spc + sl] * 4, b);
#pragma unroll
                for (int i = 0; i < 4; i++) fa[i] *= b[i];
            }
            if (sa) {
                /* undo the node's own rescaling: fn_a * 2^(256 s_a) keeps fn .* L at O(1) */
                const double sc = __hiloint2double((1023 + PLF_SCALE_BITS * sa) << 20, 0);
#pragma unroll
                for (int i = 0; i < 4; i++) fa[i] *= sc;
            }
            /* 2. edge vectors em_j = P_j L_j and y_j = F_j L_j */
            double em[F4_MAXD][4], y[F4_MAXD][4];
#pragma unroll
            for (int j = 0; j < F4_MAXD; j++) {
                if (kinds[j] == F4_KIND_TIP) {
                    f4_ld4(TPc + (mats[j] * a.K + bcs[j]) * 4, em[j]);
                    f4_ld4(TFc + (mats[j] * a.K + bcs[j]) * 4, y[j]);
                } else if (kinds[j] >= 0) {
                    f4_matvec(Pc + mats[j] * 16, lv[j], em[j]);
                    f4_matvec(Fc + mats[j] * 16, lv[j], y[j]);
                    if (bcs[j]) {
#pragma unroll
                        for (int i = 0; i < 4; i++) { em[j][i] = lv[j][i]; if (a.f_zero_rowsum) y[j][i] = 0.0; }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) { em[j][i] = 1.0; y[j][i] = 0.0; }
                }
            }
            /* 3. per child: fe_j, x_j, fn_j */
            double xe[F4_MAXD];
#pragma unroll
            for (int j = 0; j < F4_MAXD; j++) {
                xe[j] = 0.0;
                if (kinds[j] >= 0) {
                    double fe[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        double f = fa[i];
#pragma unroll
                        for (int j2 = 0; j2 < F4_MAXD; j2++) if (j2 != j) f *= em[j2][i];
                        fe[i] = f;
                    }
                    double xv = fe[0] * y[j][0];
                    xv = fma(fe[1], y[j][1], xv);
                    xv = fma(fe[2], y[j][2], xv);
                    xv = fma(fe[3], y[j][3], xv);
                    xe[j] = xv;
                    if (kinds[j] != F4_KIND_TIP) {
                        double fb[4];
                        f4_matvec_t(Pc + mats[j] * 16, fe, fb);
                        if (kinds[j] == F4_KIND_CUR) {
#pragma unroll
                            for (int i = 0; i < 4; i++) curF[i] = fb[i];
                        } else {
                            slab[cso[j]] = make_double4(fb[0], fb[1], fb[2], fb[3]);
                        }
                    }
                }
            }
            /* 4. reductions */
            if (a.edge_site_out) {
                /* per-site output: sum the categories of each site in lane order (deterministic) */
#pragma unroll
                for (int j = 0; j < F4_MAXD; j++) {
                    if (kinds[j] >= 0 && (!a.edge_mask || a.edge_mask[edges[j]])) {
                        double tot = 0.0;
                        for (int k = 0; k < C; k++) tot += __shfl_sync(0xffffffffu, xe[j], (base_lane + k) & 31);
                        if (valid && c == 0) a.edge_site_out[(size_t)edges[j] * a.S_total + a.s0 + site] = tot;
                    }
                }
            } else {
                const bool m0 = kinds[0] >= 0 && (!a.edge_mask || a.edge_mask[edges[0]]);
                const bool m1 = kinds[1] >= 0 && (!a.edge_mask || a.edge_mask[edges[1]]);
                const bool m2 = kinds[2] >= 0 && (!a.edge_mask || a.edge_mask[edges[2]]);
                if (m0 || m1) {
                    const double r = f4_warp_sum2(m0 ? xe[0] : 0.0, m1 ? xe[1] : 0.0, lane);
                    if (lane == 0 && m0) accW[edges[0]] += r;
                    if (lane == 16 && m1) accW[edges[1]] += r;
                }
                if (m2) {
                    const double r = f4_warp_sum(xe[2]);
                    if (lane == 0) accW[edges[2]] += r;
                }
            }
        }
    }

    /* ---------------- CTA-level reductions ---------------- */
    {
        __shared__ double red[32];
        double xx = f4_warp_sum(ll_acc);
        if (lane == 0) red[warp] = xx;
        __syncthreads();
        if (tid == 0) {
            double tsum = 0.0;
            for (int i = 0; i < BD / 32; i++) tsum += red[i];
            a.block_ll[blockIdx.x] = tsum;
        }
        if (EDGE && !a.edge_site_out) {
            for (int e = tid; e < a.E; e += BD) {
                double tsum = 0.0;
                for (int wv = 0; wv < BD / 32; wv++) tsum += s.accE[wv * a.E + e];
                a.block_edge[(size_t)blockIdx.x * a.E + e] = tsum;
            }
        }
    }
}

/* second stage: out[j] = sum over rows of part[row][j], fixed order (Kahan) */
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

/* gather the matrices of internal-child edges: out[c][ie] = M[c][edge_of[ie]] (16 doubles each) */
__global__ void compact_matrices_kernel(const double *M, const int *edge_of, int C, int E, int Ei, double *out)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int total = C * Ei * 16;
    if (idx >= total) return;
    int k = idx & 15, r = idx >> 4;
    int ie = r % Ei, c = r / Ei;
    out[idx] = M[((size_t)c * E + edge_of[ie]) * 16 + k];
}
