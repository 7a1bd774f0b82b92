/*
 * plf.h -- device seam of the B200 phylogenetic likelihood engine (C ABI).
 *
 * This is the boundary "host C calls CUDA through a thin C-ABI layer".  It
 * replaces, for the hot path, the reference's in-process workspaces:
 *
 *   plf_set_tree      <- csr_graph_t + navigation_t built by _validate_edges
 *                        (parsemodel.c:210-368, csr_graph.h:17-51, model.c:23-45)
 *   plf_set_model     <- cross_site_ws_update_with_edge_rates
 *                        (cross_site_ws.c:199-242; transition matrices :150-168)
 *   plf_set_data      <- pmat_t / pmat_update_base_node_vectors
 *                        (model.h:41-48, model.c:181-199)
 *   plf_ll            <- _nd_accum_update of arbplfll.c:110-177
 *                        (evaluate_site_lhood.c:6-63 inside the site x category loop)
 *   plf_deriv         <- _nd_accum_update of arbplfderiv.c:210-371
 *   plf_marginal      <- _nd_accum_update of arbplfmarginal.c:111-264
 *   plf_edge_expect   <- _update_site of arbplfdwell.c:201-312 / arbplftrans.c:222-346
 *                        with the Frechet matrices of util.c:500-548
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from plf_last_error().  Nothing aborts the host process.
 *   - all pointers are HOST pointers owned by the caller unless the name ends
 *     in _dev.  Arrays are C-contiguous.
 *   - edges are always in CSR order ("csr idx", csr_graph.c:28-46): position
 *     of the edge in indices[]; node labels are the user's.
 *   - there is no CPU implementation behind this interface: if no CUDA device
 *     is usable plf_create fails.
 *   - a handle is not thread safe; use one handle per host thread / per GPU.
 */
/*
 * Threading and sharing: an engine handle is to be used from one host thread at a time; different
 * handles (also on the same device) may be used from different threads -- the one device-wide resource,
 * the constant bank that holds the matrices of the small-tree kernels, is serialised inside the library.
 * Queries are synchronous: when a plf_* query returns, its host outputs are final.
 */
#ifndef PLF_H
#define PLF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plf_engine plf_engine;

/* root prior modes, model.h (enum root_prior_mode) */
enum { PLF_ROOT_NONE = 0, PLF_ROOT_UNIFORM = 1, PLF_ROOT_EQUILIBRIUM = 2, PLF_ROOT_CUSTOM = 3 };
/* edge expectation kinds */
enum { PLF_KIND_DWELL = 0, PLF_KIND_TRANS = 1 };
/* kernel path selection (PLF_PATH_AUTO picks the fused 4-state kernel when it applies) */
enum { PLF_PATH_AUTO = 0, PLF_PATH_GENERIC = 1, PLF_PATH_FUSED4 = 2 };

int  plf_create(plf_engine **out, int device);
void plf_destroy(plf_engine *e);
const char *plf_last_error(const plf_engine *e);

/* Force a kernel path (tests / benchmarks); default PLF_PATH_AUTO. */
int plf_set_path(plf_engine *e, int path);

/*
 * Tree in the reference's CSR form.  preorder is the BFS level order from the
 * root (csr_graph_get_tree_topo_sort); preorder[0] is the root.
 */
int plf_set_tree(plf_engine *e, int node_count,
                 const int *indptr /*[N+1]*/, const int *indices /*[E]*/,
                 const int *preorder /*[N]*/);

/*
 * Model.  Q is the *scaled* rate matrix (divisor applied, diagonal = -row sum),
 * row major, given as a double-double pair (q_lo may be NULL).  Transition
 * matrices exp(cat_rates[c]*edge_rates[e]*Q) are computed on the device.
 * root_vec: custom prior (PLF_ROOT_CUSTOM) or equilibrium (PLF_ROOT_EQUILIBRIUM);
 * ignored otherwise.
 */
int plf_set_model(plf_engine *e, int state_count, int category_count,
                  const double *q_hi /*[n*n]*/, const double *q_lo /*[n*n] or NULL*/,
                  const double *edge_rates /*[E] csr order*/,
                  const double *cat_rates /*[C]*/, const double *cat_prior /*[C]*/,
                  int root_mode, const double *root_vec /*[n] or NULL*/);

/* Creates the CUDA context of a device ahead of the first plf_create (no engine, no message; 0 or -1): a process that
 * answers one document (the arbplf-* executables, runjson.c:117-157) calls it from a second thread while it reads and
 * parses its input. */
int plf_warmup(int device);

/* Change only the edge rate coefficients (csr order); recomputes matrices lazily. */
int plf_set_edge_rates(plf_engine *e, const double *edge_rates /*[E]*/);

/*
 * Data as character codes: codes[site][node] indexes rows of defs[K][n]
 * (parsemodel.c:514-628).  code_bytes is 1 (uint8), 4 (int32) or
 * PLF_CODES_PACKED4: two codes per byte (def_count <= 16, the nucleotide case),
 * rows of (N + 1) / 2 bytes, node 2j in the low and node 2j + 1 in the high
 * nibble of byte j -- half the host-to-device traffic of uint8 codes.  A dense
 * probability_array is passed by first de-duplicating its rows into defs.
 * The copy to the device happens inside this call (pinned staging).
 */
#define PLF_CODES_PACKED4 0
int plf_set_data(plf_engine *e, int64_t site_count, int def_count,
                 const double *defs /*[K][n]*/, const void *codes /*[S][N]*/, int code_bytes);

/*
 * The same, without waiting for the copy: the codes (and, if given, the site
 * weights) travel in chunks on a second stream and the next query starts on
 * the first chunk while the later ones are still in flight (4-state fused
 * path; other paths simply wait).  `codes` and `site_weights` must stay valid
 * and unchanged until the next query or plf_synchronize returns; use pinned
 * host memory, otherwise the copy is staged and nothing overlaps.
 */
int plf_set_data_async(plf_engine *e, int64_t site_count, int def_count,
                       const double *defs /*[K][n]*/, const void *codes /*[S][N]*/, int code_bytes,
                       const double *site_weights /*[S] or NULL*/);

/*
 * Site-pattern compression (the reference leaves it to its input generators, examples/BEAST.GTRG/mknuc.py:57-66):
 * identical rows of codes[S][N] are merged.  Patterns come out in order of first occurrence: codes_out[P][N] (room for
 * S rows), counts_out[P] their multiplicities (the site weights of the compressed alignment), site_to_pattern[S] the
 * pattern of every site (scatter-back map when the site axis is kept).  Outputs other than n_patterns may be NULL.
 * Hash + full row comparison on the device; integer results, identical to the oracle's.
 */
int plf_compress_patterns(plf_engine *e, int64_t site_count, int node_count, const void *codes, int code_bytes,
                          int64_t *n_patterns, void *codes_out, int64_t *counts_out, int64_t *site_to_pattern);

/* Per-site weights used by the *_sum outputs (NULL = all ones).  reduction.c:24-118. */
int plf_set_site_weights(plf_engine *e, const double *w /*[S] or NULL*/);

/* log-likelihood.  site_ll[S] and/or sum = sum_s w_s ll_s; either may be NULL. */
int plf_ll(plf_engine *e, double *site_ll, double *sum);

/*
 * d log L / d edge_rate for the edges with edge_mask[idx] != 0 (NULL = all).
 * site_deriv is [S][E] (csr edge order), sum_deriv[E] = sum_s w_s d[s][e].
 * Any output may be NULL.  Unrequested edges are returned as 0.
 */
int plf_deriv(plf_engine *e, const unsigned char *edge_mask,
              double *site_ll, double *sum_ll, double *site_deriv, double *sum_deriv);

/*
 * Certified mode: enclosures site_lo[s] <= log L_s <= site_hi[s] computed with directed rounding (interval arithmetic
 * on the pruning recursion and on the matrix exponentials; csrc/certified.cuh), in place of the reference's Arb balls
 * (util.c:12-50, arbplfll.c:206-224).  Rigorous for inputs within the stated relative uncertainties: delta_rate on
 * rate_c * t_e and on an equilibrium root prior (the host's gamma quantiles / linear solve are accurate to 1e-14, not
 * certified; 4e-15 is the documented default), delta_q on the scaled rate matrix and the derived category priors
 * (2^-60).  sum_lo / sum_hi enclose sum_s w_s log L_s.  Any output may be NULL.
 */
int plf_ll_certified(plf_engine *e, double delta_rate, double delta_q, double *site_lo, double *site_hi,
                     double *sum_lo, double *sum_hi);

/*
 * Second order (arbplfhess.c:502-829): sum_ll = sum_s w_s log L_s, sum_deriv[E] its gradient and sum_hess[E][E] its
 * Hessian with respect to the edge rate coefficients (csr edge order, full symmetric matrix).  sum_ll and sum_deriv
 * may be NULL.  A rate category with a non-zero rate and zero likelihood at a weighted site is an error
 * ("infeasible", arbplfhess.c:640-660).
 */
int plf_hess(plf_engine *e, double *sum_ll, double *sum_deriv /*[E]*/, double *sum_hess /*[E][E]*/);

/* posterior marginals: site_marg [S][N][n], sum_marg [N][n] = sum_s w_s marg. */
int plf_marginal(plf_engine *e, double *site_marg, double *sum_marg);

/*
 * Posterior edge expectations from Frechet matrices with direction L (n x n,
 * double-double pair, l_lo may be NULL):
 *   PLF_KIND_DWELL: sum_c prior_c fe^T F L_b / L_site
 *   PLF_KIND_TRANS: sum_c prior_c rate_c t_e fe^T F L_b / L_site
 */
int plf_edge_expect(plf_engine *e, int kind,
                    const double *l_hi /*[n*n]*/, const double *l_lo,
                    const unsigned char *edge_mask,
                    double *site_out /*[S][E]*/, double *sum_out /*[E]*/);

/* Copies of device-computed matrices, for tests: P[C][E][n][n] (row major). */
int plf_get_transition_matrices(plf_engine *e, double *p_out);
/* rate_c * Q * P, the matrices used by plf_deriv */
int plf_get_derivative_matrices(plf_engine *e, double *d_out);
/* Frechet matrices for a direction L (unscaled: the top-right block of util.c:500-548) */
int plf_get_frechet_matrices(plf_engine *e, const double *l_hi, const double *l_lo, double *f_out);

/*
 * Timing / accounting of the most recent query (CUDA events on the engine's
 * stream): milliseconds for the matrix kernels (expm etc.) and the per-site
 * kernels, and the number of kernel launches since the last reset.
 */
int plf_last_timing(plf_engine *e, float *ms_matrices, float *ms_sites);
/* duration of the dominant per-site kernel alone (the fused 4-state kernel), 0 if it did not run */
int plf_last_kernel_ms(plf_engine *e, float *ms_kernel);
int64_t plf_launch_count(plf_engine *e, int reset);
/* which instantiation ran as the dominant per-site kernel of the most recent query, spelled as the profiler spells it
 * (e.g. "fused4_kernel<4,1,512,2,1,1>": categories, mode, block size, staging, packed codes, constant-memory matrices;
 * the fused kernel's configuration is picked by timing, so this is the only way to know) */
const char *plf_last_kernel_name(const plf_engine *e);

/*
 * Multi-GPU: one engine per rank, sites sharded by the caller.  After
 * plf_comm_init the *_sum outputs are all-reduced (ncclAllReduce, sum, fp64) in
 * stream, right after the block reduction.  id is an ncclUniqueId (128 bytes).
 */
int plf_comm_unique_id(char id[128]);
int plf_comm_init(plf_engine *e, int nranks, int rank, const char id[128]);
/* paused != 0: the *_sum outputs stay local (this rank's sites only) until resumed; for cross-checks of the collective */
int plf_comm_pause(plf_engine *e, int paused);

/* The CUDA stream the engine launches on (cudaStream_t as void*), for event timing. */
void *plf_stream(plf_engine *e);
int plf_synchronize(plf_engine *e);

#ifdef __cplusplus
}
#endif
#endif
