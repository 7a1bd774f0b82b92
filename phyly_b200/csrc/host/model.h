/*
 * Host-side input model: the JSON "model_and_data" object turned into plain C
 * arrays.  Restates validate_model_and_data (parsemodel.c:786-913) with the
 * same acceptance rules; all arithmetic of the reference's cross_site_ws that
 * is site independent and tiny (mixture rates, equilibrium, rate divisor,
 * scaled rate matrix) is done here in extended / double-double precision and
 * handed to the device seam (include/plf.h).
 */
#ifndef PLF_MODEL_H
#define PLF_MODEL_H

#include <stdint.h>
#include "json.h"

enum { MIX_NONE = 0, MIX_UNIFORM, MIX_CUSTOM, MIX_GAMMA, MIX_GAMMA_MEDIAN };

typedef struct {
    /* tree (csr_graph.c, model.c:23-45, util.c:369-389) */
    int N, E, root;
    int *indptr, *indices;
    int *order;            /* user edge -> csr idx */
    int *idx_to_user;      /* csr idx -> user edge */
    int *preorder;
    /* rates */
    int n;
    double *rate_matrix;   /* raw user matrix n*n */
    double *edge_rate_user;/* [E] user order */
    int use_eq_divisor;
    double rate_divisor;
    int root_mode;         /* PLF_ROOT_* */
    double *root_custom;
    int mix_mode, mix_n;
    double *mix_rates, *mix_prior;
    double gamma_shape, invariable_prior;
    int gamma_categories;
    /* data as codes into a table of distinct rows */
    int64_t S;
    int K;
    double *defs;          /* [K][n] */
    void *codes;           /* [S][N], uint8 if K <= 256 else int32 */
    int code_bytes;
    struct jv_codes *codes_shared;   /* non-NULL: `codes` belongs to the reader's data cache (json.h); one reference held */
} plf_model;

void plf_model_init(plf_model *m);
void plf_model_clear(plf_model *m);
int plf_model_parse(plf_model *m, const jv *md);
int plf_model_category_count(const plf_model *m);

/* site independent derived quantities (cross_site_ws.c:199-242) */
typedef struct {
    int n, C, E;
    double *cat_rates, *cat_prior;   /* [C] */
    double expect;
    double *equilibrium;             /* [n] or NULL */
    double *q_hi, *q_lo;             /* scaled rate matrix, double-double */
    double *edge_rates_csr;          /* [E] */
    double *root_vec;                /* [n] or NULL */
} plf_derived;

int plf_derive(plf_derived *d, const plf_model *m);
void plf_derived_clear(plf_derived *d);

/* gamma discretisation (gamma_discretization.c:298-370); exposed for tests */
int plf_gamma_rates(double *rates, int ncat, double shape, int median);

#endif
