/*
 * Axis reductions: selection / aggregation semantics of parsereduction.c and
 * reduction.c, and the table emitter of ndaccum.c:300-437.
 */
#ifndef PLF_REDUCE_H
#define PLF_REDUCE_H

#include "json.h"

enum { AGG_NONE = 0, AGG_AVG, AGG_SUM, AGG_WEIGHTED_SUM, AGG_ONLY };

typedef struct {
    int *selection;
    double *weights;
    int agg_mode;
    int selection_len;
    int *first_idx, *second_idx;    /* pair reductions only */
} reduction;

void reduction_init(reduction *r);
void reduction_clear(reduction *r);
/* validate_column_reduction, parsereduction.c:161-195 */
int reduction_parse(reduction *r, int k, const char *name, const jv *root);
/* validate_column_pair_reduction, parsereduction.c:290-392 */
int reduction_parse_pairs(reduction *r, int k, const char *name, const jv *root);

typedef struct {
    const char *name;
    int n;                       /* total number of indices on the axis */
    const reduction *r;
    int aggregated;
    double *w;                   /* [n]: agg weight / divisor (reduction.c:24-118), NULL if not aggregated */
    unsigned char *requested;    /* [n] */
    int ncomp;                   /* compound axis (trans): 2, else 0 */
    const char *comp_name[2];
    const int *comp_idx[2];
} axis;

int axis_init(axis *a, const char *name, int n, const reduction *r);
void axis_clear(axis *a);

/*
 * Emit {"columns": [...], "data": [[...], ...]} for a dense value array laid
 * out over the non-aggregated axes (extent n each, aggregated axes have extent
 * 1), in selection order, row major over the axes (ndaccum.c:300-437).
 */
char *table_to_json(const axis *axes, int ndim, const double *values);

#endif
