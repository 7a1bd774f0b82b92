/*
 * Double-double ("dd", ~106-bit significand) arithmetic shared by the host C
 * code and the CUDA kernels.
 *
 * Why it exists: the reference computes transition matrices with Arb ball
 * arithmetic at whatever precision makes the *output* correct to 53 bits
 * (cross_site_ws.c:150-168, arbplfll.c:206-224).  The engine replaces the
 * growing precision with fixed fp64 for the per-site work, but the tiny
 * site-independent matrices (P = exp(r t Q), Q.P, Frechet blocks) are built in
 * dd so that entries that are small relative to the matrix norm (short
 * branches) and differences from the stationary limit (long branches, see
 * examples/JC.long.branch) keep full *relative* accuracy after rounding to fp64.
 */
#ifndef PLF_DD_H
#define PLF_DD_H

#include <math.h>

#ifdef __CUDACC__
#define DD_FN __host__ __device__ __forceinline__
#else
#define DD_FN static inline
#endif

typedef struct { double hi, lo; } dd_t;

/* exact (unfused, uncontracted) primitives */
#if defined(__CUDA_ARCH__)
#define DD_ADD(a, b) __dadd_rn((a), (b))
#define DD_SUB(a, b) __dsub_rn((a), (b))
#define DD_MUL(a, b) __dmul_rn((a), (b))
#define DD_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
/* host: compiled with -ffp-contract=off (see Makefile) */
#define DD_ADD(a, b) ((a) + (b))
#define DD_SUB(a, b) ((a) - (b))
#define DD_MUL(a, b) ((a) * (b))
#define DD_FMA(a, b, c) fma((a), (b), (c))
#endif

DD_FN dd_t dd_make(double hi, double lo) { dd_t r; r.hi = hi; r.lo = lo; return r; }
DD_FN dd_t dd_from_d(double a) { return dd_make(a, 0.0); }
DD_FN double dd_to_d(dd_t a) { return DD_ADD(a.hi, a.lo); }

DD_FN dd_t dd_quick_two_sum(double a, double b)
{
    double s = DD_ADD(a, b);
    double e = DD_SUB(b, DD_SUB(s, a));
    return dd_make(s, e);
}

DD_FN dd_t dd_two_sum(double a, double b)
{
    double s = DD_ADD(a, b);
    double bb = DD_SUB(s, a);
    double e = DD_ADD(DD_SUB(a, DD_SUB(s, bb)), DD_SUB(b, bb));
    return dd_make(s, e);
}

DD_FN dd_t dd_two_prod(double a, double b)
{
    double p = DD_MUL(a, b);
    double e = DD_FMA(a, b, -p);
    return dd_make(p, e);
}

DD_FN dd_t dd_neg(dd_t a) { return dd_make(-a.hi, -a.lo); }

DD_FN dd_t dd_add(dd_t a, dd_t b)
{
    dd_t s = dd_two_sum(a.hi, b.hi);
    dd_t t = dd_two_sum(a.lo, b.lo);
    s.lo = DD_ADD(s.lo, t.hi);
    s = dd_quick_two_sum(s.hi, s.lo);
    s.lo = DD_ADD(s.lo, t.lo);
    return dd_quick_two_sum(s.hi, s.lo);
}

DD_FN dd_t dd_sub(dd_t a, dd_t b) { return dd_add(a, dd_neg(b)); }

DD_FN dd_t dd_add_d(dd_t a, double b)
{
    dd_t s = dd_two_sum(a.hi, b);
    s.lo = DD_ADD(s.lo, a.lo);
    return dd_quick_two_sum(s.hi, s.lo);
}

DD_FN dd_t dd_mul(dd_t a, dd_t b)
{
    dd_t p = dd_two_prod(a.hi, b.hi);
    p.lo = DD_ADD(p.lo, DD_ADD(DD_MUL(a.hi, b.lo), DD_MUL(a.lo, b.hi)));
    return dd_quick_two_sum(p.hi, p.lo);
}

DD_FN dd_t dd_mul_d(dd_t a, double b)
{
    dd_t p = dd_two_prod(a.hi, b);
    p.lo = DD_ADD(p.lo, DD_MUL(a.lo, b));
    return dd_quick_two_sum(p.hi, p.lo);
}

/* a*b + c */
DD_FN dd_t dd_fma(dd_t a, dd_t b, dd_t c) { return dd_add(dd_mul(a, b), c); }

DD_FN dd_t dd_div(dd_t a, dd_t b)
{
    double q1 = a.hi / b.hi;
    dd_t r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    dd_t q = dd_quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}

DD_FN dd_t dd_div_d(dd_t a, double b) { return dd_div(a, dd_from_d(b)); }

/* multiply by an exact power of two */
DD_FN dd_t dd_mul_pwr2(dd_t a, double p) { return dd_make(DD_MUL(a.hi, p), DD_MUL(a.lo, p)); }

DD_FN int dd_is_zero(dd_t a) { return a.hi == 0.0 && a.lo == 0.0; }
DD_FN dd_t dd_abs(dd_t a) { return (a.hi < 0.0 || (a.hi == 0.0 && a.lo < 0.0)) ? dd_neg(a) : a; }
DD_FN int dd_lt(dd_t a, dd_t b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

#endif
