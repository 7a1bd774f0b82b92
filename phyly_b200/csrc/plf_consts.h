/* Per-site rescaling constants shared by every kernel family: exponents are counted in units of 2^256. */
#pragma once
#define PLF_SCALE_BITS 256
#define PLF_TWO_P256 1.157920892373162e+77      /* 2^256  */
#define PLF_TWO_M256 8.636168555094445e-78      /* 2^-256 */
