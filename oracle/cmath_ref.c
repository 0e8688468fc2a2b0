/* cmath_ref.c -- TEST INFRASTRUCTURE (oracle): the one libm call whose result the reference depends on bit-for-bit.
 *
 * core/fusion.py:537 evaluates `(la.norm(pos - dg_v)/(2*dg_w))**2` on a numpy float32 SCALAR; numpy's scalar power
 * calls libm's powf(x, 2.0f), and glibc's powf is not correctly rounded (it differs from the rounded product x*x by
 * one ulp for ~0.06 % of inputs).  numpy's ARRAY power takes a square fast-path instead, so the vectorised oracle
 * routes this single operation through the same libm entry point to stay bit-identical to the reference.
 * Built by oracle/cmath.py (gcc) into oracle/libdfb_oracle_cmath.so.
 */
#include <math.h>

void dfb_oracle_pow2(const double* x, double* out, long n) {
    volatile double two = 2.0; /* numpy float64 scalar `** 2` is libm pow(x, 2.0), also not always the rounded product */
    for (long i = 0; i < n; ++i) out[i] = pow(x[i], two);
}

void dfb_oracle_powf2(const float* x, float* out, long n) {
    volatile float two = 2.0f; /* volatile: keep the compiler from folding powf(x, 2) into x*x */
    for (long i = 0; i < n; ++i) out[i] = powf(x[i], two);
}
