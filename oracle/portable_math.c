/* portable_math.c -- host build of die_b200/csrc/die_math.h for the oracle's "portable" math
 * backend (TEST INFRASTRUCTURE: lets the numpy oracle evaluate sin/cos/atan2 with the same
 * bit-reproducible routines the CUDA kernels use, so free-running trajectories can be compared
 * bit-for-bit).  Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/build_oracle.py). */
#include "../die_b200/csrc/die_math.h"
#include "../die_b200/csrc/die_turn.h"

void die_sincos_array(const double* x, double* sn, double* cs, long n) {
    for (long i = 0; i < n; ++i) die_sincos(x[i], sn + i, cs + i);
}

void die_atan2_array(const double* y, const double* x, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = die_atan2(y[i], x[i]);
}

void die_atan2_fast_array(const double* y, const double* x, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = die_atan2_fast(y[i], x[i]);
}

/* the guard-banded float32 sin / cos of the forward kernel (tests only: its error bound is checked against die_sincos) */
void die_sincosf_approx_array(const double* x, float* sn, float* cs, long n) {
    for (long i = 0; i < n; ++i) die_sincosf_approx(x[i], sn + i, cs + i);
}

void die_sincos_angle_array(const double* x, double* sn, double* cs, double* ang, long n) {
    for (long i = 0; i < n; ++i) die_sincos_angle(x[i], sn + i, cs + i, ang + i);
}

/* renormalize_radians (core/utils.py:177-179) as the kernels spell it (die_turn.h): checked against numpy's own
 * (r - pi) % (-2 pi) + pi bit for bit (tests/test_portable_math.py) */
void die_renormalize_radians_array(const double* r, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = die_renormalize_radians(r[i]);
}
