/* die_turn.h -- PhysarumAgent._choose_turn (core/agent/gradient.py:168-193) for one slot, twice:
 *
 *   die_turn_exact  the reference's own arithmetic, operation by operation: normalise the sampled
 *                   gradient (:59-65), phi = angle(gx + 1j gy) (core/utils.py:167-168), delta =
 *                   renormalize_radians(theta - phi), the three isclose / threshold tests.
 *   die_turn_quick  the same DECISION from the raw gradient and (sin theta, cos theta) -- which the
 *                   caller already has for the sense offset -- in float32, without sqrt, division or
 *                   arctangent:  cos(delta) |g| = c gx + s gy,  sin(delta) |g| = s gx - c gy.
 *                   Every comparison carries a guard band that is ~100x wider than the float32 error
 *                   (and ~10^9 x wider than the rounding of the exact path), and a slot that falls
 *                   inside ANY band, or whose gradient is out of float32 range, is reported as
 *                   undecided: the caller then runs die_turn_exact.  Decided slots therefore get
 *                   exactly the result of the reference arithmetic; only the cost differs
 *                   (about 70 float64 instructions per slot less on the GPU).
 *
 * Outputs are the three things _choose_turn's callers use: the turn (-1, +1, or 0 = "take the coin"),
 * the deposit mask (:190) and, exact path only, the normalised gradient.
 *
 * C99 / C++ / CUDA; compiled for the host too (tests/csrc/turn_check.c, tests/test_turn_quick.py) so the guard-band logic is
 * checked against the exact path on the CPU over adversarial inputs.
 */
#ifndef DIE_TURN_H
#define DIE_TURN_H

#include <float.h>
#include "die_math.h"

#define DIE_PI      3.141592653589793    /* np.pi */
#define DIE_TWO_PI  6.283185307179586    /* 2 * np.pi (exact doubling) */

/* np.remainder(a, b) = fmod(a, b) moved into the sign of b; an exact zero takes the sign of b.
 * fmod is exact; for |a| < 2|b| it is a or |a| - |b| (Sterbenz), which is all the hot path ever
 * sees, so the generic fmod() sits behind an unlikely branch. */
DIE_MATH_FN double die_np_remainder(double a, double b) {
    const double fa = fabs(a), fb = fabs(b);
    double m;
    if (fa < 2.0 * fb) m = (fa < fb) ? a : copysign(DIE_SUB(fa, fb), a);
    else m = fmod(a, b);
    if (m == 0.0) return copysign(0.0, b);
    return ((b < 0.0) != (m < 0.0)) ? DIE_ADD(m, b) : m;
}

/* renormalize_radians (core/utils.py:177-179): (r - pi) % (-2 pi) + pi, in (-pi, pi].
 * np.remainder(a, -2pi) spelled out for |a| < 4 pi: fmod(a, -2pi) is a, a - 2pi or a + 2pi
 * (exact), and a positive remainder is moved into the divisor's sign by one ROUNDED add of -2pi.
 * The sign of a zero remainder is irrelevant here (+-0 + pi = pi). */
DIE_MATH_FN double die_renormalize_radians(double r) {
    const double a = DIE_SUB(r, DIE_PI);
    double m;
    if (fabs(a) < 2.0 * DIE_TWO_PI) {
        /* f = a - 2pi for a >= 2pi, a + 2pi for a <= -2pi, a otherwise; m = f - 2pi for f > 0, f otherwise -- written
         * as additions of a selected constant (x - c and x + (-c) are the same operation; x + 0.0 = x except that it
         * turns -0. into +0., which the final + pi absorbs): two selects of a constant instead of four of a variable */
        const double wrap = (fabs(a) >= DIE_TWO_PI) ? DIE_TWO_PI : 0.0;
        const double f = DIE_ADD(a, copysign(wrap, -a));
        m = DIE_ADD(f, (f > 0.0) ? -DIE_TWO_PI : 0.0);
    } else {
        m = die_np_remainder(a, -DIE_TWO_PI);
    }
    return DIE_ADD(m, DIE_PI);
}

/* np.angle(x + np.multiply(1j, y))  (core/utils.py:158-168).  The complex construction yields
 * re = x + (0*y - 0), im = 0 + y, which is what makes (-0., -0.) -> +pi and every other all-zero
 * pair -> 0.  fast != 0: die_atan2_fast (angles only compared against thresholds). */
DIE_MATH_FN double die_angle_xy(double x, double y, int fast) {
    const double re = DIE_ADD(x, DIE_SUB(DIE_MUL(0.0, y), 0.0));
    const double im = DIE_ADD(0.0, y);
    return fast ? die_atan2_fast(im, re) : die_atan2(im, re);
}

/* np.nan_to_num(a / n)  (core/agent/gradient.py:62) */
DIE_MATH_FN double die_div_nan_to_num(double a, double n) {
    const double q = DIE_DIV(a, n);
    if (q != q) return 0.0;
    if (fabs(q) > DBL_MAX) return copysign(DBL_MAX, q);
    return q;
}

/* GradientAgent._get_gradient per sample (core/agent/gradient.py:59-65): scipy.linalg.norm(axis=0,
 * ord=2) == sqrt(gx*gx + gy*gy) (no hypot scaling); grad = nan_to_num(grad / norm); grad *= (norm >=
 * clip) keeps signed zeros. */
DIE_MATH_FN void die_normalize_gradient(double* gx_io, double* gy_io, int normalized, int use_clip, double clip) {
    double gx = *gx_io, gy = *gy_io;
    const double norm = sqrt(DIE_ADD(DIE_MUL(gx, gx), DIE_MUL(gy, gy)));
    const int clipped = use_clip && !(norm >= clip);
    if (normalized) {
        if (clipped) {              /* (+-q) * 0.0: only the zero's sign survives; 0/0 -> nan -> +0 */
            gx = (norm == 0.0 && gx == 0.0) ? 0.0 : copysign(0.0, gx);
            gy = (norm == 0.0 && gy == 0.0) ? 0.0 : copysign(0.0, gy);
        } else if (use_clip && clip > 0.0) {   /* norm >= clip > 0: the quotient is finite */
            gx = DIE_DIV(gx, norm);
            gy = DIE_DIV(gy, norm);
        } else {
            gx = die_div_nan_to_num(gx, norm);
            gy = die_div_nan_to_num(gy, norm);
        }
    } else if (clipped) {
        gx = DIE_MUL(gx, 0.0);
        gy = DIE_MUL(gy, 0.0);
    }
    *gx_io = gx;
    *gy_io = gy;
}

typedef struct die_turn {
    int turn;           /* -1 / +1: forced by the gradient; 0: undetermined, the coin decides */
    int deposit_mask;   /* not (undetermined_grad or undetermined_turn), :190 */
} die_turn_t;

/* _choose_turn from the gradient's polar angle (core/agent/gradient.py:168-193), the reference's arithmetic. */
DIE_MATH_FN die_turn_t die_turn_from_angle(double drads, double theta, double atol, double sense_radians) {
    die_turn_t o;
    double dd = die_renormalize_radians(DIE_SUB(theta, drads));
    const int und_grad = fabs(DIE_SUB(0.0, drads)) <= DIE_ADD(1e-8, DIE_MUL(1e-5, fabs(drads)));
    const int und_turn = fabs(DIE_SUB(0.0, dd)) <= DIE_ADD(atol, DIE_MUL(1e-2, fabs(dd)));
    const int unseen = fabs(dd) > sense_radians;
    if (und_grad || und_turn || unseen) dd = DIE_MUL(dd, 0.0);
    o.turn = 0;
    if (dd > atol) o.turn = -1;
    if (dd < -atol) o.turn = 1;
    o.deposit_mask = !(und_grad || und_turn);
    return o;
}

/* (gx, gy) is the PROCESSED gradient (die_normalize_gradient); xy2polar's fast angle (core/utils.py:167-168). */
DIE_MATH_FN die_turn_t die_turn_exact(double gx, double gy, double theta, double atol, double sense_radians) {
    return die_turn_from_angle(die_angle_xy(gx, gy, 1), theta, atol, sense_radians);
}

/* Guard-banded thresholds of the quick path, built once per launch on the host (die_turn_plan). */
typedef struct die_turn_plan {
    int   enabled;
    float clip2_lo, clip2_hi;     /* n2 <= lo: clipped for sure; n2 >= hi: not clipped for sure */
    float n2_max;                 /* above: out of the float32 comfort zone */
    float ka_lo, ka_hi;           /* signed squares of cos(atol / 0.99) -+ band */
    float ks_lo, ks_hi;           /* signed squares of cos(sense_radians) -+ band */
    float phi0_ratio;             /* |gy| <= ratio * gx with gx > 0: phi may be "close to 0" */
} die_turn_plan_t;

#define DIE_TURN_BAND 2e-5        /* on the cosine; float32 evaluation errs by < 5e-7 */

static inline float die_signed_square_(double v) { return (float)(v * fabs(v)); }

/* Host side.  The quick path is only offered where its case analysis holds:
 * normalised + clipped gradients (the Physarum defaults), 0 < atol/0.99 < sense < pi. */
static inline die_turn_plan_t die_turn_plan(int normalized, int use_clip, double clip,
                                            double atol, double sense_radians) {
    die_turn_plan_t p;
    memset(&p, 0, sizeof p);
    const double a99 = atol / 0.99;
    if (!normalized || !use_clip || !(clip >= 1e-12 && clip <= 1e6)) return p;
    if (!(atol > 1e-3 && a99 < 1.5 && a99 + 1e-3 < sense_radians && sense_radians < DIE_PI - 1e-3)) return p;
    p.enabled = 1;
    p.clip2_lo = (float)(clip * clip * (1.0 - 1e-4));
    p.clip2_hi = (float)(clip * clip * (1.0 + 1e-4));
    p.n2_max = 1e30f;
    p.ka_lo = die_signed_square_(cos(a99) - DIE_TURN_BAND);
    p.ka_hi = die_signed_square_(cos(a99) + DIE_TURN_BAND);
    p.ks_lo = die_signed_square_(cos(sense_radians) - DIE_TURN_BAND);
    p.ks_hi = die_signed_square_(cos(sense_radians) + DIE_TURN_BAND);
    p.phi0_ratio = 4e-8f;
    return p;
}

/* Returns 1 and fills *out when the decision is safe, 0 when the caller must run the exact path.
 * (gx, gy): RAW sampled gradient; (sn, cs) = sin / cos of the heading theta.
 *   n2 = |g|^2, cd = |g| cos(delta), sd = |g| sin(delta), q = cd |cd| = |g|^2 cos|cos|(delta):
 *   |delta| <= A  <=>  cos(delta) >= cos(A)  <=>  q >= cos|cos|(A) n2      (A in (0, pi)).
 * A clipped gradient (|g| < grad_clip: every slot far from any trail) is settled exactly instead: its
 * processed value is (+-0, +-0), so phi is 0 (undetermined_grad) or, for two negative zeros, exactly pi
 * -- and delta = renormalize(theta - pi) then sits ON the sense threshold for headings of +-90 degrees. */
/* The decision proper on the float32 image of the gradient.  neg_both / any_zero: both components of the float64
 * gradient carry a sign bit / one of them is a zero (only looked at for a clipped gradient). */
#define DIE_TURN_QUICK_BODY_(NEG_BOTH, ANY_ZERO)                                                                     \
    const float n2 = fmaf(gxf, gxf, gyf * gyf);                                                                      \
    if (!(n2 < p->n2_max)) return 0;                        /* huge, inf or nan */                                   \
    if (!(n2 >= p->clip2_hi)) {                                                                                      \
        if (!(n2 <= p->clip2_lo)) return 0;                 /* too close to the clip threshold */                    \
        if (!(NEG_BOTH)) {                                  /* a positive zero survives: phi = 0 */                  \
            out->turn = 0;                                                                                           \
            out->deposit_mask = 0;                                                                                   \
            return 1;                                                                                                \
        }                                                                                                            \
        /* both negative: -0 only keeps its sign if norm != 0 (die_normalize_gradient), which takes the              \
         * exact norm when a component IS zero */                                                                    \
        if (ANY_ZERO) return 0;                                                                                      \
        *out = die_turn_from_angle(DIE_PI, theta, atol, sense_radians);     /* angle(-0. - 0.j) = pi */              \
        return 1;                                                                                                    \
    }                                                                                                                \
    if (gxf > 0.0f && fabsf(gyf) <= p->phi0_ratio * gxf) return 0;    /* phi within ~1e-8 of 0: isclose(0, phi) */   \
    const float cd = fmaf(cf, gxf, sf * gyf);                                                                        \
    const float sd = fmaf(sf, gxf, -(cf * gyf));                                                                     \
    const float q = cd * fabsf(cd);                                                                                  \
    int unseen, und_turn;                                                                                            \
    if (q < p->ks_lo * n2) unseen = 1;                                                                               \
    else if (q > p->ks_hi * n2) unseen = 0;                                                                          \
    else return 0;                                                                                                   \
    if (q > p->ka_hi * n2) und_turn = 1;                                                                             \
    else if (q < p->ka_lo * n2) und_turn = 0;                                                                        \
    else return 0;                                                                                                   \
    out->deposit_mask = !und_turn;                                                                                   \
    /* seen and determined: atol/0.99 < |delta| < sense < pi, so sin(delta) is far from 0 and has                    \
     * the sign of delta; delta > atol turns by -1, delta < -atol by +1 (:186-187) */                                \
    out->turn = (unseen || und_turn) ? 0 : (sd > 0.0f ? -1 : 1);                                                     \
    return 1;

/* The gradient as published in float32 (the env's float32 cache: the value IS the float64 gradient rounded once, so
 * its signs and zeros are those of the float64 value unless that underflowed -- n2 then sits far below the clip). */
DIE_MATH_FN int die_turn_quick_ff(const die_turn_plan_t* p, float gxf, float gyf, float sf, float cf,
                                  double theta, double atol, double sense_radians, die_turn_t* out) {
    DIE_TURN_QUICK_BODY_(signbit(gxf) && signbit(gyf), gxf == 0.0f || gyf == 0.0f)
}

DIE_MATH_FN int die_turn_quick_f(const die_turn_plan_t* p, double gx, double gy, float sf, float cf,
                                 double theta, double atol, double sense_radians, die_turn_t* out) {
    const float gxf = (float)gx, gyf = (float)gy;
    DIE_TURN_QUICK_BODY_(signbit(gx) && signbit(gy), gx == 0.0 || gy == 0.0)
}

DIE_MATH_FN int die_turn_quick(const die_turn_plan_t* p, double gx, double gy, double sn, double cs,
                               double theta, double atol, double sense_radians, die_turn_t* out) {
    return die_turn_quick_f(p, gx, gy, (float)sn, (float)cs, theta, atol, sense_radians, out);
}

#endif /* DIE_TURN_H */
