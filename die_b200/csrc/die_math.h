/* die_math.h -- portable, bit-reproducible float64 sincos / atan2.
 *
 * Why: the reference evaluates np.sin / np.cos / np.arctan2 (core/utils.py:154-168), whose
 * last bit depends on the host's libm / SIMD build; CUDA's libdevice differs from it again by
 * up to 2 ulp.  The Physarum turn rule compares angles that sit EXACTLY on a threshold
 * (|theta - phi| == sense_angle when the gradient is axis aligned, core/agent/gradient.py:179),
 * so a 1-ulp difference flips a turn and the (chaotic) trajectories part for good.  These
 * routines use only IEEE-754 +,-,*,/ and fma in a fixed order, so the SAME source gives the
 * SAME bits under nvcc (device) and gcc (host, -ffp-contract=off): the CUDA path and the
 * oracle's "portable" math backend agree bit-for-bit, step after step.
 *
 * Accuracy (tests/test_portable_math.py, against mpmath at 200 bits): sin/cos <= 0.6 ulp,
 * atan2 <= 1.1 ulp (<= 0.7 ulp for |result| >= 1): the class of glibc (<= 1 ulp) and better
 * than CUDA libdevice (<= 2 ulp).  Constants come from
 * tools/gen_math_coeffs.py.  Domain: finite inputs; die_sincos assumes |x| < 2^20.
 *
 * die_atan2 is the compensated version (the division and the reduction carry their rounding
 * errors); die_atan2_fast drops the compensation (<= 1.8 ulp) for angles that are only compared
 * against thresholds.
 *
 * C99 / C++ / CUDA compatible; no dependencies beyond <math.h>.
 */
#ifndef DIE_MATH_H
#define DIE_MATH_H

#include <math.h>

#if defined(__CUDACC__)
#define DIE_MATH_FN __host__ __device__ __forceinline__
#else
#define DIE_MATH_FN static inline
#endif

/* single-rounding primitives: never contracted, never reassociated */
#if defined(__CUDA_ARCH__)
#define DIE_FMA(a, b, c) __fma_rn((a), (b), (c))
#define DIE_MUL(a, b)    __dmul_rn((a), (b))
#define DIE_ADD(a, b)    __dadd_rn((a), (b))
#define DIE_SUB(a, b)    __dsub_rn((a), (b))
#define DIE_DIV(a, b)    __ddiv_rn((a), (b))
#else
#define DIE_FMA(a, b, c) fma((a), (b), (c))
#define DIE_MUL(a, b)    ((a) * (b))      /* host: compile with -ffp-contract=off */
#define DIE_ADD(a, b)    ((a) + (b))
#define DIE_SUB(a, b)    ((a) - (b))
#define DIE_DIV(a, b)    ((a) / (b))
#endif

/* ---- constants (tools/gen_math_coeffs.py) ------------------------------------------------ */
#define DIE_2OPI    0x1.45f306dc9c883p-1      /* 2/pi */
#define DIE_PIO2_1  0x1.921fb54400000p+0      /* pi/2, first 33 bits  */
#define DIE_PIO2_T  0x1.0b4611a626331p-34     /* fl(pi/2 - PIO2_1)            */
#define DIE_PIO2_T2 0x1.1701b839a2520p-88    /* fl(pi/2 - PIO2_1 - PIO2_T)   */
#define DIE_PI_HI   0x1.921fb54442d18p+1
#define DIE_PI_LO   0x1.1a62633145c07p-53
#define DIE_PIO2_HI 0x1.921fb54442d18p+0
#define DIE_PIO2_LO 0x1.1a62633145c07p-54

/* sin(r) = r + r z S(z), z = r*r, |r| <= pi/4: near-minimax, degree 6 */
#define DIE_S0 (-0x1.5555555555555p-3)
#define DIE_S1 (0x1.1111111111110p-7)
#define DIE_S2 (-0x1.a01a01a019926p-13)
#define DIE_S3 (0x1.71de3a545e700p-19)
#define DIE_S4 (-0x1.ae64540f0ba18p-26)
#define DIE_S5 (0x1.61217c1864fa1p-33)
#define DIE_S6 (-0x1.ab161757f0c63p-41)
/* cos(r) = 1 - z/2 + z^2 C(z): degree 5 */
#define DIE_C0 (0x1.5555555555555p-5)
#define DIE_C1 (-0x1.6c16c16c16960p-10)
#define DIE_C2 (0x1.a01a019f4d0edp-16)
#define DIE_C3 (-0x1.27e4fa15bd814p-22)
#define DIE_C4 (0x1.1eeb66b683028p-29)
#define DIE_C5 (-0x1.907c0e63adbb6p-37)
/* atan(t) = t + t z A(z), z = t*t, |t| <= tan(pi/24): degree 6 */
#define DIE_A0 (-0x1.5555555555555p-2)
#define DIE_A1 (0x1.9999999999661p-3)
#define DIE_A2 (-0x1.24924923df925p-3)
#define DIE_A3 (0x1.c71c6ff5c5531p-4)
#define DIE_A4 (-0x1.745bf6698bc93p-4)
#define DIE_A5 (0x1.3ab76ce8f3222p-4)
#define DIE_A6 (-0x1.025e08eeb3e61p-4)
/* atan2 reduction: ratio intervals [0,B0) [B0,B1) [B1,B2) [B2,1] use centres 0, K0, K1, 1 */
#define DIE_ATAN_B0  0x1.126e978d4fdf4p-3     /* 0.134 >= K0/2 (tan(pi/24) = 0.1317) */
#define DIE_ATAN_B1  0x1.a827999fcef32p-2     /* tan(3pi/24) */
#define DIE_ATAN_B2  0x1.88df153d6a676p-1     /* tan(5pi/24) */
#define DIE_ATAN_K0  0x1.126145e9ecd56p-2     /* tan(pi/12)  */
#define DIE_ATAN_K1  0x1.279a74590331cp-1     /* tan(pi/6)   */
#define DIE_ATAN_D1  (-0x1.6f5580ddfaeadp-57)  /* atan(K0) -   pi/12 */
#define DIE_ATAN_D2  (-0x1.cec95d0b5c1e3p-56)  /* atan(K1) - 2 pi/12 */
#define DIE_PI12_HI  0x1.0c152382d7366p-2      /* pi/12 = hi + lo */
#define DIE_PI12_LO  (-0x1.ee6913347c2a6p-56)

/* sin and cos of x, |x| < 2^20.  Cody-Waite reduction by pi/2: x - k P1 is exact (33-bit
 * constant times a < 2^20 integer), the rest of pi/2 is subtracted as a double-double, leaving
 * r + rl = x - k pi/2 to ~2^-100; the tail rl and the rounding error of z = r*r are folded into
 * the polynomials so that the only full-size rounding is the final addition (~0.52 ulp). */
DIE_MATH_FN void die_sincos(double x, double* sn_out, double* cs_out) {
    const double kd = rint(DIE_MUL(x, DIE_2OPI));
    const double r1 = DIE_FMA(-kd, DIE_PIO2_1, x);              /* exact */
    const double w = DIE_MUL(kd, DIE_PIO2_T);
    const double wl = DIE_FMA(kd, DIE_PIO2_T2, DIE_FMA(kd, DIE_PIO2_T, -w));   /* w + wl = kd (pi/2 - P1) */
    const double r = DIE_SUB(r1, w);
    const double bb = DIE_SUB(r, r1);                            /* TwoSum(r1, -w): r + re = r1 - w */
    const double re = DIE_SUB(DIE_SUB(r1, DIE_SUB(r, bb)), DIE_ADD(w, bb));
    const double rl = DIE_SUB(re, wl);                           /* r + rl = x - kd pi/2 */
    const int k = (int)kd;
    const double z = DIE_MUL(r, r);
    const double zl = DIE_FMA(r, r, -z);                         /* z + zl = r*r exactly */

    /* sin(r + rl) = r + [ r z S(z) + rl (1 - z/2) ] */
    double ps = DIE_S6;
    ps = DIE_FMA(ps, z, DIE_S5);
    ps = DIE_FMA(ps, z, DIE_S4);
    ps = DIE_FMA(ps, z, DIE_S3);
    ps = DIE_FMA(ps, z, DIE_S2);
    ps = DIE_FMA(ps, z, DIE_S1);
    ps = DIE_FMA(ps, z, DIE_S0);
    const double hz = DIE_MUL(0.5, z);
    const double s_corr = DIE_FMA(DIE_MUL(r, z), ps, DIE_FMA(-hz, rl, rl));
    const double sn = DIE_ADD(r, s_corr);

    /* cos(r + rl) = (1 - z/2) + [ z^2 C(z) - r rl - zl/2 ], with 1 - z/2 = cw + ce exactly */
    double pc = DIE_C5;
    pc = DIE_FMA(pc, z, DIE_C4);
    pc = DIE_FMA(pc, z, DIE_C3);
    pc = DIE_FMA(pc, z, DIE_C2);
    pc = DIE_FMA(pc, z, DIE_C1);
    pc = DIE_FMA(pc, z, DIE_C0);
    const double cw = DIE_SUB(1.0, hz);
    const double ce = DIE_SUB(DIE_SUB(1.0, cw), hz);
    const double c_corr = DIE_SUB(DIE_FMA(DIE_MUL(z, z), pc, ce), DIE_FMA(r, rl, DIE_MUL(0.5, zl)));
    const double cs = DIE_ADD(cw, c_corr);

    const double s_sel = (k & 1) ? cs : sn;
    const double c_sel = (k & 1) ? sn : cs;
    *sn_out = (x == 0.0) ? x : ((k & 2) ? -s_sel : s_sel);       /* sin(-0.) = -0. */
    *cs_out = ((k + 1) & 2) ? -c_sel : c_sel;
}

/* atan2(y, x) for finite arguments, IEEE signed-zero conventions (atan2(+0,-0) = pi ...).
 * With mn = min(|x|,|y|), mx = max: the ratio interval picks a centre K_j ~ tan(j pi/12) and
 *   atan(mn/mx) = atan(K_j) + atan(t),  t = (mn - K_j mx) / (mx + K_j mn),  |t| <= tan(pi/24)
 * (ONE division).  The octant then gives  result = m pi/12 +- (atan(t) + delta_j)  with an
 * integer m in 0..12 and delta_j = atan(K_j) - j pi/12; m pi/12 is formed as an exact hi + lo
 * pair, so the only full-size rounding is the last addition. */
DIE_MATH_FN double die_atan2_fast(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const int swapped = ay > ax;
    const double mx = swapped ? ay : ax;
    const double mn = swapped ? ax : ay;
    double at = 0.0, jd = 0.0;
    if (mx != 0.0) {
        double c, dj;
        if (mn < DIE_MUL(DIE_ATAN_B0, mx))      { c = 0.0;         jd = 0.0; dj = 0.0; }
        else if (mn < DIE_MUL(DIE_ATAN_B1, mx)) { c = DIE_ATAN_K0; jd = 1.0; dj = DIE_ATAN_D1; }
        else if (mn < DIE_MUL(DIE_ATAN_B2, mx)) { c = DIE_ATAN_K1; jd = 2.0; dj = DIE_ATAN_D2; }
        else                                    { c = 1.0;         jd = 3.0; dj = 0.0; }
        const double num = DIE_FMA(-c, mx, mn);
        const double den = DIE_FMA(c, mn, mx);
        const double t = DIE_DIV(num, den);
        const double z = DIE_MUL(t, t);
        double p = DIE_A6;
        p = DIE_FMA(p, z, DIE_A5);
        p = DIE_FMA(p, z, DIE_A4);
        p = DIE_FMA(p, z, DIE_A3);
        p = DIE_FMA(p, z, DIE_A2);
        p = DIE_FMA(p, z, DIE_A1);
        p = DIE_FMA(p, z, DIE_A0);
        at = DIE_ADD(DIE_FMA(DIE_MUL(t, z), p, t), dj);
    }
    const int neg = signbit(x) != 0;
    /* octant:  A (!swapped,!neg)  j pi/12 + at      B (swapped,!neg)  (6-j) pi/12 - at
     *          C (!swapped, neg)  (12-j) pi/12 - at  D (swapped, neg)  (6+j) pi/12 + at     */
    const double md = swapped ? (neg ? DIE_ADD(6.0, jd) : DIE_SUB(6.0, jd))
                              : (neg ? DIE_SUB(12.0, jd) : jd);
    const double sat = (swapped != neg) ? -at : at;
    const double hi = DIE_MUL(md, DIE_PI12_HI);
    const double lo = DIE_FMA(md, DIE_PI12_LO, DIE_FMA(md, DIE_PI12_HI, -hi));
    const double res = DIE_ADD(hi, DIE_ADD(lo, sat));
    return copysign(res, y);
}

/* Compensated atan2: as die_atan2_fast, but num = mn - K mx and den = mx + K mn are carried as
 * double-doubles (mn - fl(K mx) is exact by Sterbenz because B0 / K0 >= 1/2), the quotient gets
 * one Newton correction from the exact remainder, and the final sum keeps its rounding error:
 * the only full-size rounding is the last addition (<= ~0.55 ulp). */
DIE_MATH_FN double die_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const int swapped = ay > ax;
    const double mx = swapped ? ay : ax;
    const double mn = swapped ? ax : ay;
    double t = 0.0, tc = 0.0, jd = 0.0;                          /* atan(t) + tc is the reduced angle */
    if (mx != 0.0) {
        double c, dj;
        if (mn < DIE_MUL(DIE_ATAN_B0, mx))      { c = 0.0;         jd = 0.0; dj = 0.0; }
        else if (mn < DIE_MUL(DIE_ATAN_B1, mx)) { c = DIE_ATAN_K0; jd = 1.0; dj = DIE_ATAN_D1; }
        else if (mn < DIE_MUL(DIE_ATAN_B2, mx)) { c = DIE_ATAN_K1; jd = 2.0; dj = DIE_ATAN_D2; }
        else                                    { c = 1.0;         jd = 3.0; dj = 0.0; }
        const double ph = DIE_MUL(c, mx), pl = DIE_FMA(c, mx, -ph);
        const double sd = DIE_SUB(mn, ph);                       /* exact */
        const double nh = DIE_SUB(sd, pl);
        const double nl = DIE_SUB(DIE_SUB(sd, nh), pl);          /* nh + nl = mn - c mx */
        const double qh = DIE_MUL(c, mn), ql = DIE_FMA(c, mn, -qh);
        const double dh = DIE_ADD(mx, qh);
        const double dl = DIE_ADD(DIE_ADD(DIE_SUB(mx, dh), qh), ql);   /* dh + dl = mx + c mn */
        t = DIE_DIV(nh, dh);
        const double rem = DIE_FMA(-t, dh, nh);                  /* exact remainder */
        const double tl = DIE_DIV(DIE_FMA(-t, dl, DIE_ADD(rem, nl)), dh);
        const double z = DIE_MUL(t, t);
        double p = DIE_A6;
        p = DIE_FMA(p, z, DIE_A5);
        p = DIE_FMA(p, z, DIE_A4);
        p = DIE_FMA(p, z, DIE_A3);
        p = DIE_FMA(p, z, DIE_A2);
        p = DIE_FMA(p, z, DIE_A1);
        p = DIE_FMA(p, z, DIE_A0);
        /* atan(t + tl) = t + [ t z A(z) + tl (1 - z) ] */
        tc = DIE_ADD(DIE_FMA(DIE_MUL(t, z), p, DIE_FMA(-z, tl, tl)), dj);
    }
    const int neg = signbit(x) != 0;
    const double md = swapped ? (neg ? DIE_ADD(6.0, jd) : DIE_SUB(6.0, jd))
                              : (neg ? DIE_SUB(12.0, jd) : jd);
    const int flip = swapped != neg;
    const double st = flip ? -t : t;
    const double stc = flip ? -tc : tc;
    const double hi = DIE_MUL(md, DIE_PI12_HI);
    const double lo = DIE_FMA(md, DIE_PI12_LO, DIE_FMA(md, DIE_PI12_HI, -hi));
    const double h = DIE_ADD(hi, st);                            /* Fast2Sum: |hi| >= |st| or hi == 0 */
    const double e = DIE_ADD(DIE_SUB(hi, h), st);
    const double res = DIE_ADD(h, DIE_ADD(e, DIE_ADD(lo, stc)));
    return copysign(res, y);
}

#endif /* DIE_MATH_H */
