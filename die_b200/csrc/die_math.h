/* die_math.h -- portable, bit-reproducible float64 sincos / atan2.
 *
 * Why: the reference evaluates np.sin / np.cos / np.arctan2 (core/utils.py:154-168), whose last
 * bit depends on the host's libm / SIMD build; CUDA's libdevice differs from it again by up to
 * 2 ulp.  The Physarum turn rule compares angles that sit EXACTLY on a threshold
 * (|theta - phi| == sense_angle when the gradient is axis aligned, core/agent/gradient.py:179),
 * so a 1-ulp difference flips a turn and the trajectories part.  These routines use only
 * IEEE-754 +,-,*,/ and fma in a fixed order, so the SAME source gives the SAME bits under nvcc
 * (device) and gcc (host, -ffp-contract=off): the CUDA path and the oracle's "portable" math
 * backend (oracle/portable_math.c) agree bit-for-bit, step after step.
 *
 * Accuracy against mpmath at 200 bits (tests/test_portable_math.py):
 *     die_sincos        <= 0.70 ulp   (98.6% of results identical to glibc's)
 *     die_atan2         <= 0.502 ulp  (compensated; glibc: 0.68)
 *     die_atan2_fast    <= 1.8 ulp    (for angles that are only compared against thresholds)
 *     die_sincos_angle  sin, cos as die_sincos, plus atan2(sin, cos) to <= 0.7 ulp at the cost
 *                       of six flops (first-order correction from the rounding errors of sin
 *                       and cos, which the evaluation already has at hand)
 * Constants come from tools/gen_math_coeffs.py.  Domain: finite inputs; |x| < 2^20 for sincos.
 *
 * C99 / C++ / CUDA compatible; no dependencies beyond <math.h>.
 */
#ifndef DIE_MATH_H
#define DIE_MATH_H

#include <math.h>
#include <string.h>

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

/* bit-level helpers: low word of a double (integer part after adding DIE_RINT_MAGIC), and a
 * conditional sign flip that is one integer XOR on the device instead of a select pair */
#define DIE_RINT_MAGIC 6755399441055744.0        /* 1.5 * 2^52: (v + M) - M == rint(v) for |v| < 2^51 */
#if defined(__CUDA_ARCH__)
#define DIE_LO32(d) __double2loint(d)
#define DIE_NEGIF(v, flag) __hiloint2double(__double2hiint(v) ^ ((flag) ? (int)0x80000000 : 0), __double2loint(v))
#else
static inline int die_lo32_(double d) {
    unsigned long long b;
    memcpy(&b, &d, sizeof b);
    return (int)(unsigned)(b & 0xffffffffu);
}
#define DIE_LO32(d) die_lo32_(d)
#define DIE_NEGIF(v, flag) ((flag) ? -(v) : (v))
#endif

/* ---- constants -------------------------------------------------------------------------------
 * One table.  On the device it is __constant__ memory, so each use is a constant-bank operand of
 * the DFMA/DMUL itself (as 64-bit literals nvcc materialises every use with two UMOVs, which
 * doubled the instruction count of the polynomials); on the host the compiler folds the
 * constant-index loads into immediates. */
#define DIE_MATH_CONSTANTS(X)                                                                   \
    X(TWO_OVER_PI, 0x1.45f306dc9c883p-1)                                                        \
    X(PIO2_1, 0x1.921fb54400000p+0)      /* pi/2, first 33 bits */                              \
    X(PIO2_T, 0x1.0b4611a626331p-34)     /* fl(pi/2 - PIO2_1) */                                \
    X(PIO2_T2, 0x1.1701b839a2520p-88)    /* fl(pi/2 - PIO2_1 - PIO2_T) */                       \
    /* sin(r) = r + r z S(z), z = r*r, |r| <= pi/4: near-minimax, degree 6 */                   \
    X(S0, -0x1.5555555555555p-3)                                                                \
    X(S1, 0x1.1111111111110p-7)                                                                 \
    X(S2, -0x1.a01a01a019926p-13)                                                               \
    X(S3, 0x1.71de3a545e700p-19)                                                                \
    X(S4, -0x1.ae64540f0ba18p-26)                                                               \
    X(S5, 0x1.61217c1864fa1p-33)                                                                \
    X(S6, -0x1.ab161757f0c63p-41)                                                               \
    /* cos(r) = 1 - z/2 + z^2 C(z): degree 5 */                                                 \
    X(C0, 0x1.5555555555555p-5)                                                                 \
    X(C1, -0x1.6c16c16c16960p-10)                                                               \
    X(C2, 0x1.a01a019f4d0edp-16)                                                                \
    X(C3, -0x1.27e4fa15bd814p-22)                                                               \
    X(C4, 0x1.1eeb66b683028p-29)                                                                \
    X(C5, -0x1.907c0e63adbb6p-37)                                                               \
    /* atan(t) = t + t z A(z), z = t*t, |t| <= 0.134: degree 6 */                               \
    X(A0, -0x1.5555555555555p-2)                                                                \
    X(A1, 0x1.9999999999661p-3)                                                                 \
    X(A2, -0x1.24924923df925p-3)                                                                \
    X(A3, 0x1.c71c6ff5c5531p-4)                                                                 \
    X(A4, -0x1.745bf6698bc93p-4)                                                                \
    X(A5, 0x1.3ab76ce8f3222p-4)                                                                 \
    X(A6, -0x1.025e08eeb3e61p-4)                                                                \
    /* atan2 reduction: ratio intervals [0,B0) [B0,B1) [B1,B2) [B2,1] use the centres below */  \
    X(ATAN_B0, 0x1.126e978d4fdf4p-3)     /* 0.134 (>= K1/2 so that mn - fl(K mx) is exact) */   \
    X(ATAN_B1, 0x1.a827999fcef32p-2)     /* tan(3 pi/24) */                                     \
    X(ATAN_B2, 0x1.88df153d6a676p-1)     /* tan(5 pi/24) */                                     \
    X(ATAN_K0, 0.0)                      /* centres K_j ~ tan(j pi/12), j = 0..3 (indexed) */   \
    X(ATAN_K1, 0x1.126145e9ecd56p-2)                                                            \
    X(ATAN_K2, 0x1.279a74590331cp-1)                                                            \
    X(ATAN_K3, 1.0)                                                                             \
    X(ATAN_D0, 0.0)                      /* delta_j = atan(K_j) - j pi/12 (indexed) */          \
    X(ATAN_D1, -0x1.6f5580ddfaeadp-57)                                                          \
    X(ATAN_D2, -0x1.cec95d0b5c1e3p-56)                                                          \
    X(ATAN_D3, 0.0)                                                                             \
    X(PI12_HI, 0x1.0c152382d7366p-2)     /* pi/12 = hi + lo */                                  \
    X(PI12_LO, -0x1.ee6913347c2a6p-56)

#define DIE_MATH_ENUM_(name, value) DIE_K_##name,
#define DIE_MATH_VALUE_(name, value) value,
enum { DIE_MATH_CONSTANTS(DIE_MATH_ENUM_) DIE_K_COUNT };
#if defined(__CUDACC__)
static __constant__ double die_kc[DIE_K_COUNT] = {DIE_MATH_CONSTANTS(DIE_MATH_VALUE_)};
#endif
static const double die_kh[DIE_K_COUNT] = {DIE_MATH_CONSTANTS(DIE_MATH_VALUE_)};
#if defined(__CUDA_ARCH__)
#define DIE_KI(index) die_kc[index]
#else
#define DIE_KI(index) die_kh[index]
#endif
#define DIE_K(name) DIE_KI(DIE_K_##name)

/* ---- sin / cos ---------------------------------------------------------------------------------
 * Cody-Waite reduction by pi/2: x - k P1 is exact (33-bit constant times a < 2^20 integer), the
 * rest of pi/2 is subtracted as a double-double, leaving r + rl = x - k pi/2 to ~2^-100.  The
 * tail rl and the rounding error of z = r*r are folded into the polynomials, so the only
 * full-size rounding of each result is its final addition: sin(r + rl) = sn + es and
 * cos(r + rl) = cs + ec with |es|, |ec| the (exactly known) errors of those additions. */
typedef struct die_sincos_parts {
    double sn, es;      /* sin of the reduced argument and the rounding error of its last add */
    double cs, ec;
    int k;              /* quadrant: x = r + k pi/2 */
} die_sincos_parts_t;

DIE_MATH_FN die_sincos_parts_t die_sincos_core(double x) {
    die_sincos_parts_t o;
    const double ku = DIE_ADD(DIE_MUL(x, DIE_K(TWO_OVER_PI)), DIE_RINT_MAGIC);
    const double kd = DIE_SUB(ku, DIE_RINT_MAGIC);                          /* rint(x 2/pi) */
    const double r1 = DIE_FMA(-kd, DIE_K(PIO2_1), x);                       /* exact */
    const double w = DIE_MUL(kd, DIE_K(PIO2_T));
    const double wl = DIE_FMA(kd, DIE_K(PIO2_T2), DIE_FMA(kd, DIE_K(PIO2_T), -w));
    const double r = DIE_SUB(r1, w);                                        /* TwoSum(r1, -w) */
    const double bb = DIE_SUB(r, r1);
    const double re = DIE_SUB(DIE_SUB(r1, DIE_SUB(r, bb)), DIE_ADD(w, bb));
    const double rl = DIE_SUB(re, wl);                                      /* r + rl = x - k pi/2 */
    o.k = DIE_LO32(ku);
    const double z = DIE_MUL(r, r);
    const double zl = DIE_FMA(r, r, -z);                                    /* z + zl = r*r exactly */
    const double hz = DIE_MUL(0.5, z);

    /* sin(r + rl) = r + [ r z S(z) + rl (1 - z/2) ] */
    double ps = DIE_K(S6);
    ps = DIE_FMA(ps, z, DIE_K(S5));
    ps = DIE_FMA(ps, z, DIE_K(S4));
    ps = DIE_FMA(ps, z, DIE_K(S3));
    ps = DIE_FMA(ps, z, DIE_K(S2));
    ps = DIE_FMA(ps, z, DIE_K(S1));
    ps = DIE_FMA(ps, z, DIE_K(S0));
    const double s_corr = DIE_FMA(DIE_MUL(r, z), ps, DIE_FMA(-hz, rl, rl));
    o.sn = DIE_ADD(r, s_corr);
    o.es = DIE_ADD(DIE_SUB(r, o.sn), s_corr);                               /* Fast2Sum: |r| >= |s_corr| */

    /* cos(r + rl) = (1 - z/2) + [ z^2 C(z) - r rl - zl/2 ], with 1 - z/2 = cw + ce exactly */
    double pc = DIE_K(C5);
    pc = DIE_FMA(pc, z, DIE_K(C4));
    pc = DIE_FMA(pc, z, DIE_K(C3));
    pc = DIE_FMA(pc, z, DIE_K(C2));
    pc = DIE_FMA(pc, z, DIE_K(C1));
    pc = DIE_FMA(pc, z, DIE_K(C0));
    const double cw = DIE_SUB(1.0, hz);
    const double ce = DIE_SUB(DIE_SUB(1.0, cw), hz);
    const double c_corr = DIE_SUB(DIE_FMA(DIE_MUL(z, z), pc, ce), DIE_FMA(r, rl, DIE_MUL(0.5, zl)));
    o.cs = DIE_ADD(cw, c_corr);
    o.ec = DIE_ADD(DIE_SUB(cw, o.cs), c_corr);                              /* Fast2Sum: cw >= |c_corr| */
    return o;
}

/* sin and cos of x, |x| < 2^20. */
DIE_MATH_FN void die_sincos(double x, double* sn_out, double* cs_out) {
    const die_sincos_parts_t o = die_sincos_core(x);
    const double s_sel = (o.k & 1) ? o.cs : o.sn;
    const double c_sel = (o.k & 1) ? o.sn : o.cs;
    *sn_out = (x == 0.0) ? x : DIE_NEGIF(s_sel, o.k & 2);                  /* sin(-0.) = -0. */
    *cs_out = DIE_NEGIF(c_sel, (o.k + 1) & 2);
}

/* ---- sqrt(s) of an s known in advance to be close to r0 * r0 ----------------------------------------------------------------
 * The forward kernel's cost hint needs sqrt(dx*dx + dy*dy) of an action (dx, dy) = scale * (cos, sin): the root is |scale| to
 * a few ulps.  IEEE sqrt in float64 is ~40 instructions on the GPU; one Newton step from r0 plus a check is 7:
 *     e = s - r0*r0 (fma),  y = r0 + e * (0.5 / r0) (fma),  d = s - y*y (fma)
 * and y is RETURNED ONLY IF it is provably the correctly rounded root: |e| < r0^2 2^-40 keeps sqrt(s) and y in r0's binade
 * (the plan refuses an r0 within 2^-30 of a power of two), where ulp u is a constant, and |d| < r0 u (1 - 2^-20) gives
 * |sqrt(s) - y| = |s - y*y| / (sqrt(s) + y) < u / 2 strictly, so y is the unique nearest double.  Everything else -- an s that
 * is not near r0^2, a root within 2^-20 u of a rounding boundary, a disabled plan -- takes sqrt().  Same result bit for bit
 * either way (tests/test_turn_quick.py::test_sqrt_near_*). */
typedef struct die_sqrt_near {
    double r0, half_inv, margin, elim;
    int enabled;
} die_sqrt_near_t;

DIE_MATH_FN die_sqrt_near_t die_sqrt_near_plan(double r0) {
    die_sqrt_near_t p;
    memset(&p, 0, sizeof(p));
    r0 = fabs(r0);
    if (!(r0 > 1e-100 && r0 < 1e100)) return p;
    int ex;
    const double mant = frexp(r0, &ex);                      /* r0 = mant 2^ex, mant in [0.5, 1) */
    if (mant < 0.5 * (1.0 + 9.4e-10) || mant > 1.0 - 9.4e-10) return p;     /* within 2^-30 of a binade edge */
    const double ulp = ldexp(1.0, ex - 53);
    p.r0 = r0;
    p.half_inv = 0.5 / r0;
    p.margin = r0 * ulp * (1.0 - 9.5367431640625e-07);       /* 1 - 2^-20 */
    p.elim = r0 * r0 * 9.094947017729282e-13;                /* 2^-40 */
    p.enabled = 1;
    return p;
}

DIE_MATH_FN double die_sqrt_near(const die_sqrt_near_t* p, double s) {
    if (p->enabled) {
        const double e = DIE_FMA(-p->r0, p->r0, s);
        const double y = DIE_FMA(e, p->half_inv, p->r0);
        const double d = DIE_FMA(-y, y, s);
        if (fabs(e) < p->elim && fabs(d) < p->margin) return y;
    }
    return sqrt(s);
}

/* sin, cos of x in float32 arithmetic, for DECISIONS WITH A GUARD BAND only: |error| <= DIE_SINCOSF_ERR absolute for
 * |x| <= DIE_SINCOSF_MAX.  Quadrant reduction in float64 (x - k pi/2 with the 33-bit head and the tail of pi/2: exact to
 * ~1e-16 for these arguments), then the classic single-precision minimax polynomials on [-pi/4, pi/4] (Cephes sinf / cosf),
 * ~25 instructions instead of die_sincos' ~60 float64 ones.  Consumers (the forward kernel: the sensed cell, the quick turn
 * decision) must tolerate +-DIE_SINCOSF_ERR and fall back to die_sincos where that could change their result.  Measured
 * against die_sincos on 4e6 arguments (tests/test_portable_math.py): max error 1.3e-7. */
#define DIE_SINCOSF_ERR 4e-7
#define DIE_SINCOSF_MAX 64.0
DIE_MATH_FN void die_sincosf_approx(double x, float* sn_out, float* cs_out) {
    const double ku = DIE_ADD(DIE_MUL(x, DIE_K(TWO_OVER_PI)), DIE_RINT_MAGIC);
    const double kd = DIE_SUB(ku, DIE_RINT_MAGIC);                          /* rint(x 2/pi) */
    const int k = DIE_LO32(ku);
    const float r = (float)DIE_SUB(DIE_FMA(-kd, DIE_K(PIO2_1), x), DIE_MUL(kd, DIE_K(PIO2_T)));
    /* explicit fmaf: the library is compiled without contraction (-fmad=false / -ffp-contract=off), and fmaf is the
     * same correctly rounded operation on the host and on the device */
    const float z = r * r;
    float ps = -1.9515295891e-4f;
    ps = fmaf(ps, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float s = fmaf(r * z, ps, r);
    float pc = 2.443315711809948e-5f;
    pc = fmaf(pc, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    const float c = fmaf(z * z, pc, fmaf(-0.5f, z, 1.0f));
    const float s_sel = (k & 1) ? c : s, c_sel = (k & 1) ? s : c;
    *sn_out = (k & 2) ? -s_sel : s_sel;
    *cs_out = ((k + 1) & 2) ? -c_sel : c_sel;
}

/* sin, cos of x as die_sincos, and ang = atan2(sin, cos) for |x| <= pi.  Mathematically
 * atan2(sin x, cos x) = x; in floating point the rounding errors ds, dc of the two results move
 * the angle by (c ds - s dc) to first order (the second-order term is < 2^-104), and both
 * errors are known here, so ang = x - (c es - s ec) costs six flops instead of a full atan2. */
DIE_MATH_FN void die_sincos_angle(double x, double* sn_out, double* cs_out, double* ang_out) {
    const die_sincos_parts_t o = die_sincos_core(x);
    const int odd = o.k & 1;
    const double s_sel = odd ? o.cs : o.sn, s_err = odd ? o.ec : o.es;
    const double c_sel = odd ? o.sn : o.cs, c_err = odd ? o.es : o.ec;
    const int s_neg = o.k & 2, c_neg = (o.k + 1) & 2;
    const double s = DIE_NEGIF(s_sel, s_neg), es = DIE_NEGIF(s_err, s_neg);   /* sin x = s + es */
    const double c = DIE_NEGIF(c_sel, c_neg), ec = DIE_NEGIF(c_err, c_neg);   /* cos x = c + ec */
    *sn_out = (x == 0.0) ? x : s;
    *cs_out = c;
    *ang_out = DIE_SUB(x, DIE_FMA(c, es, -DIE_MUL(s, ec)));
}

/* ---- atan2 -------------------------------------------------------------------------------------
 * With mn = min(|x|,|y|), mx = max: the ratio interval picks a centre K_j ~ tan(j pi/12) and
 *   atan(mn/mx) = atan(K_j) + atan(t),  t = (mn - K_j mx) / (mx + K_j mn),  |t| <= 0.134
 * (ONE division).  The octant then gives  result = m pi/12 +- (atan(t) + delta_j)  with an
 * integer m in 0..12 and delta_j = atan(K_j) - j pi/12; m pi/12 is formed as an exact hi + lo
 * pair.  Finite arguments, IEEE signed-zero conventions (atan2(+0, -0) = pi ...). */
DIE_MATH_FN int die_atan2_interval(double mn, double mx) {
    return (int)(mn >= DIE_MUL(DIE_K(ATAN_B0), mx)) + (int)(mn >= DIE_MUL(DIE_K(ATAN_B1), mx)) +
           (int)(mn >= DIE_MUL(DIE_K(ATAN_B2), mx));
}

DIE_MATH_FN double die_atan_poly(double z) {
    double p = DIE_K(A6);
    p = DIE_FMA(p, z, DIE_K(A5));
    p = DIE_FMA(p, z, DIE_K(A4));
    p = DIE_FMA(p, z, DIE_K(A3));
    p = DIE_FMA(p, z, DIE_K(A2));
    p = DIE_FMA(p, z, DIE_K(A1));
    return DIE_FMA(p, z, DIE_K(A0));
}

/* octant bookkeeping shared by both variants:
 *   A (!swapped,!neg)  j pi/12 + a       B (swapped,!neg)  (6-j) pi/12 - a
 *   C (!swapped, neg)  (12-j) pi/12 - a  D (swapped, neg)  (6+j) pi/12 + a */
DIE_MATH_FN double die_atan2_m(int swapped, int neg, int j) {
    return (double)(swapped ? (neg ? 6 + j : 6 - j) : (neg ? 12 - j : j));
}

/* <= 1.8 ulp: no compensation. */
DIE_MATH_FN double die_atan2_fast(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const int swapped = ay > ax;
    const double mx = swapped ? ay : ax;
    const double mn = swapped ? ax : ay;
    double at = 0.0;
    int j = 0;
    if (mx != 0.0) {
        j = die_atan2_interval(mn, mx);
        const double c = DIE_KI(DIE_K_ATAN_K0 + j);
        const double t = DIE_DIV(DIE_FMA(-c, mx, mn), DIE_FMA(c, mn, mx));
        const double z = DIE_MUL(t, t);
        at = DIE_ADD(DIE_FMA(DIE_MUL(t, z), die_atan_poly(z), t), DIE_KI(DIE_K_ATAN_D0 + j));
    }
    const int neg = signbit(x) != 0;
    const double md = die_atan2_m(swapped, neg, j);
    const double sat = DIE_NEGIF(at, swapped != neg);
    const double hi = DIE_MUL(md, DIE_K(PI12_HI));
    const double lo = DIE_FMA(md, DIE_K(PI12_LO), DIE_FMA(md, DIE_K(PI12_HI), -hi));
    return copysign(DIE_ADD(hi, DIE_ADD(lo, sat)), y);
}

/* <= 0.502 ulp: num = mn - K mx and den = mx + K mn are carried as double-doubles (mn - fl(K mx)
 * is exact by Sterbenz because B0 >= K1/2), the quotient gets one correction from the exact
 * remainder, and the final sum keeps its rounding error: the only full-size rounding is the
 * last addition. */
DIE_MATH_FN double die_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const int swapped = ay > ax;
    const double mx = swapped ? ay : ax;
    const double mn = swapped ? ax : ay;
    double t = 0.0, tc = 0.0;                                    /* reduced angle = t + tc */
    int j = 0;
    if (mx != 0.0) {
        j = die_atan2_interval(mn, mx);
        const double c = DIE_KI(DIE_K_ATAN_K0 + j);
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
        /* atan(t + tl) = t + [ t z A(z) + tl (1 - z) ] */
        tc = DIE_ADD(DIE_FMA(DIE_MUL(t, z), die_atan_poly(z), DIE_FMA(-z, tl, tl)), DIE_KI(DIE_K_ATAN_D0 + j));
    }
    const int neg = signbit(x) != 0;
    const double md = die_atan2_m(swapped, neg, j);
    const int flip = swapped != neg;
    const double st = DIE_NEGIF(t, flip);
    const double stc = DIE_NEGIF(tc, flip);
    const double hi = DIE_MUL(md, DIE_K(PI12_HI));
    const double lo = DIE_FMA(md, DIE_K(PI12_LO), DIE_FMA(md, DIE_K(PI12_HI), -hi));
    const double h = DIE_ADD(hi, st);                            /* Fast2Sum: |hi| >= |st| or hi == 0 */
    const double e = DIE_ADD(DIE_SUB(hi, h), st);
    return copysign(DIE_ADD(h, DIE_ADD(e, DIE_ADD(lo, stc))), y);
}

#endif /* DIE_MATH_H */
