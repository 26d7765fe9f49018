#!/usr/bin/env python
"""Generates the constants of die_b200/csrc/die_math.h (portable sincos / atan2) with mpmath:
near-minimax (Chebyshev-fit) polynomial coefficients, the Cody-Waite split of pi/2, and the
hi/lo parts of atan at the reduction breakpoints.  Prints C initialisers (hex floats)."""
import mpmath as mp

mp.mp.prec = 200


def to_double(x):
    return float(mp.mpf(x))


def hexf(x):
    return float(x).hex()


def chebfit(f, a, b, n):
    """coefficients c[0..n] (ascending powers) of a near-minimax degree-n fit of f on [a, b]."""
    coeffs = mp.chebyfit(f, [a, b], n + 1)          # descending powers
    return [c for c in reversed(coeffs)]


def maxerr(f, p, a, b, npts=4001):
    worst = mp.mpf(0)
    for k in range(npts):
        z = a + (b - a) * mp.mpf(k) / (npts - 1)
        val = sum(mp.mpf(to_double(c)) * z ** i for i, c in enumerate(p))
        worst = max(worst, abs(val - f(z)))
    return worst


def main():
    eps = mp.mpf(2) ** -60
    # ---- sin / cos on |r| <= pi/4 (+ slack) ----
    zmax = (mp.pi / 4 * mp.mpf('1.001')) ** 2

    def S(z):
        z = mp.mpf(z)
        if z < eps:
            return -mp.mpf(1) / 6 + z / 120
        r = mp.sqrt(z)
        return (mp.sin(r) - r) / (z * r)

    def Cf(z):
        z = mp.mpf(z)
        if z < eps:
            return mp.mpf(1) / 24 - z / 720
        r = mp.sqrt(z)
        return (mp.cos(r) - 1 + z / 2) / (z * z)

    for name, f, deg in (("SIN", S, 6), ("COS", Cf, 5)):
        p = chebfit(f, 0, zmax, deg)
        print(f"/* {name}: degree {deg} in z = r*r, max abs err of the fit {mp.nstr(maxerr(f, p, 0, zmax), 3)} */")
        print("static const double DIE_%s_C[%d] = {%s};" % (name, deg + 1, ", ".join(hexf(to_double(c)) for c in p)))

    # ---- atan on |t| <= tan(pi/24) (+ slack) ----
    tmax = mp.tan(mp.pi / 24) * mp.mpf('1.02')
    zmax = tmax ** 2

    def A(z):
        z = mp.mpf(z)
        if z < eps:
            return -mp.mpf(1) / 3 + z / 5
        t = mp.sqrt(z)
        return (mp.atan(t) - t) / (z * t)

    for deg in (6,):
        p = chebfit(A, 0, zmax, deg)
        print(f"/* ATAN: degree {deg} in z = t*t, max abs err of the fit {mp.nstr(maxerr(A, p, 0, zmax), 3)} "
              f"(relative to atan(t)/t^3 scale; times t^2 <= {mp.nstr(zmax, 3)} for the relative error) */")
        print("static const double DIE_ATAN_C%d[%d] = {%s};" % (deg, deg + 1, ", ".join(hexf(to_double(c)) for c in p)))

    # ---- Cody-Waite split of pi/2: 33 + 33 + 53 bits ----
    def trunc_bits(x, bits):
        m, e = mp.frexp(x)
        return mp.ldexp(mp.floor(mp.ldexp(m, bits)), e - bits)

    p = mp.pi / 2
    p1 = trunc_bits(p, 33)
    p2 = trunc_bits(p - p1, 33)
    p3 = p - p1 - p2
    print("static const double DIE_PIO2_1 = %s, DIE_PIO2_2 = %s, DIE_PIO2_3 = %s;" %
          (hexf(to_double(p1)), hexf(to_double(p2)), hexf(to_double(p3))))
    print("static const double DIE_2OPI = %s;" % hexf(to_double(2 / mp.pi)))

    # ---- hi/lo of pi, pi/2 and atan at the breakpoints ----
    def hilo(x):
        hi = mp.mpf(to_double(x))
        lo = mp.mpf(to_double(x - hi))
        return hexf(hi), hexf(lo)

    print("static const double DIE_PI_HI = %s, DIE_PI_LO = %s;" % hilo(mp.pi))
    print("static const double DIE_PIO2_HI = %s, DIE_PIO2_LO = %s;" % hilo(mp.pi / 2))
    cs = [to_double(mp.tan(mp.pi / 12)), to_double(mp.tan(mp.pi / 6)), 1.0]
    bs = [to_double(mp.tan(mp.pi / 24)), to_double(mp.tan(3 * mp.pi / 24)), to_double(mp.tan(5 * mp.pi / 24))]
    print("static const double DIE_ATAN_BREAK[3] = {%s};" % ", ".join(hexf(b) for b in bs))
    print("static const double DIE_ATAN_CENTRE[3] = {%s};" % ", ".join(hexf(c) for c in cs))
    his, los = zip(*[hilo(mp.atan(mp.mpf(c))) for c in cs])
    print("static const double DIE_ATAN_HI[3] = {%s};" % ", ".join(his))
    print("static const double DIE_ATAN_LO[3] = {%s};" % ", ".join(los))


if __name__ == "__main__":
    main()
