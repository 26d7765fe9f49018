"""die_b200/csrc/die_math.h (the kernels' bit-reproducible sin/cos/atan2): accuracy of the host
build against mpmath, IEEE special cases, and (on the GPU) bit-equality of device and host."""
import numpy as np
import pytest

from oracle import portable_math as P


def _ulps(vals, exact):
    import mpmath as mp
    out = []
    for v, e in zip(vals, exact):
        ed = float(e)
        u = np.spacing(abs(ed)) if ed != 0 else 5e-324
        out.append(float(abs(mp.mpf(float(v)) - e) / mp.mpf(u)))
    return np.array(out)


def test_sincos_accuracy_vs_mpmath():
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-np.pi, np.pi, 6000), rng.uniform(-7, 7, 1500),
                         np.arange(-12, 13) * np.radians(30), np.arange(-8, 9) * np.pi / 2,
                         rng.uniform(-1e-3, 1e-3, 500)])
    s, c = P.sincos(xs)
    es = _ulps(s, [mp.sin(mp.mpf(float(x))) for x in xs])
    ec = _ulps(c, [mp.cos(mp.mpf(float(x))) for x in xs])
    assert es.max() < 0.75 and ec.max() < 0.75, (es.max(), ec.max())
    # within 1 ulp of numpy everywhere, identical almost everywhere
    assert np.max(np.abs(s - np.sin(xs)) / np.spacing(np.abs(np.sin(xs)))) <= 1.0
    assert np.mean(s == np.sin(xs)) > 0.97 and np.mean(c == np.cos(xs)) > 0.97


def test_atan2_accuracy_vs_mpmath():
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(1)
    ang = rng.uniform(-np.pi, np.pi, 8000)
    rad = np.exp(rng.uniform(-3, 3, 8000))
    y, x = rad * np.sin(ang), rad * np.cos(ang)
    ex = [mp.atan2(mp.mpf(float(a)), mp.mpf(float(b))) for a, b in zip(y, x)]
    assert _ulps(P.atan2(y, x), ex).max() < 0.52            # compensated: ~correctly rounded
    assert _ulps(P.atan2(y, x, fast=True), ex).max() < 2.0   # thresholds-only variant


def test_atan2_ieee_special_cases():
    y = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 1, 1, -1, -1, 0.0, 1e-300, 3.0, 0.0, -0.0, 2.0, -2.0])
    x = np.array([0.0, 0.0, -0.0, -0.0, 0.0, 0.0, 1, -1, 1, -1, -2.0, 1e-300, -0.0, 5.0, 5.0, -0.0, 0.0])
    ref = np.arctan2(y, x)
    for fast in (False, True):
        got = P.atan2(y, x, fast=fast)
        assert np.array_equal(got, ref)
        assert np.array_equal(np.signbit(got), np.signbit(ref))


def test_sincos_special_cases():
    x = np.array([0.0, -0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi, np.pi / 6])
    s, c = P.sincos(x)
    assert np.array_equal(s, np.sin(x)) and np.array_equal(c, np.cos(x))
    assert np.signbit(s[1]) and not np.signbit(s[0])


def test_round_trip_is_identity_on_the_primary_turn_lattice():
    """theta' = k * 30 deg: angle(cos theta' + i sin theta') must return theta' exactly, as it does
    with numpy/glibc -- this is what keeps Physarum headings on the lattice step after step."""
    tr = np.radians(30)
    th = np.array([k * tr for k in range(-5, 7)])
    s, c = P.sincos(th)
    assert np.array_equal(P.atan2(s, c), th)
    assert np.array_equal(np.arctan2(np.sin(th), np.cos(th)), th)


@pytest.mark.gpu
def test_device_math_equals_host_build_bit_for_bit():
    import torch
    from die_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    n = 1_000_000
    xs = np.concatenate([rng.uniform(-2 * np.pi, 2 * np.pi, n), np.arange(-12, 13) * np.radians(30),
                         [0.0, -0.0, np.pi, -np.pi, np.pi / 2]])
    xd = torch.from_numpy(xs).cuda()
    sd, cd = torch.empty_like(xd), torch.empty_like(xd)
    _lib.check(lib.die_math_sincos(xd.data_ptr(), sd.data_ptr(), cd.data_ptr(), xd.numel(), None))
    torch.cuda.synchronize()
    s, c = P.sincos(xs)
    assert np.array_equal(sd.cpu().numpy(), s) and np.array_equal(cd.cpu().numpy(), c)
    y = np.concatenate([rng.normal(size=n) * np.exp(rng.uniform(-20, 20, n)), s[:100000], [0.0, -0.0, 0.0, -0.0, 1.0]])
    x = np.concatenate([rng.normal(size=n) * np.exp(rng.uniform(-20, 20, n)), c[:100000], [0.0, 0.0, -0.0, -0.0, 0.0]])
    yd, xd2 = torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda()
    od = torch.empty_like(yd)
    for fast in (0, 1):
        _lib.check(lib.die_math_atan2(yd.data_ptr(), xd2.data_ptr(), od.data_ptr(), yd.numel(), fast, None))
        torch.cuda.synchronize()
        ref = P.atan2(y, x, fast=bool(fast))
        got = od.cpu().numpy()
        assert np.array_equal(got, ref) and np.array_equal(np.signbit(got), np.signbit(ref))


def test_float32_sincos_stays_within_its_stated_error():
    """die_sincosf_approx (the forward kernel's guard-banded float32 sin / cos): |error| <= DIE_SINCOSF_ERR = 4e-7 for
    |x| <= 64 -- every consumer budgets exactly that."""
    import ctypes
    from oracle import build_oracle
    lib = ctypes.CDLL(build_oracle.build())
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-np.pi, np.pi, 1000000), rng.uniform(-64, 64, 500000), np.arange(-24, 25) * np.pi / 12,
                        np.arange(-40, 41) * np.pi / 2 + 1e-9, np.linspace(-64, 64, 200001)])
    s, c = np.empty(x.size, np.float32), np.empty(x.size, np.float32)
    sd, cd = np.empty(x.size), np.empty(x.size)
    P = lambda v, t: v.ctypes.data_as(ctypes.POINTER(t))
    lib.die_sincosf_approx_array(P(x, ctypes.c_double), P(s, ctypes.c_float), P(c, ctypes.c_float), ctypes.c_long(x.size))
    lib.die_sincos_array(P(x, ctypes.c_double), P(sd, ctypes.c_double), P(cd, ctypes.c_double), ctypes.c_long(x.size))
    assert np.abs(s - sd).max() <= 2e-7 and np.abs(c - cd).max() <= 2e-7        # half the budget


def test_renormalize_radians_equals_numpy_bit_for_bit():
    """die_renormalize_radians (die_turn.h; the select-light form of round 2) against the reference's own expression
    (core/utils.py:177-179) evaluated by numpy: headings around every multiple of pi / 12 and of 2 pi, +-ulps, +-0, the
    branch bounds at |r - pi| = 2 pi and 4 pi, large arguments (the generic fmod branch)."""
    import ctypes
    from oracle import build_oracle
    lib = ctypes.CDLL(build_oracle.build())
    rng = np.random.default_rng(1)
    base = np.concatenate([np.arange(-96, 97) * np.pi / 12, np.arange(-8, 9) * np.pi, np.array([0.0, -0.0]),
                           np.pi + np.array([-4, -2, 2, 4]) * (2 * np.pi), rng.uniform(-4 * np.pi, 4 * np.pi, 200000),
                           rng.uniform(-1e3, 1e3, 20000), rng.uniform(-1e9, 1e9, 2000)])
    xs = [base]
    for k in (1, 2, 3, 17):
        up, dn = base.copy(), base.copy()
        for _ in range(k):
            up, dn = np.nextafter(up, np.inf), np.nextafter(dn, -np.inf)
        xs += [up, dn]
    turn = np.radians(30)
    xs += [base + turn, base - turn]
    x = np.ascontiguousarray(np.concatenate(xs))
    out = np.empty_like(x)
    P = lambda v: v.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    lib.die_renormalize_radians_array(P(x), P(out), ctypes.c_long(x.size))
    want = (x - np.pi) % (-2 * np.pi) + np.pi
    assert np.array_equal(out.view(np.int64), want.view(np.int64))
