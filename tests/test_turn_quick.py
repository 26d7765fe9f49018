"""die_b200/csrc/die_turn.h on the host: the guard-banded float32 turn decision of the Physarum forward
kernel (die_turn_quick) must agree with the reference arithmetic (die_turn_exact) on every input it
claims to decide -- random, on the turn lattice, axis aligned, at and around every threshold, clipped,
signed zeros, denormals, huge -- and must decide the overwhelming majority of ordinary inputs."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "turn_check.c")
OUT = os.path.join(HERE, "_build", "libturn_check.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                    "-o", OUT, SRC, "-lm"], check=True)
    return ctypes.CDLL(OUT)


def run(lib, gx, gy, th, normalized=1, use_clip=1, clip=1e-5, turn_deg=30., rtol=0.1, sense_deg=90.):
    gx, gy, th = (np.ascontiguousarray(np.broadcast_to(a, np.broadcast(gx, gy, th).shape), dtype=np.float64).ravel()
                  for a in (gx, gy, th))
    n = gx.size
    out = np.zeros((n, 5), dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    enabled = lib.die_turn_check(ctypes.c_long(n), gx.ctypes.data_as(dp), gy.ctypes.data_as(dp), th.ctypes.data_as(dp),
                                 normalized, use_clip, ctypes.c_double(clip),
                                 ctypes.c_double(np.radians(turn_deg) * rtol), ctypes.c_double(np.radians(sense_deg)),
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return enabled, out


def assert_consistent(out):
    dec = out[:, 0] == 1
    bad = dec & ((out[:, 1] != out[:, 3]) | (out[:, 2] != out[:, 4]))
    assert not bad.any(), (int(bad.sum()), out[bad][:5])
    return dec.mean()


def lattice(n, rng, turn_deg=30.):
    k = rng.integers(-6, 7, n)
    return k * np.radians(turn_deg)


def test_random_gradients_and_headings(lib):
    rng = np.random.default_rng(0)
    n = 400_000
    ang = rng.uniform(-np.pi, np.pi, n)
    mag = np.exp(rng.uniform(np.log(1e-7), np.log(10.), n))
    th = np.where(rng.random(n) < 0.5, lattice(n, rng) + rng.normal(0, 1e-15, n), rng.uniform(-np.pi, np.pi, n))
    enabled, out = run(lib, mag * np.cos(ang), mag * np.sin(ang), th)
    assert enabled == 1
    frac = assert_consistent(out)
    assert frac > 0.995, frac           # the exact path is the rare one
    # all three outcomes and both masks occur among the decided ones
    dec = out[out[:, 0] == 1]
    assert set(np.unique(dec[:, 1])) == {-1, 0, 1} and set(np.unique(dec[:, 2])) == {0, 1}


def test_at_and_around_every_threshold(lib):
    """delta = theta - phi swept through +-atol, +-atol/0.99, +-sense, 0, +-pi with offsets from 0 to 1e-3."""
    rng = np.random.default_rng(1)
    atol = np.radians(30.) * 0.1
    marks = np.array([0., atol, -atol, atol / 0.99, -atol / 0.99, np.pi / 2, -np.pi / 2, np.pi, -np.pi])
    offs = np.concatenate([[0.], 10. ** np.arange(-17, -2.5, 0.5)])
    offs = np.concatenate([offs, -offs])
    phi = rng.uniform(-np.pi, np.pi, 300)
    mag = np.exp(rng.uniform(np.log(2e-5), np.log(3.), 300))
    P, Mk, O = np.meshgrid(phi, marks, offs, indexing='ij')
    Mg = np.broadcast_to(mag[:, None, None], P.shape)
    th = P + Mk + O
    th = (th + np.pi) % (2 * np.pi) - np.pi
    enabled, out = run(lib, Mg * np.cos(P), Mg * np.sin(P), th)
    frac = assert_consistent(out)
    assert 0.2 < frac < 0.9             # the near-threshold half must have been deferred, the rest decided


def test_axis_aligned_gradients_on_the_turn_lattice(lib):
    """The structural knife edges: gradient along an axis or a diagonal, heading k*30 degrees (+- ulps)."""
    g = np.array([[1, 0], [-1, 0], [0, 1], [0, -1], [1, 1], [1, -1], [-1, 1], [-1, -1],
                  [1, 1e-9], [1, -1e-9], [1, 1e-8], [1, 1.00001e-8], [1, 0.99999e-8], [1, 3e-8], [1, 5e-8]], dtype=float)
    k = np.arange(-12, 13)
    th0 = k * np.radians(30.)
    th0 = (th0 + np.pi) % (2 * np.pi) - np.pi
    ulps = np.array([0, 1, -1, 2, -2, 8, -8])
    G, T, U = np.meshgrid(np.arange(len(g)), th0, ulps, indexing='ij')
    th = T + U * np.spacing(np.abs(T) + 1e-300)
    for mag in (1e-4, 0.37, 2.5):
        enabled, out = run(lib, g[G, 0] * mag, g[G, 1] * mag, th)
        assert_consistent(out)


def test_clipped_and_zero_gradients(lib):
    rng = np.random.default_rng(2)
    n = 100_000
    ang = rng.uniform(-np.pi, np.pi, n)
    mag = np.exp(rng.uniform(np.log(1e-320), np.log(2e-5), n))         # denormal .. just above the clip
    th = lattice(n, rng)
    enabled, out = run(lib, mag * np.cos(ang), mag * np.sin(ang), th)
    frac = assert_consistent(out)
    assert frac > 0.95
    # signed zeros and one-zero components, every sign combination
    z = np.array([0.0, -0.0, 1e-7, -1e-7, 1e-200, -1e-200, 5e-324, -5e-324, 1e-5, -1e-5, 0.99999e-5, 1.00001e-5])
    GX, GY, TH = np.meshgrid(z, z, np.arange(-6, 7) * np.radians(30.), indexing='ij')
    enabled, out = run(lib, GX, GY, TH)
    assert_consistent(out)
    # the all-zero gradient (empty chem field: the bulk of the ghost slots) must be decided, as "coin, masked"
    enabled, out = run(lib, np.zeros(64), np.zeros(64), lattice(64, rng))
    assert (out[:, 0] == 1).all() and (out[:, 1] == 0).all() and (out[:, 2] == 0).all()


def test_out_of_range_inputs_defer(lib):
    big = np.array([1e20, 1e38, 1e200, np.inf, np.nan, -1e300])
    GX, GY = np.meshgrid(big, np.concatenate([big, [1.0, 0.0]]), indexing='ij')
    enabled, out = run(lib, GX, GY, 0.5)
    assert_consistent(out)
    assert (out[:, 0] == 0).all()


@pytest.mark.parametrize("kw", [dict(normalized=0), dict(use_clip=0), dict(clip=0.0), dict(sense_deg=180.),
                                dict(sense_deg=2.), dict(rtol=0.0), dict(turn_deg=0.01)])
def test_plan_disabled_outside_its_case_analysis(lib, kw):
    enabled, out = run(lib, np.array([0.3]), np.array([-0.2]), np.array([1.0]), **kw)
    assert enabled == 0 and out[0, 0] == 0


@pytest.mark.parametrize("kw", [dict(turn_deg=45., rtol=0.2, sense_deg=60.), dict(turn_deg=15., rtol=0.3, sense_deg=120.),
                                dict(clip=1e-3), dict(clip=1e-9, sense_deg=170.)])
def test_other_enabled_parameter_sets(lib, kw):
    rng = np.random.default_rng(3)
    n = 200_000
    ang = rng.uniform(-np.pi, np.pi, n)
    mag = np.exp(rng.uniform(np.log(1e-11), np.log(10.), n))
    th = rng.uniform(-np.pi, np.pi, n)
    enabled, out = run(lib, mag * np.cos(ang), mag * np.sin(ang), th, **kw)
    assert enabled == 1
    assert assert_consistent(out) > 0.98


def test_dense_sweep_across_the_guard_band_edges(lib):
    """delta swept densely (2000 points per threshold and side) from well inside to well outside every guard band,
    for gradient magnitudes from just above the clip threshold to O(1): whatever the quick path decides must be the
    exact path's answer, and it must start deciding again right outside the bands."""
    rng = np.random.default_rng(7)
    atol = np.radians(30.) * 0.1
    marks = np.array([atol / 0.99, -atol / 0.99, np.pi / 2, -np.pi / 2, atol, -atol])
    offs = np.concatenate([np.linspace(-3e-3, 3e-3, 2001), np.geomspace(1e-9, 3e-3, 400), -np.geomspace(1e-9, 3e-3, 400)])
    for mag in (1.0002e-5, 1.5e-5, 3e-4, 0.02, 0.9, 7.0):
        phi = rng.uniform(-np.pi, np.pi, 12)
        P, Mk, O = np.meshgrid(phi, marks, offs, indexing='ij')
        th = (P + Mk + O + np.pi) % (2 * np.pi) - np.pi
        enabled, out = run(lib, mag * np.cos(P), mag * np.sin(P), th)
        assert enabled == 1
        assert_consistent(out)
        far = np.abs(O.ravel()) > 2.5e-3                    # beyond the widest band (3.8e-4 rad) with margin
        banded = np.isin(Mk.ravel(), marks[:4])             # +-atol itself is no threshold of the quick path
        near = (np.abs(O.ravel()) < 1e-6) & banded
        assert out[far, 0].mean() > 0.99 and out[near, 0].mean() < 0.01


def test_magnitudes_around_the_clip_threshold(lib):
    rng = np.random.default_rng(8)
    n = 200_000
    ang = rng.uniform(-np.pi, np.pi, n)
    mag = 1e-5 * (1 + rng.uniform(-3e-4, 3e-4, n))          # |g| within 0.03 % of grad_clip, both sides
    mag[:64] = 1e-5 * (1 + np.linspace(-1e-15, 1e-15, 64))  # ... and within a few ulp of it
    th = np.where(rng.random(n) < 0.5, lattice(n, rng), rng.uniform(-np.pi, np.pi, n))
    enabled, out = run(lib, mag * np.cos(ang), mag * np.sin(ang), th)
    frac = assert_consistent(out)
    assert 0.2 < frac < 0.95                                 # the band around the threshold defers, the rest decides


# ------------------------------------------------------------------------------------------
# die_sqrt_near (die_math.h): the cost hint's square root -- one Newton step from |scale|, accepted only where it is
# provably the correctly rounded root
# ------------------------------------------------------------------------------------------
def _sqrt_near(lib, s, r0):
    s = np.ascontiguousarray(s, dtype=np.float64).ravel()
    y = np.empty_like(s)
    fast = np.zeros(s.size, dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    enabled = lib.die_sqrt_near_check(ctypes.c_long(s.size), s.ctypes.data_as(dp), ctypes.c_double(r0), y.ctypes.data_as(dp),
                                      fast.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return enabled, y, fast.astype(bool)


@pytest.mark.parametrize("scale", [0.007, 0.005, 0.0075, 0.01, 0.02, -0.03, 0.3, 123.456, 1e-9])
def test_sqrt_near_is_the_ieee_square_root(lib, scale):
    """Every result equals np.sqrt bit for bit -- on the kernel's own inputs (dx*dx + dy*dy of scale * (cos, sin)), on every
    double within +-2000 ulps of scale^2, on perturbed and on unrelated arguments -- and the Newton step is the common path."""
    rng = np.random.default_rng(1)
    th = np.concatenate([rng.uniform(-np.pi, np.pi, 400_000), np.arange(-12, 13) * np.radians(30.), np.arange(-8, 9) * np.radians(45.)])
    dx, dy = np.cos(th) * scale, np.sin(th) * scale
    real = dx * dx + dy * dy
    enabled, y, fast = _sqrt_near(lib, real, scale)
    assert enabled == 1
    assert np.array_equal(y, np.sqrt(real)) and fast.mean() > 0.9999
    r2 = scale * scale
    near = r2 + np.arange(-2000, 2001) * np.spacing(r2)
    _, y, fast = _sqrt_near(lib, near, scale)
    assert np.array_equal(y, np.sqrt(near)) and fast.mean() > 0.99
    wide = r2 * (1.0 + rng.uniform(-1e-9, 1e-9, 200_000))              # straddles the 2^-40 acceptance window
    _, y, fast = _sqrt_near(lib, wide, scale)
    assert np.array_equal(y, np.sqrt(wide)) and 0 < fast.mean() < 1
    other = np.concatenate([np.exp(rng.uniform(-40, 40, 100_000)), [0.0, r2 * 4, r2 / 4, np.inf, 5e-324, 1e308]])
    _, y, fast = _sqrt_near(lib, other, scale)
    assert np.array_equal(y, np.sqrt(other)) and not fast.any()


def test_sqrt_near_refuses_what_it_cannot_prove(lib):
    """An r0 next to a power of two (the ulp changes inside the acceptance window), zero, a non-finite or absurd one: the plan
    is disabled and sqrt() answers; adversarial arguments whose root sits next to a rounding boundary take sqrt() too."""
    s = np.random.default_rng(2).uniform(0.1, 4.0, 1000)
    for r0 in (1.0, 0.5, 2.0 ** -7, 1.0 + 1e-12, 1.0 - 1e-12, 0.0, np.inf, np.nan, 1e-200, 1e200):
        enabled, y, fast = _sqrt_near(lib, s, r0)
        assert enabled == 0 and not fast.any() and np.array_equal(y, np.sqrt(s))
    # roots half an ulp off a double: s = (r + u/2)^2 rounded -- the check must not accept a neighbour of the true rounding
    scale = 0.007
    r = scale + np.arange(-50, 51) * np.spacing(scale)
    u = np.spacing(scale)
    mid = np.array([float((np.longdouble(a) + np.longdouble(u) / 2) ** 2) for a in r])
    mid = np.concatenate([mid, np.nextafter(mid, 0), np.nextafter(mid, 1)])
    _, y, fast = _sqrt_near(lib, mid, scale)
    assert np.array_equal(y, np.sqrt(mid))
