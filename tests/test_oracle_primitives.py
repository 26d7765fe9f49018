"""Pins every third-party primitive the oracle leans on against the library the
reference's arithmetic bottoms out in."""
import numpy as np
import pytest
import scipy.ndimage

from oracle import die_ref as R


@pytest.mark.parametrize("n", [2, 5, 6, 256, 4096])
def test_nearest_rule_matches_pandas(n):
    rng = np.random.default_rng(n)
    g = R.grid_coords(n)
    mids = (g[:-1] + g[1:]) / 2
    pts = np.concatenate([
        rng.uniform(-0.1, 1.1, 200_000), g, mids,
        np.nextafter(g, 2), np.nextafter(g, -2), np.nextafter(mids, 2), np.nextafter(mids, -2),
        [-1e-18, -0.0, 0.0, 1.0, 1.0 + 1e-16, -5.0, 5.0, 0.5],
    ])
    assert np.array_equal(R.nearest_index(pts, n), R.nearest_index_pandas(pts, n))


def test_nearest_ties_go_up():
    assert R.nearest_index(np.array([.125, .375, .625, .875]), 5).tolist() == [1, 2, 3, 4]


@pytest.mark.parametrize("sigma", [0.5, 0.8, 1.0, 2.0])
@pytest.mark.parametrize("shape", [(37, 53), (8, 8), (5, 64)])
def test_blur_taps_equal_scipy(sigma, shape):
    rng = np.random.default_rng(1)
    f = rng.random(shape) * (rng.random(shape) < 0.3)
    ref = scipy.ndimage.gaussian_filter(f, sigma=sigma, mode='wrap')
    assert np.array_equal(R.gaussian_blur_taps(f, sigma), ref)
    assert np.array_equal(R.gaussian_blur(f, sigma), ref)


def test_blur_weights_sigma_half():
    w = R.gaussian_kernel1d(0.5)
    assert len(w) == 5
    np.testing.assert_allclose(w, [2.63865083e-4, 1.06450772e-1, 7.86570726e-1, 1.06450772e-1, 2.63865083e-4],
                               rtol=1e-8)


def test_mod1_semantics():
    a = np.array([-1e-18, -0.25, 1.25, 1.0, 0.0, -0.0, 2.0, -1.0])
    m = a % 1.
    assert m[0] == 1.0            # tiny negative wraps to exactly 1.0
    assert m.tolist()[1:] == [0.75, 0.25, 0.0, 0.0, 0.0, 0.0, 0.0]


def test_renormalize_radians_range():
    r = np.linspace(-10, 10, 100001)
    out = R.renormalize_radians(r)
    assert (out > -np.pi - 1e-12).all() and (out <= np.pi).all()
    np.testing.assert_allclose(np.cos(out), np.cos(r), atol=1e-12)
    assert R.renormalize_radians(np.array([np.pi]))[0] == np.pi
    assert R.renormalize_radians(np.array([-np.pi]))[0] == np.pi


def test_angle_signed_zero_quirk():
    """only (-0., -0.) gives +pi."""
    x = np.array([0.0, -0.0, 0.0, -0.0])
    y = np.array([0.0, 0.0, -0.0, -0.0])
    ang = R.get_radians((x, y))
    assert ang.tolist() == [0.0, 0.0, 0.0, np.pi]
    # a clipped gradient keeps the raw signs as signed zeros
    g = np.array([[-3e-7, 3e-7], [-4e-7, -4e-7]])
    clipped = g * (np.hypot(*g) >= 1e-5)
    assert R.get_radians(clipped).tolist() == [np.pi, 0.0]


def test_duplicate_assign_last_writer_wins():
    """SURVEY Q2 (numpy core of core/env.py:211)."""
    a = np.zeros(4)
    ix = np.array([1, 1, 1, 3])
    a[ix] = a[ix] + np.array([10., 20., 30., 5.])
    assert a.tolist() == [0., 30., 0., 5.]


def test_gradient_edges_nonperiodic():
    f = np.arange(20.).reshape(4, 5) ** 2
    g = R.gradient_field(f, normalized=False, grad_clip=None)
    assert g[0, 0, 0] == f[1, 0] - f[0, 0]
    assert g[0, 3, 2] == f[3, 2] - f[2, 2]
    assert g[0, 1, 2] == (f[2, 2] - f[0, 2]) / 2
    assert g[1, 1, 0] == f[1, 1] - f[1, 0]
    assert g[1, 1, 4] == f[1, 4] - f[1, 3]


def test_norm_is_plain_sqrt_of_squares():
    rng = np.random.default_rng(0)
    g = rng.normal(size=(2, 100000)) * 10.0 ** rng.integers(-8, 8, 100000)
    import scipy.linalg
    assert np.array_equal(scipy.linalg.norm(g, axis=0, ord=2), np.sqrt(g[0] * g[0] + g[1] * g[1]))
    assert np.array_equal(np.linalg.norm(g, axis=0), np.sqrt(g[0] * g[0] + g[1] * g[1]))


def test_polar2xy_is_r_cos_sin():
    rng = np.random.default_rng(0)
    th = rng.uniform(-np.pi, np.pi, 100000)
    x, y = R.polar2xy(0.04, th)
    assert np.array_equal(x, 0.04 * np.cos(th)) and np.array_equal(y, 0.04 * np.sin(th))
    x, y = R.polar2xy(1., th)
    assert np.array_equal(x, np.cos(th)) and np.array_equal(y, np.sin(th))


def test_round3_is_rint():
    u = np.random.default_rng(0).random(1_000_000)
    assert np.array_equal(u.round(3), np.rint(u * 1000.0) / 1000.0)


def test_bool_clip():
    m = np.array([True, False])
    assert m.clip(0.1, 1.0).tolist() == [1.0, 0.1]


def test_isclose_forms():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(size=1000) * 1e-7, rng.normal(size=1000), [0.0, 1e-8, -1e-8]])
    assert np.array_equal(np.isclose(0, v, rtol=1e-5), np.abs(v) <= 1e-8 + 1e-5 * np.abs(v))
    atol = np.radians(30) * 0.1
    assert np.array_equal(np.isclose(0, v, rtol=1e-2, atol=atol), np.abs(v) <= atol + 1e-2 * np.abs(v))
