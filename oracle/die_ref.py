"""die_ref -- CPU restatement (oracle) of gkirgizov/die's per-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``die_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and only as the checker / the timed CPU baseline.

PARITY STATUS: the reference cannot be imported as is (xarray, skimage, perlin_noise, gymnasium,
matplotlib, evotorch are absent, there is no network) and its own tests (``test/unit/agent.py``)
never touch this path, so there are no golden vectors shipped by the reference: formally
"parity unpinned".  What pins this file instead:
  1. ``tests/test_golden_oracle.py::test_oracle_equals_reference_executed_live`` runs the
     reference's OWN source files (``/root/reference/core/*.py``, unmodified) over minimal stand-ins
     for the missing packages (``oracle/shims``: xarray.DataArray on numpy + pandas.get_indexer,
     skimage.filters.gaussian on scipy.ndimage) side by side with this restatement: actions,
     rewards, info, medium, agents and headings are IDENTICAL BIT FOR BIT at every step, for
     Brownian / Const / Physarum agents and default / non-default dynamics.
  2. ``tests/golden/*.npz`` are vectors recorded from those runs
     (``oracle/make_golden_from_reference.py``); this oracle and the CUDA path are both checked
     against them (the reference itself does not travel to the GPU box, the vectors do).
  3. ``tests/test_oracle_primitives.py`` checks every third-party primitive bit-for-bit against the
     library the reference's arithmetic bottoms out in (``scipy.ndimage.gaussian_filter`` for
     ``skimage.filters.gaussian``, ``pandas.Index.get_indexer(method='nearest')`` for
     ``xarray .sel(method='nearest')``, numpy for ``%``, ``np.gradient``, ``np.angle``,
     ``np.isclose``, duplicate fancy-assign).
What remains restated rather than executed: xarray's own semantics (vectorised pointwise
indexing, get-add-set through ``.loc`` with repeated cells) -- see ``oracle/shims/README.md``.

Everything is float64 and in *index form*: xarray label lookups are replaced by the
integer cell indices they resolve to.  All citations are ``file:line`` relative to
``/root/reference``.

Layouts (core/base_types.py:31-36, core/data_init.py:94-130):
    medium  [3, H, W]  channels (agents, env_food, chem1); dim 'x' = axis 1, 'y' = axis 2
    agents  [4, M]     channels (x, y, alive, agent_food)
    action  [3, M]     channels (dx, dy, deposit1)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Tuple

import numpy as np
import scipy.linalg
import scipy.ndimage

CH_OCC, CH_FOOD, CH_CHEM = 0, 1, 2            # core/base_types.py:32
AG_X, AG_Y, AG_ALIVE, AG_FOOD = 0, 1, 2, 3    # core/base_types.py:33
AC_DX, AC_DY, AC_DEP = 0, 1, 2                # core/base_types.py:36


# --------------------------------------------------------------------------------------
# Index / angle primitives (core/utils.py)
# --------------------------------------------------------------------------------------

def grid_coords(n: int) -> np.ndarray:
    """Cell-centre coordinates along one axis: ``np.linspace(0, 1, n)``
    (core/data_init.py:99-100, core/utils.py:130)."""
    return np.linspace(0., 1., n)


def nearest_index_pandas(c: np.ndarray, n: int) -> np.ndarray:
    """Literal form of ``field.sel(x=c, method='nearest')`` (core/utils.py:53): xarray
    resolves it with ``pandas.Index.get_indexer(c, method='nearest')`` on the axis
    coordinate index."""
    import pandas as pd
    return pd.Index(grid_coords(n)).get_indexer(np.asarray(c, dtype=np.float64), method='nearest')


def nearest_index(c: np.ndarray, n: int) -> np.ndarray:
    """Closed form of :func:`nearest_index_pandas`, bit-exact against it
    (tests/test_oracle_primitives.py).  With g = linspace(0,1,n): L = last g <= c,
    R = first g >= c; L wins iff (c - g[L]) < (g[R] - c) *strictly* (ties go up);
    out-of-range clamps to 0 / n-1."""
    g = grid_coords(n)
    c = np.asarray(c, dtype=np.float64)
    left = np.searchsorted(g, c, side='right') - 1
    right = np.searchsorted(g, c, side='left')
    lc = np.clip(left, 0, n - 1)
    rc = np.clip(right, 0, n - 1)
    take_left = (np.abs(g[lc] - c) < np.abs(g[rc] - c)) | (right >= n)
    take_left &= left >= 0
    return np.where(take_left, lc, rc).astype(np.int64)


# Math backend.  'numpy' (default) evaluates sin / cos / arctan2 exactly as the reference does
# (numpy -> the host's libm).  'portable' routes ONLY those three functions through the host
# build of die_b200/csrc/die_math.h (oracle/portable_math.c): bit-reproducible routines, the
# same ones the CUDA kernels use, so that free-running trajectories can be compared bit-for-bit
# (the turn rule has exact knife edges, see die_math.h).  Everything else is unchanged.
_MATH = 'numpy'


def set_math_backend(name: str) -> None:
    global _MATH
    if name not in ('numpy', 'portable'):
        raise ValueError(name)
    _MATH = name


def get_math_backend() -> str:
    return _MATH


def polar2xy(r, theta):
    """core/utils.py:154-164 (via complex exp): (r cos(theta), r sin(theta))."""
    if _MATH == 'portable':
        from oracle import portable_math
        s, c = portable_math.sincos(theta)
        # numpy promotes a real r to r + 0j and multiplies complex numbers: (r c - 0 s) + 1j (r s + 0 c).  Only the
        # signs of exact zeros (r == 0) differ from (r c, r s), but those decide angle() = 0 or pi
        return r * c - 0.0 * s, r * s + 0.0 * c
    z = r * np.exp(np.multiply(1j, theta))
    return np.real(z), np.imag(z)


def xy2polar(x, y, fast_angle=False):
    """core/utils.py:158-168.  NB the complex construction turns a (-0., -0.) pair into
    angle +pi and every other all-zero pair into 0.  ``fast_angle`` only selects
    die_atan2_fast in the portable backend (angles that are merely thresholded)."""
    z = x + np.multiply(1j, y)
    if _MATH == 'portable':
        from oracle import portable_math
        return abs(z), portable_math.atan2(np.imag(z), np.real(z), fast=fast_angle)
    return abs(z), np.angle(z)


def get_radians(coords):
    """core/utils.py:171-174."""
    x, y = coords
    return xy2polar(x, y)[1]


def renormalize_radians(rads):
    """core/utils.py:177-179: into (-pi, pi]."""
    return (rads - np.pi) % (-2 * np.pi) + np.pi


def discretize(value, step):
    """core/utils.py:182-183."""
    return (value // step) * step


def get_random(size, a=0., b=1.) -> np.ndarray:
    """core/data_init.py:167-169 -- GLOBAL legacy numpy RNG, quantised to 3 decimals
    before scaling."""
    return (b - a) * np.random.random_sample(size).round(3) + a


def quantised_affine(u: np.ndarray, a: float, b: float) -> np.ndarray:
    """:func:`get_random` on already-drawn uniforms ``u`` (draw injection)."""
    return (b - a) * np.asarray(u, dtype=np.float64).round(3) + a


# --------------------------------------------------------------------------------------
# Field primitives
# --------------------------------------------------------------------------------------

def gaussian_kernel1d(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """Weights scipy builds for ``gaussian_filter1d`` (what skimage.filters.gaussian
    calls, core/env.py:140-143): radius = int(truncate*sigma + .5),
    w = exp(-.5/sigma^2 * k^2) / sum."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return phi / phi.sum()


def gaussian_blur(chem: np.ndarray, sigma: float, mode: str = 'wrap') -> np.ndarray:
    """``skimage.filters.gaussian(chem, sigma, mode, preserve_range=True)`` on a float64
    2-D image == ``scipy.ndimage.gaussian_filter(chem, sigma, mode=mode, truncate=4.0)``
    (core/env.py:140-143)."""
    return scipy.ndimage.gaussian_filter(np.asarray(chem, dtype=np.float64), sigma=sigma, mode=mode,
                                         cval=0, truncate=4.0)


def gaussian_blur_taps(chem: np.ndarray, sigma: float) -> np.ndarray:
    """Explicit periodic form of :func:`gaussian_blur` in scipy's op order (axis 0 then
    axis 1; ``x0*w0 + (x[-r]+x[+r])*w[-r] + ... + (x[-1]+x[+1])*w[-1]``).  This is the
    arithmetic the CUDA kernel restates; tests check it equals scipy bit-for-bit."""
    w = gaussian_kernel1d(sigma)
    r = (len(w) - 1) // 2
    out = np.asarray(chem, dtype=np.float64)
    for axis in (0, 1):
        src = out
        acc = src * w[r]
        for k in range(r, 0, -1):
            acc = acc + (np.roll(src, k, axis=axis) + np.roll(src, -k, axis=axis)) * w[r - k]
        out = acc
    return out


def gradient_field(chem: np.ndarray, normalized: bool = True,
                   grad_clip: Optional[float] = 1e-5) -> np.ndarray:
    """core/agent/gradient.py:55-71 -> [2, H, W] (d/d axis0, d/d axis1), non-periodic."""
    grad = np.stack(np.gradient(chem))
    norm = scipy.linalg.norm(grad, axis=0, ord=2)
    if normalized:
        with np.errstate(divide='ignore', invalid='ignore'):
            grad = np.nan_to_num(np.true_divide(grad, norm))
    if grad_clip is not None:
        grad *= (norm >= grad_clip)
    return grad


def linear_action_cost(action: np.ndarray, weights=(0.02, 0.01)) -> np.ndarray:
    """core/env.py:29-35."""
    dist = np.linalg.norm(action[[AC_DX, AC_DY]], axis=0)
    deposit = np.abs(action[AC_DEP])
    return weights[0] * deposit + weights[1] * dist


def zero_cost(action: np.ndarray) -> np.ndarray:
    """core/env.py:38-39."""
    return np.zeros(action.shape[1:])


# --------------------------------------------------------------------------------------
# State construction (core/data_init.py, core/utils.py:140-151)
# --------------------------------------------------------------------------------------

def gradient_noise(field_size: Tuple[int, int], periods: int = 8, seed: int = 0) -> np.ndarray:
    """Stand-in for ``PerlinNoise(octaves=8)`` sampled on linspace(0,1)^2
    (core/data_init.py:190-196).  The reference's noise is unseeded, hence not
    reproducible even by the reference: this is a distribution-level restatement
    (single-frequency gradient noise, ``periods`` lattice cells across [0,1], quintic
    fade, rounded to 3 dp)."""
    h, w = field_size
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0, 2 * np.pi, size=(periods + 2, periods + 2))
    gx, gy = np.cos(ang), np.sin(ang)
    xs = grid_coords(h) * periods
    ys = grid_coords(w) * periods
    X, Y = np.meshgrid(xs, ys, indexing='ij')
    x0 = np.floor(X).astype(int)
    y0 = np.floor(Y).astype(int)
    fx, fy = X - x0, Y - y0

    def fade(t):
        return t * t * t * (t * (t * 6 - 15) + 10)

    def corner(ix, iy, dx, dy):
        return gx[ix, iy] * dx + gy[ix, iy] * dy

    n00 = corner(x0, y0, fx, fy)
    n10 = corner(x0 + 1, y0, fx - 1, fy)
    n01 = corner(x0, y0 + 1, fx, fy - 1)
    n11 = corner(x0 + 1, y0 + 1, fx - 1, fy - 1)
    u, v = fade(fx), fade(fy)
    nx0 = n00 + u * (n10 - n00)
    nx1 = n01 + u * (n11 - n01)
    return (nx0 + v * (nx1 - nx0)).round(3)


def _mask(sampled, mask_below=0.0, mask_above=1.0):
    """core/data_init.py:181-185."""
    return sampled * ((mask_below <= sampled) & (sampled <= mask_above))


def init_medium(field_size: Tuple[int, int], agent_ratio: float = 0.1,
                food: Optional[np.ndarray] = None, noise_seed: int = 0) -> np.ndarray:
    """core/env.py:74-79: with_const(0.5) -> with_food_perlin(threshold=1) ->
    with_agents(ratio) -> build.  Draw order on the global RNG: one ``random_sample``
    of the field shape (with_agents, core/data_init.py:222-224)."""
    if food is None:
        food = gradient_noise(field_size, 8, noise_seed)
    medium = np.zeros((3, *field_size))
    medium[CH_FOOD] = _mask(food, mask_above=1.0)
    medium[CH_OCC] = np.ceil(_mask(get_random(field_size), mask_above=agent_ratio))
    return medium


def agents_from_medium(medium: np.ndarray, max_agents: Optional[int] = None,
                       food_ratio: float = 1.0) -> np.ndarray:
    """core/data_init.py:132-150 + core/utils.py:140-151: alive agents in row-major
    nonzero order in slots 0..A-1, sitting exactly on grid coordinates; the remaining
    slots are all-zero 'ghosts' at (0, 0)."""
    h, w = medium.shape[-2:]
    ix, iy = (medium[CH_OCC] > 0).nonzero()
    coords = np.stack([grid_coords(h)[ix], grid_coords(w)[iy]])
    n_alive = coords.shape[-1]
    init = np.vstack([coords, np.ones(n_alive), get_random(n_alive, 0.1, food_ratio)])
    if not max_agents:
        max_agents = h * w
    agents = np.zeros((4, max_agents))
    agents[:, :n_alive] = init
    return agents


# --------------------------------------------------------------------------------------
# Time-dependent food fields (core/data_init.py:16-89)
# --------------------------------------------------------------------------------------

def get_meshgrid(field_size) -> np.ndarray:
    """core/utils.py:113-118: dim order reversed; np.meshgrid's default 'xy' indexing, so for a field
    (H, W) both outputs have shape (H, W), [0] varying along axis 1 and [1] along axis 0."""
    xcs = [np.linspace(0., 1., num=size) for size in reversed(field_size)]
    return np.stack(np.meshgrid(*xcs))


class WaveSequence:
    """core/data_init.py:16-51 (FieldSequence) + :71-89 (WaveSequence)."""

    def __init__(self, field_size, dt: float = 0.01, t_bounds=(0, 10)):
        self._size = tuple(field_size)
        self._grid = get_meshgrid(self._size)
        self._ts = np.arange(*t_bounds, dt)

    def __len__(self):
        return len(self._ts)

    def __getitem__(self, t: float) -> np.ndarray:
        pi = np.pi
        x, y = (self._grid - 0.5) * 2
        r = np.linalg.norm((x, y), axis=0)
        rwave = r + np.cos(pi * x) + np.sin(0.4 * pi * y)
        phase = 1 * pi * (rwave + t)
        if _MATH == 'portable':         # the one transcendental the CUDA field kernel evaluates per cell
            from oracle import portable_math
            z_waves = portable_math.sincos(phase)[1].reshape(phase.shape)
        else:
            z_waves = np.cos(phase)
        sx, sy = 3, 3
        z_islands = (np.sin(pi * x * sx + t) + np.cos(pi * y * sy + t))
        mix = 0.25
        return (1 - mix) * z_waves + mix * z_islands

    def get_flow_operator(self, scale: float = 1.0, decay: float = 0.0, k0: int = 0):
        """core/data_init.py:29-38: ``it = iter(self)`` cycles over the time steps, one per call
        (``k0``: position the iterator starts at -- replay aid, 0 in the reference)."""
        state = {'k': int(k0)}

        def food_flow(current):
            t = self._ts[state['k'] % len(self._ts)]
            state['k'] += 1
            return scale * self[t] + (1 - decay) * current

        return food_flow


class FrameSequence:
    """core/data_init.py:16-51 (FieldSequence) for a sequence given by its frames [T, H, W] -- what any subclass
    (PerlinNoiseSequence :54-68, a user's own) amounts to once ``self[t]`` has been evaluated for every time step."""

    def __init__(self, frames: np.ndarray):
        self._frames = np.asarray(frames, dtype=np.float64)

    def __len__(self):
        return self._frames.shape[0]

    def get_flow_operator(self, scale: float = 1.0, decay: float = 0.0, k0: int = 0):
        """core/data_init.py:29-38: scale * next(it) + (1 - decay) * current, the iterator cycling."""
        state = {'k': int(k0)}

        def food_flow(current):
            frame = self._frames[state['k'] % len(self)]
            state['k'] += 1
            return scale * frame + (1 - decay) * current

        return food_flow


# --------------------------------------------------------------------------------------
# Environment (core/env.py)
# --------------------------------------------------------------------------------------

@dataclass
class Dynamics:
    """core/env.py:42-61."""
    op_action_cost: Callable = linear_action_cost
    op_food_flow: Callable = field(default=lambda x: x)
    rate_feed: float = 0.1
    rate_decay_chem: float = 0.1
    boundary: str = 'wrap'           # BoundaryCondition.wrap / .limit (core/env.py:24-26)
    diffuse_mode: str = 'wrap'
    diffuse_sigma: float = .5
    apply_sense_mask: bool = False
    strict_cost: bool = True         # declared, unused (core/env.py:55)
    food_infinite: bool = False
    agents_die: bool = False         # lifecycle is a no-op at defaults (core/env.py:245-261)
    agents_born: bool = False
    init_agent_ratio: float = 0.1


class Env:
    """core/env.py:64-298 in index form."""

    def __init__(self, field_size: Tuple[int, int], dynamics: Optional[Dynamics] = None,
                 medium: Optional[np.ndarray] = None, agents: Optional[np.ndarray] = None,
                 noise_seed: int = 0, use_pandas: bool = False):
        self._field_size = tuple(field_size)
        self.dynamics = dynamics or Dynamics()
        self._nearest = nearest_index_pandas if use_pandas else nearest_index
        if medium is None:
            medium = init_medium(self._field_size, self.dynamics.init_agent_ratio, noise_seed=noise_seed)
        self.medium = np.array(medium, dtype=np.float64)
        if agents is None:
            agents = agents_from_medium(self.medium)
        self.agents = np.array(agents, dtype=np.float64)
        self.last_cells: Optional[Tuple[np.ndarray, np.ndarray]] = None   # (ix, iy) of all M slots after move

    # -- helpers ----------------------------------------------------------------------
    def cells_of(self, coords: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """AgentIndexer.field_by_agents's lookup (core/utils.py:39-54): per-axis nearest."""
        h, w = self._field_size
        return self._nearest(coords[0], h), self._nearest(coords[1], w)

    @property
    def num_alive(self) -> int:
        """core/env.py:263-265."""
        return int(np.count_nonzero(self.agents[AG_ALIVE] > 0))

    @property
    def _get_current_obs(self):
        """core/env.py:296-298: (live agents object, fresh copy of the medium)."""
        return self.agents, self._get_sensed_medium()

    def _get_sensed_medium(self) -> np.ndarray:
        """core/env.py:275-294."""
        if self.dynamics.apply_sense_mask:
            mask = np.ceil(gaussian_blur(self.medium[CH_OCC], 2.0, 'nearest').round(3))
        else:
            mask = np.ones_like(self.medium[CH_OCC])
        return np.where(mask.astype(bool), self.medium, 0.)

    # -- the step ---------------------------------------------------------------------
    def step(self, action: np.ndarray):
        """core/env.py:101-131."""
        self._agent_move(action)
        self._agent_deposit_and_layout(action)
        gained = self._agent_feed(action)
        self._agent_lifecycle()
        self._medium_resource_dynamics()
        self._medium_diffuse_decay()

        num_agents = self.num_alive
        reward = float(np.nansum(gained))               # xarray .sum() skips NaN on floats
        mean_gain = reward / num_agents if num_agents > 0 else 0.
        info = {'num_agents': num_agents,
                'reward': np.round(reward, 3),
                'mean_reward': np.round(mean_gain, 5)}
        return self._get_current_obs, reward, num_agents == 0, False, info

    def _agent_lifecycle(self):
        """core/env.py:245-261.  agents_die: ``agents = agents.where(agent_food > 1e-4, 0)`` -- EVERY channel (x, y, alive,
        agent_food) of a slot without stock becomes 0, for all M slots: an agent that starved dies and is put back at
        (0, 0), and so is every ghost slot, every step (their stock is never above 0).  NaN stock fails the comparison too.
        The reference's own version rebinds ``self.agents`` to the new array while its AgentIndexer keeps the old one
        (core/utils.py:22), so from the next step on it moves one array and looks cells up in another; the semantics
        restated here are those of that code with the indexer following ``self.agents`` -- pinned by running the
        reference with exactly that one-line re-binding (tests/test_golden_oracle.py).  agents_born is unimplemented there."""
        if self.dynamics.agents_die:
            have_food = self.agents[AG_FOOD] > 1e-4
            self.agents[:, ~have_food] = 0.

    def _agent_move(self, action):
        """core/env.py:152-172 -- ALL M slots, no alive mask."""
        new = self.agents[[AG_X, AG_Y]] + action[[AC_DX, AC_DY]]
        if self.dynamics.boundary == 'wrap':
            new = new % 1.
        elif self.dynamics.boundary == 'limit':
            new = new.clip(0., 1.)
        self.agents[[AG_X, AG_Y]] = new

    def _agent_deposit_and_layout(self, action):
        """core/env.py:204-215.  ``medium.loc[cells, 'chem1'] += deposit`` is
        get -> add -> set with repeated cells: the LAST (highest-slot) alive agent on a
        cell wins."""
        alive = (self.agents[AG_ALIVE] > 0).nonzero()[0]
        ix, iy = self.cells_of(self.agents[[AG_X, AG_Y]][:, alive])
        deposit = action[AC_DEP, alive]
        chem = self.medium[CH_CHEM]
        chem[ix, iy] = chem[ix, iy] + deposit
        self.medium[CH_OCC] = 0
        self.medium[CH_OCC][ix, iy] = 1.

    def _agent_feed(self, action):
        """core/env.py:220-243 -- gather for ALL M slots."""
        d = self.dynamics
        consumed_field = d.rate_feed * self.medium[CH_FOOD] * (self.medium[CH_OCC] > 0)
        ix, iy = self.cells_of(self.agents[[AG_X, AG_Y]])
        self.last_cells = (ix, iy)
        consumed = consumed_field[ix, iy]
        if not d.food_infinite:
            self.medium[CH_FOOD] -= consumed_field
        burned = d.op_action_cost(action)
        gained = consumed - burned
        self.agents[AG_FOOD] += gained
        return gained

    def _medium_resource_dynamics(self):
        """core/env.py:147-150."""
        self.medium[CH_FOOD] = self.dynamics.op_food_flow(self.medium[CH_FOOD])

    def _medium_diffuse_decay(self):
        """core/env.py:136-145."""
        d = self.dynamics
        diffused = gaussian_blur(self.medium[CH_CHEM], d.diffuse_sigma, d.diffuse_mode)
        diffused *= (1. - d.rate_decay_chem)
        self.medium[CH_CHEM] = diffused


# --------------------------------------------------------------------------------------
# Agents (core/agent/static.py, core/agent/gradient.py)
# --------------------------------------------------------------------------------------

class ConstAgent:
    """core/agent/static.py:9-28 -- NOT alive-masked."""

    def __init__(self, delta_xy: Tuple[float, float], deposit: float = 0.):
        self._data = (delta_xy[0], delta_xy[1], deposit)

    def forward(self, obs):
        agents, _ = obs
        action = np.zeros((3, agents.shape[-1]))
        for ch, v in enumerate(self._data):
            action[ch] = v
        return action


class BrownianAgent:
    """core/agent/static.py:31-51."""

    def __init__(self, move_scale: float = 0.01, deposit_scale: float = 0.5):
        self._scale = move_scale
        self._dep_scale = deposit_scale

    def forward(self, obs, u: Optional[np.ndarray] = None):
        """``u`` [3, M]: injected uniforms in draw order (dx, dy, deposit1); ``None``
        draws them from the global legacy RNG exactly as core/data_init.py:218-220."""
        agents, _ = obs
        m = agents.shape[-1]
        s = self._scale
        if u is None:
            u = np.stack([np.random.random_sample(m) for _ in range(3)])
        chans = [quantised_affine(u[0], -s, s),
                 quantised_affine(u[1], -s, s),
                 quantised_affine(u[2], 0., self._dep_scale)]
        return np.stack(chans) * agents[AG_ALIVE]          # build_agents, core/data_init.py:248-253


class GradientAgent:
    """core/agent/gradient.py:13-124."""

    def __init__(self, max_agents: int = 10 ** 6, scale: float = 0.01, deposit: float = 4.0,
                 inertia: float = 0.9, sense_offset: float = 0., noise_scale: float = 0.025,
                 normalized_grad: bool = True, grad_clip: Optional[float] = 1e-5,
                 prev_grad: Optional[np.ndarray] = None, use_pandas: bool = False):
        self._size = max_agents
        self._rng = np.random.default_rng()
        self._noise_scale = noise_scale
        self._scale = scale
        self._deposit = deposit
        self._inertia = inertia
        self._sense_offset_scale = sense_offset
        self._normalized = normalized_grad
        self._grad_clip = grad_clip
        self._nearest = nearest_index_pandas if use_pandas else nearest_index
        self._prev_grad = self._get_some_noise() if prev_grad is None else np.array(prev_grad, dtype=np.float64)
        self._direction_rads = get_radians(self._prev_grad)
        self.last_sense_cells = None

    def _get_some_noise(self):
        """core/agent/gradient.py:50-53 (private, unseeded PCG64)."""
        return self._rng.normal(loc=0., scale=0.4, size=(2, self._size))

    def _sense_offset(self):
        """core/agent/gradient.py:73-76."""
        return np.stack(polar2xy(self._sense_offset_scale, self._direction_rads))

    def _process_gradient(self, grad, coin=None):
        return grad

    def _process_momentum(self, grad, noise=None):
        """core/agent/gradient.py:82-91."""
        grad = (1 - self._inertia) * grad + self._inertia * self._prev_grad
        if noise is None:
            noise = self._get_some_noise()
        grad += self._noise_scale * noise
        self._prev_grad = grad
        return grad

    def _process_deposit(self, agents, sensed_food):
        return self._deposit * sensed_food

    def forward(self, obs, coin: Optional[np.ndarray] = None, noise: Optional[np.ndarray] = None):
        """core/agent/gradient.py:96-124.  ``coin`` [M] in {0,1}: injected
        ``np.random.randint(0, 2, M)`` (Physarum); ``noise`` [2, M]: injected N(0,.4)."""
        agents, medium = obs
        h, w = medium.shape[1:]
        chem = medium[CH_CHEM]
        grad_field = gradient_field(chem, self._normalized, self._grad_clip)
        pos = agents[[AG_X, AG_Y]] + self._sense_offset()           # clamped, not wrapped (Q4)
        sx, sy = self._nearest(pos[0], h), self._nearest(pos[1], w)
        self.last_sense_cells = (sx, sy)
        grad = grad_field[:, sx, sy]
        self._fused_angle = None
        grad = self._process_gradient(grad, coin)
        grad = self._process_momentum(grad, noise)
        self._direction_rads = get_radians(grad) if self._fused_angle is None else self._fused_angle

        ix, iy = self._nearest(agents[AG_X], h), self._nearest(agents[AG_Y], w)
        deposit = self._process_deposit(agents, medium[CH_FOOD][ix, iy])

        action = np.zeros((3, agents.shape[-1]))
        action[[AC_DX, AC_DY]] = grad * self._scale
        action[AC_DEP] = deposit
        return action                                                # unmasked (Q8)


class PhysarumAgent(GradientAgent):
    """core/agent/gradient.py:138-219."""

    def __init__(self, max_agents: int = 10 ** 6, scale: float = 0.005, deposit: float = 4.0,
                 inertia: float = 0.0, sense_offset: float = 0.03, noise_scale: float = 0.0,
                 normalized_grad: bool = True, grad_clip: Optional[float] = 1e-5,
                 turn_angle: float = 30, sense_angle: float = 90, turn_tolerance: float = 0.1,
                 prev_grad: Optional[np.ndarray] = None, use_pandas: bool = False):
        super().__init__(max_agents, scale, deposit, inertia, sense_offset, noise_scale,
                         normalized_grad, grad_clip, prev_grad=prev_grad, use_pandas=use_pandas)
        self._turn_radians = np.radians(turn_angle)
        self._sense_radians = np.radians(sense_angle)
        self._rtol = turn_tolerance
        self._direction_rads = discretize(get_radians(self._prev_grad), self._turn_radians)
        self._deposit_mask = 1.

    def _choose_turn(self, drads, coin=None):
        """core/agent/gradient.py:168-193."""
        dir_delta = renormalize_radians(self._direction_rads - drads)
        atol = self._turn_radians * self._rtol
        undetermined_grad = np.isclose(0, drads, rtol=1e-5)
        undetermined_turn = np.isclose(0, dir_delta, rtol=1e-2, atol=atol)
        unseen_grad = abs(dir_delta) > self._sense_radians
        undetermined = undetermined_grad | undetermined_turn | unseen_grad
        if coin is None:
            coin = np.random.randint(0, 2, undetermined.shape)       # global legacy RNG (Q10)
        rand_choice = (np.asarray(coin) - 0.5) * 2
        dir_delta *= np.logical_not(undetermined)
        turn = rand_choice
        turn[dir_delta > atol] = -1
        turn[dir_delta < -atol] = 1
        turn *= self._turn_radians
        self._deposit_mask = np.logical_not(undetermined_grad | undetermined_turn)
        return turn

    def _process_gradient(self, grad, coin=None):
        """_discrete_turn, core/agent/gradient.py:195-208,216-219."""
        dx, dy = grad
        dr, drads = xy2polar(dx, dy, fast_angle=True)
        turn = self._choose_turn(drads, coin)
        directions = renormalize_radians(self._direction_rads + turn)
        dr = 1. if self._normalized else dr
        self._fused_angle = None
        if _MATH == 'portable' and self._normalized and self._inertia == 0. and self._noise_scale == 0.:
            # mirror of the kernel: with an identity momentum step the new heading
            # angle(cos d + i sin d) comes from die_sincos_angle (sin, cos and their angle in one go)
            from oracle import portable_math
            s, c, self._fused_angle = portable_math.sincos_angle(directions)
            return np.stack([dr * c, dr * s])
        return np.stack(polar2xy(dr, directions))

    def _process_deposit(self, agents, sensed_food):
        """core/agent/gradient.py:210-214."""
        mask = np.asarray(self._deposit_mask).clip(0.1, 1.0)
        return self._deposit * sensed_food * mask


class JonesAgent:
    """The classic three-sensor Physarum particle (Jones 2010, "Characteristics of pattern formation and evolution in
    approximations of Physarum transport networks"), on the reference's Env / Agent protocol.

    NOT a restatement: the reference has no such class ("optional ... not parity-checkable").  This
    class IS the specification that die_b200.JonesAgent / jones_forward_kernel are checked against, written with the
    reference's own building blocks so that it composes with its Env exactly as PhysarumAgent does:
      * sense offsets as core/agent/gradient.py:73-76 forms them (polar2xy(r, theta), core/utils.py:154-164), one per sensor:
        FL at heading + SA, F at heading, FR at heading - SA;
      * the field is read at the nearest cell of pos + offset, clamped not wrapped (core/utils.py:39-54);
      * F > FL and F > FR: keep the heading; F < FL and F < FR: turn +-RA by a coin (np.random.randint(0, 2, M), drawn
        for every slot like core/agent/gradient.py:181); FL < FR: turn by -RA (towards FR); FR < FL: by +RA; else keep it;
      * heading' = renormalize_radians(heading + turn) (core/utils.py:177-179);
      * action = (scale cos heading', scale sin heading', deposit * food under the agent), all M slots, not alive-masked
        (core/agent/gradient.py:113-124)."""

    def __init__(self, max_agents: int = 10 ** 6, scale: float = 0.005, deposit: float = 4.0,
                 sense_offset: float = 0.03, turn_angle: float = 45, sense_angle: float = 45,
                 theta0: Optional[np.ndarray] = None):
        self._size = max_agents
        self._scale, self._deposit, self._sense_offset_scale = scale, deposit, sense_offset
        self._turn_radians = np.radians(turn_angle)
        self._sense_radians = np.radians(sense_angle)
        if theta0 is None:
            prev = np.random.default_rng().normal(loc=0., scale=0.4, size=(2, max_agents))
            theta0 = discretize(get_radians(prev), self._turn_radians)
        self._direction_rads = np.array(theta0, dtype=np.float64)

    @staticmethod
    def _sincos(theta):
        if _MATH == 'portable':
            from oracle import portable_math
            return portable_math.sincos(theta)
        return np.sin(theta), np.cos(theta)

    def forward(self, obs, coin: Optional[np.ndarray] = None):
        agents, medium = obs
        h, w = medium.shape[1:]
        chem = medium[CH_CHEM]
        th = self._direction_rads
        sensed = []
        for ang in (th + self._sense_radians, th, th - self._sense_radians):       # FL, F, FR
            s, c = self._sincos(ang)
            sx = nearest_index(agents[AG_X] + self._sense_offset_scale * c, h)
            sy = nearest_index(agents[AG_Y] + self._sense_offset_scale * s, w)
            sensed.append(chem[sx, sy])
        fl, f, fr = sensed
        if coin is None:
            coin = np.random.randint(0, 2, th.shape)
        random_turn = np.where(np.asarray(coin) != 0, self._turn_radians, -self._turn_radians)
        turn = np.zeros_like(th)
        ahead = (f > fl) & (f > fr)
        behind = (f < fl) & (f < fr) & ~ahead
        right = ~ahead & ~behind & (fl < fr)
        left = ~ahead & ~behind & ~right & (fr < fl)
        turn[behind] = random_turn[behind]
        turn[right] = -self._turn_radians
        turn[left] = self._turn_radians
        self._direction_rads = renormalize_radians(th + turn)
        s, c = self._sincos(self._direction_rads)
        ix, iy = nearest_index(agents[AG_X], h), nearest_index(agents[AG_Y], w)
        action = np.zeros((3, agents.shape[-1]))
        action[AC_DX] = c * self._scale
        action[AC_DY] = s * self._scale
        action[AC_DEP] = self._deposit * medium[CH_FOOD][ix, iy]
        return action


# --------------------------------------------------------------------------------------
# Rendering (core/render.py) -- the frames before any matplotlib colour map
# --------------------------------------------------------------------------------------

class NeuralAutomataAgent:
    """core/agent/evo.py:117-209 (+ ConvolutionModel, :45-118) in index form.  The convolutions are torch's own
    (the library the reference's arithmetic bottoms out in here, as numpy / scipy / pandas are elsewhere): circular padding
    by k // 2 on every side, ``conv2d`` without bias per layer, Tanh once at the end, all in float32; then every slot's
    action = model output at the agent's nearest cell (``tensor_by_agents(only_alive=False)``, core/utils.py:56-65) times
    (scale, scale, deposit) as float32 products (``_rescale``, core/agent/evo.py:183-186).  ``weights``: the layers'
    Conv2d weights [cout, cin, k, k]."""

    def __init__(self, weights, scale: float = 0.1, deposit: float = 1.0, with_agent_channel: bool = True):
        import torch
        self._torch = torch
        self.weights = [torch.as_tensor(np.asarray(w), dtype=torch.float32) for w in weights]
        self.with_agent_channel = with_agent_channel
        self.action_coefs = torch.tensor([scale, scale, deposit]).reshape((-1, 1))
        self.sense_output = None

    def model(self, x):
        """ConvolutionModel.forward (core/agent/evo.py:112-118) with p_agent_dropout = 0 on a [1, C, H, W] float32 tensor."""
        F = self._torch.nn.functional
        for w in self.weights:
            p = w.shape[-1] // 2
            x = F.conv2d(F.pad(x, (p, p, p, p), mode='circular') if p else x, w)
        return self._torch.tanh(x)

    def forward(self, obs) -> np.ndarray:
        agents, medium = obs
        torch = self._torch
        med = medium if self.with_agent_channel else medium[1:]
        x = torch.as_tensor(np.ascontiguousarray(med), dtype=torch.float32)[None]      # medium2tensor, :187-199
        sense = self.model(x)
        self.sense_output = sense
        h, w = medium.shape[-2:]
        ix = nearest_index(agents[AG_X], h)
        iy = nearest_index(agents[AG_Y], w)
        per_agent = sense[0][:, torch.as_tensor(ix), torch.as_tensor(iy)]              # [3, M]
        per_agent = per_agent * self.action_coefs
        return per_agent.detach().numpy()                                              # float32, as the reference's action


class EnvRenderer:
    """core/render.py:76-132 + FieldTrace (:9-29) + RendererBase._set_colors (:47-58)."""

    field_colors = {'rgb': None, 'one': [0.19, -0.3, 0.74], 'two': [-0.45, 0.65, 0.83]}

    def __init__(self, field_size, field_colors_id: str = 'rgb', trace_steps: int = 8):
        self.field_size = tuple(field_size)
        self._decay = 1 - 1 / trace_steps
        self._trace = np.zeros(self.field_size)
        color = self.field_colors.get(field_colors_id)
        self._color = None
        if color is not None:
            color = np.array(color, dtype=np.float64)
            color /= np.linalg.norm(color)
            self._color = color

    def render(self, medium: np.ndarray, agents: np.ndarray):
        """-> [medium frame (H, W, 3), trace (H, W) (the reference colour-maps it), agents frame (W, M/W, 4)]."""
        rgb = np.stack([medium[CH_OCC], medium[CH_FOOD], medium[CH_CHEM]], axis=-1)
        if self._color is not None:
            rgb = np.cross(self._color, rgb, axisb=-1)
        self._trace = self._trace * self._decay + medium[CH_OCC]            # FieldTrace.update
        width, height = self.field_size
        data = agents[[AG_ALIVE, AG_FOOD]].reshape((2, height, -1)).transpose((1, 2, 0))
        alive_mask = data[:, :, 0].astype(bool)
        zero = np.zeros(alive_mask.shape)
        img_agents = np.stack([zero, data[:, :, 1], zero, alive_mask], axis=-1)
        return [rgb, self._trace, img_agents]


def run(env: Env, agent, iters: int):
    """The canonical caller, examples/minimal_run.py:14-29 without plotting."""
    total = 0.
    obs = env._get_current_obs
    for _ in range(iters):
        action = agent.forward(obs)
        obs, reward, _, _, _ = env.step(action)
        total += reward
    return total
