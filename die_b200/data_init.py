"""Initial-state construction on the host (reference: core/data_init.py:92-253,
core/utils.py:140-151, core/env.py:74-86).  Init-time only, not part of the per-step path.

Draw order on the GLOBAL legacy numpy RNG follows the reference, so ``np.random.seed(s)``
before ``Env(...)`` reproduces the same occupancy / agent_food draws:
    with_agents:         random_sample(field_size).round(3)          core/data_init.py:222-224
    agents_from_medium:  0.9 * random_sample(A).round(3) + 0.1       core/data_init.py:141
The food texture is a vectorised single-frequency gradient noise standing in for the
reference's unseeded pure-Python ``PerlinNoise(octaves=8)`` (not reproducible even there).
"""
from typing import Optional, Tuple

import numpy as np


def get_random(size, a=0., b=1.) -> np.ndarray:
    """core/data_init.py:167-169."""
    return (b - a) * np.random.random_sample(size).round(3) + a


def gradient_noise(field_size: Tuple[int, int], periods: int = 8, seed: Optional[int] = None,
                   ang: Optional[np.ndarray] = None) -> np.ndarray:
    """Perlin-style gradient noise on linspace(0,1)^2: `periods` lattice cells per axis,
    quintic fade, rounded to 3 dp (core/data_init.py:190-196).  `ang` ([periods+2, periods+2]) injects the
    lattice's gradient directions (default: drawn from default_rng(seed))."""
    h, w = field_size
    if ang is None:
        rng = np.random.default_rng(seed)
        ang = rng.uniform(0., 2. * np.pi, size=(periods + 2, periods + 2))
    gx, gy = np.cos(ang), np.sin(ang)
    xs = np.linspace(0., 1., h) * periods
    ys = np.linspace(0., 1., w) * periods
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    fx = (xs - x0)[:, None]
    fy = (ys - y0)[None, :]
    x0 = x0[:, None]
    y0 = y0[None, :]

    def fade(t):
        return t * t * t * (t * (t * 6. - 15.) + 10.)

    def corner(ix, iy, dx, dy):
        return gx[ix, iy] * dx + gy[ix, iy] * dy

    n00 = corner(x0, y0, fx, fy)
    n10 = corner(x0 + 1, y0, fx - 1., fy)
    n01 = corner(x0, y0 + 1, fx, fy - 1.)
    n11 = corner(x0 + 1, y0 + 1, fx - 1., fy - 1.)
    u, v = fade(fx), fade(fy)
    nx0 = n00 + u * (n10 - n00)
    nx1 = n01 + u * (n11 - n01)
    return (nx0 + v * (nx1 - nx0)).round(3)


def _mask(sampled, mask_below=0.0, mask_above=1.0):
    """core/data_init.py:181-185."""
    return sampled * ((mask_below <= sampled) & (sampled <= mask_above))


def init_medium(field_size: Tuple[int, int], agent_ratio: float,
                noise_seed: Optional[int] = None, periods: int = 8) -> np.ndarray:
    """core/env.py:74-79 -> float64 [3, H, W] (agents, env_food, chem1)."""
    medium = np.zeros((3, *field_size))
    medium[1] = _mask(gradient_noise(field_size, periods, noise_seed), mask_above=1.0)
    medium[0] = np.ceil(_mask(get_random(field_size), mask_above=agent_ratio))
    return medium


def agents_from_medium(medium: np.ndarray, max_agents: Optional[int] = None,
                       food_ratio: float = 1.0) -> np.ndarray:
    """core/data_init.py:132-150: alive agents in row-major nonzero order in slots 0..A-1,
    exactly on grid coordinates; the remaining slots are all-zero ghosts at (0, 0)."""
    h, w = medium.shape[-2:]
    ix, iy = (medium[0] > 0).nonzero()
    n_alive = ix.shape[0]
    if not max_agents:
        max_agents = h * w
    agents = np.zeros((4, max_agents))
    agents[0, :n_alive] = np.linspace(0., 1., h)[ix]
    agents[1, :n_alive] = np.linspace(0., 1., w)[iy]
    agents[2, :n_alive] = 1.
    agents[3, :n_alive] = get_random(n_alive, 0.1, food_ratio)
    return agents


# --------------------------------------------------------------------------------------------
# time-dependent food fields (core/data_init.py:16-89)
# --------------------------------------------------------------------------------------------
def get_meshgrid(field_size) -> np.ndarray:
    """core/utils.py:113-118 (dim order reversed, as there)."""
    xcs = [np.linspace(0., 1., num=size) for size in reversed(field_size)]
    return np.stack(np.meshgrid(*xcs))


class FoodFlowOperator:
    """What ``FieldSequence.get_flow_operator`` returns (core/data_init.py:29-38):
    ``food -> scale * F_t + (1 - decay) * food`` with t advancing by one entry of the sequence per call and
    cycling.  ``Env`` recognises the object and evaluates it inside the field kernel
    (``die_env_set_food_flow``); calling it on a host array evaluates it with numpy."""

    def __init__(self, sequence: 'FieldSequence', scale: float = 1.0, decay: float = 0.0):
        self.sequence, self.scale, self.decay = sequence, float(scale), float(decay)
        self.calls = 0                       # position of the iterator

    def __call__(self, current):
        t = self.sequence.ts[self.calls % len(self.sequence)]
        self.calls += 1
        return self.scale * self.sequence[t] + (1 - self.decay) * np.asarray(current)


class FieldSequence:
    """core/data_init.py:16-51."""

    def __init__(self, field_size, dt: float = 0.01, t_bounds: Tuple[float, float] = (0, 10)):
        self._size = tuple(int(s) for s in field_size)
        self._tbounds = t_bounds
        self._grid = get_meshgrid(self._size)
        self._ts = np.arange(*t_bounds, dt)

    @property
    def ts(self) -> np.ndarray:
        return self._ts

    def get_flow_operator(self, scale: float = 1.0, decay: float = 0.0) -> FoodFlowOperator:
        return FoodFlowOperator(self, scale, decay)

    def __len__(self):
        return len(self._ts)

    def __contains__(self, t: float):
        t0, t_end = self._tbounds
        return t0 <= t < t_end

    def __getitem__(self, t: float) -> np.ndarray:
        raise NotImplementedError

    def frames(self) -> np.ndarray:
        """``self[t]`` for every time step, float64 [T, H, W] (cached): what the field kernel reads when the sequence
        has no closed form on the device (``die_env_set_food_frames``)."""
        if getattr(self, '_frames', None) is None:
            self._frames = np.ascontiguousarray(np.stack([np.asarray(self[t], dtype=np.float64) for t in self._ts]))
        return self._frames


class TabulatedSequence(FieldSequence):
    """A FieldSequence given by its frames [T, H, W] (time step k <-> ``ts[k]``)."""

    def __init__(self, frames: np.ndarray, dt: float = 0.01):
        frames = np.ascontiguousarray(frames, dtype=np.float64)
        super().__init__(frames.shape[1:], dt, (0, dt * frames.shape[0]))
        self._ts = self._ts[:frames.shape[0]]
        self._frames = frames

    def __getitem__(self, t: float) -> np.ndarray:
        return self._frames[int(np.argmin(np.abs(self._ts - t)))]


class PerlinNoiseSequence(FieldSequence):
    """core/data_init.py:54-68: 3-D Perlin noise sampled on the grid at every time step, rounded to 3 decimals.
    ``noise``: a callable ``noise((x, y, t)) -> float``; default ``perlin_noise.PerlinNoise(octaves)`` as in the
    reference (a third-party, unseeded, pure-Python package -- slow, and absent from this image)."""

    def __init__(self, field_size, dt: float = 0.01, t_bounds: Tuple[float, float] = (0, 1), octaves: int = 8, noise=None):
        super().__init__(field_size, dt, t_bounds)
        if noise is None:
            try:
                from perlin_noise import PerlinNoise
            except ImportError as exc:
                raise ImportError("PerlinNoiseSequence needs the `perlin_noise` package (or pass noise=callable)") from exc
            noise = PerlinNoise(octaves=octaves)
        self._noise = noise
        self._xs = np.linspace(0, 1, self._size[0])
        self._ys = np.linspace(0, 1, self._size[1])

    def __getitem__(self, t: float) -> np.ndarray:
        return np.array([[self._noise((x, y, t)) for y in self._ys] for x in self._xs]).round(3)


class WaveSequence(FieldSequence):
    """core/data_init.py:71-89: running waves + moving islands."""

    def __getitem__(self, t: float) -> np.ndarray:
        pi = np.pi
        x, y = (self._grid - 0.5) * 2
        r = np.linalg.norm((x, y), axis=0)
        rwave = r + np.cos(pi * x) + np.sin(0.4 * pi * y)
        z_waves = np.cos(1 * pi * (rwave + t))
        sx, sy = 3, 3
        z_islands = (np.sin(pi * x * sx + t) + np.cos(pi * y * sy + t))
        mix = 0.25
        return (1 - mix) * z_waves + mix * z_islands

    def device_tables(self):
        """The parts of ``self[t]`` that do not need a per-cell transcendental at run time, tabulated with the
        reference's own numpy expressions: rwave [H, W]; sin(pi x 3 + t_k) per column [T, W]; cos(pi y 3 + t_k)
        per row [T, H]  (x varies along axis 1 and y along axis 0 of the field: np.meshgrid's 'xy' indexing)."""
        pi = np.pi
        x, y = (self._grid - 0.5) * 2
        r = np.linalg.norm((x, y), axis=0)
        rwave = r + np.cos(pi * x) + np.sin(0.4 * pi * y)
        ts = self._ts[:, None]
        col = np.sin(pi * x[0][None, :] * 3 + ts)
        row = np.cos(pi * y[:, 0][None, :] * 3 + ts)
        return np.ascontiguousarray(rwave), np.ascontiguousarray(col), np.ascontiguousarray(row)
