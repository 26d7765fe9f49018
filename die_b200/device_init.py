"""Initial-state construction ON THE DEVICE.

The reference builds the initial medium with one pure-Python ``PerlinNoise`` call per cell
(core/data_init.py:190-196) -- minutes at 4096^2 -- and compacts the occupied cells into agent slots on
the host (core/data_init.py:132-150, core/utils.py:140-151).  Here the same state is produced by a few
batched tensor operations on the GPU (init-time plumbing, not the per-step path):

    food      = gradient noise (`periods` lattice cells per axis, quintic fade) * [0 <= p <= 1], 3 dp
    occupancy = ceil(u * [0 < round(u, 3) <= ratio])                       core/data_init.py:222-231
    chem1     = 0
    agents    : occupied cells in ROW-MAJOR order -> slots 0..A-1, positioned exactly on the grid
                coordinates np.linspace(0, 1, n)[i]; alive = 1; agent_food = 0.9 * round(u, 3) + 0.1;
                slots A..M-1 all-zero "ghosts" at (0, 0)                   core/data_init.py:132-150

The arithmetic (grid coordinates, rounding, masks, slot order) is the reference's; the random draws
come from torch's device generator instead of numpy's global MT19937, so a seeded run is reproducible
but is not the stream ``np.random.seed`` would give (use the host initialiser, ``Env(init='host')``, for
that -- the Perlin texture is unseeded and irreproducible even in the reference).
"""
from typing import Optional, Tuple

import numpy as np
import torch


def grid_coords(n: int, lo: int, hi: int, device) -> torch.Tensor:
    """np.linspace(0, 1, n)[lo:hi] with numpy's arithmetic: i * (1/(n-1)), last element exactly 1."""
    idx = torch.arange(lo, hi, dtype=torch.float64, device=device)
    g = idx * (1.0 / (n - 1))
    if hi == n:
        g[-1] = 1.0
    return g


def _round3(t: torch.Tensor) -> torch.Tensor:
    """np.round(t, 3) == rint(t * 1000) / 1000."""
    return torch.round(t * 1000.0) / 1000.0


def lattice_angles(periods: int, seed: int, device, batch: Optional[int] = None) -> torch.Tensor:
    """Gradient directions of the noise lattice, [(B,) periods+2, periods+2]; a function of (seed, periods) only."""
    gen = torch.Generator(device='cpu')
    gen.manual_seed(int(seed))
    shape = (periods + 2, periods + 2) if batch is None else (batch, periods + 2, periods + 2)
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * (2 * np.pi)).to(device)


def gradient_noise_rows(H: int, W: int, row_lo: int, row_hi: int, periods: int, ang: torch.Tensor) -> torch.Tensor:
    """Rows [row_lo, row_hi) of the Perlin-style texture of die_b200.data_init.gradient_noise for the lattice
    `ang` ([P+2, P+2], or [B, P+2, P+2] for B textures at once) -> [(B,) rows, W], rounded to 3 dp."""
    device = ang.device
    gx, gy = torch.cos(ang), torch.sin(ang)
    xs = grid_coords(H, row_lo, row_hi, device) * periods
    ys = grid_coords(W, 0, W, device) * periods
    x0, y0 = torch.floor(xs).long(), torch.floor(ys).long()
    fx, fy = (xs - x0)[:, None], (ys - y0)[None, :]
    x0, y0 = x0[:, None], y0[None, :]

    def fade(t):
        return t * t * t * (t * (t * 6. - 15.) + 10.)

    def corner(ix, iy, dx, dy):
        return gx[..., ix, iy] * dx + gy[..., ix, iy] * dy

    n00 = corner(x0, y0, fx, fy)
    n10 = corner(x0 + 1, y0, fx - 1., fy)
    n01 = corner(x0, y0 + 1, fx, fy - 1.)
    n11 = corner(x0 + 1, y0 + 1, fx - 1., fy - 1.)
    u, v = fade(fx), fade(fy)
    nx0 = n00 + u * (n10 - n00)
    nx1 = n01 + u * (n11 - n01)
    return _round3(nx0 + v * (nx1 - nx0))


def device_gradient_noise(H: int, W: int, row_lo: int, row_hi: int, periods: int, seed: int, device) -> torch.Tensor:
    """One texture, rows [row_lo, row_hi) (the slab environment generates its rows rank by rank)."""
    return gradient_noise_rows(H, W, row_lo, row_hi, periods, lattice_angles(periods, seed, device))


def init_medium_device(field_size: Tuple[int, int], agent_ratio: float, seed: int, device,
                       batch: int = 1, periods: Optional[int] = None,
                       gen: Optional[torch.Generator] = None) -> torch.Tensor:
    """core/env.py:74-79 for `batch` independent environments -> float64 [B, 3, H, W] (agents, env_food, chem1)."""
    H, W = int(field_size[0]), int(field_size[1])
    periods = periods or max(8, 8 * H // 256)
    if gen is None:
        gen = torch.Generator(device=device)
        gen.manual_seed(int(seed) * 1000003 + 17)
    medium = torch.zeros((batch, 3, H, W), dtype=torch.float64, device=device)
    ang = lattice_angles(periods, seed, device, batch=batch)
    # bound the temporaries of the noise evaluation (~12 live [b, rows, W] doubles)
    env_chunk = max(1, min(batch, (1 << 26) // (H * W))) if H * W <= (1 << 26) else 1
    row_chunk = max(1, min(H, (1 << 25) // W))
    for b0 in range(0, batch, env_chunk):
        b1 = min(batch, b0 + env_chunk)
        for r0 in range(0, H, row_chunk):
            r1 = min(H, r0 + row_chunk)
            food = gradient_noise_rows(H, W, r0, r1, periods, ang[b0:b1])
            medium[b0:b1, 1, r0:r1] = food * ((food >= 0.0) & (food <= 1.0))          # mask_below=0, mask_above=1
            u = _round3(torch.rand((b1 - b0, r1 - r0, W), generator=gen, dtype=torch.float64, device=device))
            masked = u * ((u >= 0.0) & (u <= agent_ratio))                          # with_agents, :222-224
            medium[b0:b1, 0, r0:r1] = torch.ceil(masked)
            del food, u, masked
    return medium


def agents_from_medium_device(medium: torch.Tensor, seed: int, max_agents: Optional[int] = None,
                              food_ratio: float = 1.0, gen: Optional[torch.Generator] = None) -> torch.Tensor:
    """DataInitializer.agents_from_medium (core/data_init.py:132-150) for [B, 3, H, W] -> float64 [B, 4, M]:
    a row-major stream compaction of the occupied cells into slots 0..A_b-1 of every environment."""
    B, _, H, W = medium.shape
    device = medium.device
    M = int(max_agents) if max_agents else H * W
    if gen is None:
        gen = torch.Generator(device=device)
        gen.manual_seed(int(seed) * 1000003 + 29)
    agents = torch.zeros((B, 4, M), dtype=torch.float64, device=device)
    gx, gy = grid_coords(H, 0, H, device), grid_coords(W, 0, W, device)
    env_chunk = max(1, (1 << 27) // (H * W))
    for b0 in range(0, B, env_chunk):
        b1 = min(B, b0 + env_chunk)
        nz = torch.nonzero(medium[b0:b1, 0] > 0)                       # sorted (env, row, col): row-major per env
        if nz.numel() == 0:
            continue
        counts = torch.bincount(nz[:, 0], minlength=b1 - b0)
        if int(counts.max()) > M:
            raise ValueError(f"{int(counts.max())} occupied cells do not fit into max_agents={M} slots")
        starts = torch.cumsum(counts, 0) - counts
        slot = torch.arange(nz.shape[0], device=device) - starts[nz[:, 0]]
        env = nz[:, 0] + b0
        agents[env, 0, slot] = gx[nz[:, 1]]
        agents[env, 1, slot] = gy[nz[:, 2]]
        agents[env, 2, slot] = 1.0
        u = _round3(torch.rand(nz.shape[0], generator=gen, dtype=torch.float64, device=device))
        agents[env, 3, slot] = (food_ratio - 0.1) * u + 0.1            # get_random(n, 0.1, food_ratio), :141,:167-169
        del nz, slot, env, u
    return agents


def init_state_device(field_size: Tuple[int, int], agent_ratio: float, seed: int, device,
                      batch: int = 1, periods: Optional[int] = None):
    """-> (medium [B, 3, H, W], agents [B, 4, H*W]) on `device`."""
    medium = init_medium_device(field_size, agent_ratio, seed, device, batch=batch, periods=periods)
    return medium, agents_from_medium_device(medium, seed)
