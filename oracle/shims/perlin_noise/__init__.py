"""Stand-in for perlin_noise.PerlinNoise: same call signature (noise(coordinates) -> float),
gradient noise with `octaves` lattice cells per unit and a quintic fade.  Seeded (the reference
constructs it unseeded, so its texture is not reproducible there either)."""
import math

import numpy as np


class PerlinNoise:
    def __init__(self, octaves=1, seed=None):
        self.octaves = float(octaves)
        self.seed = 0 if seed is None else int(seed)
        self._cache = {}

    def _grad(self, cell):
        g = self._cache.get(cell)
        if g is None:
            rng = np.random.default_rng([self.seed, *[int(c) & 0xffffffff for c in cell]])
            v = rng.normal(size=len(cell))
            g = v / np.linalg.norm(v)
            self._cache[cell] = g
        return g

    def __call__(self, coordinates):
        if not isinstance(coordinates, (list, tuple)):
            coordinates = [coordinates]
        x = [c * self.octaves for c in coordinates]
        base = [math.floor(c) for c in x]
        frac = [c - b for c, b in zip(x, base)]
        n = len(x)
        total = 0.0
        for corner in range(1 << n):
            off = [(corner >> k) & 1 for k in range(n)]
            g = self._grad(tuple(b + o for b, o in zip(base, off)))
            w, dot = 1.0, 0.0
            for k in range(n):
                d = frac[k] - off[k]
                t = 1.0 - abs(d)
                w *= t * t * t * (t * (t * 6 - 15) + 10)
                dot += g[k] * d
            total += w * dot
        return total
