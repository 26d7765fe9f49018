"""Host replica (numpy) of the in-kernel random draws -- a DIAGNOSTIC, not a compute path.

``rng='philox'`` agents draw inside their kernels: Philox4x32-10 keyed on (seed, call counter, slot)
(die_b200/csrc/die_device.cuh: philox4x32_10, philox_draw).  The functions here reproduce those draws bit for bit on
the host, so a free run with in-kernel randomness can be replayed by anything that accepts injected draws -- the
oracle in the parity tests (tests/test_gpu_philox_replay.py), or the reference itself:

    coin = physarum_coins(seed, step, B, M)          # what PhysarumAgent.forward's kernel used at call `step`
    u    = brownian_uniforms(seed, step, B, M)       # what BrownianAgent.forward's kernel used at call `step`

Reference sites these draws replace: core/agent/gradient.py:181 (np.random.randint(0, 2, M)),
core/data_init.py:167-169 (np.random.random_sample(M)).
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)

FWD_ITEMS = 8            # kFwdItems: slots per thread of gradient_forward_kernel
AGENT_THREADS = 256      # kAgentThreads


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Ten rounds of Philox4x32 on arrays of 32-bit counters (uint64 arrays holding 32-bit values)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _MASK, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_draw(seed: int, step: int, slot, stream: int):
    """philox_draw(seed, step, slot, stream) of die_device.cuh for an array of slots -> four uint64 arrays (32-bit values)."""
    slot = np.asarray(slot, dtype=np.uint64)
    seed, step = int(seed) & 0xFFFFFFFFFFFFFFFF, int(step) & 0xFFFFFFFFFFFFFFFF
    c0 = slot & _MASK
    c1 = slot >> np.uint64(32)
    c2 = np.full(slot.shape, step & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(slot.shape, ((step >> 32) ^ (int(stream) << 24)) & 0xFFFFFFFF, dtype=np.uint64)
    return philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, seed >> 32)


def physarum_coins(seed: int, step: int, B: int, M: int, b0: int = 0) -> np.ndarray:
    """The coin of every slot as gradient_forward_kernel draws it at call ``step``: a CTA owns 2048 consecutive slots
    of one environment, thread t its slots t + 256 k, and coin k is bit k of ONE 32-bit Philox word per thread
    (die_agent_kernels.cuh: coin_bits).  -> uint8 [B, M] in {0, 1}."""
    nchunk = (M + AGENT_THREADS * FWD_ITEMS - 1) // (AGENT_THREADS * FWD_ITEMS)
    i = np.arange(M, dtype=np.int64)
    chunk, rest = i // (AGENT_THREADS * FWD_ITEMS), i % (AGENT_THREADS * FWD_ITEMS)
    k, t = rest // AGENT_THREADS, rest % AGENT_THREADS
    out = np.empty((B, M), dtype=np.uint8)
    for b in range(B):
        block = (b + b0) * nchunk + chunk
        word = philox_draw(seed, step, (block * AGENT_THREADS + t).astype(np.uint64), 2)[0]
        out[b] = ((word >> k.astype(np.uint64)) & np.uint64(1)).astype(np.uint8)
    return out


def brownian_uniforms(seed: int, step: int, B: int, M: int) -> np.ndarray:
    """The three uniforms (dx, dy, deposit1 order) of every slot as brownian_forward_kernel draws them at call ``step``:
    32 random bits each, u = w / 2^32.  -> float64 [B, 3, M]."""
    slot = (np.arange(B, dtype=np.uint64)[:, None] * np.uint64(M) + np.arange(M, dtype=np.uint64)[None, :])
    x, y, z, _ = philox_draw(seed, step, slot, 0)
    return np.stack([x, y, z], axis=1).astype(np.float64) * (1.0 / 4294967296.0)
