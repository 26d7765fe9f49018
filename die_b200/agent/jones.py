"""JonesAgent -- the classic three-sensor Physarum particle (Jones 2010, "Characteristics of pattern formation and
evolution in approximations of Physarum transport networks") on the reference's Env / Agent protocol.

SURVEY.md 8(f) rank 4 lists it as "optional ... not parity-checkable": the reference's PhysarumAgent
(core/agent/gradient.py:138-219) senses ONE point ahead and thresholds the angle of the chem gradient there; it has no
three-sensor mode.  The specification is therefore ``oracle/die_ref.py:JonesAgent`` -- written in the reference's own
terms (``polar2xy`` sense offsets core/utils.py:154-164, clamped nearest-cell lookups core/utils.py:39-54,
``renormalize_radians`` core/utils.py:177-179, the unmasked action of core/agent/gradient.py:113-124) -- and
``jones_forward_kernel`` (die_b200/csrc/die_agent_kernels.cuh) repeats it operation for operation: bit-exact.
"""
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..base_types import ActType, ObsType
from .static import _DeviceAgent, _split_obs


class JonesAgent(_DeviceAgent):
    """Sensors FL / F / FR at heading + sense_angle, heading, heading - sense_angle, ``sense_offset`` away, each reading
    chem1 at its nearest cell.  F largest: straight on; F smallest: +-turn_angle by a coin; otherwise turn towards the larger
    of FL / FR; then one step of length ``scale`` along the new heading and a deposit of ``deposit`` x the food under the
    agent.  Every slot acts (ghosts too), as with PhysarumAgent.

    ``rng='philox'``: coins drawn in-kernel, keyed on (seed, call counter, env, slot); ``rng='numpy'``: M coins per env
    from the global legacy numpy RNG (``np.random.randint(0, 2, M)``, as core/agent/gradient.py:181 draws them)."""

    def __init__(self, max_agents: int = 10 ** 6, scale: float = 0.005, deposit: float = 4.0,
                 sense_offset: float = 0.03, turn_angle: float = 45, sense_angle: float = 45,
                 *, rng: str = 'philox', seed: int = 0):
        super().__init__()
        if rng not in ('philox', 'numpy'):
            raise ValueError("rng must be 'philox' or 'numpy'")
        if not (0 <= turn_angle <= 180 and 0 <= sense_angle <= 180):
            raise ValueError("turn_angle and sense_angle are degrees in [0, 180]")
        self._init_params = dict(max_agents=max_agents, scale=scale, deposit=deposit, sense_offset=sense_offset,
                                 turn_angle=turn_angle, sense_angle=sense_angle)
        self._size = int(max_agents)
        self._rng_mode = rng
        self._seed = int(seed)
        self._rng = np.random.default_rng(seed if rng == 'philox' else None)
        self._turn_radians = float(np.radians(turn_angle))
        p = self._p = _lib.DieJonesParams()
        p.scale, p.deposit, p.sense_offset = float(scale), float(deposit), float(sense_offset)
        p.sense_radians, p.turn_radians = float(np.radians(sense_angle)), self._turn_radians
        self._theta = None
        self._coin_host = self._coin_dev = self._coin_event = None
        self._step_dev = None           # device-resident call counter (set by die_b200.graph.GraphedLoop)

    # -- state ---------------------------------------------------------------------------------
    def _lazy_init(self, B: int, M: int, device):
        if self._theta is not None and tuple(self._theta.shape) == (B, M):
            return
        if M != self._size:
            raise ValueError(f"agent was built with max_agents={self._size} but the env has {M} slots")
        # headings on the turn lattice, from normal draws, as PhysarumAgent starts (core/agent/gradient.py:162,165-166)
        prev = self._rng.normal(loc=0., scale=0.4, size=(B, 2, M))
        rads = np.angle(prev[:, 0] + np.multiply(1j, prev[:, 1]))
        tr = self._turn_radians if self._turn_radians > 0 else 1.0
        self.set_state(theta=(rads // tr) * tr, device=device)

    def set_state(self, theta, device=None):
        device = device or (self._theta.device if self._theta is not None else
                            torch.device('cuda', torch.cuda.current_device()))
        t = np.asarray(theta, dtype=np.float64)
        self._theta = torch.from_numpy(np.ascontiguousarray(t.reshape(-1, t.shape[-1]))).to(device)

    def get_state(self):
        th = self._theta.cpu().numpy()
        return (th[0] if th.shape[0] == 1 else th,)

    # -- the policy ----------------------------------------------------------------------------
    def forward(self, obs: ObsType, coin: Optional[np.ndarray] = None) -> ActType:
        """``coin`` ([B,] M in {0, 1}): explicitly injected draws (override ``rng``)."""
        if isinstance(obs[0], np.ndarray):
            return self._forward_host(obs)
        agents, medium, _, B, M = _split_obs(obs)
        self._check(agents)
        self._check_medium(medium)
        H, W = medium.shape[-2:]
        self._lazy_init(B, M, agents.device)
        action = self._action_for(agents)
        coin_ptr = None
        if coin is None and self._rng_mode == 'numpy':
            coin = np.stack([np.random.randint(0, 2, M) for _ in range(B)])
        if coin is not None:
            if self._coin_host is None or tuple(self._coin_host.shape) != (B, M):
                self._coin_host = torch.empty((B, M), dtype=torch.uint8).pin_memory()
                self._coin_dev = torch.empty((B, M), dtype=torch.uint8, device=agents.device)
            if self._coin_event is not None:      # the previous call's copy out of the pinned buffer may still be queued
                self._coin_event.synchronize()
            self._coin_host.numpy()[...] = np.asarray(coin).astype(np.uint8).reshape(B, M)
            self._coin_dev.copy_(self._coin_host, non_blocking=True)
            if self._coin_event is None:
                self._coin_event = torch.cuda.Event()
            self._coin_event.record(torch.cuda.current_stream(agents.device))
            coin_ptr = self._coin_dev.data_ptr()
        on_dev = self._step_dev is not None and coin_ptr is None
        with _lib.on_device(agents.device):
            _lib.check(self._lib.die_jones_forward(
                _lib.C.byref(self._p), H, W, M, B, agents.data_ptr(), medium.data_ptr(),
                int(medium.dtype == torch.float32), self._theta.data_ptr(), action.data_ptr(), coin_ptr,
                self._seed, self._step_dev.data_ptr() if on_dev else self._step, int(on_dev),
                torch.cuda.current_stream().cuda_stream))
        self._step += 1
        torch.autograd.graph.increment_version(action)      # written through its raw pointer: tell torch's version counter
        return action
