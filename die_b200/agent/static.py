"""ConstAgent / BrownianAgent -- drop-ins for core/agent/static.py:9-51."""
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _hints
from .. import _lib
from ..base_types import ActType, ObsType
from .base import Agent


def _split_obs(obs: ObsType):
    agents, medium = obs
    batched = agents.dim() == 3
    B = agents.shape[0] if batched else 1
    M = agents.shape[-1]
    return agents, medium, batched, B, M


class _DeviceAgent(Agent):
    """Shared plumbing: action buffer ownership and the host-buffer path."""

    def __init__(self):
        self._lib = _lib.load()
        self._action = None
        self._step = 0
        self._host = None

    def _action_for(self, agents: torch.Tensor) -> torch.Tensor:
        """The action buffer is owned by the agent and re-used every call (the env only
        borrows it for the duration of ``step``; the reference allocates a fresh array,
        core/data_init.py:152-157)."""
        shape = (*agents.shape[:-2], 3, agents.shape[-1])
        if self._action is None or tuple(self._action.shape) != shape or self._action.device != agents.device:
            self._action = torch.empty(shape, dtype=torch.float64, device=agents.device)
        return self._action

    @staticmethod
    def _check(agents: torch.Tensor):
        if not (isinstance(agents, torch.Tensor) and agents.is_cuda and agents.dtype == torch.float64
                and agents.is_contiguous()):
            raise TypeError("obs must hold contiguous float64 CUDA tensors (or numpy arrays for the host path); "
                            "die_b200 has no CPU fallback")

    @staticmethod
    def _check_medium(medium: torch.Tensor):
        """The medium may be float32 (an Env in its float32 field mode); agents, actions and headings never are."""
        if not (isinstance(medium, torch.Tensor) and medium.is_cuda and medium.dtype in (torch.float64, torch.float32)
                and medium.is_contiguous()):
            raise TypeError("obs[1] must be a contiguous float64 (or float32) CUDA tensor; die_b200 has no CPU fallback")

    # -- host-buffer path: numpy obs in, numpy action out ------------------------------------
    def _forward_host(self, obs) -> np.ndarray:
        agents_np, medium_np = obs
        dev = torch.device('cuda', torch.cuda.current_device())
        if self._host is None or self._host['agents'].shape != agents_np.shape \
                or self._host['medium'].shape != medium_np.shape:
            self._host = {
                'agents': torch.empty(agents_np.shape, dtype=torch.float64, device=dev),
                'medium': torch.empty(medium_np.shape, dtype=torch.float64, device=dev),
                'action': torch.empty((*agents_np.shape[:-2], 3, agents_np.shape[-1]),
                                      dtype=torch.float64, pin_memory=True),
            }
        hb = self._host
        hb['agents'].copy_(torch.from_numpy(agents_np), non_blocking=True)
        hb['medium'].copy_(torch.from_numpy(medium_np), non_blocking=True)
        action = self.forward((hb['agents'], hb['medium']))
        hb['action'].copy_(action, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = hb['action'].numpy()
        out.flags.writeable = False            # Env.step may then use the device copy instead of uploading it again
        _hints.register_host_action(out, action)
        return out


class ConstAgent(_DeviceAgent):
    """core/agent/static.py:9-28 -- writes the constant vector to every slot (not alive-masked)."""

    def __init__(self, delta_xy: Tuple[float, float], deposit: float = 0.):
        super().__init__()
        self._init_params = {'delta_xy': tuple(delta_xy), 'deposit': deposit}
        self._data = (float(delta_xy[0]), float(delta_xy[1]), float(deposit))

    def forward(self, obs: ObsType) -> ActType:
        if isinstance(obs[0], np.ndarray):
            return self._forward_host(obs)
        agents, _, _, B, M = _split_obs(obs)
        self._check(agents)
        action = self._action_for(agents)
        with _lib.on_device(agents.device):
            _lib.check(self._lib.die_const_forward(action.data_ptr(), M, B, *self._data,
                                                   torch.cuda.current_stream().cuda_stream))
        torch.autograd.graph.increment_version(action)      # written through its raw pointer: tell torch's version counter
        return action


class BrownianAgent(_DeviceAgent):
    """core/agent/static.py:31-51.

    ``rng='philox'`` (default): uniforms are drawn in-kernel (Philox4x32-10 keyed on
    ``seed``, call counter, slot).  ``rng='numpy'`` (validation mode): the three M-vectors are
    drawn on the host from the GLOBAL legacy numpy RNG in the reference's order (dx, dy,
    deposit1; core/data_init.py:218-220) and uploaded, so ``np.random.seed(s)`` reproduces
    the reference's stream bit-for-bit.
    """

    def __init__(self, move_scale: float = 0.01, deposit_scale: float = 0.5,
                 *, rng: str = 'philox', seed: int = 0):
        super().__init__()
        self._init_params = {'move_scale': move_scale, 'deposit_scale': deposit_scale}
        self._scale = float(move_scale)
        self._dep_scale = float(deposit_scale)
        if rng not in ('philox', 'numpy'):
            raise ValueError("rng must be 'philox' or 'numpy'")
        self._rng = rng
        self._seed = int(seed)
        self._u_host = None
        self._u_dev = None
        self._u_event = None
        self._step_dev = None           # device-resident call counter (set by die_b200.graph.GraphedLoop)

    def forward(self, obs: ObsType, u: Optional[np.ndarray] = None) -> ActType:
        """``u`` ([B,] 3, M): explicitly injected uniforms (overrides ``rng``)."""
        if isinstance(obs[0], np.ndarray):
            return self._forward_host(obs)
        agents, _, _, B, M = _split_obs(obs)
        self._check(agents)
        action = self._action_for(agents)
        u_ptr = None
        if u is None and self._rng == 'numpy':
            u = np.stack([np.stack([np.random.random_sample(M) for _ in range(3)]) for _ in range(B)])
        if u is not None:
            if self._u_host is None or self._u_host.shape != (B, 3, M):
                self._u_host = torch.empty((B, 3, M), dtype=torch.float64).pin_memory()
                self._u_dev = torch.empty((B, 3, M), dtype=torch.float64, device=agents.device)
            # the previous call's asynchronous copy out of the pinned buffer may still be queued: wait for it first
            if self._u_event is not None:
                self._u_event.synchronize()
            self._u_host.numpy()[...] = np.asarray(u, dtype=np.float64).reshape(B, 3, M)
            self._u_dev.copy_(self._u_host, non_blocking=True)
            if self._u_event is None:
                self._u_event = torch.cuda.Event()
            self._u_event.record(torch.cuda.current_stream(agents.device))
            u_ptr = self._u_dev.data_ptr()
        with _lib.on_device(agents.device):
            if self._step_dev is not None and u_ptr is None:
                # (die_b200/graph.py) the call counter lives on the device, so a captured launch can be replayed
                _lib.check(self._lib.die_brownian_forward_dev(
                    agents.data_ptr(), action.data_ptr(), M, B, self._scale, self._dep_scale,
                    self._seed, self._step_dev.data_ptr(), torch.cuda.current_stream().cuda_stream))
            else:
                _lib.check(self._lib.die_brownian_forward(
                    agents.data_ptr(), action.data_ptr(), M, B, self._scale, self._dep_scale,
                    u_ptr, self._seed, self._step, torch.cuda.current_stream().cuda_stream))
        self._step += 1
        torch.autograd.graph.increment_version(action)      # written through its raw pointer: tell torch's version counter
        return action
