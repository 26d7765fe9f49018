"""Env / Dynamics -- drop-in for the reference's ``core/env.py`` on the per-step path.

Same class names, constructor, ``step`` signature and return tuple as the reference
(core/env.py:42-131).  State lives in HBM as float64 torch tensors in the reference's
channel-major layout and every ``step`` is four kernel launches of ``libdie_sm100a.so``
through its C ABI (``include/die_b200.h``): move+claim, the fused field pass (deposit, occupancy,
food, diffusion*decay), agent feed+reduce, and the stats finalisation.  There is no CPU fallback.

Differences a caller of the reference can observe (all documented in DESIGN.md):
  * obs / action are ``torch`` CUDA tensors (or pinned ``numpy`` arrays on the host-buffer
    path) instead of ``xarray.DataArray``; ``die_b200.base_types.channel`` replaces
    ``.sel(channel=...)``.
  * ``obs[1]`` (the medium) is a zero-copy view of the env's current buffer, valid until the
    next-but-one ``step`` (two buffers ping-pong; the reference makes a fresh copy,
    core/env.py:290-294); with ``apply_sense_mask`` it is the masked copy the reference hands out.
    ``obs[0]`` is the env's live ``agents`` tensor exactly as in the reference (core/env.py:298).
  * ``batch=B`` runs B independent environments in one set of launches (leading axis B).
  * ``init='device'`` builds the initial state on the GPU; ``render()`` returns device tensors
    (``render(host=True)``: numpy arrays) instead of drawing with matplotlib.
Also on the path: ``Dynamics.op_food_flow = WaveSequence(...).get_flow_operator(...)`` (evaluated in
the field kernel), every ``diffuse_mode`` of scipy.ndimage, ``boundary=limit``, ``food_infinite``.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from enum import Enum
from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch

from . import _hints
from . import _lib
from . import data_init
from .base_types import ActType, ObsType


class BoundaryCondition(Enum):
    """core/env.py:24-26."""
    wrap = 'wrap'
    limit = 'limit'


def linear_action_cost(action, weights=(0.02, 0.01)):
    """core/env.py:29-35.  Inside ``Env.step`` this is evaluated in the feed kernel; the
    Python body exists so callers can still evaluate the operator on an action tensor."""
    act = torch.as_tensor(action)
    dist = torch.sqrt(act[..., 0, :] * act[..., 0, :] + act[..., 1, :] * act[..., 1, :])
    return weights[0] * act[..., 2, :].abs() + weights[1] * dist


def zero_cost(action):
    """core/env.py:38-39."""
    act = torch.as_tensor(action)
    return torch.zeros_like(act[..., 0, :])


def identity_food_flow(food):
    """Default ``op_food_flow`` (core/env.py:45)."""
    return food


linear_action_cost._die_weights = (0.02, 0.01)
zero_cost._die_weights = (0.0, 0.0)


@dataclass
class Dynamics:
    """core/env.py:42-61 (same fields, same defaults)."""
    op_action_cost: Callable = linear_action_cost
    op_food_flow: Callable = identity_food_flow
    rate_feed: float = 0.1
    rate_decay_chem: float = 0.1
    boundary: BoundaryCondition = BoundaryCondition.wrap
    diffuse_mode: str = 'wrap'
    diffuse_sigma: float = .5

    apply_sense_mask: bool = False
    strict_cost: bool = True        # declared but unused by the reference (core/env.py:55)
    food_infinite: bool = False
    agents_die: bool = False
    agents_born: bool = False       # unimplemented in the reference (core/env.py:256-261)

    init_agent_ratio: float = 0.1


def gaussian_kernel1d(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """The weights ``scipy.ndimage.gaussian_filter1d`` builds (what ``skimage.filters.gaussian``
    uses, core/env.py:140-143): radius = int(truncate * sigma + .5)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _is_wave_flow(op) -> bool:
    return isinstance(op, data_init.FoodFlowOperator) and isinstance(op.sequence, data_init.WaveSequence)


def _is_sequence_flow(op) -> bool:
    """Any ``FieldSequence.get_flow_operator(...)``: the wave sequence has a closed form in the kernel, every other
    sequence is tabulated by the host (``FieldSequence.frames``) and read per cell."""
    return isinstance(op, data_init.FoodFlowOperator) and isinstance(op.sequence, data_init.FieldSequence)


def _dynamics_to_c(d: Dynamics) -> _lib.DieDynamics:
    weights = getattr(d.op_action_cost, '_die_weights', None)
    if weights is None:
        raise NotImplementedError(
            "op_action_cost must be die_b200.env.linear_action_cost or zero_cost: the cost is "
            "evaluated inside the CUDA feed kernel, arbitrary Python operators are not supported")
    if d.op_food_flow is not identity_food_flow and d.op_food_flow is not None \
            and not _is_sequence_flow(d.op_food_flow):
        raise NotImplementedError("op_food_flow must be the identity or <FieldSequence>.get_flow_operator(...) "
                                  "(WaveSequence, PerlinNoiseSequence, TabulatedSequence, a subclass): the flow is "
                                  "evaluated inside the CUDA field kernel, arbitrary Python operators are not supported")
    if d.diffuse_mode not in _lib.DIFFUSE_MODES:
        raise ValueError(f"diffuse_mode must be one of {sorted(_lib.DIFFUSE_MODES)} (scipy.ndimage's modes)")
    c = _lib.DieDynamics()
    c.rate_feed = d.rate_feed
    c.rate_decay_chem = d.rate_decay_chem
    c.cost_w_deposit, c.cost_w_dist = weights
    w = gaussian_kernel1d(d.diffuse_sigma)
    radius = (len(w) - 1) // 2
    if radius > _lib.DIE_MAX_RADIUS:
        raise NotImplementedError(f"diffuse_sigma={d.diffuse_sigma} needs blur radius {radius} > {_lib.DIE_MAX_RADIUS}")
    for k, v in enumerate(w):
        c.blur_w[k] = float(v)
    c.blur_radius = radius
    if d.boundary == BoundaryCondition.wrap:
        c.boundary = _lib.BOUNDARY_WRAP
    elif d.boundary == BoundaryCondition.limit:
        c.boundary = _lib.BOUNDARY_LIMIT
    else:
        logging.warning(f'Unfamiliar boundary condition: {d.boundary}! Doing nothing with boundary...')
        c.boundary = _lib.BOUNDARY_NONE
    c.food_infinite = int(bool(d.food_infinite))
    c.diffuse_mode = _lib.DIFFUSE_MODES[d.diffuse_mode]
    # core/env.py:245-250 with the AgentIndexer following the rebound array (the reference's own version does not):
    # semantics pinned in oracle/die_ref.py:_agent_lifecycle and tests/test_golden_oracle.py
    c.agents_die = int(bool(d.agents_die))
    return c


def _check_f32_dynamics(c: _lib.DieDynamics) -> None:
    if c.blur_radius > 4:
        raise NotImplementedError(f"float32 fields support blur radius <= 4 (this diffuse_sigma gives {c.blur_radius})")


class Env:
    """core/env.py:64-298."""

    def __init__(self,
                 field_size: Tuple[int, int],
                 dynamics: Optional[Dynamics] = None,
                 *,
                 batch: Optional[int] = None,
                 device: Optional[Union[int, str, torch.device]] = None,
                 noise_seed: Optional[int] = None,
                 init_state: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                 init: str = 'host',
                 seed: Optional[int] = None,
                 field_dtype: Union[torch.dtype, str] = torch.float64,
                 verify_caches: bool = False):
        """``init='host'`` (default) builds the initial state with numpy in the reference's draw order
        (``np.random.seed`` reproduces it); ``init='device'`` builds it on the GPU (die_b200/device_init.py:
        same arithmetic and slot order, torch's device generator seeded by ``seed``) -- the only practical
        choice for 4096^2 and beyond, where the reference's per-cell Python loop takes minutes."""
        if not torch.cuda.is_available():
            raise RuntimeError("die_b200.Env needs a CUDA device: there is no CPU fallback")
        # verify_caches (a DEBUG mode, slow: it reads every tensor twice per step and synchronises): the caches below are
        # validated by torch's version counters, which a write through tensor.data, a DLPack / CuPy / Numba view or a
        # user kernel does not bump.  With this switch every cache carries a checksum of the bytes it was built from and
        # a mismatch raises instead of silently stepping on stale cells / alive bits (see invalidate_caches).
        self._verify_caches = bool(verify_caches)
        self.use_cost_hint = True          # a gradient agent's forward may leave the action cost for the feed kernel (DIE_FWD_WRITE_COST)
        self._checksums = None
        # float64 as in the reference (default), or the float32 FIELD mode: the medium (and the library's per-cell
        # scratch) in float32, agents / actions / headings still float64 -- half the field bytes, results within float32
        # rounding of the float64 ones per step (tests/test_gpu_f32_fields.py) instead of bit-exact
        self._field_dtype = {torch.float64: torch.float64, 'float64': torch.float64, 'f64': torch.float64,
                             torch.float32: torch.float32, 'float32': torch.float32, 'f32': torch.float32}.get(field_dtype)
        if self._field_dtype is None:
            raise ValueError("field_dtype must be torch.float64 or torch.float32")
        if self._field_dtype == torch.float32 and (dynamics is not None and dynamics.apply_sense_mask):
            raise NotImplementedError("apply_sense_mask is implemented for float64 fields only")
        self._lib = _lib.load()
        self._field_size = (int(field_size[0]), int(field_size[1]))
        self.dynamics = dynamics or Dynamics()
        self._batched = batch is not None
        self._B = int(batch) if batch is not None else 1
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self._noise_seed = noise_seed
        if init not in ('host', 'device'):
            raise ValueError("init must be 'host' or 'device'")
        self._init_mode = init
        self._seed = 0 if seed is None else int(seed)
        self._resets = 0
        self._handle = None
        self._host = None
        self._init_data(self._field_size, init_state)

    # -- construction ---------------------------------------------------------------------
    def _init_data(self, field_size, init_state=None, in_place=False):
        """core/env.py:74-86.  in_place (reset()): when the new state has the shape of the old one, it is written into the
        existing tensors -- the library handle, every buffer address and therefore a captured GraphedLoop stay valid."""
        h, w = field_size
        B = self._B
        owned = False
        if init_state is None and self._init_mode == 'device':
            from . import device_init
            owned = True
            with torch.cuda.device(self.device):
                init_state = device_init.init_state_device(field_size, self.dynamics.init_agent_ratio,
                                                           self._seed + 7919 * self._resets, self.device, batch=B)
            self._resets += 1
        if init_state is None:
            mediums, agentss = [], []
            for b in range(B):
                seed = None if self._noise_seed is None else self._noise_seed + b
                m = data_init.init_medium(field_size, self.dynamics.init_agent_ratio, noise_seed=seed)
                mediums.append(m)
                agentss.append(data_init.agents_from_medium(m))
            medium, agents = np.stack(mediums), np.stack(agentss)
        elif isinstance(init_state[0], torch.Tensor):
            # device (or host) tensors: no numpy round trip (large batches are assembled on the GPU)
            medium = init_state[0].to(dtype=self._field_dtype).reshape(B, 3, h, w)
            agents = init_state[1].to(dtype=torch.float64).reshape(B, 4, -1)
        else:
            medium, agents = (np.asarray(a, dtype=np.float64) for a in init_state)
            medium = medium.reshape(B, 3, h, w)
            agents = agents.reshape(B, 4, -1)
        if in_place and self._handle is not None and int(agents.shape[-1]) == self._M \
                and self._dynamics_key == self._dynamics_snapshot():
            with torch.cuda.device(self.device):
                to_t = (lambda a: a) if isinstance(medium, torch.Tensor) else (lambda a: torch.from_numpy(np.ascontiguousarray(a)))
                self._medium_buf[0].copy_(to_t(medium))
                self._agents.copy_(to_t(agents))
                self._cur = 0
                self._reward_dev.zero_()
                self._alive_dev.zero_()
                self.invalidate_caches()
                self.last_step_fused = False
                if self._obs_buf is not None:
                    self._refresh_sensed_medium()
                _hints.publish(self, self._medium_buf[0])
            return
        self._M = int(agents.shape[-1])
        with torch.cuda.device(self.device):
            if isinstance(medium, torch.Tensor):
                first = medium.to(self.device, dtype=self._field_dtype).contiguous()
                self._agents = agents.to(self.device).contiguous()
                if not owned:               # never alias the caller's tensors
                    first = first.clone() if first.data_ptr() == init_state[0].data_ptr() else first
                    self._agents = self._agents.clone() if self._agents.data_ptr() == init_state[1].data_ptr() \
                        else self._agents
            else:
                first = torch.from_numpy(np.ascontiguousarray(medium)).to(self.device, dtype=self._field_dtype)
                self._agents = torch.from_numpy(np.ascontiguousarray(agents)).to(self.device)
            self._medium_buf = [first, torch.empty((B, 3, h, w), dtype=self._field_dtype, device=self.device)]
            self._cur = 0
            self._reward_dev = torch.zeros(B, dtype=torch.float64, device=self.device)
            self._alive_dev = torch.zeros(B, dtype=torch.int64, device=self.device)
            self._reward_host = torch.zeros(B, dtype=torch.float64).pin_memory()
            self._alive_host = torch.zeros(B, dtype=torch.int64).pin_memory()
            self._reward_np, self._alive_np = self._reward_host.numpy(), self._alive_host.numpy()
            if self._handle is not None:
                _lib.check(self._lib.die_env_destroy(self._handle))
                self._handle = None
            handle = _lib.C.c_void_p()
            cdyn = _dynamics_to_c(self.dynamics)
            if self._field_dtype == torch.float32:
                _check_f32_dynamics(cdyn)
            _lib.check(self._lib.die_env_create(h, w, self._M, B, _lib.C.byref(cdyn), _lib.C.byref(handle)))
            self._handle = handle
            self._generation = getattr(self, '_generation', 0) + 1      # reset() re-creates every buffer
            if self._field_dtype == torch.float32:
                _lib.check(self._lib.die_env_set_field_dtype(handle, _lib.FIELD_F32))
            self._dynamics_key = self._dynamics_snapshot()
            self._install_food_flow()
            # Dynamics.apply_sense_mask: the observation is a masked COPY of the medium (core/env.py:275-294)
            self._obs_buf = None
            if self.dynamics.apply_sense_mask:
                self._obs_buf = [torch.empty_like(first), torch.empty_like(first)]
                self._sense_w = np.ascontiguousarray(gaussian_kernel1d(2.0), dtype=np.float64)
                self._refresh_sensed_medium()
            self._publish_grad = False
            self._hint_state = None         # (medium ptr, medium version, agents version, grad published)
            self._speculation = None        # (action ptr, action version, agents version) of a pending fused move
            self._cost_hint = None          # (action ptr, action version) the library's cost array was written for
            self._alive_version = None      # agents._version the library's alive bitmask was built for
            self.last_step_fused = False
            _hints.publish(self, first)

    def _install_food_flow(self) -> None:
        """Dynamics.op_food_flow = WaveSequence flow operator -> device tables for the field kernel."""
        op = self.dynamics.op_food_flow
        self._flow_tables = None
        if not _is_sequence_flow(op):
            return
        if tuple(op.sequence._size) != tuple(self._field_size):
            raise ValueError(f"the {type(op.sequence).__name__} was built for field {op.sequence._size}, the env is {self._field_size}")
        if not _is_wave_flow(op):
            # no closed form on the device: the frames of every time step, tabulated by the sequence's own code.
            # T x H x W float64 on the host AND the device: refuse by name what cannot fit (the base-class defaults,
            # t_bounds = (0, 10) and dt = 0.01, mean T = 1000: 0.5 GB at 256^2, 134 GB at 4096^2), and keep the device
            # copy across reset() (the frames do not change)
            seq = op.sequence
            nbytes = 8 * len(seq._ts) * self._field_size[0] * self._field_size[1]
            cached = getattr(seq, '_die_device_frames', None)
            if cached is None or cached.device != self.device:
                free, _total = torch.cuda.mem_get_info(self.device)
                if nbytes > 0.5 * free or nbytes > (32 << 30):
                    raise MemoryError(
                        f"{type(seq).__name__}: tabulating {len(seq._ts)} frames of {self._field_size[0]}x{self._field_size[1]} "
                        f"needs {nbytes / 2**30:.1f} GiB on the host and on the device ({free / 2**30:.1f} GiB free); use a "
                        f"larger dt / shorter t_bounds, a TabulatedSequence of fewer frames, or the WaveSequence (closed form)")
                seq._die_device_frames = torch.from_numpy(seq.frames()).to(self.device)
            frames = seq._die_device_frames
            self._flow_tables = [frames]               # borrowed by the library until the handle dies
            _lib.check(self._lib.die_env_set_food_frames(
                self._handle, frames.data_ptr(), frames.shape[0], op.calls % frames.shape[0], op.scale, op.decay))
            return
        rwave, col, row = op.sequence.device_tables()
        ts = np.ascontiguousarray(op.sequence.ts, dtype=np.float64)
        tabs = [torch.from_numpy(a).to(self.device) for a in (rwave, col, row)]
        self._flow_tables = tabs                   # borrowed by the library until the handle dies
        _lib.check(self._lib.die_env_set_food_flow(
            self._handle, tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(),
            ts.ctypes.data, len(ts), op.calls % len(ts), op.scale, op.decay))

    def _dynamics_snapshot(self):
        d = self.dynamics
        return (d.op_action_cost, d.op_food_flow, d.rate_feed, d.rate_decay_chem, d.boundary, d.diffuse_mode,
                d.diffuse_sigma, d.apply_sense_mask, d.food_infinite, d.agents_die)

    def _sync_dynamics(self) -> None:
        """The reference reads ``self.dynamics`` on every step (core/env.py:136-150, 220-250), so a caller may change a
        field between steps (or assign a new Dynamics).  The library holds a snapshot: compare, and on a change hand it the
        new parameters (die_env_set_dynamics), re-install the food flow tables and (de)allocate the masked observation."""
        key = self._dynamics_snapshot()
        if key == self._dynamics_key:
            return
        cdyn = _dynamics_to_c(self.dynamics)            # raises for what the kernels do not implement
        if self.dynamics.apply_sense_mask and self._field_dtype != torch.float64:
            raise NotImplementedError("apply_sense_mask is implemented for float64 fields only")
        if self._field_dtype == torch.float32:
            _check_f32_dynamics(cdyn)
        _lib.check(self._lib.die_env_set_dynamics(self._handle, _lib.C.byref(cdyn)))
        if key[1] is not self._dynamics_key[1]:
            self._install_food_flow()
        if key[7] != self._dynamics_key[7]:
            if key[7]:
                first = self._medium_buf[0]
                self._obs_buf = [torch.empty_like(first), torch.empty_like(first)]
                self._sense_w = np.ascontiguousarray(gaussian_kernel1d(2.0), dtype=np.float64)
                self._refresh_sensed_medium()
            else:
                self._obs_buf = None
        self._dynamics_key = key
        self.invalidate_caches()

    def invalidate_caches(self) -> None:
        """Forget everything the Env caches about its own tensors: the alive bitmask, the per-slot cell cache / published
        gradient offered to agents as hints, and a pending speculative move.  The caches are validated by torch's version
        counters, which in-place torch operations bump -- but a write through ``tensor.data``, a DLPack / CuPy / Numba
        view or a user kernel does NOT.  Call this after such a write (the reference re-reads its arrays every step)."""
        self._alive_version = None
        self._hint_state = None
        self._speculation = None
        self._cost_hint = None
        self._checksums = None
        if getattr(self, '_host', None) is not None:
            self._host['alive_of'] = None          # the host path downloads the alive channel again
        with _lib.on_device(self.device):
            _lib.check(self._lib.die_env_discard_move(self._handle, torch.cuda.current_stream().cuda_stream))

    def _refresh_sensed_medium(self) -> None:
        """obs medium = medium.where(sense_mask, 0.) for the current medium (die_sense_mask)."""
        w = self._sense_w
        h, wd = self._field_size
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_sense_mask(h, wd, self._B, w.ctypes.data, (len(w) - 1) // 2,
                                                self._medium_buf[self._cur].data_ptr(),
                                                self._obs_buf[self._cur].data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))

    def __del__(self):
        try:
            if getattr(self, '_handle', None) is not None:
                self._lib.die_env_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[ObsType, dict]:
        """core/env.py:94-99 (``seed`` is ignored there too).  The new state is written into the env's existing tensors
        (tensors handed out earlier as observations now show the new state, as the reference's views of its xarray would
        not -- copy what must survive a reset)."""
        self._init_data(self._field_size, in_place=True)
        return self._get_current_obs, {}

    def _hints_are_current(self) -> bool:
        """The caches offered to agents (cells, gradient, food under the agent) describe the env's CURRENT tensors."""
        st = self._hint_state
        buf = self._medium_buf[self._cur]
        return st is not None and st[0] == buf.data_ptr() and st[1] == buf._version and st[2] == self._agents._version

    # -- state access ---------------------------------------------------------------------
    def _unbatch(self, t):
        return t if self._batched else t[0]

    @property
    def medium(self) -> torch.Tensor:
        return self._unbatch(self._medium_buf[self._cur])

    @property
    def agents(self) -> torch.Tensor:
        return self._unbatch(self._agents)

    @property
    def field_size(self) -> Tuple[int, int]:
        return self._field_size

    @property
    def field_dtype(self) -> torch.dtype:
        return self._field_dtype

    @property
    def max_agents(self) -> int:
        return self._M

    @property
    def batch(self) -> int:
        return self._B

    @property
    def _num_alive_agents(self):
        """core/env.py:263-265."""
        n = (self._agents[:, 2] > 0).sum(dim=1)
        return n.cpu().numpy() if self._batched else int(n.item())

    @property
    def _get_current_obs(self) -> ObsType:
        """core/env.py:296-298: (the live agents, the sensed medium)."""
        if self._obs_buf is not None:
            return self.agents, self._unbatch(self._obs_buf[self._cur])
        return self.agents, self.medium

    def get_state(self) -> Tuple[np.ndarray, np.ndarray]:
        """Host copies (medium, agents) -- checkpoint / parity aid (the reference never saves Env state)."""
        return self.medium.cpu().numpy(), self.agents.cpu().numpy()

    def set_state(self, medium=None, agents=None) -> None:
        if medium is not None:
            src = torch.as_tensor(np.asarray(medium, dtype=np.float64)).reshape(self._medium_buf[0].shape)
            self._medium_buf[self._cur].copy_(src)          # (rounds to float32 in the float32 field mode)
            if self._obs_buf is not None:
                self._refresh_sensed_medium()
        if agents is not None:
            src = torch.as_tensor(np.asarray(agents, dtype=np.float64)).reshape(self._agents.shape)
            self._agents.copy_(src)

    def last_cells(self) -> torch.Tensor:
        """int32 [B, M] (or [M]) linear cell index ix*W+iy of every slot after the last move
        (validation aid; a copy of the library's per-slot cell cache)."""
        view = _DevicePtrView(self._lib.die_env_cells(self._handle), (self._B, self._M), '<i4')
        with torch.cuda.device(self.device):
            out = torch.as_tensor(view, device=self.device).clone()
        return self._unbatch(out)

    # -- the step -------------------------------------------------------------------------
    def _check_action(self, action: torch.Tensor) -> torch.Tensor:
        if not isinstance(action, torch.Tensor):
            raise TypeError("action must be a torch CUDA tensor (device path) or a numpy array (host path)")
        expect = (self._B, 3, self._M) if self._batched else (3, self._M)
        if tuple(action.shape) != expect:
            raise ValueError(f"action shape {tuple(action.shape)} != {expect}")
        if action.dtype != torch.float64 or action.device != self.device:
            raise ValueError("action must be float64 on the env's device")
        return action if action.is_contiguous() else action.contiguous()

    def step_async(self, action: ActType):
        """``step`` without the host synchronisation: returns (obs, reward_dev[B], alive_dev[B])
        as device tensors, everything enqueued on the current stream."""
        action = self._check_action(action)
        self._sync_dynamics()
        if self._verify_caches:
            self._verify()
        nxt = 1 - self._cur
        # a gradient agent may have evaluated this very action's move + claims already (see _forward_flags):
        # adopt them iff the action tensor and the agents are provably untouched since
        spec, self._speculation = self._speculation, None
        fused = (spec is not None and spec[:3] == (action.data_ptr(), action._version, self._agents._version))
        if spec is not None and spec[3] and not fused:
            self._speculation = spec       # (still pending: the caller may yet pass the right action)
            raise RuntimeError("Env.step: the agent committed the move of its action (fuse_move='commit'), so this step "
                               "must receive exactly the tensor that forward() returned, unmodified -- the positions in "
                               "env.agents are already those of that action")
        if self.dynamics.agents_die:
            fused = False                  # (the library refuses the combination; a pending speculation is discarded)
        cost, self._cost_hint = self._cost_hint, None
        use_cost = cost is not None and cost == (action.data_ptr(), action._version)
        flags = (_lib.STEP_ADOPT_MOVE if fused else 0) | (0 if self.dynamics.agents_die else _lib.STEP_ALIVE_BITS) \
            | (_lib.STEP_USE_COST if use_cost else 0)
        with _lib.on_device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            if not self.dynamics.agents_die:
                self._refresh_alive(stream)
            _lib.check(self._lib.die_env_step_flags(
                self._handle, self._medium_buf[self._cur].data_ptr(), self._medium_buf[nxt].data_ptr(),
                self._agents.data_ptr(), action.data_ptr(),
                self._reward_dev.data_ptr(), self._alive_dev.data_ptr(), flags, stream))
        self.last_step_fused = fused
        self._cur = nxt
        if self._flow_tables is not None:
            self.dynamics.op_food_flow.calls += 1      # the operator's iterator advances once per step
        self._after_step()
        return self._get_current_obs, self._reward_dev, self._alive_dev

    # -- fast-path hints for Agent.forward (see die_b200/_hints.py) ---------------------------------
    def _after_step(self) -> None:
        if self._obs_buf is not None:
            self._refresh_sensed_medium()
        buf = self._medium_buf[self._cur]
        self._hint_state = (buf.data_ptr(), buf._version, self._agents._version, self._publish_grad)
        if self._verify_caches:
            self._checksums = (buf._version, self._agents._version, self._checksum(buf), self._checksum(self._agents))
        _hints.publish(self, buf)

    @staticmethod
    def _checksum(t: torch.Tensor) -> int:
        return int(t.view(torch.int64 if t.element_size() == 8 else torch.int32).sum().item())

    def _verify(self) -> None:
        """verify_caches: the env's tensors still hold the bytes the caches were built from, unless torch saw the write."""
        cs = self._checksums
        if cs is None:
            return
        buf = self._medium_buf[self._cur]
        if buf._version == cs[0] and self._checksum(buf) != cs[2]:
            raise RuntimeError("verify_caches: the env's medium was written without torch noticing (tensor.data, a DLPack / "
                               "CuPy / Numba view, a user kernel): call env.invalidate_caches() after such a write")
        if self._agents._version == cs[1] and self._checksum(self._agents) != cs[3]:
            raise RuntimeError("verify_caches: the env's agents tensor was written without torch noticing (tensor.data, a "
                               "DLPack / CuPy / Numba view, a user kernel): call env.invalidate_caches() after such a write")

    def _hints_for(self, agents, medium, want_gradient: bool):
        if self._verify_caches:
            self._verify()
        if want_gradient and not self._publish_grad:
            # a gradient agent is acting on this env: publish np.gradient(chem1) from the next step on
            _lib.check(self._lib.die_env_publish_gradient(self._handle, 1))
            self._publish_grad = True
        st = self._hint_state
        if st is None:
            return None, None
        buf = self._medium_buf[self._cur]
        grad_ptr = cells_ptr = None
        if (medium.data_ptr() == st[0] == buf.data_ptr() and medium._version == st[1]
                and medium.shape[-2:] == buf.shape[-2:] and medium.numel() == buf.numel()):
            if agents.data_ptr() == self._agents.data_ptr() and agents._version == st[2] \
                    and agents.numel() == self._agents.numel():
                cells_ptr = self._lib.die_env_cells(self._handle)
            if want_gradient and st[3]:
                grad_ptr = self._lib.die_env_gradient_kind(self._handle) or None      # 1 float64, 2 float32 cache
        return grad_ptr, cells_ptr

    def _forward_flags(self, agents, medium, want_gradient: bool, speculate) -> int:
        """Flags for ``die_env_forward_gradient`` (include/die_b200.h) given the observation an agent
        received: which of this env's caches are provably valid for it, and whether the agent may
        evaluate the move of its action speculatively (``obs[0]`` must be this env's own agents tensor;
        the alive bitmask is rebuilt here whenever the agents tensor was edited)."""
        if self._speculation is not None and self._speculation[3]:
            raise RuntimeError("Agent.forward: a committed move (fuse_move='commit') is waiting for its Env.step; the "
                               "positions in env.agents are already those of the next step")
        grad_ptr, cells_ptr = self._hints_for(agents, medium, want_gradient)
        flags = (_lib.FWD_USE_GRADIENT if grad_ptr else 0) | (_lib.FWD_USE_CELLS if cells_ptr else 0)
        # cost hint (DIE_FWD_WRITE_COST): the forward launch leaves linear_action_cost of its action for the feed kernel;
        # Env.step uses it iff it receives that very tensor unmodified (_note_cost / step_async)
        if not self.dynamics.agents_die and self._field_dtype == torch.float64 and self.use_cost_hint:
            flags |= _lib.FWD_WRITE_COST
        # 'commit': the run-loop contract (include/die_b200.h, DIE_FWD_COMMIT_MOVE) -- only on the env's own, valid caches,
        # i.e. in the steady state of `action = agent.forward(obs); obs, ... = env.step(action)`; elsewhere the plain path
        commit = speculate == 'commit'
        if commit and not (grad_ptr and cells_ptr and not self._verify_caches):
            return flags
        if speculate and not self.dynamics.agents_die and self._field_dtype == torch.float64 \
                and agents.data_ptr() == self._agents.data_ptr() \
                and agents.numel() == self._agents.numel():
            with _lib.on_device(self.device):
                self._refresh_alive(torch.cuda.current_stream().cuda_stream)
            flags |= _lib.FWD_SPECULATE_MOVE | (_lib.FWD_COMMIT_MOVE if commit else 0)
        return flags

    def _refresh_alive(self, stream) -> None:
        """(Re)build the library's one-bit-per-slot alive mask iff the agents tensor was edited since it was
        built (torch's version counter; the step kernels never change the alive channel)."""
        if self._alive_version != self._agents._version:
            _lib.check(self._lib.die_env_refresh_alive(self._handle, self._agents.data_ptr(), stream))
            self._alive_version = self._agents._version

    def _note_cost(self, action: torch.Tensor) -> None:
        self._cost_hint = (action.data_ptr(), action._version)

    def _note_speculation(self, action: torch.Tensor, committed: bool = False) -> None:
        self._speculation = (action.data_ptr(), action._version, self._agents._version, committed)

    def step(self, action: ActType):
        """core/env.py:101-131 -> (obs, reward, terminated, truncated, info)."""
        if isinstance(action, np.ndarray):
            return self._step_host(action)
        obs, _, _ = self.step_async(action)
        with _lib.on_device(self.device):
            _lib.check(self._lib.die_env_read_stats(
                self._handle, self._reward_dev.data_ptr(), self._alive_dev.data_ptr(),
                self._reward_host.data_ptr(), self._alive_host.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return (obs, *self._summarise(self._reward_np.copy(), self._alive_np.copy()))

    def _summarise(self, reward: np.ndarray, alive: np.ndarray):
        """core/env.py:117-131."""
        if self._batched:
            mean = np.divide(reward, alive, out=np.zeros_like(reward), where=alive > 0)
            info = {'num_agents': alive, 'reward': np.round(reward, 3), 'mean_reward': np.round(mean, 5)}
            return reward, bool((alive == 0).all()), False, info
        r, n = float(reward[0]), int(alive[0])
        mean = r / n if n > 0 else 0.
        info = {'num_agents': n, 'reward': np.round(r, 3), 'mean_reward': np.round(mean, 5)}
        return r, n == 0, False, info

    # -- host-buffer path (what a host-side caller of the reference would bind) -----------------
    def host_buffers(self):
        """Pinned host arrays (action, agents, medium x2) used by the host path.  Filling
        ``action`` in place and passing it to ``step`` avoids one host memcpy."""
        if self._host is None:
            B, M = self._B, self._M
            h, w = self._field_size
            mk = lambda *s: torch.empty(s, dtype=torch.float64, pin_memory=True)
            self._host = {
                'action': mk(B, 3, M), 'agents': mk(B, 4, M),
                'medium': [mk(B, 3, h, w), mk(B, 3, h, w)], 'flip': 0,
                'reward': mk(B), 'alive': torch.empty(B, dtype=torch.int64, pin_memory=True),
            }
        return self._host

    def _step_host(self, action: np.ndarray):
        if self._field_dtype != torch.float64:
            raise NotImplementedError("the host-buffer path (numpy actions) runs float64 fields only")
        if self._speculation is not None and self._speculation[3]:
            raise RuntimeError("Env.step (host buffers): a committed move (fuse_move='commit') is waiting for its device step")
        hb = self.host_buffers()
        B, M = self._B, self._M
        self._sync_dynamics()
        # the library copies straight from the caller's array (cudaMemcpyAsync: full speed if it is pinned --
        # e.g. the array an Agent's host path returned -- staged by the driver if it is pageable); no host memcpy
        src = np.ascontiguousarray(np.asarray(action, dtype=np.float64).reshape(B, 3, M))
        # the very (read-only) array an Agent's host path just returned, its device copy untouched since: no re-upload
        dev_action = _hints.device_copy_of(action, self.device) if isinstance(action, np.ndarray) else None
        if dev_action is not None and dev_action.numel() != 3 * B * M:
            dev_action = None
        self.last_step_reused_device_action = dev_action is not None
        hb['flip'] ^= 1
        med_t = hb['medium'][hb['flip']]
        nxt = 1 - self._cur
        self._speculation = None          # die_env_step_host runs the plain step (and discards a pending move)
        self.last_step_fused = False
        # the pinned agents buffer already holds the current alive channel (a full download into it since the agents tensor
        # was last edited, and Env.step never changes that channel without a lifecycle): 24 instead of 32 B per slot.  The
        # array handed out is read-only, so the caller cannot have scribbled on it either.
        keep_alive = (not self.dynamics.agents_die and hb.get('alive_of') == (self._agents._version, self._generation))
        self.last_step_kept_alive_channel = keep_alive
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self._lib.die_env_step_host_flags(
                self._handle, self._medium_buf[self._cur].data_ptr(), self._medium_buf[nxt].data_ptr(),
                self._agents.data_ptr(), src.ctypes.data if dev_action is None else None,
                None if dev_action is None else dev_action.data_ptr(),
                hb['agents'].data_ptr(), med_t.data_ptr() if self._obs_buf is None else None,
                hb['reward'].data_ptr(), hb['alive'].data_ptr(),
                _lib.HOST_KEEP_ALIVE_CHANNEL if keep_alive else 0, stream))
        hb['alive_of'] = None if self.dynamics.agents_die else (self._agents._version, self._generation)
        self._cur = nxt
        if self._flow_tables is not None:
            self.dynamics.op_food_flow.calls += 1
        self._after_step()
        if self._obs_buf is not None:              # the host observation is the masked medium, too
            with torch.cuda.device(self.device):
                med_t.copy_(self._obs_buf[self._cur], non_blocking=True)
                torch.cuda.current_stream().synchronize()
        agents_np = hb['agents'].numpy()
        agents_np.flags.writeable = False      # (a host COPY of the env's agents: writes to it never reached the env anyway)
        obs = (self._unbatch(agents_np), self._unbatch(med_t.numpy()))
        return (obs, *self._summarise(hb['reward'].numpy().copy(), hb['alive'].numpy().copy()))

    def host_io_bytes_per_step(self) -> Tuple[int, int]:
        """(H2D, D2H) bytes the last host-path ``step`` moved (no H2D when it re-used the agent's device action)."""
        B, M = self._B, self._M
        h, w = self._field_size
        h2d = 0 if getattr(self, 'last_step_reused_device_action', False) else 8 * B * 3 * M
        agent_channels = 3 if getattr(self, 'last_step_kept_alive_channel', False) else 4
        return h2d, 8 * B * (agent_channels * M + 3 * h * w) + 16 * B

    # -- measurement aid ----------------------------------------------------------------------
    STEP_KERNELS = ('move_claim', 'field_step', 'agent_feed', 'finalize_stats')

    def set_profiling(self, on: bool) -> None:
        """Record CUDA events between the step's kernels (bench.py's per-kernel roofline)."""
        _lib.check(self._lib.die_env_set_profiling(self._handle, int(on)))

    def kernel_times(self):
        """-> ({kernel: accumulated ms}, profiled steps) since profiling was enabled."""
        ms = (_lib.C.c_double * len(self.STEP_KERNELS))()
        n = _lib.C.c_int64()
        _lib.check(self._lib.die_env_kernel_times(self._handle, ms, _lib.C.byref(n)))
        return dict(zip(self.STEP_KERNELS, list(ms))), int(n.value)

    def render(self, host: bool = False):
        """core/env.py:133-134: the frames of EnvRenderer.render(medium, agents) (die_b200/render.py), device tensors
        by default, numpy arrays in pinned memory with ``host=True``."""
        if self._field_dtype != torch.float64:
            raise NotImplementedError("render() is implemented for float64 fields only")
        if getattr(self, '_renderer', None) is None or self._renderer._B != self._B:
            from .render import EnvRenderer
            self._renderer = EnvRenderer(self._field_size, field_colors_id='rgb', batch=self._B, device=self.device)
        med, ag = self._medium_buf[self._cur], self._agents
        frames = self._renderer.render_host(med, ag) if host else self._renderer.render(med, ag)
        return frames if self._batched else [f[0] if f is not None else None for f in frames]


class _DevicePtrView:
    """Zero-copy torch view of library-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr,
                                         'data': (int(ptr), False), 'version': 2}
