"""GradientAgent / PhysarumAgent -- drop-ins for core/agent/gradient.py:13-219.

Agent-private state (``_direction_rads`` = theta, ``_prev_grad``) lives on the device; one
``forward`` is one launch of ``die_gradient_forward``.  The reference materialises the
normalised gradient of the whole chem1 field every call; the kernel evaluates the same
np.gradient stencil only at each slot's sensed cell.
"""
from typing import Optional

import numpy as np
import torch

from .. import _hints
from .. import _lib
from ..base_types import ActType, ObsType
from .static import _DeviceAgent, _split_obs


def _renormalize_radians(r):
    return (r - np.pi) % (-2 * np.pi) + np.pi


class GradientAgent(_DeviceAgent):
    """core/agent/gradient.py:13-124.

    ``rng='philox'`` (default): coin flips / momentum noise are drawn in-kernel.
    ``rng='numpy'`` (validation mode): drawn on the host exactly where the reference draws them
    -- ``np.random.randint(0, 2, M)`` from the GLOBAL legacy RNG for Physarum turns
    (core/agent/gradient.py:181) and ``rng.normal(0, .4, (2, M))`` from the agent's private
    generator for the momentum noise (:50-53) -- and uploaded.
    """
    _discrete_turn = False

    def __init__(self,
                 max_agents: int = 10 ** 6,
                 scale: float = 0.01,
                 deposit: float = 4.0,
                 inertia: float = 0.9,
                 sense_offset: float = 0.,
                 noise_scale: float = 0.025,
                 normalized_grad: bool = True,
                 grad_clip: Optional[float] = 1e-5,
                 *, rng: str = 'philox', seed: int = 0):
        super().__init__()
        self._init_params = dict(max_agents=max_agents, scale=scale, deposit=deposit, inertia=inertia,
                                 sense_offset=sense_offset, noise_scale=noise_scale,
                                 normalized_grad=normalized_grad, grad_clip=grad_clip)
        if rng not in ('philox', 'numpy'):
            raise ValueError("rng must be 'philox' or 'numpy'")
        self._size = int(max_agents)
        self._rng_mode = rng
        self._seed = int(seed)
        self._rng = np.random.default_rng(None if rng == 'numpy' else seed)
        self._noise_scale = float(noise_scale)
        self._scale = float(scale)
        self._deposit = float(deposit)
        self._inertia = float(inertia)
        self._sense_offset_scale = float(sense_offset)
        self._normalized = bool(normalized_grad)
        self._grad_clip = grad_clip
        self._turn_radians = 0.0
        self._sense_radians = 0.0
        self._rtol = 0.0
        self._theta = None          # [B, M] device
        self._prev_grad = None      # [B, 2, M] device (only when it can matter)
        self._coin_host = self._coin_dev = None
        self._noise_host = self._noise_dev = None
        self._step_dev = None           # device-resident call counter (set by die_b200.graph.GraphedLoop)
        self._upload_events = {}        # name -> CUDA event recorded after the last copy out of that pinned buffer
        self._sense_cells = None
        self.record_sense_cells = False
        self.use_env_hints = True           # die_b200/_hints.py: cached cells + published gradient
        # opt-in: evaluate the move + claim of the action in the same launch, for Env.step to adopt
        # (bit-identical; measured SLOWER on B200 at 4096^2 -- 1.084 vs 1.033 ms per step -- because the
        # forward kernel is instruction bound and move_claim already runs at 80 % of the HBM peak)
        # 'commit': the run-loop contract -- `action = agent.forward(obs); obs, ... = env.step(action)` with nothing in
        # between (examples/minimal_run.py:21-25 of the reference).  The forward launch then also MOVES the agents (the
        # env's positions are advanced when forward returns), Env.step runs the field pass and the feed kernel only:
        # one launch and 56 B per slot less per iteration, results bit-identical.  Anything else than stepping exactly
        # the returned tensor raises.  Active only once the env's caches are valid (from the second iteration on).
        self.fuse_move = False
        self.last_hints = (False, False)
        self.last_speculated = False
        self.last_committed = False

    # -- state ------------------------------------------------------------------------------
    def _needs_prev(self) -> bool:
        """Whether ``_process_momentum`` (core/agent/gradient.py:82-91) can change anything.  With inertia = noise_scale
        = 0 it computes 1.0 * g' + 0.0 * prev + 0.0 * noise, which is g' -- except for a component of g' that is
        exactly -0.0: (-0.0) + (+0.0) = +0.0, and the new heading angle(gx + 1j gy) is pi only for (-0.0, -0.0)
        (SURVEY Q6).  A Physarum turn on a NORMALISED gradient yields (cos d, sin d), never a zero; every other
        configuration (unnormalised: 0 * cos d; GradientAgent: the clipped gradient's signed zeros) can, so there the
        momentum arithmetic is carried out with its real operands."""
        return self._inertia != 0.0 or self._noise_scale != 0.0 or not (self._discrete_turn and self._normalized)

    def _initial_theta(self, prev_grad: np.ndarray) -> np.ndarray:
        """get_radians(prev_grad), core/agent/gradient.py:42-43."""
        return np.angle(prev_grad[..., 0, :] + np.multiply(1j, prev_grad[..., 1, :]))

    def _lazy_init(self, B: int, M: int, device):
        if self._theta is not None and tuple(self._theta.shape) == (B, M):
            return
        if M != self._size:
            raise ValueError(f"agent was built with max_agents={self._size} but the env has {M} slots "
                             f"(the reference needs them equal too, core/agent/gradient.py:105)")
        prev = self._rng.normal(loc=0., scale=0.4, size=(B, 2, M))       # _get_some_noise, :50-53
        self.set_state(theta=self._initial_theta(prev), prev_grad=prev, device=device)

    def set_state(self, theta=None, prev_grad=None, device=None):
        """Inject agent-private state (validation / checkpoint aid)."""
        device = device or (self._theta.device if self._theta is not None else
                            torch.device('cuda', torch.cuda.current_device()))
        if theta is not None:
            t = np.asarray(theta, dtype=np.float64)
            t = t.reshape(-1, t.shape[-1])
            self._theta = torch.from_numpy(np.ascontiguousarray(t)).to(device)
        if prev_grad is not None and self._needs_prev():
            p = np.asarray(prev_grad, dtype=np.float64)
            p = p.reshape(-1, 2, p.shape[-1])
            self._prev_grad = torch.from_numpy(np.ascontiguousarray(p)).to(device)

    def get_state(self):
        th = self._theta.cpu().numpy()
        pg = self._prev_grad.cpu().numpy() if self._prev_grad is not None else None
        if th.shape[0] == 1:
            th = th[0]
            pg = pg[0] if pg is not None else None
        return th, pg

    @property
    def sense_cells(self) -> Optional[torch.Tensor]:
        return self._sense_cells

    def _params_c(self) -> _lib.DieGradientParams:
        p = _lib.DieGradientParams()
        p.scale = self._scale
        p.deposit = self._deposit
        p.inertia = self._inertia
        p.sense_offset = self._sense_offset_scale
        p.noise_scale = self._noise_scale
        p.grad_clip = 0.0 if self._grad_clip is None else float(self._grad_clip)
        p.turn_radians = self._turn_radians
        p.sense_radians = self._sense_radians
        p.turn_tolerance = self._rtol
        p.normalized_grad = int(self._normalized)
        p.use_grad_clip = int(self._grad_clip is not None)
        p.discrete_turn = int(self._discrete_turn)
        return p

    def _host_staging(self, agents_np: np.ndarray, medium_np: np.ndarray):
        """Device staging tensors + pinned action buffer + stream context of the host-buffer path."""
        dev = torch.device('cuda', torch.cuda.current_device())
        hb = self._host
        if hb is None or hb['agents'].shape != agents_np.shape or hb['medium'].shape != medium_np.shape:
            hb = self._host = {
                'agents': torch.empty(agents_np.shape, dtype=torch.float64, device=dev),
                'medium': torch.empty(medium_np.shape, dtype=torch.float64, device=dev),
                'action': torch.empty((*agents_np.shape[:-2], 3, agents_np.shape[-1]), dtype=torch.float64, pin_memory=True),
            }
        if getattr(self, '_host_ctx', None) is None:
            ctx = _lib.C.c_void_p()
            _lib.check(self._lib.die_host_ctx_create(_lib.C.byref(ctx)))
            self._host_ctx = ctx
        return hb['agents'], hb['medium']

    def __del__(self):
        try:
            if getattr(self, '_host_ctx', None) is not None:
                self._lib.die_host_ctx_destroy(self._host_ctx)
                self._host_ctx = None
        except Exception:
            pass

    def _params_cached(self) -> _lib.DieGradientParams:
        """The C struct is rebuilt only when one of the (private) constructor parameters was changed."""
        key = (self._scale, self._deposit, self._inertia, self._sense_offset_scale, self._noise_scale, self._grad_clip,
               self._turn_radians, self._sense_radians, self._rtol, self._normalized, self._discrete_turn)
        if getattr(self, '_params_key', None) != key:
            self._params_key, self._params_struct = key, self._params_c()
        return self._params_struct

    # -- forward ------------------------------------------------------------------------------
    @staticmethod
    def host_io_bytes_per_forward(B: int, M: int, C: int):
        """(H2D, D2H) bytes of one host-path ``forward`` over B environments of M slots and C cells: the channels the
        policy reads -- x, y and env_food, chem1 (die_gradient_forward_host) -- up, the action down."""
        return 8 * B * (2 * M + 2 * C), 8 * B * 3 * M

    def _upload(self, name: str, arr: np.ndarray, shape, dtype, device):
        host, dev = getattr(self, f'_{name}_host'), getattr(self, f'_{name}_dev')
        if host is None or tuple(host.shape) != tuple(shape):
            host = torch.empty(shape, dtype=dtype).pin_memory()
            dev = torch.empty(shape, dtype=dtype, device=device)
            setattr(self, f'_{name}_host', host)
            setattr(self, f'_{name}_dev', dev)
        # the previous call's asynchronous copy out of this pinned buffer may still be queued (step_async loops never
        # synchronise): wait for it before overwriting the buffer, or the GPU could read the NEXT call's draws
        ev = self._upload_events.get(name)
        if ev is not None:
            ev.synchronize()
        host.numpy()[...] = np.asarray(arr).reshape(shape)
        dev.copy_(host, non_blocking=True)
        if ev is None:
            ev = self._upload_events[name] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        return dev.data_ptr()

    def forward(self, obs: ObsType, coin: Optional[np.ndarray] = None,
                noise: Optional[np.ndarray] = None) -> ActType:
        """core/agent/gradient.py:96-124.  ``coin`` ([B,] M in {0,1}) / ``noise`` ([B,] 2, M):
        explicitly injected draws (override ``rng``)."""
        host = isinstance(obs[0], np.ndarray)
        if host:
            # host-buffer path: the observation is uploaded into device staging tensors, chunk by chunk, inside
            # die_gradient_forward_host; everything else (state, injected draws) is handled as on the device path
            agents_np, medium_np = (np.ascontiguousarray(a, dtype=np.float64) for a in obs)
            agents, medium = self._host_staging(agents_np, medium_np)
        else:
            agents, medium = obs
        _, _, _, B, M = _split_obs((agents, medium))
        self._check(agents)
        self._check_medium(medium)
        H, W = medium.shape[-2:]
        self._lazy_init(B, M, agents.device)
        action = self._action_for(agents)

        coin_ptr = noise_ptr = None
        if self._discrete_turn:
            if coin is None and self._rng_mode == 'numpy':
                coin = np.stack([np.random.randint(0, 2, M) for _ in range(B)])
            if coin is not None:
                coin_ptr = self._upload('coin', np.asarray(coin).astype(np.uint8), (B, M), torch.uint8, agents.device)
        if self._needs_prev():
            if noise is None and self._rng_mode == 'numpy':
                noise = self._rng.normal(loc=0., scale=0.4, size=(B, 2, M))
            if noise is not None:
                noise_ptr = self._upload('noise', noise, (B, 2, M), torch.float64, agents.device)
        cells_ptr = None
        if self.record_sense_cells:
            if self._sense_cells is None or tuple(self._sense_cells.shape) != (B, M):
                self._sense_cells = torch.empty((B, M), dtype=torch.int32, device=agents.device)
            cells_ptr = self._sense_cells.data_ptr()

        p = self._params_cached()
        prev_ptr = self._prev_grad.data_ptr() if self._prev_grad is not None else None
        if host:
            hb = self._host
            with _lib.on_device(agents.device):
                _lib.check(self._lib.die_gradient_forward_host(
                    self._host_ctx, _lib.C.byref(p), H, W, M, B, agents_np.ctypes.data, medium_np.ctypes.data,
                    agents.data_ptr(), medium.data_ptr(), self._theta.data_ptr(), prev_ptr, action.data_ptr(),
                    hb['action'].data_ptr(), coin_ptr, noise_ptr, cells_ptr, self._seed, self._step,
                    torch.cuda.current_stream().cuda_stream))
            self.last_hints, self.last_speculated = (False, False), False
            self._step += 1
            torch.autograd.graph.increment_version(action)     # the kernel wrote it through its raw pointer
            out = hb['action'].numpy()
            out.flags.writeable = False        # Env.step may then use the device copy instead of uploading it again
            _hints.register_host_action(out, action)
            return out
        # The Env that produced this observation, if it provably did (die_b200/_hints.py): its cached cells and
        # published gradient replace gathers, and the move of the action is evaluated in the same launch.
        env = _hints.find_env(agents, medium) if self.use_env_hints else None
        with _lib.on_device(agents.device):
            stream = torch.cuda.current_stream().cuda_stream
            if env is not None:
                flags = env._forward_flags(agents, medium, want_gradient=True, speculate=self.fuse_move)
                step_arg = self._step
                if self._step_dev is not None:       # (die_b200/graph.py) the call counter lives on the device
                    flags |= _lib.FWD_STEP_ON_DEVICE
                    step_arg = self._step_dev.data_ptr()
                _lib.check(self._lib.die_env_forward_gradient(
                    env._handle, _lib.C.byref(p), agents.data_ptr(), medium.data_ptr(), self._theta.data_ptr(),
                    prev_ptr, action.data_ptr(), coin_ptr, noise_ptr, cells_ptr, flags,
                    self._seed, step_arg, stream))
            else:
                flags = 0
                if self._step_dev is not None:
                    raise RuntimeError("the device-resident call counter (die_b200.GraphedLoop) needs the producing Env: "
                                       "pass the observation that Env.step returned and keep use_env_hints on")
                if medium.dtype == torch.float32:           # a float32 medium without its env (hints off, or a copy)
                    _lib.check(self._lib.die_gradient_forward_f32(
                        _lib.C.byref(p), H, W, M, B, agents.data_ptr(), medium.data_ptr(), self._theta.data_ptr(),
                        prev_ptr, action.data_ptr(), coin_ptr, noise_ptr, cells_ptr, self._seed, self._step, stream))
                else:
                    _lib.check(self._lib.die_gradient_forward(
                        _lib.C.byref(p), H, W, M, B, agents.data_ptr(), medium.data_ptr(), self._theta.data_ptr(),
                        prev_ptr, action.data_ptr(), coin_ptr, noise_ptr, cells_ptr, None, None,
                        self._seed, self._step, stream))
        # the kernel wrote `action` through its raw pointer: make that visible to torch's version counter, so a
        # stale speculation can never be mistaken for this one (two envs sharing one agent, repeated forwards)
        torch.autograd.graph.increment_version(action)
        self.last_hints = (bool(flags & _lib.FWD_USE_GRADIENT), bool(flags & _lib.FWD_USE_CELLS))
        self.last_speculated = bool(flags & _lib.FWD_SPECULATE_MOVE)
        self.last_committed = bool(flags & _lib.FWD_COMMIT_MOVE)
        if self.last_speculated:
            env._note_speculation(action, self.last_committed)
        if flags & _lib.FWD_WRITE_COST:
            env._note_cost(action)
        self._step += 1
        return action


class PhysarumAgent(GradientAgent):
    """core/agent/gradient.py:138-219."""
    _discrete_turn = True

    def __init__(self,
                 max_agents: int = 10 ** 6,
                 scale: float = 0.005,
                 deposit: float = 4.0,
                 inertia: float = 0.0,
                 sense_offset: float = 0.03,
                 noise_scale: float = 0.0,
                 normalized_grad: bool = True,
                 grad_clip: Optional[float] = 1e-5,
                 turn_angle: float = 30,
                 sense_angle: float = 90,
                 turn_tolerance: float = 0.1,
                 *, rng: str = 'philox', seed: int = 0):
        super().__init__(max_agents, scale, deposit, inertia, sense_offset, noise_scale,
                         normalized_grad, grad_clip, rng=rng, seed=seed)
        self._init_params.update(turn_angle=turn_angle, sense_angle=sense_angle, turn_tolerance=turn_tolerance)
        self._turn_radians = float(np.radians(turn_angle))
        self._sense_radians = float(np.radians(sense_angle))
        self._rtol = float(turn_tolerance)

    def _initial_theta(self, prev_grad: np.ndarray) -> np.ndarray:
        """_discretize_grad, core/agent/gradient.py:162,165-166."""
        rads = super()._initial_theta(prev_grad)
        return (rads // self._turn_radians) * self._turn_radians
