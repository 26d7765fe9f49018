"""NeuralAutomataAgent / ConvolutionModel -- drop-ins for core/agent/evo.py:45-209 on the per-step path.

The reference's learned policy is a stack of small circular convolutions over the medium (float32, no bias, a Tanh at
the end) whose output, sampled at every slot's cell and scaled by (scale, scale, deposit), is the action.  Here the
stack runs as shared-memory tiled stencil kernels (``die_conv_policy_forward``, die_b200/csrc/die_conv_kernels.cuh) on
the medium where it lives; the weights are float32 CUDA tensors in torch's Conv2d layout, ``state_dict`` keys are the
reference's (``kernels.<l>.weight``), so a model saved by the reference's ``TorchAgent.save`` loads here and vice versa.

What is NOT here (the reference has it, off the per-step path): autograd through the model (the reference detaches the
output before building the action, core/agent/evo.py:176), dropout (``p_agent_dropout`` must be 0: the reference's own
TODO says it is unfinished), boundaries other than 'circular'."""
from __future__ import annotations

import io
import os
from typing import Any, Dict, Optional, Sequence, Union

import numpy as np
import torch

from .. import _hints
from .. import _lib
from ..base_types import ActType, ObsType
from .static import _DeviceAgent, _split_obs

MAX_KERNEL = 7
MAX_CHANNELS = 4


class ConvolutionModel:
    """core/agent/evo.py:45-118: ``len(kernel_sizes)`` Conv2d layers (obs channels -> obs channels, the last one
    -> action channels; padding 'same', circular; no bias) followed by Tanh."""

    def __init__(self, num_obs_channels: int = 3, num_act_channels: int = 3, kernel_sizes: Sequence[int] = (3,),
                 boundary: str = 'circular', p_agent_dropout: float = 0., requires_grad: bool = True, device=None):
        if boundary != 'circular':
            raise NotImplementedError("ConvolutionModel: only the reference's default boundary 'circular' is compiled")
        if p_agent_dropout != 0.:
            raise NotImplementedError("ConvolutionModel: p_agent_dropout must be 0 (unfinished in the reference, too)")
        self.kernel_sizes = tuple(int(k) for k in kernel_sizes)
        if not self.kernel_sizes or any(k < 1 or k > MAX_KERNEL or k % 2 == 0 for k in self.kernel_sizes):
            raise NotImplementedError(f"ConvolutionModel: kernel sizes must be odd and <= {MAX_KERNEL}")
        if not (1 <= num_obs_channels <= MAX_CHANNELS and 1 <= num_act_channels <= MAX_CHANNELS):
            raise NotImplementedError(f"ConvolutionModel: at most {MAX_CHANNELS} channels")
        self.num_obs_channels, self.num_act_channels = int(num_obs_channels), int(num_act_channels)
        self._init_params = dict(num_obs_channels=num_obs_channels, num_act_channels=num_act_channels,
                                 kernel_sizes=list(self.kernel_sizes), boundary=boundary,
                                 p_agent_dropout=p_agent_dropout, requires_grad=requires_grad)
        self._lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        n = len(self.kernel_sizes)
        outs = [self.num_obs_channels] * (n - 1) + [self.num_act_channels]
        # torch's own Conv2d default initialisation, drawn in the reference's order from torch's global generator
        # (th.manual_seed reproduces the reference's weights)
        self.kernels = []
        for k, cout in zip(self.kernel_sizes, outs):
            conv = torch.nn.Conv2d(self.num_obs_channels, cout, k, bias=False)
            self.kernels.append(conv.weight.detach().to(torch.float32))
        self._flat = None
        self._population = None         # [P, num_parameters] float32 on the device: one model per environment of a batch
        self._scratch = None
        self._upload()

    # -- weights ------------------------------------------------------------------------------
    def _upload(self) -> None:
        self._flat = torch.cat([w.reshape(-1) for w in self.kernels]).to(self.device, dtype=torch.float32).contiguous()

    def init_weights(self) -> None:
        """core/agent/evo.py:107-110: xavier_uniform on every kernel."""
        for w in self.kernels:
            torch.nn.init.xavier_uniform_(w)
        self._upload()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {f'kernels.{l}.weight': w.clone() for l, w in enumerate(self.kernels)}

    def load_state_dict(self, state: Dict[str, torch.Tensor]) -> None:
        for l, w in enumerate(self.kernels):
            src = torch.as_tensor(state[f'kernels.{l}.weight'], dtype=torch.float32)
            if tuple(src.shape) != tuple(w.shape):
                raise ValueError(f"kernels.{l}.weight has shape {tuple(src.shape)}, expected {tuple(w.shape)}")
            w.copy_(src)
        self._upload()

    @property
    def num_parameters(self) -> int:
        return sum(int(w.numel()) for w in self.kernels)

    def set_population(self, vectors) -> None:
        """One weight vector per environment of a batch (``[P, num_parameters]``, the layout of ``parameters_vector``):
        ``forward`` on a batch of P environments then evaluates model p on environment p -- what a neuro-evolution
        search scores (the reference's examples/learning_agents.py runs its candidates one after the other).  A
        population of the same size is written IN PLACE, so a captured CUDA graph (die_b200.GraphedLoop) sees it.
        ``None`` returns to the single model in ``kernels``."""
        if vectors is None:
            self._population = None
            return
        v = torch.as_tensor(vectors, dtype=torch.float32)
        if v.dim() != 2 or v.shape[1] != self.num_parameters:
            raise ValueError(f"population must be [P, {self.num_parameters}], got {tuple(v.shape)}")
        if self._population is not None and tuple(self._population.shape) == tuple(v.shape):
            self._population.copy_(v, non_blocking=False)
        else:
            self._population = v.to(self.device).contiguous().clone()

    @property
    def population(self) -> Optional[torch.Tensor]:
        return self._population

    def parameters_vector(self) -> torch.Tensor:
        """All weights as one float32 vector (what a neuro-evolution search perturbs); see ``set_parameters_vector``."""
        return torch.cat([w.reshape(-1) for w in self.kernels])

    def set_parameters_vector(self, vec) -> None:
        vec = torch.as_tensor(vec, dtype=torch.float32).reshape(-1).cpu()
        pos = 0
        for w in self.kernels:
            w.copy_(vec[pos:pos + w.numel()].reshape(w.shape))
            pos += w.numel()
        if pos != vec.numel():
            raise ValueError(f"expected {pos} parameters, got {vec.numel()}")
        self._upload()

    # -- evaluation -----------------------------------------------------------------------------
    def _buffers(self, B: int, H: int, W: int):
        ch = max(self.num_obs_channels, self.num_act_channels)
        shape = (B, ch, H, W)
        if self._scratch is None or tuple(self._scratch[0].shape) != shape:
            self._scratch = [torch.empty(shape, dtype=torch.float32, device=self.device) for _ in range(2)]
        return self._scratch

    def _run(self, medium: torch.Tensor, in_total: int, in_ch0: int, B: int, H: int, W: int, M: int = 1,
             agents=None, cells_ptr=None, coefs=None, action=None) -> torch.Tensor:
        sa, sb = self._buffers(B, H, W)
        ks = (_lib.C.c_int32 * len(self.kernel_sizes))(*self.kernel_sizes)
        cf = (_lib.C.c_float * 3)(*(coefs or (1.0, 1.0, 1.0)))
        final = _lib.C.c_int32(0)
        dtype = _lib.FIELD_F32 if medium.dtype == torch.float32 else _lib.FIELD_F64
        weights, stride = self._flat, 0
        if self._population is not None:
            if self._population.shape[0] != B:
                raise ValueError(f"a population of {self._population.shape[0]} models needs a batch of as many "
                                 f"environments, got {B}")
            weights, stride = self._population, self.num_parameters
        with _lib.on_device(self.device):
            _lib.check(self._lib.die_conv_policy_forward_population(
                H, W, M, B, dtype, medium.data_ptr(), in_total, in_ch0, self.num_obs_channels, self.num_act_channels,
                len(self.kernel_sizes), ks, weights.data_ptr(), stride, sa.data_ptr(), sb.data_ptr(),
                agents.data_ptr() if agents is not None else None, cells_ptr, cf,
                action.data_ptr() if action is not None else None, _lib.C.byref(final),
                torch.cuda.current_stream().cuda_stream))
        out = (sa, sb)[final.value]
        return out[:, :self.num_act_channels]

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        """core/agent/evo.py:112-118 on a [B, C, H, W] (or [C, H, W]) float32 / float64 CUDA tensor -> [B, A, H, W] float32."""
        x = input if input.dim() == 4 else input[None]
        if not x.is_cuda:
            x = x.to(self.device)
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float32)
        B, Cin, H, W = x.shape
        if Cin != self.num_obs_channels:
            raise ValueError(f"input has {Cin} channels, the model takes {self.num_obs_channels}")
        return self._run(x, Cin, 0, B, H, W).clone()

    __call__ = forward


class NeuralAutomataAgent(_DeviceAgent):
    """core/agent/evo.py:117-209."""

    def __init__(self, scale: float = 0.1, deposit: float = 1.0, with_agent_channel: bool = True,
                 initial_obs: Optional[ObsType] = None, **model_kwargs):
        super().__init__()
        self._init_params = dict(scale=scale, deposit=deposit, with_agent_channel=with_agent_channel, initial_obs=None,
                                 model_kwargs=dict(model_kwargs))
        self._with_agent_channel = bool(with_agent_channel)
        n_obs = 3 if with_agent_channel else 2                     # DataChannels.medium or medium[1:] (:128-129)
        self._model = ConvolutionModel(num_obs_channels=n_obs, num_act_channels=3, **model_kwargs)
        self.action_coefs = (float(np.float32(scale)), float(np.float32(scale)), float(np.float32(deposit)))
        self._sense_output = None
        self.use_env_hints = True
        self._step_dev = None           # (no random draws: nothing to keep on the device for a captured graph, see GraphedLoop)
        self._rng = 'philox'

    @property
    def model(self) -> ConvolutionModel:
        return self._model

    @property
    def init_params(self) -> Dict[str, Any]:
        """A property here, as in the reference's TorchAgent family (core/agent/evo.py:146-148)."""
        return dict(self._init_params)

    def forward(self, obs: ObsType) -> ActType:
        """core/agent/evo.py:150-178: model(medium) in float32, sampled at every slot's cell (all M slots), rescaled."""
        if isinstance(obs[0], np.ndarray):
            return self._forward_host(obs)
        agents, medium, _, B, M = _split_obs(obs)
        self._check(agents)
        self._check_medium(medium)
        H, W = medium.shape[-2:]
        action = self._action_for(agents)
        env = _hints.find_env(agents, medium) if self.use_env_hints else None
        cells_ptr = None
        if env is not None:
            _grad, cells_ptr = env._hints_for(agents, medium, want_gradient=False)
        sense = self._model._run(medium, 3, 0 if self._with_agent_channel else 1, B, H, W, M=M, agents=agents,
                                 cells_ptr=cells_ptr, coefs=self.action_coefs, action=action)
        self._sense_output = sense
        torch.autograd.graph.increment_version(action)
        self._step += 1
        return action

    def set_population(self, vectors) -> None:
        """See ``ConvolutionModel.set_population``."""
        self._model.set_population(vectors)

    def render(self):
        """core/agent/evo.py:180-185: the model output with channels last."""
        if self._sense_output is None:
            return [None]
        data = self._sense_output[0] if self._sense_output.dim() == 4 else self._sense_output
        return [torch.moveaxis(data, 0, -1).cpu().numpy()]

    # -- persistence in the reference's format (core/agent/evo.py:24-42) -----------------------------------------------
    def save(self, file: Union[str, os.PathLike, io.IOBase]):
        params = {k: v for k, v in self._init_params.items() if k != 'initial_obs'}
        torch.save(dict(params_dict=params, model_state=self._model.state_dict()), file)

    @classmethod
    def load(cls, file: Union[str, os.PathLike, io.IOBase]) -> 'NeuralAutomataAgent':
        loaded = torch.load(file, weights_only=False)
        params = dict(loaded['params_dict'])
        kwargs = params.pop('model_kwargs', {}) or {}
        params.pop('initial_obs', None)
        agent = cls(**params, **kwargs)
        agent.model.load_state_dict(loaded['model_state'])
        return agent
