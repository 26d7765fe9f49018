"""EnvRenderer / FieldTrace -- drop-ins for core/render.py:9-29, 76-132 with the frames computed on the device
(the step either side of the hot path).  One launch of ``die_render_frames`` per call
produces the medium frame, updates the agent trace and builds the agents frame; ``render_host`` additionally copies
them to pinned host memory (asynchronously, one synchronisation) for matplotlib / GIF writers.

The reference colour-maps the trace with matplotlib (``cm.get_cmap('magma')``); matplotlib is not a dependency of
this package, so the trace frame is returned as the scalar field unless a ``colormap`` callable is given."""
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


class FieldTrace:
    """core/render.py:9-29; the field lives on the device and is updated by ``EnvRenderer.render``."""

    def __init__(self, field_size: Tuple[int, int], trace_steps: int = 8, batch: int = 1, device=None):
        self._decay = 1 - 1 / trace_steps
        self._trace_field = torch.zeros((batch, *field_size), dtype=torch.float64, device=device)

    @property
    def trace(self) -> torch.Tensor:
        return self._trace_field

    def as_mask(self, inverse=False) -> torch.Tensor:
        return 1. - self._trace_field if inverse else self._trace_field


class EnvRenderer:
    """core/render.py:76-132."""

    field_colors = {'rgb': None, 'one': [0.19, -0.3, 0.74], 'two': [-0.45, 0.65, 0.83], 'random': None}

    def __init__(self, field_size: Tuple[int, int], is_trace_colored: bool = True, field_colors_id: str = 'rgb',
                 *, batch: int = 1, device=None, colormap: Optional[Callable[[np.ndarray], np.ndarray]] = None):
        self.field_size = (int(field_size[0]), int(field_size[1]))
        self._lib = _lib.load()
        self._is_trace_colored = is_trace_colored
        self._colormap = colormap
        self._B = int(batch)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        # RendererBase._set_colors (core/render.py:47-58)
        if field_colors_id == 'random':
            color = (np.random.random(3) - 0.5) * 2
        else:
            color = self.field_colors.get(field_colors_id)
        self._color = None
        if color is not None:
            color = np.asarray(color, dtype=np.float64)
            self._color = np.ascontiguousarray(color / np.linalg.norm(color))
        self._agent_trace = FieldTrace(self.field_size, batch=self._B, device=self.device)
        h, w = self.field_size
        self._img_medium = torch.empty((self._B, h, w, 3), dtype=torch.float64, device=self.device)
        self._img_agents = None
        self._host = None

    @property
    def img_shape(self) -> Tuple[int, int]:
        width, height = self.field_size
        return height, width

    def render(self, medium: torch.Tensor, agents: torch.Tensor) -> List[torch.Tensor]:
        """-> [medium frame (H, W, 3), trace (H, W), agents frame (W, M/W, 4)] as device tensors (leading batch
        axis for a batched env).  The tensors are re-used by the next call."""
        h, w = self.field_size
        B = self._B
        M = int(agents.shape[-1])
        batched = medium.dim() == 4
        if not (medium.is_cuda and medium.dtype == torch.float64 and medium.is_contiguous()
                and agents.is_cuda and agents.dtype == torch.float64 and agents.is_contiguous()):
            raise TypeError("render needs contiguous float64 CUDA tensors (there is no CPU fallback)")
        if medium.numel() != B * 3 * h * w:
            raise ValueError(f"medium has {medium.numel()} elements, expected {B}x3x{h}x{w}")
        with_agents = M == h * w                       # the reference's reshape (2, height, -1) needs it too
        if with_agents and (self._img_agents is None or self._img_agents.shape[1] != w):
            self._img_agents = torch.empty((B, w, M // w, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.die_render_frames(
                h, w, M, B, medium.data_ptr(), agents.data_ptr(), self._agent_trace.trace.data_ptr(),
                self._agent_trace._decay, self._color.ctypes.data if self._color is not None else None,
                self._img_medium.data_ptr(), self._img_agents.data_ptr() if with_agents else None,
                torch.cuda.current_stream().cuda_stream))
        frames = [self._img_medium, self._agent_trace.as_mask(), self._img_agents if with_agents else None]
        return frames if batched else [f[0] if f is not None else None for f in frames]

    def render_host(self, medium: torch.Tensor, agents: torch.Tensor) -> List[np.ndarray]:
        """``render`` + D2H into pinned buffers; the trace frame goes through ``colormap`` if one was given."""
        frames = self.render(medium, agents)
        if self._host is None:
            self._host = [torch.empty(f.shape, dtype=f.dtype).pin_memory() if f is not None else None for f in frames]
        with torch.cuda.device(self.device):
            for dst, src in zip(self._host, frames):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = [d.numpy() if d is not None else None for d in self._host]
        if self._colormap is not None:
            out[1] = self._colormap(out[1])
        return out
