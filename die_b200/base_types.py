"""Layout contract of the hot path (reference: core/base_types.py:8-36).

The reference passes labelled ``xarray.DataArray`` objects; here the same channel-major
arrays are ``torch`` CUDA tensors (float64) -- or pinned host ``numpy`` arrays on the
host-buffer path -- in the identical channel order:

    medium  [3, H, W]   (agents, env_food, chem1)
    agents  [4, M]      (x, y, alive, agent_food)
    action  [3, M]      (dx, dy, deposit1)

A batch of B independent environments adds a leading axis ([B, 3, H, W], ...).
"""
from typing import Sequence, Tuple, Union

import numpy as np
import torch

Array = Union[torch.Tensor, np.ndarray]
ActType = Array
AgtType = Array
MediumType = Array
ObsType = Tuple[AgtType, MediumType]
Channels = Sequence[str]


class DataChannels:
    medium: Channels = ('agents', 'env_food', 'chem1')
    agents: Channels = ('x', 'y', 'alive', 'agent_food')
    actions: Channels = ('dx', 'dy', 'deposit1')


def channel(array: Array, kind: str, name: str) -> Array:
    """``array.sel(channel=name)`` of the reference: ``channel(medium, 'medium', 'chem1')``."""
    idx = getattr(DataChannels, kind).index(name)
    return array[..., idx, :, :] if kind == 'medium' else array[..., idx, :]
