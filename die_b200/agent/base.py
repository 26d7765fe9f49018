"""Agent base (reference: core/agent/base.py:12-62).  Only the ``forward`` contract is on
the hot path; json ``save``/``load`` of constructor parameters is kept because it is
host-side and trivial."""
import io
import json
import os
from abc import ABC, abstractmethod
from typing import Any, Dict, Union

from ..base_types import ActType, AgtType, ObsType


class Agent(ABC):
    @abstractmethod
    def forward(self, obs: ObsType) -> ActType:
        """Act in the environment given observations (core/agent/base.py:13-16)."""

    def render(self):
        return [None]

    def init_params(self) -> Dict[str, Any]:
        """Parameters from which the agent can be reconstructed (core/agent/base.py:22-26)."""
        return dict(self._init_params)

    def save(self, file: Union[str, os.PathLike, io.IOBase]):
        data = json.dumps(self.init_params())
        if isinstance(file, (str, os.PathLike)):
            with open(file, 'w') as f:
                f.write(data)
        else:
            file.write(data)

    @classmethod
    def load(cls, file: Union[str, os.PathLike, io.IOBase]) -> 'Agent':
        if isinstance(file, (str, os.PathLike)):
            with open(file, 'r') as f:
                params = json.load(f)
        else:
            params = json.load(file)
        return cls(**params)

    # -- action post-processing helpers (core/agent/base.py:45-62); not used by the built-in policies there either -----
    @staticmethod
    def postprocess_action(agents: AgtType, action: ActType) -> ActType:
        return Agent._rescale_outputs(Agent._masked_alive(agents, action))

    @staticmethod
    def _rescale_outputs(action: ActType) -> ActType:
        """core/agent/base.py:52-55: identity."""
        return action

    @staticmethod
    def _masked_alive(agents: AgtType, action: ActType) -> ActType:
        """core/agent/base.py:58-62: action * (alive > 0), broadcast over the action channels."""
        alive = agents[..., 2, :] > 0
        return action * alive.unsqueeze(-2).to(action.dtype)
