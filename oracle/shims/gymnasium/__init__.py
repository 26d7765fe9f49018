"""Stand-in: the reference only subclasses gym.Env[ObsType, ActType] (core/env.py:64)."""
from typing import Generic, TypeVar

_O = TypeVar("_O")
_A = TypeVar("_A")


class Env(Generic[_O, _A]):
    pass
