from .base import Agent
from .static import ConstAgent, BrownianAgent
from .gradient import GradientAgent, PhysarumAgent

__all__ = ['Agent', 'ConstAgent', 'BrownianAgent', 'GradientAgent', 'PhysarumAgent']
