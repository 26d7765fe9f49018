from .base import Agent
from .static import ConstAgent, BrownianAgent
from .gradient import GradientAgent, PhysarumAgent
from .evo import ConvolutionModel, NeuralAutomataAgent

__all__ = ['Agent', 'ConstAgent', 'BrownianAgent', 'GradientAgent', 'PhysarumAgent', 'ConvolutionModel', 'NeuralAutomataAgent']
