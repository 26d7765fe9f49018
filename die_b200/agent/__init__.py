from .base import Agent
from .static import ConstAgent, BrownianAgent
from .gradient import GradientAgent, PhysarumAgent
from .evo import ConvolutionModel, NeuralAutomataAgent
from .jones import JonesAgent

__all__ = ['Agent', 'ConstAgent', 'BrownianAgent', 'GradientAgent', 'PhysarumAgent', 'ConvolutionModel', 'NeuralAutomataAgent', 'JonesAgent']
