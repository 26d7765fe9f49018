"""die_b200 -- the per-step environment + agent dynamics of gkirgizov/die as hand-written
sm_100a CUDA kernels behind the reference's own Python API (Env.step / Agent.forward).

    from die_b200 import Env, Dynamics, BrownianAgent, PhysarumAgent

Importing the package needs the built ``libdie_sm100a.so`` (``python -m die_b200._build``);
there is no CPU or PyTorch fallback."""
from . import _lib
from .base_types import DataChannels, channel
from .env import Env, Dynamics, BoundaryCondition, linear_action_cost, zero_cost
from .data_init import WaveSequence, FieldSequence, TabulatedSequence, PerlinNoiseSequence
from .agent import (Agent, ConstAgent, BrownianAgent, GradientAgent, PhysarumAgent, ConvolutionModel, NeuralAutomataAgent,
                    JonesAgent)
from .graph import GraphedLoop
from .evolve import PopulationEvaluator, PGPE

_lib.load()     # fail loudly at import time if the CUDA library is missing

__all__ = ['Env', 'Dynamics', 'BoundaryCondition', 'linear_action_cost', 'zero_cost',
           'Agent', 'ConstAgent', 'BrownianAgent', 'GradientAgent', 'PhysarumAgent', 'ConvolutionModel', 'NeuralAutomataAgent',
           'JonesAgent', 'GraphedLoop', 'PopulationEvaluator', 'PGPE', 'DataChannels', 'channel', 'WaveSequence', 'FieldSequence', 'TabulatedSequence', 'PerlinNoiseSequence']
__version__ = '0.1.0'
