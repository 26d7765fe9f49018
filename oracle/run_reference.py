"""Imports the reference's own source files from /root/reference over the stand-in packages in
oracle/shims (see oracle/shims/README.md).  TEST INFRASTRUCTURE; only usable where /root/reference
exists (this container, not the GPU box)."""
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("DIE_REFERENCE_ROOT", "/root/reference")
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


def load():
    """-> namespace with the reference's Env, Dynamics, BrownianAgent, PhysarumAgent, GradientAgent,
    ConstAgent, BoundaryCondition, zero_cost (the real classes, unmodified)."""
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} not present")
    if not hasattr(np, "float"):
        np.float = float                      # core/data_init.py:215 uses the alias NumPy 1.24 removed
    for p in (SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import core.env as env
    import core.agent.static as static
    import core.agent.gradient as gradient
    import core.data_init as data_init

    class NS:
        pass

    ns = NS()
    ns.Env, ns.Dynamics, ns.BoundaryCondition, ns.zero_cost = env.Env, env.Dynamics, env.BoundaryCondition, env.zero_cost
    ns.BrownianAgent, ns.ConstAgent = static.BrownianAgent, static.ConstAgent
    ns.PhysarumAgent, ns.GradientAgent = gradient.PhysarumAgent, gradient.GradientAgent
    ns.WaveSequence = data_init.WaveSequence
    return ns
