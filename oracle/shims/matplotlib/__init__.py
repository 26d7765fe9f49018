"""Stand-in: imported at module top by the reference (core/env.py:8-9, core/utils.py:7,
core/render.py:4), never used by Env.step / Agent.forward."""
from . import pyplot, animation, cm, image  # noqa: F401
