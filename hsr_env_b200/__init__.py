"""hsr_env_b200 - B200-native batched physics backend for the HSR block-manipulation environment.

Only the hot path of the reference is rebuilt here: the ``sim.step()`` x300 loop inside
``HSREnv.step`` (/root/reference/hsr/env.py:115-135) for N independent environments, behind the same env API.
"""
from .model import Model  # noqa: F401
