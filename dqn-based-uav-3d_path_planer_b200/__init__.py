"""uavrl_b200 -- B200-native hot path of the RLGF UAV path planner (env step + DQN-family learner).

The directory name carries the reference repository's name and is not a valid Python identifier:
import it through the `uavrl_b200` shim at the repository root.
"""
from . import _build, _lib  # noqa: F401
from ._lib import (ACT_CONT_F32, ACT_CONT_F64, ACT_DISCRETE27, ALGO_DDQN, ALGO_DQN, ALGO_DUELING,  # noqa: F401
                   INFO_NAMES, OBS_DIM, UavrlError)


def build(force=False, verbose=False):
    return _build.build(force=force, verbose=verbose)
