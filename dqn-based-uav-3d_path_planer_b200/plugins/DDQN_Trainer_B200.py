"""Trainer plug-in: double DQN (Trainer/DDQN_Trainer.py) on the B200 library."""
import uavrl_b200  # noqa: F401  (repository root must be on sys.path)
from uavrl_b200 import engine
from uavrl_b200.plugins._trainer_base import TrainerB200


class DDQN_Trainer_B200(TrainerB200):
    ALGO = engine.ALGO_DDQN
    TAG = "DDQN_"
