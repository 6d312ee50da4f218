"""``from envs.po_walking_quad import POWalkingQuadrupedEnv`` (/root/reference/src/envs/po_walking_quad.py:8;
imported by train_quadruped.py:8 and eval_quadruped.py:3)."""
from quadruped_gym_b200.envs.single import POWalkingQuadrupedEnv  # noqa: F401
from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv  # noqa: F401
from quadruped_gym_b200.envs.sb3 import SB3VecEnv  # noqa: F401
