"""``from envs.walking_quad import WalkingQuadrupedEnv`` (/root/reference/src/envs/walking_quad.py:9)."""
from quadruped_gym_b200.envs.single import WalkingQuadrupedEnv  # noqa: F401
from quadruped_gym_b200.envs.walking_quad import VecWalkingQuadrupedEnv  # noqa: F401
