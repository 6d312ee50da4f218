"""``from envs.quadruped import QuadrupedEnv`` (/root/reference/src/envs/quadruped.py:9)."""
from quadruped_gym_b200.envs.quadruped import QuadrupedEnv, VecQuadrupedEnv  # noqa: F401
