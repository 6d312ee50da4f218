"""``from envs.control_inputs import VelocityHeadingControls`` (/root/reference/src/envs/control_inputs.py:3): the
device-backed command object of one environment."""
from quadruped_gym_b200.envs.single import SingleControls as VelocityHeadingControls  # noqa: F401
