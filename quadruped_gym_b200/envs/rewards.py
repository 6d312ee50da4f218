"""Fused reward / termination specs for ``VecQuadrupedEnv.reward_fns`` / ``termination_fns``.

The reference's ``reward_fns`` dict maps names to zero-argument Python callables
(/root/reference/src/envs/quadruped.py:97,170-175).  Here a dict value may be

* a ``FusedTerm`` -- evaluated inside the step kernel's epilogue (no extra launch), or
* any zero-argument callable returning a ``[N]`` torch tensor (evaluated on device tensors after
  the kernel, exactly where the reference evaluates its callables).

Each constructor cites the reference formula it implements.
"""
from __future__ import annotations

from dataclasses import dataclass

from .._lib import TERM_IDS


@dataclass(frozen=True)
class FusedTerm:
    term: str
    weight: float = 1.0
    param: float = 0.0

    @property
    def term_id(self) -> int:
        return TERM_IDS[self.term]


def alive_bonus(weight: float = 1.0) -> FusedTerm:
    """``return 1`` (walking_quad.py:286-290; README.md:71-72)."""
    return FusedTerm("alive", weight)


def ctrl_sq(weight: float = -0.1) -> FusedTerm:
    """``weight * np.sum(np.square(env.data.ctrl))`` (README.md:68-69)."""
    return FusedTerm("ctrl_sq", weight)


def forward_velocity(weight: float = 1.0) -> FusedTerm:
    """``env.data.qvel[0]`` -- state after the step (README.md:65-66)."""
    return FusedTerm("qvel_x", weight)


def forward_reward(weight: float = 1.0) -> FusedTerm:
    """``body_linvel[0] * body_pos[0]`` from sensordata (dummy_walking_quad.py:11-13)."""
    return FusedTerm("forward", weight)


def no_drift_reward(weight: float = 1.0) -> FusedTerm:
    """``abs(body_linvel[1] * body_pos[1])`` (dummy_walking_quad.py:15-17)."""
    return FusedTerm("drift", weight)


def control_cost(weight: float = 1.0, alpha: float = 0.8) -> FusedTerm:
    """``alpha * first_cost + (1-alpha) * sum((ctrl-previous_ctrl)^2)`` with the reference's quirk that
    ``previous_ctrl_cost`` is set on the first ever call and never updated (walking_quad.py:255-270)."""
    return FusedTerm("control_cost", weight, alpha)


def orientation_reward(weight: float = 1.0) -> FusedTerm:
    """``body_zaxis[2]`` (walking_quad.py:237-241)."""
    return FusedTerm("orientation", weight)


def body_height_cost(weight: float = 1.0, height: float = 0.12) -> FusedTerm:
    """``abs(body_pos[2] - height)`` (walking_quad.py:243-247)."""
    return FusedTerm("height_cost", weight, height)


def joint_posture_cost(weight: float = 1.0) -> FusedTerm:
    """``np.linalg.norm((ctrl - joint_centers) / nu)`` (walking_quad.py:249-253)."""
    return FusedTerm("posture_cost", weight)


def exp_orientation(weight: float = 1.0) -> FusedTerm:
    """``exp_dist(orientation_reward())`` (walking_quad.py:368; math_utils.py:4-5)."""
    return FusedTerm("exp_orientation", weight)


def exp_body_height(weight: float = 1.0, height: float = 0.13) -> FusedTerm:
    """``exp_dist(body_height_cost(height))`` (walking_quad.py:369)."""
    return FusedTerm("exp_height", weight, height)


@dataclass(frozen=True)
class FusedTermination:
    kind: str  # "time_limit" | "flip"


def time_limit() -> FusedTermination:
    """``data.time >= max_time`` (quadruped.py:149-151)."""
    return FusedTermination("time_limit")


def flip_termination() -> FusedTermination:
    """``sensordata[body_zaxis + 2] < 0`` (walking_quad.py:152-156)."""
    return FusedTermination("flip")
