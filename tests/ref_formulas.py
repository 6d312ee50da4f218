"""numpy (float64) restatement of the reference's reward / termination formulas used by the tests.

Each function cites the reference lines it restates; tests/test_golden_traces.py pins every one of them
against traces produced by the reference's own Python (tests/golden/make_env_golden.py), so the GPU
tests can apply them to device-produced sensordata / ctrl.
"""
import numpy as np

JOINT_CENTERS = np.array([0.0, 0.0, -0.5] * 4, dtype=np.float32)   # walking_quad.py:36-39


def exp_dist(x):                                   # math_utils.py:4-5
    return np.exp(x) - 1


def readme_reward(qvel_after, ctrl):
    """README.md:65-78 trio summed in dict order by QuadrupedEnv.step (quadruped.py:170-175)."""
    comps = [qvel_after[0], -0.1 * np.sum(np.square(ctrl)), 1.0]
    total = 0.0
    for c in comps:
        total += c
    return total, comps


class ControlCost:
    """walking_quad.py:255-270 incl. the never-updated previous_ctrl_cost."""

    def __init__(self):
        self.previous_ctrl = JOINT_CENTERS.copy()
        self.previous_ctrl_cost = None

    def reset(self):                               # walking_quad.py:106 (previous_ctrl_cost survives)
        self.previous_ctrl = JOINT_CENTERS.copy()

    def __call__(self, ctrl, alpha=0.8):
        diff = ctrl - self.previous_ctrl
        self.previous_ctrl = np.copy(ctrl)
        cost = np.sum(np.square(diff))
        if self.previous_ctrl_cost is None:
            self.previous_ctrl_cost = cost
        return alpha * self.previous_ctrl_cost + (1 - alpha) * cost


def orientation_reward(obs):                       # walking_quad.py:237-241
    return obs[29]


def body_height_cost(obs, height=0.12):            # walking_quad.py:243-247
    return np.abs(obs[20] - height)


def joint_posture_cost(ctrl):                      # walking_quad.py:249-253
    return np.linalg.norm((ctrl - JOINT_CENTERS) / 12)


def forward_reward(obs):                           # dummy_walking_quad.py:11-13
    return obs[21] * obs[18]


def no_drift_reward(obs):                          # dummy_walking_quad.py:15-17
    return np.abs(obs[22] * obs[19])


def flip_termination(obs):                         # walking_quad.py:152-156
    return obs[29] < 0


def time_limit(time, max_time):                    # quadruped.py:149-151
    return time >= max_time
