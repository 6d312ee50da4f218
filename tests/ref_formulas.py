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


class WalkingRewardRef:
    """WalkingQuadrupedEnv.step bookkeeping + input_control_reward (walking_quad.py:128-148,352-422) with the
    estimator of math_utils.py:11-133, restated for one environment."""

    def __init__(self, timestep=0.002, frame_skip=4, velocity=(0, 0, 0), heading=(0, 0, 0), global_velocity=(0, 0, 0)):
        self.timestep, self.frame_skip = timestep, frame_skip
        self.dt = timestep * frame_skip
        self.window = int(np.ceil(2 / (1 * self.dt)))
        self.velocity, self.heading = np.array(velocity, dtype=float), np.array(heading, dtype=float)
        self.global_velocity = np.array(global_velocity, dtype=float)
        self.ideal_position = np.zeros(3)
        self.cc = ControlCost()
        self.prev_derive = None
        self.prev_ctrl_for_estimator = JOINT_CENTERS.astype(np.float64)      # data.ctrl after reset
        n = 12
        self.cross_buf = np.zeros((self.window, n), dtype=int)
        self.sig_buf = np.zeros((self.window, n))
        self.idx, self.cross_count, self.sample_count = 0, np.zeros(n, dtype=int), 0
        self.prev_sample, self.prev_sign = None, None
        self.f_est, self.a_est = np.zeros(n), np.zeros(n)

    def reset(self):                                  # walking_quad.py:96-126 (estimator and first cost survive)
        self.ideal_position = np.zeros(3)
        self.cc.reset()
        self.prev_derive = None
        self.prev_ctrl_for_estimator = JOINT_CENTERS.astype(np.float64)

    def _estimator_update(self, x):
        x = np.asarray(x, dtype=float)
        if self.prev_sample is None:
            self.prev_sample = x.copy()
            self.sig_buf[self.idx] = x
            self.sample_count = 1
            self.idx = (self.idx + 1) % self.window
            return
        diff = x - self.prev_sample
        sign = np.sign(diff)
        if self.prev_sign is not None:
            z = sign == 0
            sign[z] = self.prev_sign[z]
            crossing = (sign != self.prev_sign).astype(int)
        else:
            crossing = np.zeros(12, dtype=int)
        if self.sample_count < self.window:
            self.sample_count += 1
        self.cross_count -= self.cross_buf[self.idx]
        self.cross_buf[self.idx] = crossing
        self.cross_count += crossing
        self.sig_buf[self.idx] = x
        self.idx = (self.idx + 1) % self.window
        self.prev_sample, self.prev_sign = x.copy(), sign.copy()
        f_cur = (self.cross_count / 2.0) / (self.sample_count * self.dt)
        self.f_est = 0.8 * self.f_est + (1 - 0.8) * f_cur
        w = self.sig_buf[: self.sample_count] if self.sample_count < self.window else self.sig_buf
        self.a_est = 0.8 * self.a_est + (1 - 0.8) * (np.max(w, axis=0) - np.min(w, axis=0))

    def step(self, obs, ctrl):
        """obs / ctrl AFTER the physics of this env.step(); returns (total, 11 values)."""
        self.ideal_position = self.ideal_position + self.global_velocity * self.timestep * self.frame_skip
        self._estimator_update(self.prev_ctrl_for_estimator)
        self.prev_ctrl_for_estimator = np.asarray(ctrl, dtype=float).copy()
        unit = lambda x: x / np.linalg.norm(x)
        with np.errstate(invalid="ignore", divide="ignore"):
            body_vel, cmd = obs[30:32], self.velocity[:2]
            vals = [
                10.0 * 1,
                -2.0 * self.cc(np.asarray(ctrl, dtype=float)),
                10.0 * np.dot(unit(body_vel), unit(cmd)),
                -50.0 * np.square(np.linalg.norm(body_vel) - np.linalg.norm(cmd)),
                10.0 * exp_dist(np.dot(obs[24:26], self.heading[:2])),
                10.0 * exp_dist(obs[29]),
                -50.0 * exp_dist(np.abs(obs[20] - 0.13)),
                -1.0 * np.linalg.norm((ctrl - JOINT_CENTERS) / 12),
                -2.5 * np.linalg.norm((self.a_est - np.array([1.5, 0.5, 0.0] * 4, dtype=np.float32)) / 12),
                -8.0 * np.linalg.norm((self.f_est - np.array([1.0, 1.0, 0.0] * 4, dtype=np.float32)) / 12),
            ]
        r = -20.0 * np.linalg.norm(obs[18:20] - self.ideal_position[:2])
        if self.prev_derive is None:
            self.prev_derive = r
        vals.append((r - self.prev_derive) / (self.timestep * self.frame_skip))
        self.prev_derive = r
        return sum(np.array(vals)), np.array(vals)


class MadgwickRef:
    """ahrs.filters.Madgwick.updateIMU + Quaternion.to_angles as restated in SURVEY.md App. G (unpinned)."""

    def __init__(self, Dt, beta=0.033):
        self.Dt, self.beta = Dt, beta

    def update(self, q, gyr, acc):
        q, gyr, acc = np.asarray(q, float), np.asarray(gyr, float), np.asarray(acc, float)
        if np.linalg.norm(gyr) == 0:
            return q
        w, x, y, z = q
        qdot = 0.5 * np.array([-x * gyr[0] - y * gyr[1] - z * gyr[2], w * gyr[0] + y * gyr[2] - z * gyr[1],
                               w * gyr[1] - x * gyr[2] + z * gyr[0], w * gyr[2] + x * gyr[1] - y * gyr[0]])
        an = np.linalg.norm(acc)
        if an > 0:
            a = acc / an
            w, x, y, z = q / np.linalg.norm(q)
            f = np.array([2 * (x * z - w * y) - a[0], 2 * (w * x + y * z) - a[1], 2 * (0.5 - x * x - y * y) - a[2]])
            if np.linalg.norm(f) > 0:
                J = np.array([[-2 * y, 2 * z, -2 * w, 2 * x], [2 * x, 2 * w, 2 * z, 2 * y], [0, -4 * x, -4 * y, 0]])
                g = J.T @ f
                qdot = qdot - self.beta * g / np.linalg.norm(g)
        q = q + qdot * self.Dt
        return q / np.linalg.norm(q)

    @staticmethod
    def to_angles(q):
        w, x, y, z = q
        return np.array([np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)), np.arcsin(np.clip(2 * (w * y - z * x), -1, 1)),
                         np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))])
