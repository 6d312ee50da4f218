"""Golden traces produced by the reference's own Python classes (tests/golden/make_env_golden.py) pin
the env-level semantics: clip, frame_skip loop, lagged sensordata, reward dict sum, fp64 time limit,
flip termination, reset.  Here the CPU oracle + the numpy restatement in tests/ref_formulas.py replay
them exactly; the GPU tests then use the same restatement against the device.  CPU only."""
import os

import numpy as np
import pytest

from oracle.oracle import OracleData
from tests import ref_formulas as F

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "env_traces.npz"))


def test_trace_A_base_env_orchestration(oracle_model):
    d = OracleData(oracle_model)
    d.ctrl[:] = [0, 0, -0.5] * 4                       # quadruped.py:124
    assert np.array_equal(G["A_obs0"], np.zeros(33))   # reset() returns the zeroed sensordata
    for t in range(len(G["A_actions"])):
        a = np.clip(G["A_actions"][t], -1.0, 1.0)      # quadruped.py:160 (float32 clip)
        d.env_step(a.astype(np.float64), 4)
        assert np.array_equal(d.sensordata, G["A_obs"][t])
        total, comps = F.readme_reward(d.qvel, d.ctrl)
        assert total == G["A_reward"][t]
        assert np.array_equal(np.array(comps, dtype=np.float64), G["A_components"][t])
        assert d.time == G["A_time"][t]
        assert F.time_limit(d.time, 10.0) == G["A_terminated"][t]
    # the observation lags the state: sensordata joint positions != qpos after the step
    assert not np.array_equal(d.sensordata[:12], d.qpos[7:])


def test_trace_B_time_limit_indices():
    assert int(G["B4_first_terminated_step"]) == 1250    # frame_skip 4, max_time 10
    assert int(G["B10_first_terminated_step"]) == 1001   # frame_skip 10, max_time 20: fp64 sum falls short at 10,000
    for fs, mt, want in ((4, 10.0, 1250), (10, 20.0, 1001)):
        t, n = 0.0, 0
        while True:
            for _ in range(fs):
                t += 0.002
            n += 1
            if t >= mt:
                break
        assert n == want


def test_trace_C_fused_terms_match_reference_values():
    keys = list(G["C_keys"])
    cc = F.ControlCost()
    for t in range(len(G["C_obs"])):
        obs, ctrl = G["C_obs"][t], G["C_ctrl"][t]
        want = dict(zip(keys, G["C_terms"][t]))
        assert 10.0 * 1 == want["alive_bonus"]
        assert -2.0 * cc(ctrl) == want["control_cost"]
        assert 10.0 * F.exp_dist(F.orientation_reward(obs)) == want["orientation_reward"]
        assert -50.0 * F.exp_dist(F.body_height_cost(obs, 0.13)) == want["body_height_cost"]
        assert -1.0 * F.joint_posture_cost(ctrl) == want["joint_posture_cost"]
        assert (obs[29] < 0) == G["C_terminated"][t] or G["C_terminated"][t] == (obs[29] < 0)


def test_trace_C_full_walking_reward_stack_restatement():
    """tests/ref_formulas.WalkingRewardRef (ideal position, estimator, 11 terms, Python sum) == the reference's
    WalkingQuadrupedEnv on every step of trace C, bit for bit."""
    w = F.WalkingRewardRef(velocity=G["C_cmd_velocity"], heading=G["C_cmd_heading"], global_velocity=G["C_global_velocity"])
    assert w.window == 250
    for t in range(len(G["C_obs"])):
        total, vals = w.step(G["C_obs"][t], G["C_ctrl"][t])
        assert np.array_equal(vals, G["C_terms"][t]) and total == G["C_reward"][t]


@pytest.mark.parametrize("tag", ["D80", "D95"])
def test_trace_D_flip_termination(oracle_model, tag):
    d = OracleData(oracle_model)
    d.ctrl[:] = [0, 0, -0.5] * 4
    half = 0.5 * np.deg2rad(float(G[tag + "_roll_deg"]))
    d.qpos[3:7] = [np.cos(half), np.sin(half), 0.0, 0.0]
    for t in range(len(G[tag + "_terminated"])):
        d.env_step(np.zeros(12), 4)
        assert d.sensordata[29] == G[tag + "_zaxis_z"][t]
        assert F.flip_termination(d.sensordata) == G[tag + "_terminated"][t]
    assert G[tag + "_terminated"].any()
