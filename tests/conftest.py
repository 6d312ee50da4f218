import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_SCENE = "/root/reference/src/models/quadruped/scene.xml"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def blob():
    from quadruped_gym_b200.model import DEFAULT_BLOB
    with open(DEFAULT_BLOB, "rb") as fh:
        return fh.read()


@pytest.fixture(scope="session")
def oracle_model(blob):
    from oracle.oracle import OracleModel
    return OracleModel(blob)


def rollout_states(oracle_model, n_envs, n_steps, seed, frame_skip=4, action_scale=1.0, hold=5):
    """States visited by oracle rollouts under piecewise-constant random actions (drop, landing,
    standing, stumbling).  Returns dict of float64 arrays sampled at random times, one per env."""
    from oracle.oracle import OracleData
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("qpos", "qvel", "act", "warm", "ctrl", "time")}
    for e in range(n_envs):
        d = OracleData(oracle_model)
        d.ctrl[:] = [0, 0, -0.5] * 4
        stop = rng.integers(1, n_steps + 1)
        a = rng.uniform(-1, 1, 12) * action_scale
        for s in range(stop):
            if s % hold == 0:
                a = rng.uniform(-1, 1, 12) * action_scale
            d.env_step(a, frame_skip)
        out["qpos"].append(d.qpos.copy()); out["qvel"].append(d.qvel.copy()); out["act"].append(d.act.copy())
        out["warm"].append(d.qacc_warmstart.copy()); out["ctrl"].append(d.ctrl.copy()); out["time"].append(d.time)
    return {k: np.array(v) for k, v in out.items()}
