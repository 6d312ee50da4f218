"""If tests/golden/mujoco_steps.npz exists (produced by make_mujoco_golden.py on a machine with the `mujoco` wheel),
pin the CPU oracle against the real mj_step offline.  Skipped otherwise -- which is the state of this repository today
("parity unpinned")."""
import os

import numpy as np
import pytest

PATH = os.path.join(os.path.dirname(__file__), "golden", "mujoco_steps.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="no MuJoCo golden dump (mujoco is not installable offline)")


def test_oracle_matches_mujoco_golden_steps():
    from oracle.oracle import OracleData, OracleModel
    G = np.load(PATH)
    om = OracleModel(G["blob"].tobytes())
    n = len(G["in_time"])
    bad = 0
    for i in range(n):
        d = OracleData(om)
        d.set_state(G["in_qpos"][i], G["in_qvel"][i], G["in_act"][i], G["in_qacc_warmstart"][i], float(G["in_time"][i]), G["in_ctrl"][i])
        d.step()
        if d.ncon != int(G["out_ncon"][i]):
            bad += 1
            continue
        tol = 1e-4 if G["out_nefc"][i] == 0 else 1e-3
        assert np.abs(d.qpos - G["out_qpos"][i]).max() <= tol * max(1.0, np.abs(G["out_qpos"][i]).max())
        assert np.abs(d.qvel - G["out_qvel"][i]).max() <= tol * max(1.0, np.abs(G["out_qvel"][i]).max())
    assert bad <= 0.02 * n
