"""Runs only where the real `mujoco` wheel imports (not in the build container, not on the GPU boxes):
the CPU oracle against mujoco.mj_step on identical states and controls, and the in-repo MJCF compiler against
MuJoCo's compiled model.  This is the test that un-pins "parity unpinned"."""
import os

import numpy as np
import pytest

mujoco = pytest.importorskip("mujoco")

from oracle.oracle import OracleData, OracleModel  # noqa: E402
from quadruped_gym_b200.model import compile_mjcf  # noqa: E402
from quadruped_gym_b200.model.export_mujoco import export  # noqa: E402

SCENE = os.environ.get("QG_SCENE", "/root/reference/src/models/quadruped/scene.xml")
pytestmark = pytest.mark.skipif(not os.path.exists(SCENE), reason="reference MJCF not present")


@pytest.fixture(scope="module")
def mj():
    m = mujoco.MjModel.from_xml_path(SCENE)
    return m, mujoco.MjData(m)


def test_sizes_and_compiler_constants(mj):
    m, _ = mj
    cm = compile_mjcf(SCENE)
    assert (m.nq, m.nv, m.nu, m.nbody, m.nsensordata) == (19, 18, 12, 14, 33)
    assert np.allclose(cm["qpos0"], m.qpos0) and np.allclose(cm["body_mass"], m.body_mass, rtol=1e-9)
    assert np.allclose(cm["dof_damping"], m.dof_damping) and np.allclose(cm["dof_armature"], m.dof_armature)
    assert np.allclose(cm["jnt_range"], m.jnt_range)
    # mesh-inertia dependent: report, then require agreement for at least one supported mode
    ok = False
    for mode in ("legacy", "convex", "exact"):
        c2 = compile_mjcf(SCENE, mesh_inertia=mode)
        if np.allclose(c2["body_ipos"], m.body_ipos, atol=1e-6):
            ok = True
            print("mesh inertia mode matching MuJoCo", mujoco.__version__, "=", mode)
    assert ok


def test_oracle_step_matches_mj_step(mj):
    m, d = mj
    om = OracleModel(export(m).to_blob())
    rng = np.random.default_rng(0)
    o = OracleData(om)
    mujoco.mj_resetData(m, d)
    worst = 0.0
    for t in range(400):
        if t % 20 == 0:
            ctrl = rng.uniform(-1, 1, 12)
        # teacher forcing: the oracle starts every step from MuJoCo's state
        o.set_state(d.qpos.copy(), d.qvel.copy(), d.act.copy(), d.qacc_warmstart.copy(), d.time, ctrl)
        d.ctrl[:] = ctrl
        mujoco.mj_step(m, d)
        o.step()
        tol = 1e-4 if d.ncon == 0 else 1e-3
        assert np.abs(o.qpos - d.qpos).max() <= tol * max(1.0, np.abs(d.qpos).max())
        assert np.abs(o.qvel - d.qvel).max() <= tol * max(1.0, np.abs(d.qvel).max())
        s = np.abs(o.sensordata - d.sensordata)
        s[12:15] /= max(1.0, np.abs(d.qacc).max())
        assert s.max() <= tol
        worst = max(worst, float(np.abs(o.qvel - d.qvel).max()))
    print("worst |qvel - mujoco| over 400 teacher-forced steps:", worst)
