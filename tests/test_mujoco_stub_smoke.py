"""Keeps the MuJoCo-facing code paths alive where the wheel cannot be installed: the exporter
``export(MjModel) -> blob``, the golden dump script and the gated parity suite all RUN against ``tests/mujoco_stub.py``
(field names / packed layouts of MjModel written from MuJoCo's public headers; the numbers underneath are the oracle's,
so these tests prove the plumbing, not parity).  Needs the reference MJCF (this container only)."""
import importlib
import importlib.util
import os
import sys

import numpy as np
import pytest

SCENE = "/root/reference/src/models/quadruped/scene.xml"
pytestmark = pytest.mark.skipif(not os.path.exists(SCENE), reason="reference MJCF not present")


@pytest.fixture()
def stub(monkeypatch):
    try:
        import mujoco  # noqa: F401
        pytest.skip("the real mujoco wheel is present: tests/test_mujoco_gated.py runs instead")
    except ImportError:
        pass
    from tests import mujoco_stub
    mj = mujoco_stub.build_module()
    monkeypatch.setitem(sys.modules, "mujoco", mj)
    return mj


def test_export_round_trips_the_compiled_model(stub):
    from quadruped_gym_b200.model import compile_mjcf
    from quadruped_gym_b200.model.export_mujoco import export
    cm = compile_mjcf(SCENE)
    ex = export(stub.MjModel.from_xml_path(SCENE))
    for k, v in cm.arrays.items():
        if k in ("mesh_cedge", "mesh_vert_cedge", "mesh_cedgeadr"):      # the climb graph is derived at load time
            continue
        assert k in ex.arrays, k
        a, b = np.asarray(v, dtype=float).ravel(), np.asarray(ex.arrays[k], dtype=float).ravel()
        assert a.shape == b.shape, k
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12), k
    # and the exported blob loads in the oracle and steps
    from oracle.oracle import OracleData, OracleModel
    d = OracleData(OracleModel(ex.to_blob()))
    d.ctrl[:] = [0, 0, -0.5] * 4
    for _ in range(80):
        d.step()
    assert np.isfinite(d.qpos).all() and d.ncon >= 0


def test_golden_dump_script_and_offline_pin_run(stub, tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    try:
        mod = importlib.import_module("make_mujoco_golden")
        importlib.reload(mod)
        monkeypatch.setattr(mod, "HERE", str(tmp_path))
        monkeypatch.setattr(sys, "argv", ["make_mujoco_golden.py", SCENE])
        mod.main()
    finally:
        sys.path.pop(0)
    out = tmp_path / "mujoco_steps.npz"
    G = np.load(out)
    assert len(G["in_time"]) >= 1000 and G["roll_obs"].shape == (32, 50, 33) and str(G["version"]) == "0.0.stub"
    import tests.test_mujoco_golden as pin
    monkeypatch.setattr(pin, "PATH", str(out))
    pin.test_oracle_matches_mujoco_golden_steps()          # oracle vs oracle through the dump: the plumbing holds


def test_gated_suite_runs_against_the_stub(stub):
    import tests.test_mujoco_gated as gated
    gated = importlib.reload(gated)
    m = stub.MjModel.from_xml_path(SCENE)
    gated.test_oracle_step_matches_mj_step((m, stub.MjData(m)))
    gated.test_sizes_and_compiler_constants((m, stub.MjData(m)))
