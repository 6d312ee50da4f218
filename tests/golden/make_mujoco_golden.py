"""Dump golden vectors from the REAL MuJoCo -- run this wherever `mujoco` imports:

    python tests/golden/make_mujoco_golden.py [/path/to/quadruped-gym/src/models/quadruped/scene.xml]

It writes tests/golden/mujoco_steps.npz: the model blob exported from MjModel, >= 1000 teacher-forcing samples
(state + ctrl in, state + sensordata + qacc + ncon/nefc out after one mj_step) drawn from flight, landing, standing and
stumbling, and 50-step open-loop rollouts.  With that file present, tests/test_mujoco_golden.py pins the CPU oracle
(and through it the CUDA path) against MuJoCo even where the wheel is not installed -- it is the missing pin that
DESIGN.md section 2 calls "parity unpinned".  This script cannot run in the build container (no wheel, no network).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    import mujoco
    from quadruped_gym_b200.model.export_mujoco import export
    scene = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/models/quadruped/scene.xml"
    m = mujoco.MjModel.from_xml_path(scene)
    d = mujoco.MjData(m)
    rng = np.random.default_rng(0)
    keys = ("qpos", "qvel", "act", "ctrl", "qacc_warmstart")
    ins = {k: [] for k in keys + ("time",)}
    outs = {k: [] for k in ("qpos", "qvel", "act", "sensordata", "qacc", "ncon", "nefc", "niter")}
    for ep in range(40):
        mujoco.mj_resetData(m, d)
        d.ctrl[:] = [0, 0, -0.5] * 4
        if ep % 4 == 3:   # tumbling starts
            q = rng.normal(size=4)
            d.qpos[3:7] = q / np.linalg.norm(q)
            d.qpos[2] = 0.25
        for t in range(300):
            if t % 10 == 0:
                ctrl = rng.uniform(-1, 1, 12)
            d.ctrl[:] = ctrl
            if rng.random() < 0.1:
                for k in keys:
                    ins[k].append(getattr(d, k).copy())
                ins["time"].append(d.time)
                mujoco.mj_step(m, d)
                for k in ("qpos", "qvel", "act", "sensordata", "qacc"):
                    outs[k].append(getattr(d, k).copy())
                outs["ncon"].append(d.ncon); outs["nefc"].append(d.nefc); outs["niter"].append(int(d.solver_niter[0]))
            else:
                mujoco.mj_step(m, d)
    roll_ctrl = rng.uniform(-1, 1, (32, 50, 12))
    roll_obs = np.zeros((32, 50, m.nsensordata))
    for s in range(32):
        mujoco.mj_resetData(m, d)
        for t in range(50):
            d.ctrl[:] = roll_ctrl[s, t]
            for _ in range(4):
                mujoco.mj_step(m, d)
            roll_obs[s, t] = d.sensordata
    blob = np.frombuffer(export(m).to_blob(), dtype=np.uint8)
    path = os.path.join(HERE, "mujoco_steps.npz")
    np.savez_compressed(path, version=np.array(mujoco.__version__), blob=blob, roll_ctrl=roll_ctrl, roll_obs=roll_obs,
                        **{"in_" + k: np.array(v) for k, v in ins.items()}, **{"out_" + k: np.array(v) for k, v in outs.items()})
    print("wrote", path, "samples", len(ins["time"]), "mujoco", mujoco.__version__)


if __name__ == "__main__":
    main()
