"""Generate golden env traces by running the REFERENCE's own Python classes
(/root/reference/src/envs/quadruped.py, walking_quad.py) in this container.

`mujoco` and `gymnasium` are not installable here, so the reference modules are imported against two
small stub modules: `gymnasium` (Env, spaces.Box) and `mujoco`, whose MjModel/MjData/mj_step/
mj_resetData are backed by the float64 CPU oracle (oracle/qg_oracle.c).  Everything ABOVE the five
MuJoCo entry points -- action clipping, the frame_skip loop, the sensordata copy, the reward_fns /
termination_fns dictionaries, WalkingQuadrupedEnv's 11-term reward with all its stateful quirks,
flip termination, the fp64 time limit, reset semantics -- is therefore executed by the unmodified
reference code.  The physics underneath is the oracle (parity unpinned vs the real wheel).

    python tests/golden/make_env_golden.py      ->  tests/golden/env_traces.npz

The script needs /root/reference and is NOT run on the GPU box; the .npz is committed.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"

from oracle.oracle import OracleData, OracleModel  # noqa: E402
from quadruped_gym_b200.model import SENSORS, compile_mjcf  # noqa: E402


def install_stubs():
    gym = types.ModuleType("gymnasium")

    class Env:
        def __init__(self, *a, **k):
            pass

    class Box:
        def __init__(self, low, high, shape, dtype=np.float32):
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
            self.shape, self.dtype = shape, dtype

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = Box
    gym.Env, gym.spaces = Env, spaces
    sys.modules["gymnasium"], sys.modules["gymnasium.spaces"] = gym, spaces

    mj = types.ModuleType("mujoco")

    class _Opt:
        timestep = 0.002

    class MjModel:
        def __init__(self, path):
            self.cm = compile_mjcf(path)
            self.om = OracleModel(self.cm.to_blob())
            self.nu, self.nsensordata = 12, 33
            self.opt = _Opt()
            self.opt.timestep = float(self.cm["opt_f"][0])
            self.sensor_names = list(SENSORS.keys())
            self.sensor_adr = np.array([SENSORS[n][0] for n in self.sensor_names])

        @staticmethod
        def from_xml_path(path):
            return MjModel(path)

    class MjData:
        def __init__(self, model):
            self.o = OracleData(model.om)
            self.qpos, self.qvel, self.ctrl, self.sensordata = self.o.qpos, self.o.qvel, self.o.ctrl, self.o.sensordata

        @property
        def time(self):
            return self.o.time

        @time.setter
        def time(self, v):
            self.o.time = v

    class _Flags(dict):
        def __getitem__(self, k):
            return False

    class MjvOption:
        def __init__(self):
            self.flags = {}
            self.frame = None
            self.geomgroup = np.zeros(6)

    class MjvCamera:
        pass

    class _Enum:
        def __getattr__(self, k):
            return k

    mj.MjModel, mj.MjData, mj.MjvOption, mj.MjvCamera = MjModel, MjData, MjvOption, MjvCamera
    mj.mjtVisFlag, mj.mjtFrame, mj.mjtObj, mj.mjtGeom = _Enum(), _Enum(), _Enum(), _Enum()
    mj.mj_resetData = lambda m, d: d.o.reset()
    mj.mj_step = lambda m, d: d.o.step()
    mj.mj_name2id = lambda m, typ, name: m.sensor_names.index(name)
    sys.modules["mujoco"] = mj


def main():
    install_stubs()
    sys.path.insert(0, REF_SRC)
    cwd = os.getcwd()
    os.chdir(REF_SRC)  # the reference resolves ./models/... relative to src/
    from envs.quadruped import QuadrupedEnv
    from envs.walking_quad import WalkingQuadrupedEnv

    out = {}
    rng = np.random.default_rng(2024)

    # --- trace A: base QuadrupedEnv, README-style reward lambdas (README.md:65-78), defaults frame_skip 4
    env = QuadrupedEnv()
    env.reward_fns = {
        "forward": lambda: env.data.qvel[0],
        "control_cost": lambda: -0.1 * np.sum(np.square(env.data.ctrl)),
        "alive_bonus": lambda: 1.0,
    }
    obs0, _ = env.reset()
    T = 300
    acts = (rng.uniform(-1.3, 1.3, (T, 12))).astype(np.float32)  # beyond +-1 to exercise the clip
    acts = np.repeat(acts[::5], 5, axis=0)
    obs, rew, term, tim, comps = [], [], [], [], []
    for t in range(T):
        o, r, te, tr, info = env.step(acts[t])
        obs.append(o); rew.append(r); term.append(te); tim.append(info["time"])
        comps.append([info["reward_components"][k] for k in ("forward", "control_cost", "alive_bonus")])
        assert tr is False
    out.update(A_actions=acts, A_obs0=obs0, A_obs=np.array(obs), A_reward=np.array(rew), A_terminated=np.array(term),
               A_time=np.array(tim), A_components=np.array(comps, dtype=np.float64))

    # --- trace B: time-limit termination index for (frame_skip, max_time) = (4, 10) and (10, 20)
    for tag, fs, mt in (("B4", 4, 10.0), ("B10", 10, 20.0)):
        env = QuadrupedEnv(frame_skip=fs, max_time=mt)
        env.reset()
        n = 0
        while True:
            _, _, te, _, _ = env.step(np.zeros(12, dtype=np.float32))
            n += 1
            if te:
                break
        out[tag + "_first_terminated_step"] = np.array(n)

    # --- trace C: WalkingQuadrupedEnv, full 11-term reward (walking_quad.py:352-422), fixed command
    wenv = WalkingQuadrupedEnv(frame_skip=4, max_time=10.0)
    wenv.control_inputs.set_orientation(0.3)
    wenv.control_inputs.set_velocity_speed_alpha(0.3, 0.1)
    wobs0, _ = wenv.reset()
    T = 400
    acts = np.repeat(rng.uniform(-1, 1, (T // 4, 12)).astype(np.float32), 4, axis=0)
    keys = WalkingQuadrupedEnv.reward_keys
    obs, rew, term, terms, ctrl = [], [], [], [], []
    for t in range(T):
        o, r, te, tr, info = wenv.step(acts[t])
        obs.append(o); rew.append(r); term.append(te); ctrl.append(wenv.data.ctrl.copy())
        terms.append([info[k] for k in keys])
    out.update(C_actions=acts, C_obs=np.array(obs), C_reward=np.array(rew), C_terminated=np.array(term),
               C_terms=np.array(terms, dtype=np.float64), C_ctrl=np.array(ctrl), C_keys=np.array(keys),
               C_cmd_velocity=wenv.control_inputs.velocity.copy(), C_cmd_heading=wenv.control_inputs.heading.copy(),
               C_global_velocity=wenv.control_inputs.global_velocity.copy())

    # --- trace D: flip termination (walking_quad.py:152-162): start rolled by 80 deg / 95 deg about x
    for tag, roll in (("D80", 80.0), ("D95", 95.0)):
        wenv = WalkingQuadrupedEnv(frame_skip=4, max_time=10.0)
        wenv.reset()
        half = 0.5 * np.deg2rad(roll)
        wenv.data.qpos[3:7] = [np.cos(half), np.sin(half), 0.0, 0.0]
        zs, term = [], []
        for t in range(150):
            o, r, te, tr, info = wenv.step(np.zeros(12, dtype=np.float32))
            zs.append(o[29]); term.append(te)
        out[tag + "_zaxis_z"] = np.array(zs)
        out[tag + "_terminated"] = np.array(term)
        out[tag + "_roll_deg"] = np.array(roll)

    os.chdir(cwd)
    path = os.path.join(HERE, "env_traces.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})
    print("B4 / B10 first terminated step:", out["B4_first_terminated_step"], out["B10_first_terminated_step"])
    for tag in ("D80", "D95"):
        print(tag, "terminated steps:", int(out[tag + "_terminated"].sum()), "first", int(np.argmax(out[tag + "_terminated"])),
              "zaxis_z min/max", out[tag + "_zaxis_z"].min(), out[tag + "_zaxis_z"].max())


if __name__ == "__main__":
    main()
