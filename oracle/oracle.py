"""ctypes face of the CPU oracle (oracle/qg_oracle.c) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs import this module.  It restates ``frame_skip x mujoco.mj_step`` + sensor readout
(/root/reference/src/envs/quadruped.py:153-167) in float64 on the CPU.  PARITY UNPINNED against the
real ``mujoco`` wheel (not installable here) -- see the header of qg_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libqgoracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("qg_oracle.c", "qg_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _parse_struct(header: str, name: str, consts):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        typ, rest = stmt.split(None, 1)
        ct = {"double": C.c_double, "int": C.c_int}[typ]
        for decl in rest.split(","):
            decl = decl.strip()
            mm = re.match(r"(\w+)\[(.+)\]$", decl)
            if mm:
                n = int(eval(mm.group(2), {}, consts))
                fields.append((mm.group(1), ct * n))
            else:
                fields.append((decl, ct))
    return fields


def _load():
    build()
    header = open(os.path.join(_HERE, "qg_oracle.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"#define (QGO_\w+) (\d+)", header)}

    class Model(C.Structure):
        _fields_ = _parse_struct(header, "qgo_model", consts)

    class Data(C.Structure):
        _fields_ = _parse_struct(header, "qgo_data", consts)

    lib = C.CDLL(_SO)
    lib.qgo_sizeof_data.restype = C.c_size_t
    lib.qgo_sizeof_model.restype = C.c_size_t
    assert lib.qgo_sizeof_data() == C.sizeof(Data), (lib.qgo_sizeof_data(), C.sizeof(Data))
    assert lib.qgo_sizeof_model() == C.sizeof(Model)
    lib.qgo_model_load.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(Model))]
    lib.qgo_model_free.argtypes = [C.POINTER(Model)]
    for fn in ("qgo_reset", "qgo_forward", "qgo_step"):
        getattr(lib, fn).argtypes = [C.POINTER(Model), C.POINTER(Data)]
        getattr(lib, fn).restype = None
    lib.qgo_env_step.argtypes = [C.POINTER(Model), C.POINTER(Data), C.POINTER(C.c_double), C.c_int]
    lib.qgo_env_step.restype = None
    lib.qgo_rollout.argtypes = [C.POINTER(Model), C.POINTER(Data), C.c_int, C.POINTER(C.c_double), C.c_int,
                                C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double)]
    lib.qgo_rollout.restype = None
    return lib, Model, Data, consts


_lib, Model, Data, CONSTS = _load()


class OracleModel:
    def __init__(self, blob: bytes):
        self._p = C.POINTER(Model)()
        rc = _lib.qgo_model_load(blob, len(blob), C.byref(self._p))
        if rc != 0:
            raise ValueError(f"qgo_model_load failed: {rc}")
        self.m = self._p.contents
        self.nq, self.nv, self.nu = self.m.nq, self.m.nv, self.m.nu

    def __del__(self):
        if getattr(self, "_p", None):
            _lib.qgo_model_free(self._p)
            self._p = None


def _view(arr, n=None):
    a = np.ctypeslib.as_array(arr)
    return a if n is None else a[:n]


class OracleData:
    """One environment's mjData-like record.  Attribute access returns numpy views."""

    def __init__(self, model: OracleModel):
        self.model = model
        self.d = Data()
        _lib.qgo_reset(model._p, C.byref(self.d))

    def __getattr__(self, name):
        d = object.__getattribute__(self, "d")
        m = object.__getattribute__(self, "model")
        v = getattr(d, name)
        if isinstance(v, C.Array):
            a = np.ctypeslib.as_array(v)
            sizes = {"qpos": m.nq, "qvel": m.nv, "act": m.nu, "ctrl": m.nu, "qacc_warmstart": m.nv,
                     "qacc": m.nv, "qacc_smooth": m.nv, "qfrc_bias": m.nv, "qfrc_smooth": m.nv,
                     "qfrc_constraint": m.nv, "qfrc_actuator": m.nv, "qfrc_passive": m.nv}
            if name in sizes:
                return a[:sizes[name]]
            if name in ("M", "L"):
                return a[:m.nv * m.nv].reshape(m.nv, m.nv)
            if name == "efc_J":
                return a[:d.nefc * m.nv].reshape(d.nefc, m.nv)
            if name.startswith("efc_"):
                return a[:d.nefc]
            if name in ("con_pos", "con_vert"):
                return a[:3 * d.ncon].reshape(d.ncon, 3)
            if name.startswith("con_"):
                return a[:d.ncon]
            return a
        return v

    @property
    def time(self):
        return self.d.time

    @time.setter
    def time(self, v):
        self.d.time = v

    def reset(self):
        _lib.qgo_reset(self.model._p, C.byref(self.d))

    def forward(self):
        _lib.qgo_forward(self.model._p, C.byref(self.d))

    def step(self):
        _lib.qgo_step(self.model._p, C.byref(self.d))

    def env_step(self, action, frame_skip: int):
        a = np.ascontiguousarray(action, dtype=np.float64)
        _lib.qgo_env_step(self.model._p, C.byref(self.d), a.ctypes.data_as(C.POINTER(C.c_double)), frame_skip)

    def set_state(self, qpos, qvel, act=None, warm=None, time=0.0, ctrl=None):
        self.qpos[:] = qpos
        self.qvel[:] = qvel
        if act is not None:
            self.act[:] = act
        if warm is not None:
            self.qacc_warmstart[:] = warm
        if ctrl is not None:
            self.ctrl[:] = ctrl
        self.d.time = time


class OracleBatch:
    """n independent environments stepped in a C loop (CPU baseline / batched parity checks)."""

    def __init__(self, model: OracleModel, n: int):
        self.model, self.n = model, n
        self.arr = (Data * n)()
        for i in range(n):
            _lib.qgo_reset(model._p, C.byref(self.arr[i]))
            for k in range(model.nu):
                self.arr[i].ctrl[k] = -0.5 if k % 3 == 2 else 0.0

    def env(self, i) -> OracleData:
        od = OracleData.__new__(OracleData)
        od.model = self.model
        od.d = self.arr[i]
        return od

    def rollout(self, actions: np.ndarray, frame_skip: int, max_time: float = 10.0, auto_reset: bool = True,
                want_obs: bool = False):
        """actions [n_steps, n, nu] float64 -> obs [n_steps, n, 33] (if want_obs)."""
        a = np.ascontiguousarray(actions, dtype=np.float64)
        n_steps = a.shape[0]
        obs = np.zeros((n_steps, self.n, 33)) if want_obs else None
        _lib.qgo_rollout(self.model._p, self.arr, self.n, a.ctypes.data_as(C.POINTER(C.c_double)), n_steps,
                         frame_skip, max_time, int(auto_reset),
                         obs.ctypes.data_as(C.POINTER(C.c_double)) if want_obs else None)
        return obs
