"""ctypes binding of libquadgym.so (include/quadgym.h).

There is no CPU fallback and no alternative backend: if the CUDA extension has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``quadruped_gym_b200/csrc/build.sh``) the
import of anything that needs it raises ``QuadGymLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QG_LIB", os.path.join(_HERE, "libquadgym.so"))  # QG_LIB: tuning builds only

QG_NQ, QG_NV, QG_NU, QG_NSENSORDATA, QG_MAX_TERMS = 19, 18, 12, 33, 16

TERM_IDS = {
    "alive": 0, "ctrl_sq": 1, "qvel_x": 2, "forward": 3, "drift": 4, "control_cost": 5,
    "orientation": 6, "height_cost": 7, "posture_cost": 8, "exp_orientation": 9, "exp_height": 10,
}

# every symbol include/quadgym.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "qg_last_error", "qg_version", "qg_model_load", "qg_model_destroy", "qg_model_info", "qg_batch_create",
    "qg_batch_destroy", "qg_batch_num_envs", "qg_set_options", "qg_set_reward_table", "qg_reset", "qg_step",
    "qg_step_host", "qg_step_host_async", "qg_host_wait", "qg_get_state", "qg_set_state", "qg_debug_step", "qg_get_counters", "qg_launch_count",
    "qg_fp32_peak", "qg_walk_enable", "qg_walk_set_sample_options", "qg_walk_reset", "qg_walk_set_commands", "qg_walk_get_commands", "qg_walk_step",
    "qg_po_enable", "qg_po_observe",
]

WALK_REWARD_KEYS = [  # WalkingQuadrupedEnv.reward_keys (/root/reference/src/envs/walking_quad.py:331-350)
    "alive_bonus", "control_cost", "progress_direction_reward_local", "progress_speed_cost_local", "heading_reward",
    "orientation_reward", "body_height_cost", "joint_posture_cost", "control_amplitude_cost", "control_frequency_cost",
    "diff_ideal_position_cost",
]


class QuadGymLibraryError(RuntimeError):
    pass


class Counters(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in (
        "physics_steps", "contacts", "efc_rows", "newton_iters", "ls_evals", "verts_tested", "diverged",
        "contact_overflow", "episodes", "active_rows")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    """Load (once) and return the C-ABI library; fail loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QuadGymLibraryError(
            f"{LIB_PATH} not found: the sm_100a CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback).")
    L = C.CDLL(LIB_PATH)
    vp, i32, u8p, f32p, f64p = C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p
    L.qg_last_error.restype = C.c_char_p
    L.qg_version.restype = C.c_char_p
    L.qg_model_load.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    L.qg_model_destroy.argtypes = [vp]
    L.qg_model_destroy.restype = None
    L.qg_model_info.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_double)]
    L.qg_batch_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.qg_batch_destroy.argtypes = [vp]
    L.qg_batch_destroy.restype = None
    L.qg_batch_num_envs.argtypes = [vp]
    L.qg_set_options.argtypes = [vp, C.c_double, i32, i32, i32, i32]
    L.qg_set_reward_table.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.qg_reset.argtypes = [vp, u8p, C.c_uint64, i32, C.c_longlong, vp]
    L.qg_step.argtypes = [vp, f32p, i32, f32p, f32p, f32p, u8p, f32p, vp]
    L.qg_step_host.argtypes = [vp, f32p, i32, f32p, f32p, f32p, u8p, f32p, vp]
    L.qg_step_host_async.argtypes = [vp, f32p, i32, f32p, f32p, f32p, u8p, f32p, vp]
    L.qg_host_wait.argtypes = [vp, vp]
    L.qg_get_state.argtypes = [vp, f32p, f32p, f32p, f32p, f64p, f32p, vp]
    L.qg_set_state.argtypes = [vp, f32p, f32p, f32p, f32p, f64p, f32p, vp]
    L.qg_debug_step.argtypes = [vp, f32p, f32p, f32p, f32p, f32p, vp, f32p, vp]
    L.qg_get_counters.argtypes = [vp, C.POINTER(Counters), i32, vp]
    L.qg_launch_count.restype = C.c_ulonglong
    L.qg_fp32_peak.argtypes = [i32, i32, C.POINTER(C.c_double)]
    L.qg_walk_enable.argtypes = [vp, i32, C.c_double, C.c_double, i32, C.c_double, i32, C.POINTER(C.c_double), C.POINTER(i32)]
    L.qg_walk_set_sample_options.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i32)]
    L.qg_walk_reset.argtypes = [vp, u8p, i32, C.c_uint64, C.c_longlong, vp]
    L.qg_walk_set_commands.argtypes = [vp, f64p, u8p, vp]
    L.qg_walk_get_commands.argtypes = [vp, f64p, f64p, f64p, f64p, f64p, f64p, vp]
    L.qg_walk_step.argtypes = [vp, f32p, f32p, u8p, f32p, f32p, f32p, f64p, f64p, i32, vp]
    L.qg_po_enable.argtypes = [vp, i32, C.c_double, C.c_double, C.c_double]
    L.qg_po_observe.argtypes = [vp, f32p, u8p, f32p, f32p, i32, i32, vp]
    _lib = L
    return L


def check(rc: int, what: str = "libquadgym"):
    if rc != 0:
        msg = lib().qg_last_error().decode("utf-8", "replace")
        names = {-1: "QG_EINVAL", -2: "QG_EBLOB", -3: "QG_EMODEL", -4: "QG_ECUDA", -5: "QG_ENOMEM"}
        exc = ValueError if rc in (-1, -2, -3) else QuadGymLibraryError
        raise exc(f"{what} failed with {names.get(rc, rc)}: {msg}")
