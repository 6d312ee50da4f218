"""quadruped_gym_b200 -- B200-native batched simulator for the hot path of antopio26/quadruped-gym:
``QuadrupedEnv.step()`` = frame_skip x mj_step -> sensordata -> reward_fns / termination_fns
(/root/reference/src/envs/quadruped.py:153-182) behind the reference's Gymnasium reset/step API.

The compute path is hand-written sm_100a CUDA behind a C ABI (include/quadgym.h, libquadgym.so);
this package is the thin Python host: model compiler, ctypes binding, env classes.
"""
from . import _lib  # noqa: F401
from .model import compile_mjcf, load_model_blob  # noqa: F401

__all__ = ["VecQuadrupedEnv", "QuadrupedEnv", "VecWalkingQuadrupedEnv", "VecPOWalkingQuadrupedEnv", "SB3VecEnvAdapter", "SB3VecEnv",
           "WalkingQuadrupedEnv", "POWalkingQuadrupedEnv",
           "RolloutBuffer", "rewards", "compile_mjcf", "load_model_blob"]


def __getattr__(name):  # torch is imported lazily so that the model compiler works without it
    if name in ("VecQuadrupedEnv", "QuadrupedEnv"):
        from .envs import quadruped
        return getattr(quadruped, name)
    if name == "VecWalkingQuadrupedEnv":
        from .envs import walking_quad
        return walking_quad.VecWalkingQuadrupedEnv
    if name == "VecPOWalkingQuadrupedEnv":
        from .envs import po_walking_quad
        return po_walking_quad.VecPOWalkingQuadrupedEnv
    if name in ("SB3VecEnvAdapter", "SB3VecEnv"):
        from .envs import sb3
        return getattr(sb3, name)
    if name in ("WalkingQuadrupedEnv", "POWalkingQuadrupedEnv"):
        from .envs import single
        return getattr(single, name)
    if name == "RolloutBuffer":
        from .rollout import RolloutBuffer
        return RolloutBuffer
    if name == "rewards":
        from .envs import rewards
        return rewards
    raise AttributeError(name)
