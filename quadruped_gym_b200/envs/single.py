"""Single-environment task classes with the reference's constructors and numpy return types, so that the reference's
scripts run against the B200 path without edits:

* ``WalkingQuadrupedEnv(settling_time, random_controls, random_init, reset_options, **base_kwargs)``
  (/root/reference/src/envs/walking_quad.py:11) and
* ``POWalkingQuadrupedEnv(obs_window, **kwargs)`` (/root/reference/src/envs/po_walking_quad.py:10),

as ``train_quadruped.py:15-27,171-193`` and ``eval_quadruped.py:11-27`` construct and drive them: ``reset() -> (obs, {})``,
``step(a) -> (obs, float, bool, False, info)`` with the 11 reward keys in ``info``, ``env.control_inputs.set_orientation /
set_velocity_speed_alpha``, ``env.data.qpos[3:7] = ...`` style writes, ``env.render()`` (MuJoCo bridge, needs the wheel).
Each object is a one-environment batch of the vectorised classes: the arithmetic is the same kernels.  The ``envs``
shadow package under ``dropin/`` re-exports these under the reference's module names.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .po_walking_quad import VecPOWalkingQuadrupedEnv
from .quadruped import _EnvBase, _gym
from .render import MujocoRenderBridge
from .walking_quad import VecWalkingQuadrupedEnv


class _Field(np.ndarray):
    """float64 copy of one ``mjData`` field of environment 0 that WRITES THROUGH to the device on item assignment
    (``env.data.qpos[3:7] = q``, ``env.data.ctrl[:] = c``: walking_quad.py:74, quadruped.py:124)."""

    def __new__(cls, arr, push):
        obj = np.asarray(arr, dtype=np.float64).view(cls)
        obj._push = push
        return obj

    def __array_finalize__(self, obj):
        self._push = None        # slices and results of arithmetic are plain values; only the field itself writes through

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, key, value)
        if self._push is not None:
            self._push(np.asarray(self))


class SingleData:
    """``env.data`` of a single environment: qpos, qvel, act, ctrl, qacc_warmstart, sensordata (numpy float64) and time."""

    _FIELDS = {"qpos": "qpos", "qvel": "qvel", "act": "act", "ctrl": "ctrl", "qacc_warmstart": "qacc_warmstart"}

    def __init__(self, vec):
        object.__setattr__(self, "_vec", vec)

    def __getattr__(self, name):
        vec = object.__getattribute__(self, "_vec")
        if name in SingleData._FIELDS:
            arr = getattr(vec.data, name)[0].double().cpu().numpy()
            return _Field(arr, lambda a, n=name: vec.set_state(**{n: torch.as_tensor(np.asarray(a, dtype=np.float32))[None]}))
        if name == "sensordata":
            return vec.data.sensordata[0].double().cpu().numpy()
        if name == "time":
            return float(vec.data.time[0])
        raise AttributeError(name)

    def __setattr__(self, name, value):
        vec = object.__getattribute__(self, "_vec")
        if name == "time":
            vec.set_state(time=np.array([float(value)]))
        elif name in SingleData._FIELDS:
            vec.set_state(**{name: torch.as_tensor(np.asarray(value, dtype=np.float32))[None]})
        else:
            raise AttributeError(f"env.data.{name} is not writable")


class SingleControls:
    """``env.control_inputs`` of one environment (control_inputs.py:3-116): velocity / heading / global_velocity vectors
    and the setters the scripts call (eval_quadruped.py:13-14), kept on the device next to the reward kernel."""

    def __init__(self, vec):
        self._vec = vec

    def _vecs(self):
        c = self._vec.control_inputs
        return c.velocity[0].cpu().numpy(), c.heading[0].cpu().numpy(), c.global_velocity[0].cpu().numpy()

    velocity = property(lambda s: s._vecs()[0])
    heading = property(lambda s: s._vecs()[1])
    global_velocity = property(lambda s: s._vecs()[2])

    def _polar(self):
        v, h, _ = self._vecs()
        return math.hypot(v[0], v[1]), math.atan2(v[1], v[0]), math.atan2(h[1], h[0])

    def set_velocity_speed_alpha(self, speed, alpha):
        _, _, theta = self._polar()
        self._vec.control_inputs.set_speed_alpha_theta(float(speed), float(alpha), theta)

    def set_velocity_xy(self, x, y):
        self.set_velocity_speed_alpha(math.hypot(x, y), math.atan2(y, x))

    def set_orientation(self, theta):
        speed, alpha, _ = self._polar()
        self._vec.control_inputs.set_speed_alpha_theta(speed, alpha, float(theta))

    def get_heading_theta(self):
        return self._polar()[2]

    def get_velocity_aplha_speed(self):      # (sic) the reference's spelling, control_inputs.py:60
        s, a, _ = self._polar()
        return s, a

    def get_global_velocity_alpha_speed(self):
        g = self._vecs()[2]
        return math.hypot(g[0], g[1]), math.atan2(g[1], g[0])

    def sample(self, options=None):
        """control_inputs.sample(options) (control_inputs.py:74-116) called by hand: heading, velocity angle and speed
        from numpy's global generator in the reference's order of draws, unless fixed by the options."""
        o = options or {}
        theta = o["fixed_heading_angle"] if o.get("fixed_heading_angle") is not None else np.random.uniform(-np.pi, np.pi)
        alpha = o["fixed_velocity_angle"] if o.get("fixed_velocity_angle") is not None else np.random.uniform(-np.pi, np.pi)
        speed = o["fixed_speed"] if o.get("fixed_speed") is not None else np.random.uniform(o.get("min_speed", 0.0), o.get("max_speed", 1.0))
        self._vec.control_inputs.set_speed_alpha_theta(float(speed), float(alpha), float(theta))


class WalkingQuadrupedEnv(_EnvBase):
    """walking_quad.py:9-428 for ONE environment on the device.  Same keywords; ``device`` is the only addition."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}
    reward_keys = _lib.WALK_REWARD_KEYS
    _VEC = VecWalkingQuadrupedEnv

    def __init__(self, settling_time=0, random_controls=False, random_init=False, reset_options=None,
                 model_path: Optional[str] = None, max_time: float = 10.0, frame_skip: int = 4, render_mode: Optional[str] = None,
                 width: int = 720, height: int = 480, render_fps: int = 30, reward_fns: Optional[dict] = None,
                 termination_fns: Optional[dict] = None, save_video: bool = False, video_path: str = "videos/simulation.mp4",
                 use_default_termination: bool = True, device="cuda:0", seed: int = 0, **vec_kwargs):
        if _gym is not None:
            super().__init__()
        self.vec = self._VEC(1, device, settling_time=settling_time, random_controls=random_controls, random_init=random_init,
                             reset_options=reset_options, model_path=model_path, max_time=max_time, frame_skip=frame_skip,
                             auto_reset=False, seed=seed, **vec_kwargs)
        v = self.vec
        self.model, self.data = v.model, SingleData(v)
        self.max_time, self.frame_skip, self.render_mode = max_time, frame_skip, render_mode
        self.settling_time, self.random_controls, self.random_init, self.reset_options = settling_time, random_controls, random_init, reset_options
        self.action_space, self.observation_space = v.action_space, v.observation_space
        self.control_inputs = SingleControls(v)
        self.joint_centers = np.array([0.0, 0.0, -0.5] * 4, dtype=np.float32)
        # user-supplied dictionaries replace the fused walking reward / termination, as in quadruped.py:97-100
        self.reward_fns = reward_fns
        self.termination_fns = termination_fns
        self.use_default_termination = use_default_termination
        self.info = {}
        self._bridge = None
        if render_mode is not None or save_video:
            self._bridge = MujocoRenderBridge(model_path, render_mode, width, height, render_fps, save_video, video_path)

    # -- reference attribute surface ---------------------------------------------------------------
    @property
    def ideal_position(self):
        return self.vec.ideal_position[0].cpu().numpy()

    @property
    def ctrl_f_est(self):
        return self.vec.ctrl_f_est[0].cpu().numpy()

    @property
    def ctrl_a_est(self):
        return self.vec.ctrl_a_est[0].cpu().numpy()

    def seed(self, seed=None):
        if seed is not None:
            self.vec.seed(seed)
        return [seed]

    def _obs(self, t: torch.Tensor) -> np.ndarray:
        return t[0].double().cpu().numpy()

    def reset(self, seed=None, options=None):
        obs, _ = self.vec.reset(seed=seed, options=options)
        if self._bridge is not None:
            self._bridge.restart()
        self.info = {}
        return self._obs(obs), self.info

    def step(self, action):
        a = torch.as_tensor(np.asarray(action, dtype=np.float32)).reshape(1, 12)
        obs, rew, term, _, info = self.vec.step(a)
        keys = self.reward_keys
        # one packed read: observation | reward | terminated | the 11 reward terms
        packed = torch.cat([obs[0].double(), rew.double(), term.double(), torch.stack([info[k][0] for k in keys]).double()]).cpu().numpy()
        d = obs.shape[1]
        reward, terminated = float(packed[d]), bool(packed[d + 1] > 0.5)
        self.info = {k: float(packed[d + 2 + i]) for i, k in enumerate(keys)}     # walking_quad.py:148,419
        if self.reward_fns is not None:          # modular dictionary given by the user (quadruped.py:170-175)
            comps = {k: fn() for k, fn in self.reward_fns.items()}
            reward = float(sum(comps.values()))
            self.info = {"time": self.data.time, "reward_components": comps}
        if self.termination_fns is not None:     # user terminations come first, "default" last (quadruped.py:98-100,178)
            terminated = any(fn() for fn in self.termination_fns.values()) or (self.use_default_termination and terminated)
        return packed[:d].copy(), reward, terminated, False, self.info

    def render(self):
        if self._bridge is None:
            return None
        v, h, g = self.control_inputs._vecs()
        origin = self.data.sensordata[18:21]
        overlays = [("arrow", origin, g, (1, 0, 0, 1), 0.1), ("arrow", origin, h, (0, 1, 0, 1), 0.05),
                    ("point", self.ideal_position, (1, 0, 1, 1), 0.0)]          # walking_quad.py:77-86
        return self._bridge.frame(self.data.qpos, self.data.time, overlays)

    def close(self):
        if self._bridge is not None:
            self._bridge.close()
            self._bridge = None
        self.vec.close()


class POWalkingQuadrupedEnv(WalkingQuadrupedEnv):
    """po_walking_quad.py:8-90 for ONE environment: 26-value frames stacked over ``obs_window`` (float64 numpy out)."""

    _VEC = VecPOWalkingQuadrupedEnv

    def __init__(self, obs_window=1, **kwargs):
        super().__init__(obs_window=obs_window, **kwargs)
        self.obs_window = obs_window
