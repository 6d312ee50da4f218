"""Vectorised WalkingQuadrupedEnv: the reference's task environment
(/root/reference/src/envs/walking_quad.py:9-428) for N environments on the device.

Per ``step`` three launches: the physics kernel (``qg_step`` with flip + time-limit termination, settling mask
fused), the walking kernel (``qg_walk_step``: ideal position, frequency/amplitude estimator, the 11 reward terms
in float64, reset bookkeeping, zero reset observation) and a masked physics reset.  Nothing leaves the device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from .. import _lib
from . import rewards as R
from .quadruped import VecQuadrupedEnv, _ptr


class VecVelocityHeadingControls:
    """Per-env command inputs (control_inputs.py:3-116) as device tensors: velocity, heading, global_velocity [N,3]."""

    def __init__(self, env: "VecWalkingQuadrupedEnv"):
        self._env = env

    def _get(self, which):
        e = self._env
        bufs = {k: None for k in ("velocity", "heading", "global_velocity", "ideal_position", "f_est", "a_est")}
        shape = (12, e.num_envs) if which in ("f_est", "a_est") else (e.num_envs, 3)
        out = torch.empty(shape, dtype=torch.float64, device=e.device)
        bufs[which] = out
        _lib.check(_lib.lib().qg_walk_get_commands(e._batch, *[_ptr(bufs[k]) for k in bufs], e._stream()), "qg_walk_get_commands")
        return out

    velocity = property(lambda s: s._get("velocity"))
    heading = property(lambda s: s._get("heading"))
    global_velocity = property(lambda s: s._get("global_velocity"))

    def set_speed_alpha_theta(self, speed, alpha, theta, mask: Optional[torch.Tensor] = None):
        """set_velocity_speed_alpha + set_orientation (control_inputs.py:36-51) for all (or masked) environments."""
        e = self._env
        sat = torch.stack([torch.as_tensor(x, dtype=torch.float64, device=e.device).expand(e.num_envs) for x in (speed, alpha, theta)], dim=1).contiguous()
        m = None if mask is None else mask.to(device=e.device, dtype=torch.uint8).contiguous()
        _lib.check(_lib.lib().qg_walk_set_commands(e._batch, _ptr(sat), _ptr(m), e._stream()), "qg_walk_set_commands")
        torch.cuda.current_stream(e.device).synchronize()

    def get_heading_theta(self):
        h = self.heading
        return torch.atan2(h[:, 1], h[:, 0])


class VecWalkingQuadrupedEnv(VecQuadrupedEnv):
    """Same keywords as the reference (walking_quad.py:11): settling_time, random_controls, random_init,
    reset_options (keys of control_inputs.py:88-92) plus the base-env keywords."""

    reward_keys = _lib.WALK_REWARD_KEYS

    def __init__(self, num_envs: int = 1, device="cuda:0", settling_time: float = 0, random_controls: bool = False,
                 random_init: bool = False, reset_options: Optional[dict] = None, **kwargs):
        kwargs.setdefault("termination_fns", {})
        super().__init__(num_envs=num_envs, device=device, random_init=random_init, **kwargs)
        # walking_quad.py:158-162: flip termination or time limit
        self.termination_fns = {"flip": R.flip_termination(), "default": R.time_limit()}
        self.settling_time, self.random_controls, self.reset_options = float(settling_time), bool(random_controls), reset_options
        ts = self.model.opt.timestep
        self.dt = ts * self.frame_skip                                  # walking_quad.py:56
        self.window_size = int(np.ceil(2 / (1 * self.dt)))              # math_utils.py:28 with min_freq = 1
        sample_opts, sample_has = self._sample_options(reset_options)
        _lib.check(_lib.lib().qg_walk_enable(self._batch, self.window_size, self.dt, ts, self.frame_skip, self.settling_time,
                                             int(self.random_controls), sample_opts, sample_has), "qg_walk_enable")
        # key the command sampler on THIS env's seed and shard offset (global env id = env_offset + i)
        _lib.check(_lib.lib().qg_walk_reset(self._batch, None, 1, self.seed_value, self.env_offset, self._stream()), "qg_walk_reset")
        self.control_inputs = VecVelocityHeadingControls(self)
        self.joint_centers = torch.tensor([0.0, 0.0, -0.5] * 4, dtype=torch.float32, device=self.device)
        n = self.num_envs
        self._wterms = torch.zeros((n, len(self.reward_keys)), dtype=torch.float32, device=self.device)
        self._wterms64 = torch.zeros((n, len(self.reward_keys)), dtype=torch.float64, device=self.device)
        self._wrew64 = torch.zeros((n,), dtype=torch.float64, device=self.device)
        self.info = {}

    @staticmethod
    def _sample_options(options: Optional[dict]):
        """control_inputs.sample's option dict (control_inputs.py:88-92) as the two C arrays of the ABI."""
        o = options or {}
        sample_opts = (C.c_double * 5)(float(o.get("min_speed", 0.0)), float(o.get("max_speed", 1.0)),
                                       float(o.get("fixed_heading_angle") or 0.0), float(o.get("fixed_velocity_angle") or 0.0),
                                       float(o.get("fixed_speed") or 0.0))
        sample_has = (C.c_int * 3)(int(o.get("fixed_heading_angle") is not None), int(o.get("fixed_velocity_angle") is not None),
                                   int(o.get("fixed_speed") is not None))
        return sample_opts, sample_has

    def _walk_reset(self, m, options):
        """WalkingQuadrupedEnv.reset bookkeeping + command resampling with `options` for this call only
        (walking_quad.py:100-103,121-122); auto-resets inside step() keep using the constructor's reset_options."""
        L = _lib.lib()
        if options is not None:
            _lib.check(L.qg_walk_set_sample_options(self._batch, *self._sample_options(options)), "qg_walk_set_sample_options")
        _lib.check(L.qg_walk_reset(self._batch, _ptr(m), 0, self.seed_value, self.env_offset, self._stream()), "qg_walk_reset")
        if options is not None:
            _lib.check(L.qg_walk_set_sample_options(self._batch, *self._sample_options(self.reset_options)), "qg_walk_set_sample_options")

    @property
    def ideal_position(self):
        return self.control_inputs._get("ideal_position")

    @property
    def ctrl_f_est(self):
        return self.control_inputs._get("f_est").T

    @property
    def ctrl_a_est(self):
        return self.control_inputs._get("a_est").T

    def reset(self, seed=None, options=None, mask: Optional[torch.Tensor] = None):
        obs, _ = super().reset(seed=seed, options=options, mask=mask)
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        self._walk_reset(m, options)
        self.info = {}
        return obs, self.info

    def step(self, action: torch.Tensor):
        L = _lib.lib()
        # physics: flip + time-limit termination, no in-kernel reset, no fused reward table
        if self._table_key != "walking":
            _lib.check(L.qg_set_reward_table(self._batch, 0, None, None, None), "qg_set_reward_table")
            _lib.check(L.qg_set_options(self._batch, self.max_time, 1, 0, 0, 0), "qg_set_options")
            self._table_key = "walking"
        a = torch.as_tensor(action, device=self.device, dtype=torch.float32).reshape(self.num_envs, 12).contiguous()
        st = self._stream()
        _lib.check(L.qg_step(self._batch, _ptr(a), self.frame_skip, _ptr(self._obs), _ptr(self._reward), None,
                             _ptr(self._terminated), None, st), "qg_step")
        _lib.check(L.qg_walk_step(self._batch, _ptr(self._obs), None, _ptr(self._terminated), _ptr(self._term_obs),
                                  _ptr(self._reward), _ptr(self._wterms), _ptr(self._wrew64), _ptr(self._wterms64),
                                  int(self.auto_reset), st), "qg_walk_step")
        if self.auto_reset:
            _lib.check(L.qg_reset(self._batch, _ptr(self._terminated), self.seed_value, int(self.random_init), self.env_offset, st), "qg_reset")
        terminated = self._terminated.view(torch.bool)
        self._sd_pending, self._sd_merge = True, bool(self.auto_reset)
        # walking_quad.py:148 returns self.info = per-term reward dict (the base info is dropped)
        self.info = {k: self._wterms[:, i] for i, k in enumerate(self.reward_keys)}
        self.info["terminal_observation"] = self._term_obs
        return self._obs, self._reward, terminated, self._truncated, self.info
