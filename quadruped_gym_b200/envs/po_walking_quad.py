"""Vectorised POWalkingQuadrupedEnv (/root/reference/src/envs/po_walking_quad.py:8-90) and an SB3-style VecEnv
adapter, so that the reference's training / evaluation scripts can swap their ``SubprocVecEnv([...]*10)``
(/root/reference/src/train_quadruped.py:49-50) for N device-resident environments.

Observation per frame (26): gyro, accel, Euler angles of a Madgwick IMU filter, body_vel xy, ctrl, command
velocity xy, heading angle -- stacked over ``obs_window`` frames.  The filter is ``ahrs``' (third party, not
installable here): restated from its published algorithm, parity unpinned.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _lib
from .quadruped import _box, _ptr
from .walking_quad import VecWalkingQuadrupedEnv


class VecPOWalkingQuadrupedEnv(VecWalkingQuadrupedEnv):
    FRAME = 26

    def __init__(self, num_envs: int = 1, device="cuda:0", obs_window: int = 1, madgwick_gain: float = 0.033, **kwargs):
        super().__init__(num_envs=num_envs, device=device, **kwargs)
        self.obs_window = int(obs_window)
        _lib.check(_lib.lib().qg_po_enable(self._batch, self.obs_window, self.dt, float(madgwick_gain), self.settling_time), "qg_po_enable")
        d = self.FRAME * self.obs_window                                       # po_walking_quad.py:22-27
        self.observation_space = _box(-np.inf, np.inf, (d,))
        self._stacked = torch.zeros((self.num_envs, d), dtype=torch.float32, device=self.device)
        self._term_stacked = torch.zeros((self.num_envs, d), dtype=torch.float32, device=self.device)

    def reset(self, seed=None, options=None, mask: Optional[torch.Tensor] = None):
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        # the reset observation is built inside the base reset, i.e. BEFORE command resampling (quadruped.py:138
        # runs inside walking_quad.py:103): fill the stack first, then do the walking / physics reset
        super(VecWalkingQuadrupedEnv, self).reset(seed=seed, options=options, mask=mask)
        _lib.check(_lib.lib().qg_po_observe(self._batch, None, _ptr(m), _ptr(self._stacked), None, 1, 1, self._stream()), "qg_po_observe")
        self._walk_reset(m, options)
        self.info = {}
        return self._stacked, self.info

    def step(self, action: torch.Tensor):
        L = _lib.lib()
        if self._table_key != "walking":
            _lib.check(L.qg_set_reward_table(self._batch, 0, None, None, None), "qg_set_reward_table")
            _lib.check(L.qg_set_options(self._batch, self.max_time, 1, 0, 0, 0), "qg_set_options")
            self._table_key = "walking"
        a = torch.as_tensor(action, device=self.device, dtype=torch.float32).reshape(self.num_envs, 12).contiguous()
        st = self._stream()
        _lib.check(L.qg_step(self._batch, _ptr(a), self.frame_skip, _ptr(self._obs), _ptr(self._reward), None,
                             _ptr(self._terminated), None, st), "qg_step")
        _lib.check(L.qg_po_observe(self._batch, _ptr(self._obs), _ptr(self._terminated), _ptr(self._stacked),
                                   _ptr(self._term_stacked), int(self.auto_reset), 0, st), "qg_po_observe")
        _lib.check(L.qg_walk_step(self._batch, _ptr(self._obs), None, _ptr(self._terminated), _ptr(self._term_obs),
                                  _ptr(self._reward), _ptr(self._wterms), _ptr(self._wrew64), _ptr(self._wterms64),
                                  int(self.auto_reset), st), "qg_walk_step")
        if self.auto_reset:
            _lib.check(L.qg_reset(self._batch, _ptr(self._terminated), self.seed_value, int(self.random_init), self.env_offset, st), "qg_reset")
        terminated = self._terminated.view(torch.bool)
        self._sd_pending, self._sd_merge = True, bool(self.auto_reset)
        self.info = {k: self._wterms[:, i] for i, k in enumerate(self.reward_keys)}
        self.info["terminal_observation"] = self._term_stacked
        return self._stacked, self._reward, terminated, self._truncated, self.info


def __getattr__(name):   # the SB3 adapter moved to envs/sb3.py; the old import path keeps working
    if name in ("SB3VecEnvAdapter", "SB3VecEnv"):
        from . import sb3
        return getattr(sb3, name)
    raise AttributeError(name)
