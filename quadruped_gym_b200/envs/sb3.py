"""stable-baselines3 ``VecEnv`` over a vectorised environment of this package, so that the reference's training script
can swap ``SubprocVecEnv([lambda: make_env(options)] * 10)`` (/root/reference/src/train_quadruped.py:49-50) for N
device-resident environments and hand the result to ``PPO("MlpPolicy", env, ...)`` (:58) unchanged.

* ``SB3VecEnv`` SUBCLASSES ``stable_baselines3.common.vec_env.VecEnv`` whenever SB3 is importable (SB3 wraps anything
  that is not a ``VecEnv`` instance into a ``DummyVecEnv``); without SB3 the same class stands on ``object``.
* numpy in / numpy out, ONE packed device-to-host transfer per step (observation | reward | done | the 11 reward terms)
  through a page-locked buffer.
* same-step auto-reset (SB3 convention): the returned observation of a finished environment is its reset observation,
  ``infos[i]["terminal_observation"]`` holds the last one, ``infos[i]["TimeLimit.truncated"]`` is False (the reference
  reports the time limit as ``terminated``, quadruped.py:149-151,178-179).
* ``infos`` is a lazy sequence of mapping views over the packed arrays: no per-environment dict is built unless asked
  for.  ``RewardCallback._on_step`` (train_quadruped.py:86-92) indexes ``info[key]`` for the 11 reward keys of every
  environment; each key's column is converted once per step.
"""
from __future__ import annotations

from collections.abc import Mapping, Sequence
from typing import Optional

import numpy as np
import torch

try:  # pragma: no cover - stable-baselines3 is not installable in the build image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:
    _VecEnvBase = object


class _InfoView(Mapping):
    """``infos[i]``: a read-only mapping with the reward keys, ``TimeLimit.truncated`` and, for finished environments,
    ``terminal_observation``.  ``dict(view)`` gives a plain dict."""

    __slots__ = ("_o", "_i")

    def __init__(self, owner, i):
        self._o, self._i = owner, i

    def __getitem__(self, key):
        o = self._o
        if key in o._col:
            return o._column(key)[self._i]
        if key == "TimeLimit.truncated":
            return False
        if key == "terminal_observation" and o.dones[self._i]:
            return o.terminal_obs[o._done_row[self._i]]
        raise KeyError(key)

    def __iter__(self):
        yield from self._o.keys
        yield "TimeLimit.truncated"
        if self._o.dones[self._i]:
            yield "terminal_observation"

    def __len__(self):
        return len(self._o.keys) + 1 + int(self._o.dones[self._i])


class LazyInfos(Sequence):
    """The ``infos`` list of one ``step_wait``: ``terms`` [N,K] float32, ``dones`` [N] bool, ``terminal_obs`` rows of
    the finished environments in env order."""

    def __init__(self, keys, terms, dones, terminal_obs):
        self.keys, self.terms, self.dones, self.terminal_obs = list(keys), terms, dones, terminal_obs
        self._col = {k: c for c, k in enumerate(self.keys)}
        self._cache = {}
        self._views = None
        self._dicts = None
        self._done_row = {int(e): j for j, e in enumerate(np.flatnonzero(dones))} if terminal_obs is not None else {}

    def _column(self, key):
        col = self._cache.get(key)
        if col is None:
            col = self._cache[key] = self.terms[:, self._col[key]].tolist()     # python floats, converted once
        return col

    def _materialise(self):
        """Plain dicts for every environment, built once in bulk: what a consumer that reads every key of every info
        (RewardCallback._on_step) is served fastest with."""
        if self._dicts is None:
            rows = self.terms.tolist()
            keys = self.keys
            self._dicts = ds = [dict(zip(keys, row)) for row in rows]
            for d in ds:
                d["TimeLimit.truncated"] = False
            for e, j in self._done_row.items():
                ds[e]["terminal_observation"] = self.terminal_obs[j]
        return self._dicts

    def __iter__(self):
        return iter(self._materialise())

    def mean(self, key) -> float:
        """Fast path for per-step logging: mean of one reward term over the environments."""
        return float(self.terms[:, self._col[key]].mean())

    def __len__(self):
        return len(self.dones)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if self._dicts is not None:
            return self._dicts[i]
        if self._views is None:
            self._views = [None] * len(self)
        if i < 0:
            i += len(self)
        v = self._views[i]
        if v is None:
            v = self._views[i] = _InfoView(self, i)
        return v


class SB3VecEnv(_VecEnvBase):
    def __init__(self, env, copy: bool = True):
        """``copy=False`` returns views of two alternating page-locked buffers instead of fresh arrays: valid for
        consumers that keep an observation for at most one further step (SB3's algorithms copy into their own buffers)."""
        self._copy = bool(copy)
        if not env.auto_reset:
            raise ValueError("SB3 VecEnv semantics need auto_reset=True")
        self.env = env
        if _VecEnvBase is not object:  # pragma: no cover
            super().__init__(env.num_envs, env.observation_space, env.action_space)
        else:
            self.num_envs, self.observation_space, self.action_space = env.num_envs, env.observation_space, env.action_space
            self.reset_infos = [{} for _ in range(env.num_envs)]
        self.render_mode = env.render_mode
        self.reward_keys = list(getattr(env, "reward_keys", []))
        self._actions = None
        d, k = int(env.observation_space.shape[0]), len(self.reward_keys)
        self._d, self._k = d, k
        self._packed_dev = torch.zeros((env.num_envs, d + 2 + k), dtype=torch.float32, device=env.device)
        self._packed_hosts = [torch.zeros((env.num_envs, d + 2 + k), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._flip = 0
        self._act_host = torch.zeros((env.num_envs, 12), dtype=torch.float32).pin_memory()

    # -- VecEnv API -------------------------------------------------------------------------------
    def reset(self):
        obs, _ = self.env.reset()
        return obs.cpu().numpy()

    def step_async(self, actions):
        self._act_host.numpy()[...] = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, 12)

    def step_wait(self):
        env, d, k = self.env, self._d, self._k
        obs, rew, term, trunc, info = env.step(self._act_host.to(env.device, non_blocking=True))
        p = self._packed_dev
        p[:, :d].copy_(obs)
        p[:, d].copy_(rew)
        p[:, d + 1].copy_(term)
        for c, key in enumerate(self.reward_keys):
            p[:, d + 2 + c].copy_(info[key])
        self._flip ^= 1
        host = self._packed_hosts[self._flip]
        host.copy_(p, non_blocking=True)
        torch.cuda.current_stream(env.device).synchronize()
        h = host.numpy()
        dones = h[:, d + 1] > 0.5
        tobs = info["terminal_observation"][term].cpu().numpy() if dones.any() else None
        if self._copy:
            return h[:, :d].copy(), h[:, d].copy(), dones, LazyInfos(self.reward_keys, h[:, d + 2:].copy(), dones, tobs)
        return h[:, :d], h[:, d], dones, LazyInfos(self.reward_keys, h[:, d + 2:], dones, tobs)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def seed(self, seed: Optional[int] = None):
        self.env.seed(seed)
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        return [indices] if isinstance(indices, int) else list(indices)

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.env, attr_name)] * len(self._indices(indices))

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        return [getattr(self.env, method_name)(*args, **kwargs)] * len(self._indices(indices))

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * len(self._indices(indices))

    def get_images(self):
        return [self.env.render(0)] + [None] * (self.num_envs - 1)

    def render(self, mode=None):
        return self.env.render(0)


SB3VecEnvAdapter = SB3VecEnv     # round-1 name
