"""Vectorised QuadrupedEnv on B200 and the single-env Gymnasium-compatible shim.

Mirrors the interface of the reference's base environment
(/root/reference/src/envs/quadruped.py:9-182): same constructor keywords, ``reset`` / ``step`` /
``close``, the ``reward_fns`` / ``termination_fns`` dictionaries, ``env.model`` / ``env.data`` attribute
surface that reward callables touch (walking_quad.py:19-20,56,93,142,253,393).  All N environments
live in device memory owned by libquadgym; one ``step`` is ONE kernel launch.

Rendering (quadruped.py:184-316) is a bridge (``render.py``): ``render()`` copies one environment's qpos to a CPU
``mujoco.MjData`` and uses MuJoCo's renderer with the reference's pacing ("human" window, "rgb_array", video); it needs
the ``mujoco`` wheel and fails loudly at the first ``render()`` call without it.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import numpy as np
import torch

from .. import _lib
from ..model import SENSORS, load_model_blob
from . import rewards as R

try:  # gymnasium is optional (absent in this image); the API is duck-typed without it
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _EnvBase = _gym.Env
except Exception:  # pragma: no cover
    _gym = None
    _spaces = None
    _EnvBase = object


class Box:
    """Minimal stand-in for ``gymnasium.spaces.Box`` when gymnasium is not installed."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


def _box(low, high, shape):
    if _spaces is not None:
        return _spaces.Box(low=low, high=high, shape=shape, dtype=np.float32)
    return Box(low, high, shape)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Opt:
    def __init__(self, timestep):
        self.timestep = timestep


class ModelView:
    """The slice of ``mujoco.MjModel`` the reference reads: nu, nq, nv, nsensordata, sensor_adr, opt.timestep."""

    def __init__(self, handle, blob: bytes):
        sizes = (C.c_int * 8)()
        ts = C.c_double()
        _lib.check(_lib.lib().qg_model_info(handle, sizes, C.byref(ts)), "qg_model_info")
        self.nq, self.nv, self.nu, self.nbody, self.njnt, self.ngeom, self.nmesh, self.nsensordata = list(sizes)
        self.opt = _Opt(ts.value)
        self.sensor_names = list(SENSORS.keys())
        self.sensor_adr = np.array([SENSORS[n][0] for n in self.sensor_names], dtype=np.int32)
        self.sensor_dim = np.array([SENSORS[n][1] for n in self.sensor_names], dtype=np.int32)
        self.blob = blob

    def sensor_name2id(self, name: str) -> int:
        """``mj_name2id(model, mjOBJ_SENSOR, name)`` (walking_quad.py:19)."""
        return self.sensor_names.index(name)


class DataView:
    """Device-tensor view of ``mujoco.MjData`` fields: qpos, qvel, act, ctrl, time, qacc_warmstart, sensordata."""

    def __init__(self, env: "VecQuadrupedEnv"):
        self._env = env

    def _get(self, which):
        e = self._env
        n, dev = e.num_envs, e.device
        bufs = {"qpos": (19, torch.float32), "qvel": (18, torch.float32), "act": (12, torch.float32),
                "warm": (18, torch.float32), "time": (0, torch.float64), "ctrl": (12, torch.float32)}
        k, dt = bufs[which]
        out = torch.empty((n, k) if k else (n,), dtype=dt, device=dev)
        args = {w: None for w in bufs}
        args[which] = out
        _lib.check(_lib.lib().qg_get_state(e._batch, _ptr(args["qpos"]), _ptr(args["qvel"]), _ptr(args["act"]),
                                           _ptr(args["warm"]), _ptr(args["time"]), _ptr(args["ctrl"]), e._stream()),
                   "qg_get_state")
        return out

    qpos = property(lambda s: s._get("qpos"))
    qvel = property(lambda s: s._get("qvel"))
    act = property(lambda s: s._get("act"))
    ctrl = property(lambda s: s._get("ctrl"))
    time = property(lambda s: s._get("time"))
    qacc_warmstart = property(lambda s: s._get("warm"))

    @property
    def sensordata(self):
        return self._env._sensordata()


class VecQuadrupedEnv:
    """N environments as CUDA tensors behind the reference's reset/step API.

    ``step(action[N,12]) -> (obs[N,33], reward[N], terminated[N] bool, truncated[N] bool (all False), info)``.
    With ``auto_reset=True`` (default, SB3 VecEnv convention) terminated environments are reset inside the
    same kernel launch; their returned observation is the reset observation (all zeros, as the reference's
    ``reset()`` returns: quadruped.py:120,138) and ``info["terminal_observation"]`` holds the last one.
    """

    def __init__(self, num_envs: int = 1, device="cuda:0", model_path: Optional[str] = None, max_time: float = 10.0,
                 frame_skip: int = 4, render_mode: Optional[str] = None, reward_fns: Optional[dict] = None,
                 termination_fns: Optional[dict] = None, use_default_termination: bool = True,
                 auto_reset: bool = True, seed: int = 0, env_offset: int = 0, random_init: bool = False,
                 mesh_inertia: str = "legacy", model_blob: Optional[bytes] = None, **unused_render_kwargs):
        if render_mode not in (None, "human", "rgb_array"):
            raise ValueError(f"unknown render_mode {render_mode!r} (quadruped.py:39: human, rgb_array)")
        self.render_mode = render_mode
        self._renderer = None
        self._render_kwargs = {k: unused_render_kwargs[k] for k in ("width", "height", "render_fps", "save_video", "video_path") if k in unused_render_kwargs}
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.QuadGymLibraryError("VecQuadrupedEnv needs a CUDA device; there is no CPU fallback")
        L = _lib.lib()
        self.num_envs = int(num_envs)
        self.model_path = model_path
        blob = model_blob if model_blob is not None else load_model_blob(model_path, mesh_inertia)
        self._model = C.c_void_p()
        _lib.check(L.qg_model_load(blob, len(blob), C.byref(self._model)), "qg_model_load")
        self.model = ModelView(self._model, blob)
        self._batch = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(L.qg_batch_create(self._model, self.num_envs, dev_index, C.byref(self._batch)), "qg_batch_create")
        self.max_time, self.frame_skip = float(max_time), int(frame_skip)
        self.auto_reset, self.seed_value, self.env_offset, self.random_init = bool(auto_reset), int(seed), int(env_offset), bool(random_init)
        self.action_space = _box(-1.0, 1.0, (self.model.nu,))
        self.observation_space = _box(-np.inf, np.inf, (self.model.nsensordata,))
        n, dev = self.num_envs, self.device
        self._obs = torch.zeros((n, 33), dtype=torch.float32, device=dev)
        self._term_obs = torch.zeros((n, 33), dtype=torch.float32, device=dev)
        self._reward = torch.zeros((n,), dtype=torch.float32, device=dev)
        self._terminated = torch.zeros((n,), dtype=torch.uint8, device=dev)
        self._terms = torch.zeros((n, _lib.QG_MAX_TERMS), dtype=torch.float32, device=dev)
        self._truncated = torch.zeros((n,), dtype=torch.bool, device=dev)   # quadruped.py:179: never truncated
        self._zero_reward = torch.zeros((n,), dtype=torch.float32, device=dev)
        # env.data.sensordata is materialised on demand (no per-step kernel for a field few callers read)
        self._sd = torch.zeros((n, 33), dtype=torch.float32, device=dev)
        self._sd_pending, self._sd_merge = False, False
        self.data = DataView(self)
        # modular reward / termination dictionaries (quadruped.py:97-100)
        self.reward_fns: Dict[str, object] = reward_fns if reward_fns is not None else {"default": self._default_reward}
        self.termination_fns: Dict[str, object] = termination_fns if termination_fns is not None else {}
        if use_default_termination:
            self.termination_fns["default"] = R.time_limit()
        self._table_key = None

    # -- plumbing ----------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _default_reward(self):
        """Default reward function that returns 0 (quadruped.py:145-147)."""
        return self._zero_reward

    def _sync_tables(self):
        """Push the fused parts of reward_fns / termination_fns to the library when the dicts change."""
        fused = [(k, v) for k, v in self.reward_fns.items() if isinstance(v, R.FusedTerm)]
        # the built-in default reward (constant 0, quadruped.py:145-147) needs no evaluation and does not block the in-kernel reset
        zero = [k for k, v in self.reward_fns.items() if v == self._default_reward]
        pyfn = [(k, v) for k, v in self.reward_fns.items() if not isinstance(v, R.FusedTerm) and k not in zero]
        kinds = [v.kind for v in self.termination_fns.values() if isinstance(v, R.FusedTermination)]
        pyterm = [(k, v) for k, v in self.termination_fns.items() if not isinstance(v, R.FusedTermination)]
        key = (tuple((k, v) for k, v in fused), tuple(kinds), len(pyterm), len(pyfn), self.max_time, self.auto_reset)
        self._fused, self._pyfn, self._pyterm, self._zero_terms = fused, pyfn, pyterm, zero
        if key == self._table_key:
            return
        L = _lib.lib()
        n = len(fused)
        ids = (C.c_int * max(n, 1))(*[v.term_id for _, v in fused])
        w = (C.c_double * max(n, 1))(*[float(v.weight) for _, v in fused])
        p = (C.c_double * max(n, 1))(*[float(v.param) for _, v in fused])
        _lib.check(L.qg_set_reward_table(self._batch, n, ids, w, p), "qg_set_reward_table")
        max_time = self.max_time if "time_limit" in kinds else float("inf")
        # Python callables (rewards and terminations) are evaluated after the launch on the terminal state, exactly
        # where the reference evaluates them (quadruped.py:170-178): the reset then happens after them, not in the kernel
        kernel_reset = self.auto_reset and not pyterm and not pyfn
        _lib.check(L.qg_set_options(self._batch, max_time, int("flip" in kinds), int(kernel_reset), 0, 0), "qg_set_options")
        self._kernel_reset = kernel_reset
        self._table_key = key

    # -- Gymnasium-style API -------------------------------------------------------------------
    def seed(self, seed=None):
        if seed is not None:
            self.seed_value = int(seed)
        return [seed]

    def reset(self, seed=None, options=None, mask: Optional[torch.Tensor] = None):
        """mj_resetData + default ctrl for all (or the masked) environments -> (obs, info) (quadruped.py:115-139)."""
        if seed is not None:
            self.seed_value = int(seed)
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(_lib.lib().qg_reset(self._batch, _ptr(m), self.seed_value, int(self.random_init), self.env_offset,
                                       self._stream()), "qg_reset")
        if m is None:
            self._obs.zero_()
            self._sd, self._sd_pending = torch.zeros_like(self._obs), False
        else:
            sd = self._sensordata()
            self._sd, self._sd_pending = (sd.clone() if sd is self._obs else sd), False
            self._obs[m.bool()] = 0
            self._sd[m.bool()] = 0
        return self._obs, {}

    def _sensordata(self) -> torch.Tensor:
        """``data.sensordata``: sensordata of the last forward pass, also for environments that were auto-reset in the
        same call (their returned observation is the reset observation).  Valid until the next ``step``."""
        if self._sd_pending:
            t = self._terminated.view(torch.bool)
            self._sd = torch.where(t[:, None], self._term_obs, self._obs) if self._sd_merge else self._obs
            self._sd_pending = False
        return self._sd

    def step(self, action: torch.Tensor):
        self._sync_tables()
        a = torch.as_tensor(action, device=self.device, dtype=torch.float32).reshape(self.num_envs, 12).contiguous()
        nterm = len(self._fused)
        _lib.check(_lib.lib().qg_step(self._batch, _ptr(a), self.frame_skip, _ptr(self._obs), _ptr(self._reward),
                                      _ptr(self._terms) if nterm else None, _ptr(self._terminated),
                                      _ptr(self._term_obs), self._stream()), "qg_step")
        terminated = self._terminated.view(torch.bool)   # 0/1 bytes: a view, no kernel
        self._sd_pending, self._sd_merge = True, self._kernel_reset
        reward = self._reward
        info = {"reward_components": {}}
        if nterm:
            tv = self._terms[:, :nterm] if nterm == _lib.QG_MAX_TERMS else self._terms.view(-1)[: self.num_envs * nterm].view(self.num_envs, nterm)
            for i, (k, _) in enumerate(self._fused):
                info["reward_components"][k] = tv[:, i]
        for k in self._zero_terms:
            info["reward_components"][k] = self._zero_reward
        for k, fn in self._pyfn:
            r = fn()
            r = torch.as_tensor(r, device=self.device, dtype=torch.float32).expand(self.num_envs)
            info["reward_components"][k] = r
            reward = reward + r
        for _, fn in self._pyterm:
            terminated = terminated | torch.as_tensor(fn(), device=self.device).bool().expand(self.num_envs)
        if self.auto_reset and not self._kernel_reset and bool(terminated.any()):
            self._term_obs = torch.where(terminated[:, None], self._obs, torch.zeros_like(self._obs))
            self.reset(mask=terminated)
        info["terminal_observation"] = self._term_obs
        return self._obs, reward, terminated, self._truncated, info

    def pinned_action_buffer(self) -> np.ndarray:
        """A page-locked [N,12] float32 array: fill it and pass it to ``step_host`` for a staging-free H2D copy."""
        if not hasattr(self, "_h_act"):
            self._h_act_t = torch.empty((self.num_envs, 12), dtype=torch.float32).pin_memory()
            self._h_act = self._h_act_t.numpy()
        return self._h_act

    def step_host(self, action: np.ndarray, want_terms: bool = False, want_terminal_obs: bool = False, wait: bool = True):
        """End-to-end call with HOST buffers (numpy in, numpy out): qg_step_host -- segments of the batch are pipelined
        so that the PCIe copies run under the kernels.  The returned arrays are views of page-locked buffers owned by
        the env (valid until the next call).  ``wait=False`` returns right after enqueueing (``host_wait()`` completes
        the step); info carries the fused reward terms / terminal observations when asked for."""
        self._sync_tables()
        if self._pyfn or self._pyterm:
            raise NotImplementedError("step_host evaluates fused reward / termination specs only (rewards.py); Python callables "
                                      "need device tensors: use step()")
        a = np.ascontiguousarray(action, dtype=np.float32).reshape(self.num_envs, 12)
        n = self.num_envs
        if not hasattr(self, "_h_obs"):
            self._h_out = [torch.empty((n, 33), dtype=torch.float32).pin_memory(), torch.empty((n,), dtype=torch.float32).pin_memory(),
                           torch.empty((n,), dtype=torch.uint8).pin_memory()]
            self._h_obs, self._h_rew, self._h_term = (t.numpy() for t in self._h_out)
            self._h_terms_t = self._h_tobs_t = None
        nterm = len(self._fused)
        if want_terms and nterm and self._h_terms_t is None:
            self._h_terms_t = torch.empty((n * _lib.QG_MAX_TERMS,), dtype=torch.float32).pin_memory()
        if want_terminal_obs and self._h_tobs_t is None:
            self._h_tobs_t = torch.zeros((n, 33), dtype=torch.float32).pin_memory()
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        terms = self._h_terms_t.numpy() if (want_terms and nterm) else None
        tobs = self._h_tobs_t.numpy() if want_terminal_obs else None
        fn = _lib.lib().qg_step_host if wait else _lib.lib().qg_step_host_async
        _lib.check(fn(self._batch, vp(a), self.frame_skip, vp(self._h_obs), vp(self._h_rew), None if terms is None else vp(terms),
                      vp(self._h_term), None if tobs is None else vp(tobs), self._stream()), "qg_step_host")
        self._sd_pending, self._sd_merge = False, False
        info = {}
        if terms is not None:
            tv = terms[: n * nterm].reshape(n, nterm)
            info["reward_components"] = {k: tv[:, i] for i, (k, _) in enumerate(self._fused)}
        if tobs is not None:
            info["terminal_observation"] = tobs
        return self._h_obs, self._h_rew, self._h_term.view(np.bool_), np.zeros(n, dtype=bool), info

    def host_wait(self):
        """Block until the outputs of the last ``step_host(wait=False)`` are in the host buffers."""
        _lib.check(_lib.lib().qg_host_wait(self._batch, self._stream()), "qg_host_wait")

    def render(self, index: int = 0, overlays=()):
        """Rendering bridge: draw ONE environment through a CPU ``mujoco`` renderer with the reference's
        camera, pacing and modes (quadruped.py:77-86,250-306).  Needs the ``mujoco`` wheel; the batched physics never
        depends on it."""
        if self.render_mode is None and not self._render_kwargs.get("save_video"):
            return None
        if self._renderer is None:
            from .render import MujocoRenderBridge
            kw = self._render_kwargs
            self._renderer = MujocoRenderBridge(self.model_path, self.render_mode, kw.get("width", 720), kw.get("height", 480),
                                                kw.get("render_fps", 30), kw.get("save_video", False), kw.get("video_path", "videos/simulation.mp4"))
        return self._renderer.frame(self.data.qpos[index].double().cpu().numpy(), float(self.data.time[index]), overlays)

    # -- state access for parity tests -----------------------------------------------------------
    def set_state(self, qpos=None, qvel=None, act=None, qacc_warmstart=None, time=None, ctrl=None):
        def prep(x, k, dt=torch.float32):
            if x is None:
                return None
            t = torch.as_tensor(x, device=self.device, dtype=dt)
            return t.reshape((self.num_envs, k) if k else (self.num_envs,)).contiguous()
        keep = [prep(qpos, 19), prep(qvel, 18), prep(act, 12), prep(qacc_warmstart, 18), prep(time, 0, torch.float64), prep(ctrl, 12)]
        _lib.check(_lib.lib().qg_set_state(self._batch, _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]),
                                           _ptr(keep[4]), _ptr(keep[5]), self._stream()), "qg_set_state")
        torch.cuda.current_stream(self.device).synchronize()

    def debug_step(self, ctrl):
        """One mj_step with stage outputs (qg_debug_step) -> dict of tensors."""
        n, dev = self.num_envs, self.device
        c = torch.as_tensor(ctrl, device=dev, dtype=torch.float32).reshape(n, 12).contiguous()
        out = {"qacc": torch.zeros((n, 18), device=dev), "qacc_smooth": torch.zeros((n, 18), device=dev),
               "qfrc_bias": torch.zeros((n, 18), device=dev), "M": torch.zeros((n, 18, 18), device=dev),
               "counts": torch.zeros((n, 4), dtype=torch.int32, device=dev), "sensordata": torch.zeros((n, 33), device=dev)}
        _lib.check(_lib.lib().qg_debug_step(self._batch, _ptr(c), _ptr(out["qacc"]), _ptr(out["qacc_smooth"]),
                                            _ptr(out["qfrc_bias"]), _ptr(out["M"]), _ptr(out["counts"]),
                                            _ptr(out["sensordata"]), self._stream()), "qg_debug_step")
        return out

    def counters(self, reset: bool = False) -> dict:
        c = _lib.Counters()
        _lib.check(_lib.lib().qg_get_counters(self._batch, C.byref(c), int(reset), self._stream()), "qg_get_counters")
        return c.as_dict()

    def close(self):
        L = _lib.lib()
        if getattr(self, "_renderer", None) is not None:
            self._renderer.close()
            self._renderer = None
        if getattr(self, "_batch", None):
            L.qg_batch_destroy(self._batch)
            self._batch = None
        if getattr(self, "_model", None):
            L.qg_model_destroy(self._model)
            self._model = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class QuadrupedEnv(_EnvBase):
    """Single-environment shim with the reference's exact signature and return types
    (/root/reference/src/envs/quadruped.py:40-52,115,153): numpy observation, float reward, bool flags.
    Reward / termination callables are zero-argument Python functions over ``env.data`` / ``env.model``
    as in the reference; fused specs from ``rewards`` are accepted too."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}

    def __init__(self, model_path: Optional[str] = None, max_time: float = 10.0, frame_skip: int = 4,
                 render_mode: str = None, width: int = 720, height: int = 480, render_fps: int = 30,
                 reward_fns: dict = None, termination_fns: dict = None, save_video: bool = False,
                 video_path: str = "videos/simulation.mp4", use_default_termination: bool = True, device="cuda:0"):
        if _gym is not None:
            super().__init__()
        self.vec = VecQuadrupedEnv(1, device=device, model_path=model_path, max_time=max_time, frame_skip=frame_skip,
                                   render_mode=render_mode, reward_fns=reward_fns, termination_fns=termination_fns,
                                   use_default_termination=use_default_termination, auto_reset=False, width=width, height=height,
                                   render_fps=render_fps, save_video=save_video, video_path=video_path)
        self.model, self.max_time, self.frame_skip, self.render_mode = self.vec.model, max_time, frame_skip, render_mode
        self.action_space, self.observation_space = self.vec.action_space, self.vec.observation_space
        from .single import SingleData
        self.data = SingleData(self.vec)

    @property
    def reward_fns(self):
        return self.vec.reward_fns

    @reward_fns.setter
    def reward_fns(self, v):
        self.vec.reward_fns = v

    @property
    def termination_fns(self):
        return self.vec.termination_fns

    @termination_fns.setter
    def termination_fns(self, v):
        self.vec.termination_fns = v

    def seed(self, seed=None):
        np.random.seed(seed)
        return [seed]

    def reset(self, seed=None, options=None):
        obs, _ = self.vec.reset()
        return obs[0].double().cpu().numpy(), {}

    def step(self, action):
        action = np.clip(action, self.action_space.low, self.action_space.high)
        obs, rew, term, trunc, info = self.vec.step(torch.as_tensor(action, dtype=torch.float32)[None])
        comps = {k: float(v[0]) for k, v in info["reward_components"].items()}
        t = float(self.vec.data.time[0])
        return obs[0].double().cpu().numpy(), float(rew[0]), bool(term[0]), False, {"time": t, "reward_components": comps}

    def render(self):
        return self.vec.render(0)

    def close(self):
        self.vec.close()
