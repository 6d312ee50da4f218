"""Drop-in surface (CPU part): the ``envs`` shadow package under dropin/, the SB3 ``VecEnv`` adapter's lazy infos and base
class, the render bridge's pacing / overlay logic against a recording stub of the MuJoCo calls it makes, and -- where the
reference checkout is present (this container) -- the reference's OWN ``train_quadruped.py`` imported unmodified against
the shadow package with stub third-party modules: its ``make_env`` must resolve our class with the reference's keywords
and get as far as the device (no CUDA device here -> the library's "no CPU fallback" error)."""
import ast
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dropin")
REF_SRC = "/root/reference/src"


@pytest.fixture()
def shadow_envs(monkeypatch):
    monkeypatch.syspath_prepend(DROPIN)
    for m in [k for k in sys.modules if k == "envs" or k.startswith("envs.")]:
        monkeypatch.delitem(sys.modules, m)
    yield
    for m in [k for k in sys.modules if k == "envs" or k.startswith("envs.")]:
        sys.modules.pop(m, None)


def test_shadow_package_exports_reference_names(shadow_envs):
    from envs.po_walking_quad import POWalkingQuadrupedEnv
    from envs.quadruped import QuadrupedEnv
    from envs.walking_quad import WalkingQuadrupedEnv
    from envs.control_inputs import VelocityHeadingControls
    from envs.math_utils import exp_dist, unit
    assert issubclass(POWalkingQuadrupedEnv, WalkingQuadrupedEnv)
    assert len(POWalkingQuadrupedEnv.reward_keys) == 11 and POWalkingQuadrupedEnv.reward_keys[0] == "alive_bonus"
    assert exp_dist(0.0) == 0.0 and np.allclose(unit(np.array([3.0, 4.0])), [0.6, 0.8])
    for name in ("set_orientation", "set_velocity_speed_alpha", "set_velocity_xy", "get_heading_theta", "sample"):
        assert hasattr(VelocityHeadingControls, name)
    assert QuadrupedEnv.metadata["render_modes"] == ["human", "rgb_array"]


def _init_keywords(path, cls):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name == "__init__":
                    return [a.arg for a in f.args.args[1:]]
    raise AssertionError(cls)


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference checkout not present")
def test_constructor_keywords_cover_the_reference(shadow_envs):
    import inspect
    from envs.po_walking_quad import POWalkingQuadrupedEnv
    from envs.quadruped import QuadrupedEnv
    from envs.walking_quad import WalkingQuadrupedEnv
    base = _init_keywords(os.path.join(REF_SRC, "envs/quadruped.py"), "QuadrupedEnv")
    walk = _init_keywords(os.path.join(REF_SRC, "envs/walking_quad.py"), "WalkingQuadrupedEnv")
    po = _init_keywords(os.path.join(REF_SRC, "envs/po_walking_quad.py"), "POWalkingQuadrupedEnv")
    ours_base = list(inspect.signature(QuadrupedEnv.__init__).parameters)
    ours_walk = list(inspect.signature(WalkingQuadrupedEnv.__init__).parameters)
    ours_po = list(inspect.signature(POWalkingQuadrupedEnv.__init__).parameters)
    assert set(base) <= set(ours_base)
    assert set(walk) | set(base) <= set(ours_walk)          # **kwargs of the reference = the base keywords
    assert set(po) <= set(ours_po)
    assert ours_walk[1:5] == walk and ours_po[1] == "obs_window"      # positional order of the leading keywords


def _stub(monkeypatch, name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    monkeypatch.setitem(sys.modules, name, m)
    return m


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference checkout not present")
def test_reference_train_script_resolves_our_env(shadow_envs, monkeypatch):
    """train_quadruped.py, unmodified, imported with stub SB3 / plotting modules: ``make_env`` builds OUR
    POWalkingQuadrupedEnv with the reference's keywords (max_time, frame_skip, obs_window, random_controls,
    reset_options).  Without a GPU the construction must stop exactly at the device boundary."""
    import torch
    class _Any:
        def __init__(self, *a, **k): pass
    sb3 = _stub(monkeypatch, "stable_baselines3", PPO=_Any, SAC=_Any, TD3=_Any)
    _stub(monkeypatch, "stable_baselines3.common")
    _stub(monkeypatch, "stable_baselines3.common.callbacks", BaseCallback=_Any)
    _stub(monkeypatch, "stable_baselines3.common.vec_env", SubprocVecEnv=_Any, VecEnv=_Any)
    _stub(monkeypatch, "utils")
    _stub(monkeypatch, "utils.plot", plot_data_line=lambda *a, **k: None, plot_reward_components=lambda *a, **k: None)
    _stub(monkeypatch, "matplotlib")
    _stub(monkeypatch, "matplotlib.pyplot")
    spec = importlib.util.spec_from_file_location("ref_train_quadruped", os.path.join(REF_SRC, "train_quadruped.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                     # the __main__ block does not run
    from quadruped_gym_b200.envs.single import POWalkingQuadrupedEnv as Ours
    assert mod.POWalkingQuadrupedEnv is Ours
    options = {"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3}
    if torch.cuda.is_available():
        env = mod.make_env(options)
        assert env.observation_space.shape == (260,) and env.frame_skip == 10 and env.max_time == 20
        env.close()
    else:
        from quadruped_gym_b200._lib import QuadGymLibraryError
        with pytest.raises(QuadGymLibraryError, match="CUDA"):
            mod.make_env(options)


def test_lazy_infos_behave_like_a_list_of_dicts():
    from quadruped_gym_b200.envs.sb3 import LazyInfos
    keys = ["alive_bonus", "control_cost"]
    terms = np.arange(10, dtype=np.float32).reshape(5, 2)
    dones = np.array([False, True, False, False, True])
    tobs = np.array([[1.0, 2.0], [3.0, 4.0]], dtype=np.float32)
    infos = LazyInfos(keys, terms, dones, tobs)
    assert len(infos) == 5 and infos[1]["control_cost"] == 3.0 and infos[-1]["alive_bonus"] == 8.0
    assert np.mean([info["alive_bonus"] for info in infos]) == terms[:, 0].mean() == infos.mean("alive_bonus")   # RewardCallback's pattern
    assert infos[0].get("terminal_observation") is None and infos[0].get("episode") is None
    assert np.array_equal(infos[1]["terminal_observation"], [1.0, 2.0]) and np.array_equal(infos[4].get("terminal_observation"), [3.0, 4.0])
    assert infos[2].get("TimeLimit.truncated", False) is False and "terminal_observation" not in infos[2]
    assert dict(infos[4]).keys() == {"alive_bonus", "control_cost", "TimeLimit.truncated", "terminal_observation"}
    assert isinstance(infos[0]["alive_bonus"], float)
    with pytest.raises(KeyError):
        infos[0]["nope"]


def test_sb3_adapter_subclasses_vecenv_when_sb3_is_importable(monkeypatch):
    class VecEnv:                                    # what stable_baselines3.common.vec_env exports
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space
    _stub(monkeypatch, "stable_baselines3")
    _stub(monkeypatch, "stable_baselines3.common")
    _stub(monkeypatch, "stable_baselines3.common.vec_env", VecEnv=VecEnv)
    import quadruped_gym_b200.envs.sb3 as sb3
    sb3 = importlib.reload(sb3)
    try:
        assert issubclass(sb3.SB3VecEnv, VecEnv) and sb3.SB3VecEnvAdapter is sb3.SB3VecEnv
        for m in ("reset", "step_async", "step_wait", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped", "seed"):
            assert callable(getattr(sb3.SB3VecEnv, m))
    finally:
        for k in ("stable_baselines3.common.vec_env", "stable_baselines3.common", "stable_baselines3"):
            monkeypatch.delitem(sys.modules, k)
        importlib.reload(sb3)


def test_render_bridge_pacing_and_overlays(monkeypatch, tmp_path):
    """The bridge against a recording stub of the MuJoCo API it uses: a frame is produced only when simulated time has
    advanced by 1/render_fps, the camera follows the base, three overlay geoms are appended, rgb_array returns pixels."""
    calls = []
    class Scene:
        def __init__(self): self.ngeom, self.maxgeom, self.geoms = 0, 10, [object() for _ in range(10)]
    class Renderer:
        def __init__(self, m, height, width): self.scene, self.hw = Scene(), (height, width)
        def update_scene(self, d, scene_option=None, camera=None): self.scene.ngeom = 0; calls.append(("update", camera.lookat.copy()))
        def render(self): return np.zeros((*self.hw, 3), np.uint8)
        def close(self): calls.append(("close",))
    class Data:
        def __init__(self, m): self.qpos, self.time = np.zeros(19), 0.0
    class Cam:
        def __init__(self): self.lookat = np.zeros(3); self.distance = self.elevation = self.azimuth = 0
    class Opt:
        def __init__(self): self.flags, self.geomgroup, self.frame = np.zeros(32, bool), np.zeros(6, int), 0
    ns = types.SimpleNamespace
    scene_file = tmp_path / "scene.xml"
    scene_file.write_text("<mujoco/>")
    _stub(monkeypatch, "mujoco", MjModel=ns(from_xml_path=lambda p: object()), MjData=Data, MjvCamera=Cam, MjvOption=Opt, Renderer=Renderer,
          mjtVisFlag=ns(mjVIS_JOINT=0, mjVIS_CONTACTPOINT=1), mjtFrame=ns(mjFRAME_SITE=3), mjtGeom=ns(mjGEOM_ARROW1=100, mjGEOM_SPHERE=2),
          mj_forward=lambda m, d: calls.append(("forward", d.qpos[:3].copy())),
          mjv_initGeom=lambda *a: calls.append(("init", a[1])), mjv_connector=lambda *a: calls.append(("connector", a[2])))
    from quadruped_gym_b200.envs.render import MujocoRenderBridge
    br = MujocoRenderBridge(str(scene_file), "rgb_array", width=64, height=48, fps=30)
    q = np.zeros(19); q[:3] = [0.1, 0.2, 0.13]
    overlays = [("arrow", q[:3], np.array([1.0, 0, 0]), (1, 0, 0, 1), 0.1), ("arrow", q[:3], np.array([0, 1.0, 0]), (0, 1, 0, 1), 0.05),
                ("point", np.zeros(3), (1, 0, 1, 1), 0.0)]
    assert br.frame(q, 0.008, overlays) is None and not calls            # 8 ms: no frame due at 30 fps
    px = br.frame(q, 0.040, overlays)
    assert px.shape == (48, 64, 3)
    assert [c[0] for c in calls] == ["forward", "update", "init", "connector", "init", "connector", "init"]
    assert np.allclose(calls[1][1], q[:3])                                # camera follows the base (quadruped.py:238-244)
    assert br.frame(q, 0.041, overlays) is None                           # same frame slot
    assert br.frame(q, 0.070, overlays) is not None
    br.restart()
    assert br.frame(q, 0.040, overlays) is not None                       # reset() restarts the frame counter
    br.close()
    assert calls[-1] == ("close",)
    with pytest.raises(ValueError):
        MujocoRenderBridge(None, "ascii")


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference checkout not present")
def test_shadow_package_mirrors_the_reference_modules():
    """Every importable module of the reference's src/envs has a counterpart in dropin/envs (dummy_walking_quad.py imports
    a module that does not exist in the reference, SURVEY 2 #8, and is left out)."""
    ref = {f for f in os.listdir(os.path.join(REF_SRC, "envs")) if f.endswith(".py")} - {"dummy_walking_quad.py"}
    ours = {f for f in os.listdir(os.path.join(DROPIN, "envs")) if f.endswith(".py")}
    assert ref <= ours, ref - ours
