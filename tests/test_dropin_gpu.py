"""Drop-in surface on the GPU: the single-environment classes driven the way the reference's scripts drive them
(/root/reference/src/train_quadruped.py:15-27,171-193, eval_quadruped.py:11-27), writable ``env.data`` views, and the
SB3 ``VecEnv`` adapter driven the way SB3's ``collect_rollouts`` and the script's ``RewardCallback._on_step``
(train_quadruped.py:86-92) drive a vector env.  The call patterns are restated here (the reference checkout does not
travel to the GPU box); tests/test_dropin.py imports the reference's own script where it is present."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def envs_pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    for m in [k for k in sys.modules if k == "envs" or k.startswith("envs.")]:
        del sys.modules[m]
    import envs.po_walking_quad as po
    yield po
    sys.path.remove(os.path.join(ROOT, "dropin"))


def test_training_script_construction_and_episode(envs_pkg):
    """make_env of train_quadruped.py:15-27 and the evaluation loop of :171-193 / eval_quadruped.py:19-27."""
    options = {"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3}
    env = envs_pkg.POWalkingQuadrupedEnv(max_time=0.3, frame_skip=10, obs_window=10, random_controls=True, reset_options=options)
    assert env.observation_space.shape == (260,) and env.action_space.shape == (12,)
    obs, info = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (260,) and obs.dtype == np.float64 and info == {}
    assert np.allclose(env.control_inputs.velocity[:2], [0.3, 0.0]) and abs(env.control_inputs.get_heading_theta()) < 1e-12
    done, steps, rewards = False, 0, []
    rng = np.random.default_rng(0)
    while not done:
        action = rng.uniform(-1.5, 1.5, 12).astype(np.float32)       # model.predict output: unclipped float32
        obs, reward, done, _, info = env.step(action)
        assert obs.shape == (260,) and isinstance(reward, float) and isinstance(done, bool) and _ is False
        assert list(info) == envs_pkg.POWalkingQuadrupedEnv.reward_keys and all(isinstance(v, float) for v in info.values())
        assert reward == pytest.approx(sum(info.values()), rel=1e-5, abs=1e-4)
        assert np.array_equal(obs[-26:][11:23], np.clip(action, -1, 1).astype(np.float64))     # newest frame carries data.ctrl
        rewards.append(reward)
        steps += 1
        assert steps <= 15
    assert steps == 15                                 # 0.3 s at 20 ms per step: the fp64 clock reads 0.3000000000000002 after step 15
    assert env.render() is None                        # render_mode None
    obs2, _ = env.reset()                              # the script resets by hand after `done`
    assert np.allclose(obs2[:6], 0) and env.data.time == 0.0      # gyro / accel zero; the Euler angles keep the stale filter state
    env.close()


def test_eval_script_sequence_and_writable_data(envs_pkg):
    """eval_quadruped.py:11-15: commands set by hand before reset; reference-style in-place writes reach the device."""
    env = envs_pkg.POWalkingQuadrupedEnv(obs_window=5)
    env.control_inputs.set_orientation(0)
    env.control_inputs.set_velocity_speed_alpha(0.2, 0)
    obs, _ = env.reset()
    assert obs.shape == (130,)
    f = obs.reshape(5, 26)
    assert np.allclose(f[:, 23:26], [0.2, 0.0, 0.0], atol=1e-7) and np.allclose(f[0, 11:23], [0, 0, -0.5] * 4)
    assert np.allclose(env.control_inputs.global_velocity, [0.2, 0.0, 0.0])
    env.control_inputs.set_orientation(np.pi / 2)
    assert np.allclose(env.control_inputs.global_velocity, [0.0, 0.2, 0.0], atol=1e-15) and np.allclose(env.control_inputs.velocity[:2], [0.2, 0.0])
    # walking_quad.py:74 / quadruped.py:124 style writes
    half = 0.3
    env.data.qpos[3:7] = np.array([np.cos(half), 0, 0, np.sin(half)])
    env.data.ctrl[:] = np.array([0.1, 0.2, -0.3] * 4)
    assert np.allclose(env.data.qpos[3:7], [np.cos(half), 0, 0, np.sin(half)], atol=1e-7)
    assert np.allclose(env.vec.data.qpos[0, 3:7].cpu().numpy(), [np.cos(half), 0, 0, np.sin(half)], atol=1e-7)
    assert np.allclose(env.data.ctrl, [0.1, 0.2, -0.3] * 4, atol=1e-7)
    env.data.time = 1.5
    assert env.data.time == 1.5
    with pytest.raises(AttributeError):
        env.data.nonsense = 1
    obs, r, done, _, info = env.step(np.zeros(12, np.float32))
    assert env.data.time == pytest.approx(1.508) and np.isfinite(obs).all()
    env.close()
    # a rendering env constructs headless and fails loudly only at render() time without the mujoco wheel
    env = envs_pkg.POWalkingQuadrupedEnv(render_mode="human", obs_window=5)
    env.reset()
    for _ in range(6):
        env.step(np.zeros(12, np.float32))
    try:
        import mujoco  # noqa: F401
    except ImportError:
        with pytest.raises((NotImplementedError, FileNotFoundError)):
            env.render()
    env.close()


def test_single_env_equals_one_row_of_the_vector_env(envs_pkg):
    from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv
    kw = dict(max_time=5.0, frame_skip=10, obs_window=4, settling_time=0.1)
    one = envs_pkg.POWalkingQuadrupedEnv(**kw)
    vec = VecPOWalkingQuadrupedEnv(3, "cuda:0", auto_reset=False, **kw)
    one.control_inputs.set_velocity_speed_alpha(0.25, 0.1)
    vec.control_inputs.set_speed_alpha_theta(0.25, 0.1, 0.0)
    o1, _ = one.reset()
    ov, _ = vec.reset()
    assert np.array_equal(o1, ov[1].double().cpu().numpy())
    rng = np.random.default_rng(3)
    for t in range(25):
        a = rng.uniform(-1, 1, 12).astype(np.float32)
        o1, r1, d1, _, i1 = one.step(a)
        ov, rv, dv, _, iv = vec.step(torch.from_numpy(np.tile(a, (3, 1))).cuda())
        assert np.array_equal(o1, ov[1].double().cpu().numpy()) and r1 == float(rv[1]) and d1 == bool(dv[1])
        assert all(i1[k] == float(iv[k][1]) for k in one.reward_keys)
    one.close(); vec.close()


def test_sb3_vecenv_rollout_pattern(envs_pkg):
    """SB3's collect_rollouts + the script's RewardCallback against SB3VecEnv over 512 device environments."""
    from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv
    n = 512
    options = {"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3}
    venv = envs_pkg.SB3VecEnv(VecPOWalkingQuadrupedEnv(n, "cuda:0", max_time=0.1, frame_skip=10, obs_window=10, random_controls=True,
                                                        reset_options=options))
    keys = venv.reward_keys
    last_obs = venv.reset()
    assert last_obs.shape == (n, 260) and last_obs.dtype == np.float32
    rng = np.random.default_rng(1)
    seen_done = 0
    buf = []
    for t in range(12):
        actions = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        new_obs, rewards, dones, infos = venv.step(actions)
        buf.append(last_obs)                                       # SB3 stores _last_obs AFTER env.step: it must not alias new_obs
        assert new_obs is not last_obs and not np.shares_memory(new_obs, last_obs)
        assert rewards.shape == (n,) and dones.dtype == bool and len(infos) == n
        # RewardCallback._on_step (train_quadruped.py:86-92)
        comps = {key: np.mean([info[key] for info in infos]) for key in keys}
        assert np.isfinite(list(comps.values())).all() and comps["alive_bonus"] == 10.0
        assert np.mean(rewards) == pytest.approx(sum(comps.values()), rel=1e-4, abs=1e-3)
        for idx, done in enumerate(dones):                         # collect_rollouts' bootstrap branch
            if done and infos[idx].get("terminal_observation") is not None and infos[idx].get("TimeLimit.truncated", False):
                raise AssertionError("the reference never truncates")
            if done:
                seen_done += 1
                assert infos[idx]["terminal_observation"].shape == (260,)
                assert np.array_equal(new_obs[idx].reshape(10, 26)[0], new_obs[idx].reshape(10, 26)[-1])   # reset stack
        last_obs = new_obs
    assert seen_done == 2 * n                                      # max_time 0.1 s at 20 ms: every env ends at steps 5 and 10
    assert not any(np.shares_memory(a, b) for a, b in zip(buf[:-1], buf[1:]))
    assert venv.get_attr("frame_skip") == [10] * n and venv.env_is_wrapped(object, indices=[0, 1]) == [False, False]
    venv.close()
