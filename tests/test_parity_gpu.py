"""GPU parity tests: the sm_100a path (through the C ABI) against the float64 CPU oracle on the same
seeded inputs.  Tolerances (fp32 device arithmetic vs fp64 oracle), stated per quantity:

  * mass matrix, bias forces:      1e-5 relative to the largest entry
  * qacc_smooth / qacc:            1e-4 relative to max|qacc| per env  (north star: 1e-4 contact-free,
                                   1e-3 with contacts) -- contact rows require the same contact set
  * next qpos / qvel / act:        1e-5 absolute on qpos/act, 1e-4 relative on qvel
  * sensordata:                    1e-5 absolute, accelerometer channels 1e-4 * max(1, |qacc|max)
  * reward terms / termination:    exact given equal state (float64 formulas on float32 inputs)
"""
import os

import numpy as np
import pytest
import torch

from oracle.oracle import OracleData
from tests import ref_formulas as F
from tests.conftest import rollout_states

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "env_traces.npz"))


@pytest.fixture(scope="module")
def Vec():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from quadruped_gym_b200 import VecQuadrupedEnv
    return VecQuadrupedEnv


def _oracle_at(om, st32, time, ctrl, e):
    d = OracleData(om)
    d.set_state(st32["qpos"][e].astype(np.float64), st32["qvel"][e].astype(np.float64), st32["act"][e].astype(np.float64),
                st32["warm"][e].astype(np.float64), time[e], ctrl[e].astype(np.float64))
    return d


def _bform(M, R):
    T = np.eye(18)
    T[:3, :3] = R
    return T.T @ M @ T


@pytest.fixture(scope="module")
def teacher_forced(Vec, oracle_model):
    """384 states from oracle rollouts (flight, landing, standing, stumbling), one debug step on the GPU."""
    n = 384
    st = rollout_states(oracle_model, n, 150, seed=11)
    st32 = {k: v.astype(np.float32) for k, v in st.items() if k != "time"}
    env = Vec(n, "cuda:0", auto_reset=False)
    env.set_state(qpos=st32["qpos"], qvel=st32["qvel"], act=st32["act"], qacc_warmstart=st32["warm"], time=st["time"], ctrl=st32["ctrl"])
    ctrl = np.random.default_rng(3).uniform(-1.2, 1.2, (n, 12)).astype(np.float32)
    out = {k: v.cpu().numpy() for k, v in env.debug_step(ctrl).items()}
    nxt = {"qpos": env.data.qpos.cpu().numpy(), "qvel": env.data.qvel.cpu().numpy(), "act": env.data.act.cpu().numpy(),
           "warm": env.data.qacc_warmstart.cpu().numpy(), "time": env.data.time.cpu().numpy()}
    ora = [_oracle_at(oracle_model, st32, st["time"], ctrl, e) for e in range(n)]
    for d in ora:
        d.forward()
    env.close()
    return n, st, st32, ctrl, out, nxt, ora


def test_smooth_dynamics_stage_parity(teacher_forced):
    n, st, st32, ctrl, out, nxt, ora = teacher_forced
    for e, d in enumerate(ora):
        R = d.xmat[9:18].reshape(3, 3)
        MB = _bform(d.M.copy(), R)
        assert np.abs(out["M"][e] - MB).max() <= 1e-5 * np.abs(MB).max()
        bias = d.qfrc_bias.copy()
        bias[:3] = R.T @ bias[:3]
        assert np.abs(out["qfrc_bias"][e] - bias).max() <= 1e-5 * max(1.0, np.abs(bias).max())
        assert np.abs(out["qacc_smooth"][e] - d.qacc_smooth).max() <= 1e-4 * max(1.0, np.abs(d.qacc_smooth).max())


def test_contact_sets_and_constrained_acceleration(teacher_forced):
    n, st, st32, ctrl, out, nxt, ora = teacher_forced
    ncon = np.array([d.ncon for d in ora])
    nefc = np.array([d.nefc for d in ora])
    same = (out["counts"][:, 0] == ncon) & (out["counts"][:, 1] == nefc)
    assert same.mean() >= 0.98            # contact-set flips only at fp32-vs-fp64 ties
    assert (ncon > 0).sum() >= 100 and (ncon == 0).sum() >= 20    # the sample covers both regimes
    flip_err = 0.0
    for e, d in enumerate(ora):
        err = np.abs(out["qacc"][e] - d.qacc).max() / max(1.0, np.abs(d.qacc).max())
        if not same[e]:        # contact set differs at an fp32-vs-fp64 tie (a vertex within round-off of the margin):
            flip_err = max(flip_err, err)       # one contact more or less, bounded instead of skipped
            continue
        tol = 1e-4 if d.nefc == 0 else 1e-3
        assert err <= tol, (e, d.ncon)
    assert flip_err <= 1.0      # never the scale of the acceleration itself
    # solver effort comparable to the oracle's
    assert out["counts"][:, 2].mean() <= np.mean([d.solver_niter for d in ora]) + 1.0


def test_collision_queue_under_load(Vec, oracle_model):
    """Robots pressed flat against the floor in random attitudes: most geoms survive the cull, so a warp's collision queue
    carries over between batches (> 32 candidates) and leg lanes hold many contacts.  Contact counts and the constrained
    acceleration must still match the oracle environment by environment."""
    n = 256
    rng = np.random.default_rng(101)
    env = Vec(n, "cuda:0", auto_reset=False)
    env.reset()
    qpos = env.data.qpos.cpu().numpy().copy()
    qpos[:, 2] = rng.uniform(0.0, 0.05, n)
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rng.uniform(0, np.pi, n)
    qpos[:, 3] = np.cos(ang / 2); qpos[:, 4:7] = ax * np.sin(ang / 2)[:, None]
    qpos[:, 7:] += rng.uniform(-0.5, 0.5, (n, 12)).astype(np.float32)
    qvel = (0.1 * rng.normal(size=(n, 18))).astype(np.float32)
    zeros12 = np.zeros((n, 12), np.float32)
    env.set_state(qpos=qpos, qvel=qvel, act=zeros12, qacc_warmstart=np.zeros((n, 18), np.float32), time=np.zeros(n), ctrl=zeros12)
    ctrl = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
    out = {k: v.cpu().numpy() for k, v in env.debug_step(ctrl).items()}
    st32 = {"qpos": qpos, "qvel": qvel, "act": zeros12, "warm": np.zeros((n, 18), np.float32)}
    ncon = np.zeros(n, int)
    ok = 0
    for e in range(n):
        d = _oracle_at(oracle_model, st32, np.zeros(n), ctrl, e)
        d.forward()
        ncon[e] = d.ncon
        if out["counts"][e, 0] == d.ncon and out["counts"][e, 1] == d.nefc:
            ok += 1
            # deep penetrations: large forces, the fp32 solve is compared relative to the acceleration scale
            assert np.abs(out["qacc"][e] - d.qacc).max() <= 2e-3 * max(1.0, np.abs(d.qacc).max()), (e, d.ncon)
    assert ncon.mean() >= 6 and ncon.max() >= 12                 # the load this test is about
    assert np.add.reduceat(ncon, np.arange(0, n, 8)).max() >= 48  # some warp (8 envs) had well over 32 candidates
    assert ok >= 0.9 * n
    assert env.counters()["contact_overflow"] == 0
    env.close()


def test_sensordata_parity_and_lag(teacher_forced):
    n, st, st32, ctrl, out, nxt, ora = teacher_forced
    same = np.array([out["counts"][e, 0] == d.ncon for e, d in enumerate(ora)])
    for e, d in enumerate(ora):
        s, g = d.sensordata.copy(), out["sensordata"][e]
        # only the accelerometer depends on the contact set; everything else is the pre-integration state
        acc_tol = (1e-3 if same[e] else 1.0) * max(1.0, np.abs(d.qacc).max())
        assert np.abs(g[12:15] - s[12:15]).max() <= acc_tol
        s[12:15] = g[12:15] = 0
        assert np.abs(g - s).max() <= 1e-5
        # sensordata is the PRE-integration state (quadruped.py:167 returns the last forward pass)
        assert np.array_equal(g[:12], st32["qpos"][e, 7:])
        assert np.array_equal(g[15:18], st32["qvel"][e, 3:6])


def test_next_state_parity(teacher_forced, oracle_model):
    n, st, st32, ctrl, out, nxt, ora = teacher_forced
    same = np.array([out["counts"][e, 0] == d.ncon for e, d in enumerate(ora)])
    h = 0.002
    for e in range(n):
        d = _oracle_at(oracle_model, st32, st["time"], ctrl, e)
        d.step()
        if not same[e]:       # flipped contact set: the error of one step is bounded by h * (acceleration scale), see above
            a = max(1.0, np.abs(ora[e].qacc).max())
            assert np.abs(nxt["qvel"][e] - d.qvel).max() <= 2 * h * a and np.abs(nxt["qpos"][e] - d.qpos).max() <= 2 * h * h * a + 1e-5
            assert nxt["time"][e] == d.time
            continue
        assert np.abs(nxt["qpos"][e] - d.qpos).max() <= 1e-5
        assert np.abs(nxt["act"][e] - d.act).max() <= 1e-5
        assert np.abs(nxt["qvel"][e] - d.qvel).max() <= 1e-4 * max(1.0, np.abs(d.qvel).max())
        assert np.abs(nxt["warm"][e] - d.qacc_warmstart).max() <= 1e-3 * max(1.0, np.abs(d.qacc_warmstart).max())
        assert nxt["time"][e] == d.time            # fp64 time accumulation, bit exact


def test_joint_limit_states(Vec, oracle_model):
    """Edge case: joints pushed beyond their ranges with servos driving further out (limit rows active)."""
    n = 32
    rng = np.random.default_rng(5)
    qpos = np.tile(np.r_[0, 0, 1.0, 1, 0, 0, 0, np.tile(np.deg2rad([-45, 37.5, 0]), 4)], (n, 1)).astype(np.float32)
    lo, hi = np.tile(np.deg2rad([-45, -45, -90]), 4), np.tile(np.deg2rad([45, 120, 90]), 4)
    below = rng.random((n, 12)) < 0.5
    qpos[:, 7:] = np.where(below, lo - rng.uniform(0.01, 0.2, (n, 12)), hi + rng.uniform(0.01, 0.2, (n, 12)))
    act = np.where(below, -1.5, 1.5).astype(np.float32)
    qvel = (rng.normal(size=(n, 18)) * 0.5).astype(np.float32)
    env = Vec(n, "cuda:0", auto_reset=False)
    env.set_state(qpos=qpos, qvel=qvel, act=act, qacc_warmstart=np.zeros((n, 18), np.float32), time=np.zeros(n), ctrl=np.zeros((n, 12), np.float32))
    ctrl = np.where(below, -1.0, 1.0).astype(np.float32)
    out = {k: v.cpu().numpy() for k, v in env.debug_step(ctrl).items()}
    active = 0
    for e in range(n):
        d = OracleData(oracle_model)
        d.set_state(qpos[e].astype(np.float64), qvel[e].astype(np.float64), act[e].astype(np.float64), np.zeros(18), 0.0, ctrl[e].astype(np.float64))
        d.forward()
        assert out["counts"][e, 1] == d.nefc == 12
        active += int((d.efc_force > 0).sum())
        assert np.abs(out["qacc"][e] - d.qacc).max() <= 1e-3 * max(1.0, np.abs(d.qacc).max())
    assert active > n * 6
    env.close()


def test_open_loop_rollout_bounded_divergence(Vec, oracle_model):
    """64 envs from reset through the 10 cm drop, landing and 40 more env steps under random actions:
    median error stays at fp32 round-off level; the worst env is bounded (contact-set flips amplify)."""
    n, T = 64, 60
    env = Vec(n, "cuda:0", auto_reset=False, frame_skip=4)
    obs0, _ = env.reset()
    assert torch.count_nonzero(obs0) == 0            # reset observation is all zeros (quadruped.py:120,138)
    rng = np.random.default_rng(21)
    acts = np.repeat(rng.uniform(-1, 1, (T // 5, n, 12)).astype(np.float32), 5, axis=0)
    ds = [OracleData(oracle_model) for _ in range(n)]
    for d in ds:
        d.ctrl[:] = [0, 0, -0.5] * 4
    med, worst = [], []
    for t in range(T):
        obs, rew, term, trunc, info = env.step(torch.from_numpy(acts[t]).cuda())
        o = obs.cpu().numpy().astype(np.float64)
        ref = np.zeros((n, 33))
        for e, d in enumerate(ds):
            d.env_step(acts[t, e].astype(np.float64), 4)
            ref[e] = d.sensordata
        err = np.abs(o - ref)
        err[:, 12:15] = 0
        med.append(np.median(err.max(1)))
        worst.append(err.max())
        assert not bool(trunc.any())
    assert max(med[:20]) < 1e-5          # contact-free flight: round-off only
    assert max(med) < 1e-4               # through landing and stance
    assert np.median(worst) < 1e-3
    qg, qo = env.data.qpos.cpu().numpy(), np.array([d.qpos.copy() for d in ds])
    assert np.median(np.abs(qg - qo).max(1)) < 1e-4
    env.close()


def test_golden_trace_A_env_step_semantics(Vec):
    """The reference's own QuadrupedEnv (run on the oracle physics) vs VecQuadrupedEnv on the same actions:
    clip, frame_skip loop, lagged sensordata, README reward trio, time, termination."""
    from quadruped_gym_b200.envs import rewards as R
    env = Vec(2, "cuda:0", auto_reset=False, reward_fns={"forward": R.forward_velocity(1.0), "control_cost": R.ctrl_sq(-0.1),
                                                          "alive_bonus": R.alive_bonus(1.0)})
    obs, _ = env.reset()
    assert np.array_equal(obs.cpu().numpy()[0], G["A_obs0"])
    T = 120   # through drop + landing; later steps diverge chaotically at fp32 level
    for t in range(T):
        a = torch.from_numpy(np.stack([G["A_actions"][t]] * 2)).cuda()     # unclipped actions: the env clips
        obs, rew, term, trunc, info = env.step(a)
        o = obs[0].cpu().numpy()
        err = np.abs(o - G["A_obs"][t])
        err[12:15] /= 100.0
        assert err.max() < (1e-4 if t < 25 else 5e-3), (t, err.max())
        # reward exact given the device's own state: float64 formulas on float32 inputs
        qv = env.data.qvel[0].cpu().numpy().astype(np.float64)
        ctrl = env.data.ctrl[0].cpu().numpy().astype(np.float64)
        assert np.array_equal(ctrl, np.clip(G["A_actions"][t], -1, 1).astype(np.float64))
        total, comps = F.readme_reward(qv, ctrl)
        assert rew[0].item() == np.float32(total)
        for k, c in zip(("forward", "control_cost", "alive_bonus"), comps):
            assert info["reward_components"][k][0].item() == np.float32(c)
        assert abs(rew[0].item() - G["A_reward"][t]) < 5e-3
        assert float(env.data.time[0]) == G["A_time"][t]
        assert bool(term[0]) == bool(G["A_terminated"][t]) and not bool(trunc[0])
        assert torch.equal(obs[0], obs[1]) and rew[0] == rew[1]           # determinism across lanes
    env.close()


@pytest.mark.parametrize("frame_skip,max_time,want", [(4, 10.0, 1250), (10, 20.0, 1001)])
def test_time_limit_termination_index(Vec, frame_skip, max_time, want):
    """fp64 time accumulation on the device: terminated first at env.step() #1250 / #1001 (golden trace B)."""
    assert int(G["B4_first_terminated_step" if frame_skip == 4 else "B10_first_terminated_step"]) == want
    env = Vec(4, "cuda:0", auto_reset=False, frame_skip=frame_skip, max_time=max_time)
    env.reset()
    a = torch.zeros((4, 12), device="cuda")
    first = None
    for i in range(1, want + 2):
        _, _, term, trunc, _ = env.step(a)
        if first is None and bool(term.any()):
            first = i
            assert bool(term.all())
    assert first == want
    env.close()


def test_fused_reward_terms_exact(Vec):
    """Every fused term equals the reference formula (tests/ref_formulas.py, pinned by golden trace C)
    evaluated in float64 on the device's own float32 sensordata / ctrl."""
    from quadruped_gym_b200.envs import rewards as R
    n = 8
    fns = {"alive": R.alive_bonus(10.0), "cc": R.control_cost(-2.0, 0.8), "ori": R.exp_orientation(10.0),
           "h": R.exp_body_height(-50.0, 0.13), "post": R.joint_posture_cost(-1.0), "fwd": R.forward_reward(5.0),
           "drift": R.no_drift_reward(-3.0), "o": R.orientation_reward(1.0), "hc": R.body_height_cost(1.0, 0.12)}
    env = Vec(n, "cuda:0", auto_reset=False, reward_fns=fns)
    env.reset()
    rng = np.random.default_rng(8)
    ccs = [F.ControlCost() for _ in range(n)]
    for t in range(40):
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        obs, rew, term, trunc, info = env.step(torch.from_numpy(a).cuda())
        o = obs.cpu().numpy().astype(np.float64)
        for e in range(n):
            ctrl = a[e].astype(np.float64)
            want = {"alive": 10.0 * 1, "cc": -2.0 * ccs[e](ctrl), "ori": 10.0 * F.exp_dist(F.orientation_reward(o[e])),
                    "h": -50.0 * F.exp_dist(F.body_height_cost(o[e], 0.13)), "post": -1.0 * F.joint_posture_cost(ctrl),
                    "fwd": 5.0 * F.forward_reward(o[e]), "drift": -3.0 * F.no_drift_reward(o[e]),
                    "o": F.orientation_reward(o[e]), "hc": F.body_height_cost(o[e], 0.12)}
            for k, v in want.items():
                got = info["reward_components"][k][e].item()
                if k in ("ori", "h", "post"):          # exp / sqrt: 1 ulp of float64 differences survive the f32 cast rarely
                    assert got == pytest.approx(np.float32(v), rel=2e-7), k
                else:
                    assert got == np.float32(v), k
            total = 0.0
            for k in fns:
                total += want[k]
            assert rew[e].item() == pytest.approx(np.float32(total), rel=1e-6)
    env.close()


@pytest.mark.parametrize("tag", ["D80", "D95"])
def test_flip_termination_and_auto_reset(Vec, tag):
    """Golden trace D: rolled start; terminated == (lagged zaxis_z < 0) exactly; auto-reset returns the zero
    observation, keeps the terminal one, and restarts time."""
    from quadruped_gym_b200.envs import rewards as R
    env = Vec(3, "cuda:0", auto_reset=True, termination_fns={"flip": R.flip_termination()})
    env.reset()
    half = 0.5 * np.deg2rad(float(G[tag + "_roll_deg"]))
    q = env.data.qpos.cpu().numpy()
    q[:, 3:7] = [np.cos(half), np.sin(half), 0, 0]
    env.set_state(qpos=q)
    a = torch.zeros((3, 12), device="cuda")
    flips = 0
    for t in range(len(G[tag + "_terminated"])):
        obs, rew, term, trunc, info = env.step(a)
        tobs = info["terminal_observation"]
        z = torch.where(term, tobs[:, 29], obs[:, 29])
        assert torch.equal(term, z < 0)                        # exact given the device's own sensordata
        if bool(term[0]):
            flips += 1
            assert torch.count_nonzero(obs[0]) == 0            # reset observation
            assert float(env.data.time[0]) == 0.0
            assert np.allclose(env.data.qpos[0].cpu().numpy()[:7], [0, 0, 0.13, 1, 0, 0, 0])
            break
        elif abs(G[tag + "_zaxis_z"][t]) > 0.02:
            assert bool(G[tag + "_terminated"][t]) is False
            assert abs(obs[0, 29].item() - G[tag + "_zaxis_z"][t]) < 0.02
    assert flips == 1
    t_flip = int(np.argmax(G[tag + "_terminated"]))
    assert abs(t - t_flip) <= 2
    env.close()


def test_shard_invariance_and_host_path(Vec):
    """Two half-batches with env_offset == one full batch (no exchange between environments), and the
    HOST-buffer C-ABI call returns the same numbers as the device call."""
    n = 64
    rng = np.random.default_rng(13)
    acts = rng.uniform(-1, 1, (30, n, 12)).astype(np.float32)
    full = Vec(n, "cuda:0", auto_reset=True)
    halves = [Vec(n // 2, "cuda:0", auto_reset=True, env_offset=i * n // 2) for i in range(2)]
    host = Vec(n, "cuda:0", auto_reset=True)
    for e in (full, host, *halves):
        e.reset()
    for t in range(30):
        o, r, te, _, _ = full.step(torch.from_numpy(acts[t]).cuda())
        parts = [h.step(torch.from_numpy(acts[t, i * n // 2:(i + 1) * n // 2]).cuda()) for i, h in enumerate(halves)]
        assert torch.equal(o, torch.cat([p[0] for p in parts]))
        assert torch.equal(r, torch.cat([p[1] for p in parts]))
        ho, hr, hte, _, _ = host.step_host(acts[t])
        assert np.array_equal(ho, o.cpu().numpy()) and np.array_equal(hr, r.cpu().numpy()) and np.array_equal(hte, te.cpu().numpy())
    c = full.counters()
    assert c["physics_steps"] == n * 30 * 4 and c["diverged"] == 0 and c["contact_overflow"] == 0
    for e in (full, host, *halves):
        e.close()


def test_non_finite_state_guard(Vec):
    """mj_checkPos analogue: a non-finite state resets that environment only and is counted."""
    env = Vec(4, "cuda:0", auto_reset=False)
    env.reset()
    q = env.data.qpos.cpu().numpy()
    q[2, 8] = np.nan
    env.set_state(qpos=q)
    obs, *_ = env.step(torch.zeros((4, 12), device="cuda"))
    assert bool(torch.isfinite(obs).all())
    assert env.counters()["diverged"] == 1
    assert torch.equal(obs[0], obs[1]) and torch.equal(obs[0], obs[3])
    env.close()


def test_single_env_shim_matches_reference_signature(Vec):
    from quadruped_gym_b200 import QuadrupedEnv
    env = QuadrupedEnv()
    env.reward_fns = {"forward": lambda: env.data.qvel[0], "alive": lambda: 1.0}     # README.md:65-78 style callables
    obs, info = env.reset()
    assert obs.shape == (33,) and obs.dtype == np.float64 and info == {} and not obs.any()
    o, r, te, tr, info = env.step(np.full(12, 3.0, dtype=np.float32))
    assert o.shape == (33,) and isinstance(r, float) and te is False and tr is False
    assert set(info) == {"time", "reward_components"} and info["time"] == pytest.approx(0.008)
    assert np.array_equal(env.data.ctrl, np.ones(12))           # clipped to the action space (quadruped.py:160)
    assert r == pytest.approx(env.data.qvel[0] + 1.0, abs=1e-6)
    assert env.action_space.shape == (12,) and env.observation_space.shape == (33,)
    env.close()
    with pytest.raises(FileNotFoundError):
        QuadrupedEnv(model_path="/nonexistent/scene.xml")        # quadruped.py:55-56


def test_walking_reward_stack_against_reference_trace(Vec):
    """Golden trace C (the reference's WalkingQuadrupedEnv): the device walking kernel fed the trace's sensordata
    (cast to float32, as the step kernel produces it) and ctrl reproduces all 11 reward terms, the total and the
    estimator state of the numpy restatement evaluated on the same float32 inputs.  Tolerance 1e-12 relative
    (float64 exp / sqrt / fused multiply-adds differ from numpy in the last bits); alive / time are exact."""
    import ctypes as C
    from quadruped_gym_b200 import _lib
    from quadruped_gym_b200.envs.walking_quad import VecWalkingQuadrupedEnv
    n = 4
    env = VecWalkingQuadrupedEnv(n, "cuda:0", auto_reset=False)
    env.reset()
    env.control_inputs.set_speed_alpha_theta(0.3, 0.1, 0.3)     # trace C: set_orientation(0.3), speed 0.3 / alpha 0.1
    assert np.allclose(env.control_inputs.velocity[0].cpu().numpy(), G["C_cmd_velocity"], atol=1e-15)
    assert np.allclose(env.control_inputs.global_velocity[0].cpu().numpy(), G["C_global_velocity"], atol=1e-15)
    ref = F.WalkingRewardRef(velocity=env.control_inputs.velocity[0].cpu().numpy(), heading=env.control_inputs.heading[0].cpu().numpy(),
                             global_velocity=env.control_inputs.global_velocity[0].cpu().numpy())
    r64 = torch.zeros(n, dtype=torch.float64, device="cuda")
    t64 = torch.zeros((n, 11), dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    for t in range(len(G["C_obs"])):
        obs32 = G["C_obs"][t].astype(np.float32)
        obs = torch.from_numpy(np.tile(obs32, (n, 1))).cuda().contiguous()
        ctrl = torch.from_numpy(np.tile(G["C_ctrl"][t].astype(np.float32), (n, 1))).cuda().contiguous()
        _lib.check(_lib.lib().qg_walk_step(env._batch, p(obs), p(ctrl), None, None, None, None, p(r64), p(t64), 0, None))
        total, vals = ref.step(obs32.astype(np.float64), G["C_ctrl"][t])
        got = t64[0].cpu().numpy()
        assert got[0] == vals[0] == 10.0
        assert np.allclose(got[:10], vals[:10], rtol=1e-12, atol=1e-13), (t, got - vals)
        # the 11th term is a finite difference (r - r_prev)/0.008 of values ~1: last-bit differences are amplified
        assert abs(got[10] - vals[10]) <= 1e-11 * max(1.0, abs(vals[10]))
        assert r64[0].item() == pytest.approx(total, rel=1e-12, abs=1e-11)
        assert abs(r64[0].item() - G["C_reward"][t]) < 1e-4 * max(1.0, abs(G["C_reward"][t]))   # vs the float64-obs reference
        assert torch.equal(t64[0], t64[n - 1])
    assert np.allclose(env.ctrl_f_est[0].cpu().numpy(), ref.f_est, rtol=1e-13, atol=1e-15)
    assert np.allclose(env.ctrl_a_est[0].cpu().numpy(), ref.a_est, rtol=1e-13, atol=1e-15)
    assert np.allclose(env.ideal_position[0].cpu().numpy(), ref.ideal_position, rtol=1e-14)
    env.close()


def test_walking_env_end_to_end(Vec):
    """VecWalkingQuadrupedEnv rollout: reward / info terms equal the restatement applied to the device's own
    sensordata and ctrl; flip / time-limit termination; reset bookkeeping (ideal position, derivative memory,
    command resampling with random_controls); settling mask."""
    from quadruped_gym_b200.envs.walking_quad import VecWalkingQuadrupedEnv
    n = 6
    env = VecWalkingQuadrupedEnv(n, "cuda:0", auto_reset=True, max_time=0.4, settling_time=0.05, random_controls=True,
                                 reset_options={"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3})
    obs, info = env.reset()
    assert info == {} and torch.count_nonzero(obs) == 0
    v0 = env.control_inputs.velocity.cpu().numpy()
    assert np.allclose(v0, np.tile([0.3, 0.0, 0.0], (n, 1)))                 # fixed_* options (train_quadruped.py:40-46)
    refs = [F.WalkingRewardRef(velocity=v0[e], heading=env.control_inputs.heading[e].cpu().numpy(),
                               global_velocity=env.control_inputs.global_velocity[e].cpu().numpy()) for e in range(n)]
    rng = np.random.default_rng(4)
    saw_term = False
    for t in range(70):
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        time_before = env.data.time.cpu().numpy()
        obs, rew, term, trunc, info = env.step(torch.from_numpy(a).cuda())
        sens = env.data.sensordata.cpu().numpy().astype(np.float64)         # last forward pass, also for reset envs
        for e in range(n):
            applied = np.where(time_before[e] < 0.05, np.array([0, 0, -0.5] * 4, np.float32), a[e])   # settling mask
            total, vals = refs[e].step(sens[e], applied.astype(np.float64))
            got = np.array([info[k][e].item() for k in env.reward_keys])
            assert np.allclose(got, vals.astype(np.float32), rtol=2e-6, atol=1e-6), (t, e)
            assert rew[e].item() == pytest.approx(np.float32(total), rel=2e-6, abs=1e-5)
            want_term = (time_before[e] + 4 * 0.002 >= 0.4 - 1e-12) or sens[e][29] < 0
            assert bool(term[e]) == bool(want_term)
            if term[e]:
                saw_term = True
                refs[e].reset()
                assert torch.count_nonzero(obs[e]) == 0 and float(env.data.time[e]) == 0.0
                assert np.array_equal(env.ideal_position[e].cpu().numpy(), np.zeros(3))
        assert set(env.reward_keys) <= set(info)
    assert saw_term
    env.close()


def test_po_walking_observation_and_sb3_adapter(Vec):
    """VecPOWalkingQuadrupedEnv: 26-value frames (gyro, accel, Madgwick Euler, body_vel xy, ctrl, command) stacked
    oldest-first over obs_window; the filter follows the restatement of ahrs' updateIMU (unpinned); the SB3 adapter
    returns numpy, auto-resets in the same step and carries the 11 reward keys in every info dict."""
    from quadruped_gym_b200.envs.po_walking_quad import SB3VecEnvAdapter, VecPOWalkingQuadrupedEnv
    n, W = 5, 4
    env = VecPOWalkingQuadrupedEnv(n, "cuda:0", obs_window=W, max_time=0.25, frame_skip=10, random_controls=True,
                                   reset_options={"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3})
    assert env.observation_space.shape == (26 * W,)
    obs, info = env.reset()
    assert obs.shape == (n, 26 * W) and info == {}
    f0 = obs[0].cpu().numpy().reshape(W, 26)
    assert np.array_equal(f0[0], f0[-1])                                           # [obs] * obs_window (po_walking_quad.py:65)
    assert np.allclose(f0[0, :9], 0) and np.allclose(f0[0, 11:23], [0, 0, -0.5] * 4)
    mw = F.MadgwickRef(Dt=0.002 * 10)
    q = [None] * n
    rng = np.random.default_rng(2)
    prev = obs.cpu().numpy().copy()
    done_seen = False
    for t in range(20):
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        quat_before_reset = None
        obs, rew, term, trunc, info = env.step(torch.from_numpy(a).cuda())
        sens = env.data.sensordata.cpu().numpy().astype(np.float64)
        o = obs.cpu().numpy()
        tobs = info["terminal_observation"].cpu().numpy()
        for e in range(n):
            stack = (tobs[e] if term[e] else o[e]).reshape(W, 26)
            assert np.array_equal(stack[:-1], prev[e].reshape(W, 26)[1:])              # FIFO shift
            fr = stack[-1]
            assert np.array_equal(fr[0:3], sens[e, 15:18].astype(np.float32)) and np.array_equal(fr[3:6], sens[e, 12:15].astype(np.float32))
            assert np.array_equal(fr[9:11], sens[e, 30:32].astype(np.float32)) and np.array_equal(fr[11:23], a[e])
            assert np.allclose(fr[23:26], [0.3, 0.0, 0.0], atol=1e-7)
            # filter: first update after a reset starts from the true base orientation (the qpos view quirk)
            if q[e] is None:
                ang = fr[6:9]
                assert np.all(np.abs(ang) < 0.5)
                q[e] = "running"
            if term[e]:
                done_seen = True
                q[e] = None
                rs = o[e].reshape(W, 26)
                assert np.array_equal(rs[0], rs[-1]) and np.allclose(rs[0, :6], 0) and np.allclose(rs[0, 11:23], [0, 0, -0.5] * 4)
                assert np.allclose(rs[0, 6:9], fr[6:9], atol=1e-6)                     # stale filter state in the reset frame
        prev = o.copy()
    assert done_seen
    env.close()

    # Madgwick restatement vs the device on a hand-made IMU sequence (one env, obs_window 1, settling 0)
    env = VecPOWalkingQuadrupedEnv(1, "cuda:0", obs_window=1, frame_skip=4, auto_reset=False)
    env.reset()
    ref = F.MadgwickRef(Dt=0.002 * 4)
    qref = None
    for t in range(40):
        obs, *_ = env.step(torch.zeros((1, 12), device="cuda"))
        sens = env.data.sensordata[0].cpu().numpy().astype(np.float64)
        if qref is None:
            qref = env.data.qpos[0, 3:7].cpu().numpy().astype(np.float64)            # view of qpos at the first update
        qref = ref.update(qref, sens[15:18], sens[12:15])
        assert np.allclose(obs[0, 6:9].cpu().numpy(), ref.to_angles(qref), atol=2e-6)
    env.close()

    sb3 = SB3VecEnvAdapter(VecPOWalkingQuadrupedEnv(8, "cuda:0", obs_window=10, max_time=0.1, frame_skip=10, random_controls=True,
                                                    reset_options={"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3}))
    o = sb3.reset()
    assert isinstance(o, np.ndarray) and o.shape == (8, 260) and o.dtype == np.float32   # train_quadruped.py:16-22 -> 26*10
    dones = 0
    for t in range(8):
        o, r, d, infos = sb3.step(rng.uniform(-1, 1, (8, 12)))
        assert o.shape == (8, 260) and r.shape == (8,) and d.dtype == bool and len(infos) == 8
        for i, inf in enumerate(infos):
            assert set(sb3.reward_keys) <= set(inf) and inf["TimeLimit.truncated"] is False
            assert ("terminal_observation" in inf) == bool(d[i])
            if d[i]:
                dones += 1
                assert inf["terminal_observation"].shape == (260,)
    assert dones == 8            # max_time 0.1 at frame_skip 10 -> every env terminates at step 5
    assert sb3.get_attr("frame_skip") == [10] * 8 and sb3.env_is_wrapped(object) == [False] * 8
    sb3.close()


@pytest.mark.parametrize("impratio", [1.0, 4.0])
def test_elliptic_cone_parity(Vec, blob, impratio):
    """Elliptic friction cones (BASELINE config 5): teacher-forced single-step parity against the oracle on states
    from elliptic-cone rollouts (sticking, sliding and separating contacts), same tolerances as the pyramidal test."""
    from oracle.oracle import OracleModel
    from quadruped_gym_b200.model import blob as qblob
    A = qblob.unpack(blob)
    A["opt_i"][1] = 1
    A["opt_f"][6] = impratio
    eb = qblob.pack(A)
    om = OracleModel(eb)
    n = 256
    st = rollout_states(om, n, 150, seed=17)
    st32 = {k: v.astype(np.float32) for k, v in st.items() if k != "time"}
    env = Vec(n, "cuda:0", auto_reset=False, model_blob=eb)
    env.set_state(qpos=st32["qpos"], qvel=st32["qvel"], act=st32["act"], qacc_warmstart=st32["warm"], time=st["time"], ctrl=st32["ctrl"])
    ctrl = np.random.default_rng(5).uniform(-1, 1, (n, 12)).astype(np.float32)
    out = {k: v.cpu().numpy() for k, v in env.debug_step(ctrl).items()}
    gq, gv = env.data.qpos.cpu().numpy(), env.data.qvel.cpu().numpy()
    same = cone_zone = 0
    for e in range(n):
        d = _oracle_at(om, st32, st["time"], ctrl, e)
        d.forward()
        if out["counts"][e, 0] != d.ncon:
            continue
        same += 1
        assert out["counts"][e, 1] == d.nefc == 3 * d.ncon + d.nlimit
        f = d.efc_force[d.nlimit:].reshape(-1, 3)
        cone_zone += int(np.any((f[:, 0] > 0) & (np.hypot(f[:, 1], f[:, 2]) > 0.999 * f[:, 0])))   # sliding contacts
        tol = 1e-4 if d.nefc == 0 else 2e-3
        assert np.abs(out["qacc"][e] - d.qacc).max() <= tol * max(1.0, np.abs(d.qacc).max()), (e, d.ncon, d.solver_niter)
        d2 = _oracle_at(om, st32, st["time"], ctrl, e)
        d2.step()
        assert np.abs(gq[e] - d2.qpos).max() <= 1e-5
        assert np.abs(gv[e] - d2.qvel).max() <= 2e-4 * max(1.0, np.abs(d2.qvel).max())
    assert same >= 0.97 * n and cone_zone >= 10
    # stability over a rollout: no divergence, forces keep the robot up
    env.reset()
    a = torch.zeros((n, 12), device="cuda"); a[:, 2::3] = -0.5
    for _ in range(300):
        obs, *_ = env.step(a)
    assert env.counters()["diverged"] == 0
    assert torch.allclose(obs[:, 20], torch.full((n,), 0.143, device="cuda"), atol=3e-3)   # standing height
    env.close()


def test_env_binning_does_not_change_results(Vec, monkeypatch):
    """The slot -> environment permutation (environments grouped by last-step solver effort) is a pure scheduling
    device: observations, rewards and flags are bit-identical with and without it."""
    n = 512
    rng = np.random.default_rng(31)
    acts = rng.uniform(-1, 1, (40, n, 12)).astype(np.float32)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("QG_BINNING", flag)
        env = Vec(n, "cuda:0", auto_reset=True, max_time=0.2)
        env.reset()
        rec = []
        for t in range(40):
            o, r, te, _, info = env.step(torch.from_numpy(acts[t]).cuda())
            rec.append((o.clone(), r.clone(), te.clone(), info["terminal_observation"].clone()))
        outs.append(rec)
        env.close()
    for a, b in zip(*outs):
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_step_replays_from_a_cuda_graph(Vec, monkeypatch):
    """qg_step keeps no host-side state per launch (the persistent kernel's chunk counter is re-armed on the device by the
    last block to leave), so a captured launch can be replayed: graph replays and eager calls give identical bits.
    Also covers more chunks than SMs (4,800 envs = 75 chunks of 64 on 148 SMs is not enough: use 12,800 = 200)."""
    monkeypatch.setenv("QG_BINNING", "1")
    n = 12800
    rng = np.random.default_rng(5)
    acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 12)).astype(np.float32)).cuda() for _ in range(6)]
    envs = [Vec(n, "cuda:0", auto_reset=True, max_time=0.5) for _ in range(2)]
    for e in envs:
        e.reset()
        for t in range(30):   # land first; the binning permutation exists after the first step
            e.step(acts[t % 6])
    a_static = acts[0].clone()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        envs[0].step(a_static)
    for t in range(12):
        a_static.copy_(acts[t % 6])
        g.replay()
        o1, r1, te1 = envs[0]._obs.clone(), envs[0]._reward.clone(), envs[0]._terminated.clone()
        o2, r2, te2, _, _ = envs[1].step(acts[t % 6])
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(te1.bool(), te2)
    assert torch.equal(envs[0].data.qpos, envs[1].data.qpos)
    for e in envs:
        e.close()


def test_full_size_properties_and_rollout_buffer(Vec):
    """BASELINE sizes (65,536 envs): size-independent properties -- identical environments stay bit-identical under
    identical actions, every value stays finite, counters add up, and the [T,N,.] rollout buffer of config 4 fills on
    the device."""
    from quadruped_gym_b200.envs import rewards as R
    from quadruped_gym_b200.rollout import RolloutBuffer
    n = 65536
    env = Vec(n, "cuda:0", auto_reset=True, termination_fns={"flip": R.flip_termination()},
              reward_fns={"forward": R.forward_velocity(1.0), "control_cost": R.ctrl_sq(-0.1), "alive_bonus": R.alive_bonus(1.0)})
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for t in range(40):      # all environments receive the same action -> they must remain identical
        a = (torch.rand((1, 12), device="cuda", generator=g) * 2 - 1).expand(n, 12).contiguous()
        obs, rew, term, _, _ = env.step(a)
    assert bool(torch.isfinite(obs).all()) and bool((obs == obs[0]).all()) and bool((rew == rew[0]).all())
    assert torch.equal(env.data.qpos, env.data.qpos[:1].expand(n, 19))
    c = env.counters(reset=True)
    assert c["physics_steps"] == n * 40 * 4 and c["contacts"] % n == 0 and c["diverged"] == 0 and c["contact_overflow"] == 0
    env.close()
    env = Vec(8192, "cuda:0", frame_skip=10, max_time=20.0, auto_reset=True, termination_fns={"flip": R.flip_termination()},
              reward_fns={"forward": R.forward_velocity(1.0), "control_cost": R.ctrl_sq(-0.1), "alive_bonus": R.alive_bonus(1.0)})
    buf = RolloutBuffer(env, 24).collect(generator=g)
    assert buf.obs.shape == (25, 8192, 33) and buf.actions.shape == (24, 8192, 12)
    assert torch.count_nonzero(buf.obs[0]) == 0 and torch.count_nonzero(buf.obs[1]) > 0
    ref = 1.0 - 0.1 * (buf.actions.double() ** 2).sum(-1)          # alive + control cost part of the fused reward
    assert torch.allclose((buf.rewards.double() - ref)[buf.dones.logical_not()].abs().max(), torch.tensor(0.0, dtype=torch.float64, device="cuda"), atol=5.0)
    assert bool(torch.isfinite(buf.rewards).all()) and buf.stats()["env_steps"] == 24 * 8192
    env.close()


@pytest.mark.parametrize("n", [1, 3, 63, 65, 1000, 148 * 64 + 1])   # the last: one chunk more than persistent blocks
def test_ragged_batch_sizes_and_state_roundtrip(Vec, oracle_model, n):
    """Batch sizes that do not fill a quad-of-quads / warp / block: tail quads must not disturb their neighbours,
    get_state(set_state(x)) == x, masked reset touches only the masked environments."""
    env = Vec(n, "cuda:0", auto_reset=False)
    env.reset()
    rng = np.random.default_rng(n)
    qpos = env.data.qpos.cpu().numpy()
    qpos[:, 7:] += rng.uniform(-0.2, 0.2, (n, 12)).astype(np.float32)
    qvel = rng.normal(size=(n, 18)).astype(np.float32)
    act = rng.uniform(-0.3, 0.3, (n, 12)).astype(np.float32)
    tm = rng.uniform(0, 5, n)
    env.set_state(qpos=qpos, qvel=qvel, act=act, time=tm)
    assert np.array_equal(env.data.qpos.cpu().numpy(), qpos) and np.array_equal(env.data.qvel.cpu().numpy(), qvel)
    assert np.array_equal(env.data.act.cpu().numpy(), act) and np.array_equal(env.data.time.cpu().numpy(), tm)
    a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
    obs, rew, term, trunc, _ = env.step(torch.from_numpy(a).cuda())
    assert obs.shape == (n, 33) and bool(torch.isfinite(obs).all()) and not bool(trunc.any())
    # last environment against the oracle (it sits next to the shadow quads of the tail)
    d = OracleData(oracle_model)
    d.set_state(qpos[-1].astype(np.float64), qvel[-1].astype(np.float64), act[-1].astype(np.float64), np.zeros(18), tm[-1], np.zeros(12))
    d.env_step(a[-1].astype(np.float64), 4)
    err = np.abs(obs[-1].cpu().numpy() - d.sensordata)
    err[12:15] /= max(1.0, np.abs(d.qacc).max())
    assert err.max() < 1e-4
    mask = torch.zeros(n, dtype=torch.bool)
    mask[::2] = True
    before = env.data.qpos.clone()
    env.reset(mask=mask.cuda())
    after = env.data.qpos
    assert torch.equal(after[~mask.cuda()], before[~mask.cuda()])
    assert np.allclose(after[mask.cuda()].cpu().numpy()[:, :7], [0, 0, 0.13, 1, 0, 0, 0])
    assert float(env.data.time[0]) == 0.0
    env.close()


def test_argument_errors(Vec):
    env = Vec(4, "cuda:0")
    env.frame_skip = 0
    with pytest.raises(ValueError):
        env.step(torch.zeros((4, 12), device="cuda"))          # QG_EINVAL: frame_skip must be >= 1
    env.frame_skip = 4
    with pytest.raises(RuntimeError):
        env.step(torch.zeros((5, 12), device="cuda"))          # wrong batch size
    with pytest.raises(ValueError):
        Vec(2, "cuda:0", render_mode="ascii")
    for mode in ("human", "rgb_array"):
        e2 = Vec(2, "cuda:0", render_mode=mode)                # construction never needs the renderer
        e2.reset()
        for _ in range(6):                                     # 48 ms of simulated time: a frame is due at 30 fps
            e2.step(torch.zeros((2, 12), device="cuda"))
        try:
            import mujoco  # noqa: F401
        except ImportError:
            with pytest.raises(NotImplementedError):
                e2.render()                                    # the bridge needs the mujoco wheel: fails loudly without it
        e2.close()
    env.close()


def test_plane_mesh_neighbour_rule_switch(Vec, blob):
    """The recalled-but-unverified rule for extra plane-mesh contacts is data (blob opt_i[4]): with "far from the first
    contact only" the device and the oracle still agree, on states where many geoms touch the floor (robot lying down)."""
    from oracle.oracle import OracleModel
    from quadruped_gym_b200.model import blob as qblob
    A = qblob.unpack(blob)
    A["opt_i"][4] = 1
    rb = qblob.pack(A)
    om = OracleModel(rb)
    n = 96
    rng = np.random.default_rng(23)
    qpos = np.tile(np.r_[0, 0, 0.03, 1, 0, 0, 0, np.tile(np.deg2rad([-45, 37.5, 0]), 4)], (n, 1))
    qpos[:, 2] = rng.uniform(0.01, 0.06, n)                     # base close to / into the floor
    ang = rng.uniform(-0.3, 0.3, (n, 3))
    qpos[:, 3:7] = np.c_[np.ones(n), 0.5 * ang]
    qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
    qpos[:, 7:] += rng.uniform(-0.4, 0.8, (n, 12))
    qpos = qpos.astype(np.float32)
    env = Vec(n, "cuda:0", auto_reset=False, model_blob=rb)
    env.set_state(qpos=qpos, qvel=np.zeros((n, 18), np.float32), act=np.zeros((n, 12), np.float32),
                  qacc_warmstart=np.zeros((n, 18), np.float32), time=np.zeros(n), ctrl=np.zeros((n, 12), np.float32))
    out = {k: v.cpu().numpy() for k, v in env.debug_step(np.zeros((n, 12), np.float32)).items()}
    same = multi = 0
    for e in range(n):
        d = OracleData(om)
        d.set_state(qpos[e].astype(np.float64), np.zeros(18), np.zeros(12), np.zeros(18), 0.0, np.zeros(12))
        d.forward()
        multi += int(d.ncon >= 8)
        if out["counts"][e, 0] == d.ncon:
            same += 1
            assert np.abs(out["qacc"][e] - d.qacc).max() <= 2e-3 * max(1.0, np.abs(d.qacc).max()), (e, d.ncon)
    assert multi >= n // 3 and same >= 0.9 * n          # many-contact states; ties at fp32 flip a few contact sets
    assert env.counters()["contact_overflow"] == 0
    env.close()
