"""GPU parity tests on the BASELINE configurations themselves (BASELINE.json configs[1] and configs[4]), on the PRODUCTION
instantiation of the step kernel, and on the random streams (reset yaw, command sampler) across shards.

  * C2  -- 4,096 envs, random actions, frame_skip 4: EVERY env.step() is also stepped by the float64 oracle from the
           device's own pre-step state and compared (qpos 1e-4 abs on >= 99.99 % of the (env, step) pairs, the rest bounded).
  * C5  -- elliptic cone, position-servo gains x5, randomised poses: teacher-forced single steps incl. active joint limits.
  * production kernel -- qg_step (frame_skip 1), not qg_debug_step, against the oracle at the stage-test tolerances.
  * random_init / random_controls with env_offset: two half-batches == one full batch bit for bit; yaw ~ U(0, 2 pi).
"""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleBatch, OracleData, OracleModel
from tests.conftest import rollout_states

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Vec():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from quadruped_gym_b200 import VecQuadrupedEnv
    return VecQuadrupedEnv


def _set_oracle_from_device(ob, env, n):
    st = {k: getattr(env.data, k).cpu().numpy() for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "ctrl")}
    for e in range(n):
        ob.env(e).set_state(st["qpos"][e].astype(np.float64), st["qvel"][e].astype(np.float64), st["act"][e].astype(np.float64),
                            st["qacc_warmstart"][e].astype(np.float64), float(st["time"][e]), st["ctrl"][e].astype(np.float64))
    return st


def test_config2_every_step_teacher_forced(Vec, oracle_model):
    """BASELINE configs[1] exactly: 4,096 envs, U(-1,1) actions, default frame_skip; every env.step() is compared with the
    oracle restarted from the device's own state (north star: per-step qpos / qvel / sensordata within fp32 tolerance,
    1e-4 relative contact-free, 1e-3 with contacts).  Pairs whose contact set flipped at an fp32 / fp64 tie are bounded,
    not skipped."""
    N, T, FS = 4096, 30, 4
    env = Vec(N, "cuda:0", frame_skip=FS, auto_reset=False)
    env.reset()
    ob = OracleBatch(oracle_model, N)
    rng = np.random.default_rng(0)
    within = total = 0
    p99_q, p99_v, worst_q, worst_v, worst_s = 0.0, 0.0, 0.0, 0.0, 0.0
    contact_steps = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 12)).astype(np.float32)
        _set_oracle_from_device(ob, env, N)
        obs, *_ = env.step(torch.from_numpy(a).cuda())
        oo = ob.rollout(a[None].astype(np.float64), FS, 1e9, False, want_obs=True)[0]
        qg, vg = env.data.qpos.cpu().numpy(), env.data.qvel.cpu().numpy()
        qo = np.array([ob.env(e).qpos.copy() for e in range(N)])
        vo = np.array([ob.env(e).qvel.copy() for e in range(N)])
        nco = np.array([ob.env(e).ncon for e in range(N)])
        contact_steps += int((nco > 0).sum())
        eq = np.abs(qg - qo).max(1)
        ev = np.abs(vg - vo).max(1) / np.maximum(1.0, np.abs(vo).max(1))
        es = np.abs(obs.cpu().numpy() - oo)
        es[:, 12:15] = 0          # accelerometer = qacc (|values| ~ 1e2): covered by the stage tests with a scaled tolerance
        es = es.max(1) / np.maximum(1.0, np.abs(vo).max(1))       # sensordata carries the velocities of the last forward pass
        good = (eq < 1e-4) & (ev < 1e-4)
        within += int((eq < 1e-4).sum()); total += N
        p99_q, p99_v = max(p99_q, float(np.percentile(eq, 99))), max(p99_v, float(np.percentile(ev, 99)))
        worst_q, worst_v, worst_s = max(worst_q, float(eq.max())), max(worst_v, float(ev.max())), max(worst_s, float(es[good].max()))
    print(f"C2: {within}/{total} pairs within 1e-4 on qpos; p99 qpos {p99_q:.2e} qvel(rel) {p99_v:.2e}; worst qpos {worst_q:.2e} "
          f"qvel(rel) {worst_v:.2e} sensordata(good pairs) {worst_s:.2e}; (env,step) pairs in contact {contact_steps}")
    assert contact_steps > 0.3 * total / 2          # the run covers landing and stance, not only the drop
    assert within >= 0.9999 * total
    assert p99_q <= 2e-6 and p99_v <= 2e-5
    assert worst_s <= 2e-4
    # contact-set flips (a hull vertex within fp32 round-off of the margin): one env.step() = 8 ms of a different contact force
    assert worst_q <= 2e-2 and worst_v <= 0.5
    assert env.counters()["diverged"] == 0 and env.counters()["contact_overflow"] == 0
    env.close()


def _c5_model(blob):
    from quadruped_gym_b200.model import blob as qblob
    A = qblob.unpack(blob)
    A["opt_i"][1] = 1                                            # elliptic cone
    A["act_gain"] = A["act_gain"] * 5.0                          # kp x5 (bench.py workload c5)
    A["act_bias"] = A["act_bias"].reshape(-1, 3) * np.array([1.0, 5.0, 1.0])
    return qblob.pack(A)


def _c5_poses(n, rng):
    """bench.py's C5 initial poses: yaw U(0,2pi), tilt <= 30 deg, height U(0.05,0.2), joints U(range)."""
    yaw, tilt, tdir = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, np.pi / 6, n), rng.uniform(0, 2 * np.pi, n)
    qt = np.c_[np.cos(tilt / 2), np.cos(tdir) * np.sin(tilt / 2), np.sin(tdir) * np.sin(tilt / 2), np.zeros(n)]
    qy = np.c_[np.cos(yaw / 2), np.zeros(n), np.zeros(n), np.sin(yaw / 2)]
    w1, x1, y1, z1 = qy.T; w2, x2, y2, z2 = qt.T
    quat = np.c_[w1*w2 - x1*x2 - y1*y2 - z1*z2, w1*x2 + x1*w2 + y1*z2 - z1*y2, w1*y2 - x1*z2 + y1*w2 + z1*x2, w1*z2 + x1*y2 - y1*x2 + z1*w2]
    lo, hi = np.tile(np.deg2rad([-45, -45, -90]), 4), np.tile(np.deg2rad([45, 120, 90]), 4)
    qpos = np.zeros((n, 19))
    qpos[:, 2] = rng.uniform(0.05, 0.2, n)
    qpos[:, 3:7] = quat
    qpos[:, 7:] = lo + (hi - lo) * rng.random((n, 12))
    return qpos


def test_config5_elliptic_high_gain_randomised_poses(Vec, blob):
    """BASELINE configs[4] as bench.py builds it (elliptic cone, kp x5, randomised poses): states a few env steps into
    oracle rollouts from those poses under random actions -- penetrating starts, sliding contacts, joints driven past
    their ranges by the stiff servos (limit rows active) -- one teacher-forced step on the device vs the oracle."""
    eb = _c5_model(blob)
    om = OracleModel(eb)
    n = 320
    rng = np.random.default_rng(55)
    qpos0 = _c5_poses(n, rng)
    st = {k: [] for k in ("qpos", "qvel", "act", "warm", "ctrl", "time")}
    for e in range(n):
        d = OracleData(om)
        d.set_state(qpos0[e], np.zeros(18), np.zeros(12), np.zeros(18), 0.0, np.array([0, 0, -0.5] * 4, float))
        bang = e % 2 == 0         # half of the envs get bang-bang actions: the stiff servos overshoot the joint ranges
        draw = lambda: np.sign(rng.uniform(-1, 1, 12)) if bang else rng.uniform(-1, 1, 12)
        a = draw()
        for s in range(int(rng.integers(0, 25))):
            if s % 5 == 0:
                a = draw()
            d.env_step(a, 4)
        for k, v in (("qpos", d.qpos), ("qvel", d.qvel), ("act", d.act), ("warm", d.qacc_warmstart), ("ctrl", d.ctrl)):
            st[k].append(v.copy())
        st["time"].append(d.time)
    st = {k: np.array(v) for k, v in st.items()}
    st32 = {k: v.astype(np.float32) for k, v in st.items() if k != "time"}
    env = Vec(n, "cuda:0", auto_reset=False, model_blob=eb)
    env.set_state(qpos=st32["qpos"], qvel=st32["qvel"], act=st32["act"], qacc_warmstart=st32["warm"], time=st["time"], ctrl=st32["ctrl"])
    ctrl = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
    out = {k: v.cpu().numpy() for k, v in env.debug_step(ctrl).items()}
    gq, gv = env.data.qpos.cpu().numpy(), env.data.qvel.cpu().numpy()
    same = with_limits = with_contacts = sliding = 0
    worst_flip = 0.0
    for e in range(n):
        d = OracleData(om)
        d.set_state(st32["qpos"][e].astype(np.float64), st32["qvel"][e].astype(np.float64), st32["act"][e].astype(np.float64),
                    st32["warm"][e].astype(np.float64), st["time"][e], ctrl[e].astype(np.float64))
        d.forward()
        scale = max(1.0, np.abs(d.qacc).max())
        err = np.abs(out["qacc"][e] - d.qacc).max() / scale
        if out["counts"][e, 0] != d.ncon or out["counts"][e, 1] != d.nefc:
            worst_flip = max(worst_flip, err)       # contact / limit set differs at a tie: bounded below, not skipped
            continue
        same += 1
        with_limits += int(d.nlimit > 0)
        with_contacts += int(d.ncon > 0)
        if d.ncon:
            f = d.efc_force[d.nlimit:].reshape(-1, 3)
            sliding += int(np.any((f[:, 0] > 0) & (np.hypot(f[:, 1], f[:, 2]) > 0.999 * f[:, 0])))
        assert out["counts"][e, 1] == 3 * d.ncon + d.nlimit
        # 5x servo gains: |qacc| reaches 1e3..1e4 and the fp32 solve's relative error on the unconstrained system is a
        # little above the stock model's (2e-4 instead of 1e-4 of max|qacc|)
        tol = 2e-4 if d.nefc == 0 else 2e-3
        assert err <= tol, (e, d.ncon, d.nlimit, err)
        d.step()
        assert np.abs(gq[e] - d.qpos).max() <= 2e-5
        assert np.abs(gv[e] - d.qvel).max() <= 4e-4 * max(1.0, np.abs(d.qvel).max())
    print(f"C5: same sets {same}/{n}, with limit rows {with_limits}, with contacts {with_contacts}, sliding {sliding}, worst flipped-set error {worst_flip:.2e}")
    assert same >= 0.95 * n and with_limits >= 10 and with_contacts >= 0.5 * n and sliding >= 10
    assert worst_flip <= 1.0            # a flipped set changes one contact's force, never the scale of the acceleration
    env.close()


def test_production_kernel_single_step_parity(Vec, oracle_model):
    """The PRODUCTION instantiation (qg_step -> qg_step_kernel<false, .>, frame_skip 1) against the oracle from the same
    states as the debug-instantiation stage tests, at test_next_state_parity's tolerances; sensordata included."""
    n = 384
    st = rollout_states(oracle_model, n, 150, seed=11)
    st32 = {k: v.astype(np.float32) for k, v in st.items() if k != "time"}
    env = Vec(n, "cuda:0", auto_reset=False, frame_skip=1)
    env.set_state(qpos=st32["qpos"], qvel=st32["qvel"], act=st32["act"], qacc_warmstart=st32["warm"], time=st["time"], ctrl=st32["ctrl"])
    a = np.random.default_rng(3).uniform(-1.2, 1.2, (n, 12)).astype(np.float32)
    obs, *_ = env.step(torch.from_numpy(a).cuda())
    o = obs.cpu().numpy()
    nxt = {k: getattr(env.data, k).cpu().numpy() for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "ctrl")}
    assert np.array_equal(nxt["ctrl"], np.clip(a, -1, 1))                       # env.step clips (quadruped.py:160)
    c = env.counters()
    ncon = nefc = flips = 0
    worst_flip_q = worst_flip_v = 0.0
    for e in range(n):
        d = OracleData(oracle_model)
        d.set_state(st32["qpos"][e].astype(np.float64), st32["qvel"][e].astype(np.float64), st32["act"][e].astype(np.float64),
                    st32["warm"][e].astype(np.float64), st["time"][e], np.clip(a[e], -1, 1).astype(np.float64))
        d.step()
        ncon += d.ncon; nefc += d.nefc
        eq, ev = np.abs(nxt["qpos"][e] - d.qpos).max(), np.abs(nxt["qvel"][e] - d.qvel).max() / max(1.0, np.abs(d.qvel).max())
        if eq > 1e-5 or ev > 1e-4:          # only a contact-set flip may exceed the stage tolerances: bounded, counted
            flips += 1
            worst_flip_q, worst_flip_v = max(worst_flip_q, eq), max(worst_flip_v, ev)
            continue
        assert np.abs(nxt["act"][e] - d.act).max() <= 1e-5
        assert np.abs(nxt["qacc_warmstart"][e] - d.qacc_warmstart).max() <= 1e-3 * max(1.0, np.abs(d.qacc_warmstart).max())
        assert nxt["time"][e] == d.time
        s = d.sensordata.copy()
        assert np.abs(o[e, 12:15] - s[12:15]).max() <= 1e-3 * max(1.0, np.abs(d.qacc).max())
        s[12:15] = 0; g = o[e].copy(); g[12:15] = 0
        assert np.abs(g - s).max() <= 1e-5
    print(f"production kernel: flips {flips}/{n}, worst flipped qpos {worst_flip_q:.2e} qvel(rel) {worst_flip_v:.2e}; "
          f"contacts device {c['contacts']} oracle {ncon}, rows device {c['efc_rows']} oracle {nefc}")
    assert flips <= 0.02 * n and worst_flip_q <= 1e-3 and worst_flip_v <= 0.2
    assert abs(c["contacts"] - ncon) <= 0.02 * ncon + 4 and abs(c["efc_rows"] - nefc) <= 0.02 * nefc + 16
    env.close()


def test_random_streams_are_shard_invariant(Vec):
    """random_init (reset yaw, walking_quad.py:68-75) and random_controls (command sampler, control_inputs.py:74-116) are
    counter-based streams keyed on (seed, GLOBAL env id, episode): two half-batches with env_offset reproduce one full
    batch bit for bit through auto-resets; another seed gives other draws; yaw ~ U(0, 2 pi)."""
    from quadruped_gym_b200.envs.walking_quad import VecWalkingQuadrupedEnv as W
    n, T = 256, 40
    kw = dict(auto_reset=True, max_time=0.1, random_init=True, random_controls=True, reset_options={"min_speed": 0.1, "max_speed": 0.6})
    full = W(n, "cuda:0", seed=7, **kw)
    halves = [W(n // 2, "cuda:0", seed=7, env_offset=i * n // 2, **kw) for i in range(2)]
    other = W(n, "cuda:0", seed=8, **kw)
    for e in (full, other, *halves):
        e.reset()
    cat = lambda f: torch.cat([f(h) for h in halves])
    assert torch.equal(full.data.qpos, cat(lambda h: h.data.qpos))
    assert torch.equal(full.control_inputs.velocity, cat(lambda h: h.control_inputs.velocity))
    assert torch.equal(full.control_inputs.heading, cat(lambda h: h.control_inputs.heading))
    assert not torch.equal(full.control_inputs.velocity, other.control_inputs.velocity)      # the seed reaches the sampler
    assert not torch.equal(full.data.qpos[:, 3:7], other.data.qpos[:, 3:7])                  # ... and the reset yaw
    assert not torch.equal(full.control_inputs.velocity[: n // 2], full.control_inputs.velocity[n // 2:])
    rng = np.random.default_rng(17)
    resets = 0
    yaws = []
    for t in range(T):
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        o, r, te, _, info = full.step(torch.from_numpy(a).cuda())
        parts = [h.step(torch.from_numpy(a[i * n // 2:(i + 1) * n // 2]).cuda()) for i, h in enumerate(halves)]
        assert torch.equal(o, torch.cat([p[0] for p in parts])) and torch.equal(r, torch.cat([p[1] for p in parts]))
        assert torch.equal(te, torch.cat([p[2] for p in parts]))
        if bool(te.any()):
            resets += int(te.sum())
            q = full.data.qpos
            assert torch.equal(q, cat(lambda h: h.data.qpos))                                # new yaw draws agree across shards
            assert torch.equal(full.control_inputs.global_velocity, cat(lambda h: h.control_inputs.global_velocity))
            qq = q[te].cpu().numpy()
            assert np.allclose(qq[:, 4:6], 0) and np.allclose(np.hypot(qq[:, 3], qq[:, 6]), 1, atol=1e-6)
            yaws.append(2 * np.arctan2(qq[:, 6], qq[:, 3]))
    assert resets >= 2 * n                       # max_time 0.1 s -> every env is reset every 13 steps
    y = np.concatenate(yaws) % (2 * np.pi)
    assert abs(y.mean() - np.pi) < 0.25 and abs(y.std() - 2 * np.pi / np.sqrt(12)) < 0.15     # U(0, 2 pi): mean pi, std 1.81
    assert np.histogram(y, bins=8, range=(0, 2 * np.pi))[0].min() > 0.5 * len(y) / 8
    sp = np.linalg.norm(full.control_inputs.velocity.cpu().numpy()[:, :2], axis=1)
    assert sp.min() >= 0.1 - 1e-12 and sp.max() <= 0.6 + 1e-12 and sp.std() > 0.05           # reset_options reach the sampler
    # reset(options=...) applies to THAT call only (walking_quad.py:100-103), later auto-resets use reset_options again
    full.reset(options={"fixed_speed": 0.25, "fixed_velocity_angle": 0.0, "fixed_heading_angle": 0.5})
    v = full.control_inputs.velocity.cpu().numpy()
    assert np.allclose(v[:, 0], 0.25) and np.allclose(v[:, 1], 0.0)
    assert np.allclose(full.control_inputs.get_heading_theta().cpu().numpy(), 0.5)
    for t in range(15):
        _, _, te, _, _ = full.step(torch.zeros((n, 12), device="cuda"))
    sp = np.linalg.norm(full.control_inputs.velocity.cpu().numpy()[:, :2], axis=1)
    assert sp.std() > 0.05 and sp.min() >= 0.1 - 1e-12
    for e in (full, other, *halves):
        e.close()


def test_python_reward_callable_sees_terminal_state(Vec):
    """A Python reward callable (README.md:65-78 pattern) is evaluated on the TERMINAL state of a terminating step, as the
    reference does (quadruped.py:170-178 run before any reset): the in-kernel reset is deferred until after it."""
    n = 8
    env = Vec(n, "cuda:0", auto_reset=True, max_time=0.05)
    seen = []
    env.reward_fns = {"t": lambda: env.data.time.float()}
    env.reset()
    for t in range(8):
        obs, rew, term, _, info = env.step(torch.zeros((n, 12), device="cuda"))
        seen.append((rew.clone(), term.clone()))
        if bool(term.any()):
            assert bool((rew[term] >= 0.05 - 1e-6).all())           # the callable saw time >= max_time, not the reset time 0
            assert float(env.data.time[0]) == 0.0                   # ... and the reset happened afterwards
            assert torch.count_nonzero(obs[term]) == 0
            assert torch.count_nonzero(info["terminal_observation"][term]) > 0
    assert any(bool(te.any()) for _, te in seen)
    env.close()


@pytest.mark.parametrize("segments,n", [("1", 700), ("3", 1000), ("4", 40000)])
def test_pipelined_host_path_equals_device_path(Vec, monkeypatch, segments, n):
    """qg_step_host cuts the batch into segments on their own streams (copies under kernels): obs, reward, flags, fused
    reward terms and terminal observations are bit-identical to qg_step's, for ragged segment sizes, with binning on,
    through auto-resets, and when the two paths alternate on one batch (the slot permutation changes layout)."""
    from quadruped_gym_b200.envs import rewards as R
    monkeypatch.setenv("QG_HOST_SEGMENTS", segments)
    monkeypatch.setenv("QG_BINNING", "1")
    fns = {"forward": R.forward_velocity(1.0), "control_cost": R.ctrl_sq(-0.1), "alive_bonus": R.alive_bonus(1.0)}
    mk = lambda: Vec(n, "cuda:0", auto_reset=True, max_time=0.12, reward_fns=dict(fns), termination_fns={"flip": R.flip_termination()})
    dev, host = mk(), mk()
    dev.reset(); host.reset()
    rng = np.random.default_rng(int(segments))
    T = 24
    resets = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        o, r, te, _, info = dev.step(torch.from_numpy(a).cuda())
        if t == 10:      # alternate: one device-path step on the host-path batch (and the same on the other)
            ho, hr, hte, _, hinfo = host.step(torch.from_numpy(a).cuda())
            assert torch.equal(ho, o) and torch.equal(hr, r) and torch.equal(hte, te)
            continue
        if t % 2:
            ho, hr, hte, _, hinfo = host.step_host(a, want_terms=True, want_terminal_obs=True, wait=False)
            host.host_wait()
        else:
            ho, hr, hte, _, hinfo = host.step_host(a, want_terms=True, want_terminal_obs=True)
        assert np.array_equal(ho, o.cpu().numpy()) and np.array_equal(hr, r.cpu().numpy()) and np.array_equal(hte, te.cpu().numpy())
        assert np.array_equal(hinfo["terminal_observation"], info["terminal_observation"].cpu().numpy())
        for k in fns:
            assert np.array_equal(hinfo["reward_components"][k], info["reward_components"][k].cpu().numpy())
        resets += int(hte.sum())
    assert resets >= n          # max_time 0.12 s: every env was auto-reset at least once
    assert torch.equal(dev.data.qpos, host.data.qpos) and torch.equal(dev.data.time, host.data.time)
    cd, ch = dev.counters(), host.counters()
    cd.pop("verts_tested"); ch.pop("verts_tested")      # a statistic that depends on which lane served a queue entry (shadow quads)
    assert cd == ch and cd["physics_steps"] == n * T * 4
    dev.close(); host.close()


def test_po_env_step_replays_from_a_cuda_graph(Vec):
    """The walking / PO launches keep no host state per step (the observation ring's head is advanced on the device), so a
    captured VecPOWalkingQuadrupedEnv.step replays: graph replays and eager calls give identical bits through auto-resets."""
    from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv as PO
    n = 300
    kw = dict(obs_window=4, frame_skip=10, max_time=0.12, random_controls=True, random_init=True, seed=3)
    a, b = PO(n, "cuda:0", **kw), PO(n, "cuda:0", **kw)
    a.reset(); b.reset()
    rng = np.random.default_rng(9)
    acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 12)).astype(np.float32)).cuda() for _ in range(5)]
    for t in range(3):                      # warm up both (lazy allocations happen outside the capture)
        a.step(acts[t]); b.step(acts[t])
    static = acts[0].clone()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        a.step(static)
    resets = 0
    for t in range(14):
        static.copy_(acts[t % 5])
        g.replay()
        o, r, te, _, info = b.step(acts[t % 5])
        assert torch.equal(a._stacked, o) and torch.equal(a._reward, r) and torch.equal(a._terminated.bool(), te)
        assert torch.equal(a._term_stacked[te], info["terminal_observation"][te])
        resets += int(te.sum())
    assert resets >= n
    assert torch.equal(a.data.qpos, b.data.qpos) and torch.equal(a.control_inputs.velocity, b.control_inputs.velocity)
    a.close(); b.close()
