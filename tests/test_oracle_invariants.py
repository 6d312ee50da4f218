"""Physics invariants of the CPU oracle (end) -- the checks available without MuJoCo.  CPU only."""
import numpy as np
import pytest

from oracle.oracle import OracleData, OracleModel
from quadruped_gym_b200.model import blob as qblob


def _variant(blob, **overrides):
    A = qblob.unpack(blob)
    for k, v in overrides.items():
        A[k] = np.asarray(v, dtype=A[k].dtype).reshape(A[k].shape)
    return OracleModel(qblob.pack(A))


def _random_state(om, rng, height=1.0):
    d = OracleData(om)
    q = rng.normal(size=4)
    d.qpos[:3] = [rng.normal(), rng.normal(), height]
    d.qpos[3:7] = q / np.linalg.norm(q)
    d.qpos[7:] = d.qpos[7:] + rng.uniform(-0.5, 0.5, 12)
    d.qvel[:] = rng.normal(size=18) * np.r_[np.ones(3), 3 * np.ones(3), 5 * np.ones(12)]
    return d


def test_mass_matrix_spd_and_matches_kinetic_energy(oracle_model):
    rng = np.random.default_rng(0)
    for _ in range(5):
        d = _random_state(oracle_model, rng)
        d.forward()
        M = d.M.copy()
        assert np.allclose(M, M.T, atol=1e-14) and np.linalg.eigvalsh(M).min() > 0
        # kinetic energy from body velocities == 1/2 v^T (M - armature) v
        m = oracle_model.m
        ke = 0.0
        cvel = np.ctypeslib.as_array(d.d.cvel).reshape(-1, 6)
        cin = np.ctypeslib.as_array(d.d.cinert).reshape(-1, 10)
        for b in range(1, m.nbody):
            I = cin[b]
            w, v = cvel[b, :3], cvel[b, 3:]
            Im = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]])
            mc = I[6:9]
            ke += 0.5 * (w @ Im @ w + I[9] * v @ v) + v @ np.cross(w, mc)
        v = d.qvel
        arm = np.array(m.dof_armature[:18])
        assert ke == pytest.approx(0.5 * v @ (M - np.diag(arm)) @ v, rel=1e-10)


def test_bias_matches_lagrangian_finite_difference(oracle_model, blob):
    """qfrc_bias = C(q,v) v + g(q): check the gravity part against dV/dq by finite differences."""
    om = oracle_model
    rng = np.random.default_rng(1)
    d = _random_state(om, rng)
    d.qvel[:] = 0
    d.forward()
    bias = d.qfrc_bias.copy()

    def potential(qpos):
        e = OracleData(om)
        e.qpos[:] = qpos
        e.forward()
        xipos = np.ctypeslib.as_array(e.d.xipos).reshape(-1, 3)
        mass = np.array(om.m.body_mass[:om.m.nbody])
        return 9.81 * np.sum(mass * xipos[:om.m.nbody, 2])

    q = d.qpos.copy()
    eps = 1e-6
    for j in range(12):  # hinge coordinates
        qp, qm = q.copy(), q.copy()
        qp[7 + j] += eps
        qm[7 + j] -= eps
        assert bias[6 + j] == pytest.approx((potential(qp) - potential(qm)) / (2 * eps), abs=1e-6)
    qp, qm = q.copy(), q.copy()
    qp[2] += eps
    qm[2] -= eps
    assert bias[2] == pytest.approx((potential(qp) - potential(qm)) / (2 * eps), abs=1e-6)


def test_free_flight_conserves_momentum_and_energy(blob):
    """No damping, no armature, no actuation, no contact: the continuous dynamics conserve linear momentum
    (up to gravity), angular momentum about the CoM and total energy; the first-order integrator must show
    errors that shrink linearly with the time step (a wrong M, Coriolis term or integrator would not)."""

    def run(h, T=0.4):
        om = _variant(blob, dof_damping=np.zeros(18), dof_armature=np.zeros(18), act_gain=np.zeros(12),
                      act_bias=np.zeros(36), jnt_limited=np.zeros(13),
                      opt_f=np.r_[h, qblob.unpack(blob)["opt_f"][1:]])
        m = om.m
        d = _random_state(om, np.random.default_rng(2), height=50.0)
        mass = np.array(m.body_mass[:m.nbody])

        def momenta():
            d.forward()
            cvel = np.ctypeslib.as_array(d.d.cvel).reshape(-1, 6).copy()
            cin = np.ctypeslib.as_array(d.d.cinert).reshape(-1, 10).copy()
            P, L, ke = np.zeros(3), np.zeros(3), 0.0
            for b in range(1, m.nbody):
                I = cin[b]
                w, v = cvel[b, :3], cvel[b, 3:]
                Im = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]])
                mc = I[6:9]
                P += I[9] * v + np.cross(w, mc)
                L += Im @ w + np.cross(mc, v)
                ke += 0.5 * (w @ Im @ w + I[9] * v @ v) + v @ np.cross(w, mc)
            xipos = np.ctypeslib.as_array(d.d.xipos).reshape(-1, 3)
            return P, L, ke + 9.81 * np.sum(mass * xipos[:m.nbody, 2])

        P0, L0, E0 = momenta()
        n = int(round(T / h))
        for _ in range(n):
            d.step()
        P1, L1, E1 = momenta()
        return (np.abs(P1 - P0 - [0, 0, -mass.sum() * 9.81 * n * h]).max(), np.abs(L1 - L0).max(), abs(E1 - E0))

    coarse, fine = run(0.002), run(0.0005)
    for c, f in zip(coarse, fine):
        assert 3.0 < c / f < 5.0      # first order: 4x smaller step -> ~4x smaller error
    assert coarse[0] < 0.05 and coarse[1] < 0.01 and coarse[2] < 0.1   # |E| ~ 545 J here


def test_static_stand_supports_weight(oracle_model):
    d = OracleData(oracle_model)
    d.ctrl[:] = [0, 0, -0.5] * 4
    for _ in range(1500):
        d.step()
    assert d.ncon == 4 and d.nefc == 16
    f = d.efc_force.reshape(-1, 4)
    total_normal = f.sum()
    assert total_normal == pytest.approx(1.110 * 9.81, rel=1e-6)
    assert np.all(f >= 0)                                  # pyramid generators push only
    assert np.allclose(d.sensordata[12:15], [0, 0, 9.81], atol=0.1)   # accelerometer at rest
    assert d.qpos[2] == pytest.approx(0.143, abs=2e-3)     # standing height
    # KKT: M qacc - qfrc_smooth - J^T f = 0
    r = d.M @ d.qacc - d.qfrc_smooth - d.efc_J.T @ d.efc_force
    assert np.abs(r).max() < 1e-8


def test_solver_optimality_on_contact_states(oracle_model):
    from tests.conftest import rollout_states
    st = rollout_states(oracle_model, 24, 80, seed=5)
    seen = 0
    for e in range(24):
        d = OracleData(oracle_model)
        d.set_state(st["qpos"][e], st["qvel"][e], st["act"][e], st["warm"][e], st["time"][e], st["ctrl"][e])
        d.forward()
        if d.nefc == 0:
            continue
        seen += 1
        jar = d.efc_J @ d.qacc - d.efc_aref
        f = np.where(jar < 0, -d.efc_D * jar, 0.0)
        grad = d.M @ d.qacc - d.qfrc_smooth - d.efc_J.T @ f
        assert np.abs(grad).max() < 1e-6 * max(1.0, np.abs(d.qfrc_smooth).max())
    assert seen >= 8


def test_joint_limit_rows(oracle_model):
    d = OracleData(oracle_model)
    d.qpos[2] = 1.0
    d.qpos[7] = -0.9       # hip 1 below its lower limit -0.785
    d.qpos[9] = 1.7        # ankle 1 above its upper limit 1.571
    d.act[0], d.act[2] = -1.5, 1.5   # servos pushing further out of range (force-clamped at 1.71)
    d.forward()
    assert d.nlimit == 2 and d.nefc == 2
    J = d.efc_J
    assert J[0, 6] == 1.0 and J[1, 8] == -1.0
    assert d.efc_pos[0] == pytest.approx(-0.9 + np.pi / 4) and d.efc_pos[1] == pytest.approx(np.pi / 2 - 1.7)
    assert np.all(d.efc_force > 0)      # pushes back into range
    assert d.qacc[6] > d.qacc_smooth[6] and d.qacc[8] < d.qacc_smooth[8]


def test_reset_state_and_time_limit_index(oracle_model):
    d = OracleData(oracle_model)
    assert np.all(d.sensordata == 0) and d.time == 0 and np.all(d.qvel == 0)   # mj_resetData
    t, n = 0.0, 0
    while t < 10.0:          # fp64 running sum of 0.002
        t += 0.002
        n += 1
    assert n == 5000
    t, n = 0.0, 0
    while t < 20.0:
        t += 0.002
        n += 1
    assert n == 10001        # -> env.step() #1001 at frame_skip 10


@pytest.mark.parametrize("impratio", [1.0, 4.0])
def test_elliptic_cone_stand_and_optimality(blob, impratio):
    """Elliptic friction cone (BASELINE config 5): static stand carries the weight with forces inside the cone,
    and the solver's point satisfies the first-order optimality of the three-zone cone cost."""
    A = qblob.unpack(blob)
    A["opt_i"][1] = 1
    A["opt_f"][6] = impratio
    om = OracleModel(qblob.pack(A))
    d = OracleData(om)
    d.ctrl[:] = [0, 0, -0.5] * 4
    for _ in range(1200):
        d.step()
    assert d.ncon == 4 and d.nefc == 12
    f = d.efc_force.reshape(-1, 3)
    assert f[:, 0].sum() == pytest.approx(1.110 * 9.81, rel=2e-3)
    assert np.all(np.hypot(f[:, 1], f[:, 2]) <= 1.0 * f[:, 0] + 1e-9)        # |f_t| <= mu f_n, mu = 1
    r = d.M @ d.qacc - d.qfrc_smooth - d.efc_J.T @ d.efc_force
    assert np.abs(r).max() < 1e-7
    rng = np.random.default_rng(3)
    for t in range(150):                                                       # sliding / stumbling states
        d.env_step(rng.uniform(-1, 1, 12), 4)
        if d.nefc:
            r = d.M @ d.qacc - d.qfrc_smooth - d.efc_J.T @ d.efc_force
            assert np.abs(r).max() < 1e-5 * max(1.0, np.abs(d.qfrc_smooth).max())
    assert d.warnings == 0
