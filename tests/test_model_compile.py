"""MJCF compiler facts and blob round trip.  CPU only."""
import os

import numpy as np
import pytest

from quadruped_gym_b200.model import DEFAULT_BLOB, blob as qblob, compile_mjcf
from tests.conftest import REF_SCENE

needs_ref = pytest.mark.skipif(not os.path.exists(REF_SCENE), reason="reference checkout not present")


@pytest.fixture(scope="module")
def cm():
    if not os.path.exists(REF_SCENE):
        pytest.skip("reference checkout not present")
    return compile_mjcf(REF_SCENE)


def test_sizes_and_options(cm):
    assert list(cm["sizes"]) == [19, 18, 12, 14, 13, 25, 5, 33]  # nq nv nu nbody njnt ngeom nmesh nsensordata
    assert cm["opt_f"][0] == 0.002 and tuple(cm["opt_f"][1:4]) == (0.0, 0.0, -9.81)
    assert cm["opt_i"][0] == 1 and cm["opt_i"][1] == 0  # implicitfast (quadruped.xml:4), pyramidal default


def test_masses_qpos0_ranges(cm):
    m = cm["body_mass"]
    assert m[1] == pytest.approx(0.018 + 4 * 0.056)          # FRAME + 4 hip servos (quadruped.xml:64-68)
    assert m[2] == pytest.approx(0.022 + 0.056) and m[3] == pytest.approx(0.013) and m[4] == pytest.approx(0.07 + 0.056)
    assert m.sum() == pytest.approx(1.110)
    q0 = cm["qpos0"]
    assert np.allclose(q0[:7], [0, 0, 0.13, 1, 0, 0, 0])
    assert np.allclose(q0[7:], np.tile(np.deg2rad([-45, 37.5, 0]), 4))
    rng = cm["jnt_range"][1:]
    assert np.allclose(rng[0], np.deg2rad([-45, 45])) and np.allclose(rng[1], np.deg2rad([-45, 120])) and np.allclose(rng[2], np.deg2rad([-90, 90]))
    assert list(cm["jnt_limited"]) == [0] + [1] * 12
    assert np.all(cm["dof_damping"] == 0.2) and np.all(cm["dof_armature"] == 0.001)  # default class reaches the free joint


def test_actuators(cm):
    assert np.all(cm["act_gear"] == 0.64) and np.all(cm["act_gain"] == 100)
    assert np.allclose(cm["act_bias"], np.tile([0, -100, -1], (12, 1)))
    assert np.allclose(cm["act_ctrlrange"][:3], [[-0.5, 0.5], [-0.91, 0.91], [-1, 1]])
    assert np.allclose(cm["act_frcrange"], np.tile([-1.71, 1.71], (12, 1))) and np.all(cm["act_tau"] == 0.01)
    assert list(cm["act_dof"]) == list(range(6, 18))


def test_contact_parameters_and_hulls(cm):
    assert np.all(cm["geom_mu"] == 1.0)        # max(0.6 robot, 1.0 floor)
    assert np.all(cm["geom_margin"] == 0.001)
    assert list(cm["mesh_vertnum"]) == [184, 70, 10, 118, 435]  # FRAME FEMA SHIN FOOT SERVO hull sizes
    # every neighbour list is -1 terminated and references valid local vertices
    for me in range(5):
        n, v0, e0 = cm["mesh_vertnum"][me], cm["mesh_vertadr"][me], cm["mesh_edgeadr"][me]
        for v in range(n):
            i = e0 + cm["mesh_vert_edge"][v0 + v]
            deg = 0
            while cm["mesh_edge"][i] >= 0:
                assert cm["mesh_edge"][i] < n
                i += 1
                deg += 1
            assert deg >= 3
    # at qpos0 the lowest hull vertex is ~9.7 cm above the floor
    from quadruped_gym_b200.model.mjcf import kinematics, quat2mat
    xpos, xmat = kinematics(cm.arrays, cm["qpos0"])
    zmin = 1e9
    for g in range(25):
        b, me = cm["geom_body"][g], cm["geom_mesh"][g]
        V = cm["mesh_vert"][cm["mesh_vertadr"][me]: cm["mesh_vertadr"][me] + cm["mesh_vertnum"][me]]
        R = xmat[b] @ quat2mat(cm["geom_quat"][g])
        zmin = min(zmin, (xpos[b] + xmat[b] @ cm["geom_pos"][g] + V @ R.T)[:, 2].min())
    assert 0.09 < zmin < 0.105


def test_invweight_and_meaninertia(cm):
    from quadruped_gym_b200.model.mjcf import mass_matrix
    M, _, _ = mass_matrix(cm.arrays, cm["qpos0"])
    assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0
    assert cm["opt_f"][8] == pytest.approx(np.trace(M) / 18)
    Minv = np.linalg.inv(M)
    assert cm["dof_invweight0"][6] == pytest.approx(Minv[6, 6])
    assert cm["dof_invweight0"][0] == pytest.approx(np.mean(np.diag(Minv)[:3]))


@needs_ref
def test_mesh_inertia_modes_change_only_shape():
    a, b = compile_mjcf(REF_SCENE, "legacy"), compile_mjcf(REF_SCENE, "exact")
    assert np.allclose(a["body_mass"], b["body_mass"])            # explicit geom masses
    assert not np.allclose(a["body_ipos"], b["body_ipos"])        # CoM moves with the mode


def test_blob_roundtrip_and_packaged_blob(cm):
    raw = cm.to_blob()
    back = qblob.unpack(raw)
    for k, v in cm.arrays.items():
        assert np.array_equal(np.asarray(v).ravel(), back[k]), k
    packaged = qblob.unpack(open(DEFAULT_BLOB, "rb").read())
    for k, v in cm.arrays.items():  # packaged blob == fresh compile of the reference MJCF
        assert np.allclose(np.asarray(v).ravel(), packaged[k], rtol=0, atol=1e-12), k


def test_missing_model_raises():
    with pytest.raises(FileNotFoundError):   # quadruped.py:55-56
        compile_mjcf("/nonexistent/scene.xml")
    with pytest.raises(ValueError):
        qblob.unpack(b"garbage-not-a-blob")
