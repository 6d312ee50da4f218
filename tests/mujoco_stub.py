"""A stand-in for the ``mujoco`` module, built from the in-repo MJCF compiler and backed by the CPU oracle, with the field
names and packed layouts of the real ``MjModel`` / ``MjData`` that this repository's MuJoCo-facing code reads
(``model/export_mujoco.py``, ``tests/golden/make_mujoco_golden.py``, ``tests/test_mujoco_gated.py``).

Purpose: keep those three runnable and exercised in an image where the wheel cannot be installed -- attribute names,
shapes, the ``mesh_graph`` packing and the control flow are smoke-tested (tests/test_mujoco_stub_smoke.py).  It proves
NOTHING about MuJoCo's numbers: the physics underneath is the oracle itself.  Field names are written from MuJoCo's
public documentation (mjmodel.h / mjdata.h) and are themselves unverified offline.  TEST INFRASTRUCTURE.
"""
import sys
import types

import numpy as np

from oracle.oracle import OracleData, OracleModel
from quadruped_gym_b200.model import SENSORS, compile_mjcf
from quadruped_gym_b200.model.mjcf import quat2mat


def _mat2quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    if w > 1e-6:
        return np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
    x = np.sqrt(max(0.0, 1 + R[0, 0] - R[1, 1] - R[2, 2])) / 2
    if x > 1e-6:
        return np.array([(R[2, 1] - R[1, 2]) / (4 * x), x, (R[0, 1] + R[1, 0]) / (4 * x), (R[0, 2] + R[2, 0]) / (4 * x)])
    y = np.sqrt(max(0.0, 1 - R[0, 0] + R[1, 1] - R[2, 2])) / 2
    if y > 1e-6:
        return np.array([(R[0, 2] - R[2, 0]) / (4 * y), (R[0, 1] + R[1, 0]) / (4 * y), y, (R[1, 2] + R[2, 1]) / (4 * y)])
    return np.array([0.0, 0.0, 0.0, 1.0])


class _NS(types.SimpleNamespace):
    pass


def build_module():
    mj = types.ModuleType("mujoco")
    mj.__version__ = "0.0.stub"
    mj.mjtGeom = _NS(mjGEOM_PLANE=0, mjGEOM_MESH=7)
    mj.mjtIntegrator = _NS(mjINT_EULER=0, mjINT_IMPLICITFAST=3)
    mj.mjtCone = _NS(mjCONE_PYRAMIDAL=0, mjCONE_ELLIPTIC=1)
    mj.mjtJoint = _NS(mjJNT_FREE=0, mjJNT_HINGE=3)

    class MjModel:
        def __init__(self, path):
            cm = compile_mjcf(path)
            A = cm.arrays
            self._cm = cm
            self._om = OracleModel(cm.to_blob())
            nq, nv, nu, nbody, njnt, ngeom, nmesh, nsens = [int(x) for x in A["sizes"]]
            self.nq, self.nv, self.nu, self.nbody, self.njnt, self.nmesh, self.nsensordata = nq, nv, nu, nbody, njnt, nmesh, nsens
            self.ngeom = ngeom + 1                                   # + the floor plane, last
            of, oi = A["opt_f"], A["opt_i"]
            self.opt = _NS(timestep=float(of[0]), gravity=np.array(of[1:4]), tolerance=float(of[4]), ls_tolerance=float(of[5]),
                           impratio=float(of[6]), integrator=3 if oi[0] == 1 else 0, cone=int(oi[1]), iterations=int(oi[2]),
                           ls_iterations=int(oi[3]))
            self.stat = _NS(meaninertia=float(of[8]))
            self.geom_type = np.r_[np.full(ngeom, 7), 0]
            self.geom_pos = np.vstack([A["geom_pos"].reshape(-1, 3), [0, 0, of[7]]])
            self.geom_quat = np.vstack([A["geom_quat"].reshape(-1, 4), [1, 0, 0, 0]])
            self.geom_bodyid = np.r_[A["geom_body"], 0]
            self.geom_dataid = np.r_[A["geom_mesh"], -1]
            self.geom_rbound = np.r_[A["geom_rbound"], 0.0]
            self.geom_margin = np.r_[A["geom_margin"], A["geom_margin"].max()]
            self.geom_friction = np.c_[np.r_[A["geom_mu"], A["geom_mu"].max()], np.full(ngeom + 1, 0.005), np.full(ngeom + 1, 1e-4)]
            self.geom_solref = np.vstack([A["geom_solref"].reshape(-1, 2), A["geom_solref"].reshape(-1, 2)[0]])
            self.geom_solimp = np.vstack([A["geom_solimp"].reshape(-1, 5), A["geom_solimp"].reshape(-1, 5)[0]])
            self.body_parentid = A["body_parent"].copy()
            self.body_pos, self.body_quat = A["body_pos"].reshape(-1, 3).copy(), A["body_quat"].reshape(-1, 4).copy()
            self.body_mass, self.body_ipos = A["body_mass"].copy(), A["body_ipos"].reshape(-1, 3).copy()
            self.body_invweight0 = A["body_invweight0"].reshape(-1, 2).copy()
            self.body_inertia, self.body_iquat = np.zeros((nbody, 3)), np.tile([1.0, 0, 0, 0], (nbody, 1))
            for b in range(nbody):                                    # principal moments + frame, as MjModel stores them
                xx, yy, zz, xy, xz, yz = A["body_inertia"].reshape(-1, 6)[b]
                I = np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])
                w, V = np.linalg.eigh(I)
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                self.body_inertia[b], self.body_iquat[b] = w, _mat2quat(V)
            jt = A["jnt_type"].copy()
            self.jnt_type = np.where(jt == 0, 0, 3)
            self.jnt_bodyid, self.jnt_qposadr, self.jnt_dofadr = A["jnt_body"].copy(), A["jnt_qposadr"].copy(), A["jnt_dofadr"].copy()
            self.jnt_axis, self.jnt_pos = A["jnt_axis"].reshape(-1, 3).copy(), A["jnt_pos"].reshape(-1, 3).copy()
            self.jnt_range, self.jnt_limited = A["jnt_range"].reshape(-1, 2).copy(), A["jnt_limited"].copy()
            self.jnt_solref = np.tile(A["jnt_solref"], (njnt, 1))
            self.jnt_solimp = np.tile(A["jnt_solimp"], (njnt, 1))
            self.qpos0 = A["qpos0"].copy()
            self.dof_damping, self.dof_armature, self.dof_invweight0 = A["dof_damping"].copy(), A["dof_armature"].copy(), A["dof_invweight0"].copy()
            self.dof_bodyid = A["dof_body"].copy()
            dof2jnt = {int(d): j for j, d in enumerate(self.jnt_dofadr)}
            self.actuator_trnid = np.array([[dof2jnt[int(d)], -1] for d in A["act_dof"]])
            self.actuator_gear = np.c_[A["act_gear"], np.zeros((nu, 5))]
            self.actuator_gainprm = np.c_[A["act_gain"], np.zeros((nu, 9))]
            self.actuator_biasprm = np.c_[A["act_bias"].reshape(-1, 3), np.zeros((nu, 7))]
            self.actuator_dynprm = np.c_[A["act_tau"], np.zeros((nu, 9))]
            self.actuator_ctrlrange, self.actuator_ctrllimited = A["act_ctrlrange"].reshape(-1, 2).copy(), A["act_ctrllimited"].copy()
            self.actuator_forcerange, self.actuator_forcelimited = A["act_frcrange"].reshape(-1, 2).copy(), A["act_frclimited"].copy()
            # meshes: here the mesh vertices ARE the hull vertices; mesh_graph = [nvert, nface, vert_edgeadr[nvert],
            # vert_globalid[nvert], edge_localid[nvert + 3 nface]] per mesh, neighbour lists -1 terminated
            V = A["mesh_vert"].reshape(-1, 3)
            self.mesh_vert, self.mesh_vertadr, self.mesh_vertnum = V.copy(), A["mesh_vertadr"].copy(), A["mesh_vertnum"].copy()
            graph, gadr = [], []
            for me in range(nmesh):
                v0, vn, e0 = int(A["mesh_vertadr"][me]), int(A["mesh_vertnum"][me]), int(A["mesh_edgeadr"][me])
                e1 = int(A["mesh_edgeadr"][me + 1]) if me + 1 < nmesh else len(A["mesh_edge"])
                edges = [int(x) for x in A["mesh_edge"][e0:e1]]
                nface = (len(edges) - vn) // 3
                assert vn + 3 * nface == len(edges), "stub: hull graph of a closed triangulated hull has nvert + 3 nface entries"
                gadr.append(len(graph))
                graph += [vn, nface] + [int(x) for x in A["mesh_vert_edge"][v0:v0 + vn]] + list(range(vn)) + edges
            self.mesh_graph, self.mesh_graphadr = np.array(graph, np.int32), np.array(gadr, np.int32)
            self.sensor_names = list(SENSORS.keys())
            self.sensor_adr = np.array([SENSORS[n][0] for n in self.sensor_names])

        @staticmethod
        def from_xml_path(path):
            return MjModel(path)

    class MjData:
        def __init__(self, model):
            self.o = OracleData(model._om)
            o = self.o
            self.qpos, self.qvel, self.act, self.ctrl, self.qacc_warmstart = o.qpos, o.qvel, o.act, o.ctrl, o.qacc_warmstart
            self.sensordata, self.qacc = o.sensordata, o.qacc
            self.solver_niter = np.zeros(1, int)

        time = property(lambda s: s.o.time, lambda s, v: setattr(s.o, "time", v))
        ncon = property(lambda s: s.o.ncon)
        nefc = property(lambda s: s.o.nefc)

    def mj_step(m, d):
        d.o.step()
        d.solver_niter[0] = d.o.solver_niter

    mj.MjModel, mj.MjData = MjModel, MjData
    mj.mj_resetData = lambda m, d: d.o.reset()
    mj.mj_step = mj_step
    mj.mj_forward = lambda m, d: d.o.forward()
    return mj


def install():
    mj = build_module()
    sys.modules["mujoco"] = mj
    return mj
