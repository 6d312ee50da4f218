"""mujoco.MjModel -> model blob (second producer of the boundary object; needs the `mujoco` wheel).

Where MuJoCo is importable this removes every compiler-side uncertainty of ``mjcf.py`` (mesh inertia mode,
qhull vertex order, invweight0): all numbers are read from the compiled ``MjModel``
(/root/reference/src/envs/quadruped.py:59).  Not usable in the build container (no wheel, no network).
"""
from __future__ import annotations

import numpy as np

from .mjcf import CompiledModel, quat2mat


def export(model) -> CompiledModel:  # pragma: no cover - needs mujoco
    import mujoco
    m = model
    A = {}
    robot_geoms = [g for g in range(m.ngeom) if m.geom_type[g] == mujoco.mjtGeom.mjGEOM_MESH]
    planes = [g for g in range(m.ngeom) if m.geom_type[g] == mujoco.mjtGeom.mjGEOM_PLANE]
    if len(planes) != 1:
        raise ValueError("exactly one floor plane is supported")
    pl = planes[0]
    A["sizes"] = np.array([m.nq, m.nv, m.nu, m.nbody, m.njnt, len(robot_geoms), m.nmesh, m.nsensordata], np.int32)
    integ = {int(mujoco.mjtIntegrator.mjINT_EULER): 0, int(mujoco.mjtIntegrator.mjINT_IMPLICITFAST): 1}[int(m.opt.integrator)]
    A["opt_f"] = np.array([m.opt.timestep, *m.opt.gravity, m.opt.tolerance, m.opt.ls_tolerance, m.opt.impratio,
                           m.geom_pos[pl][2], m.stat.meaninertia])
    A["opt_i"] = np.array([integ, int(m.opt.cone == mujoco.mjtCone.mjCONE_ELLIPTIC), m.opt.iterations, m.opt.ls_iterations, 0], np.int32)
    A["body_parent"] = m.body_parentid.astype(np.int32)
    A["body_pos"], A["body_quat"], A["body_mass"] = m.body_pos.copy(), m.body_quat.copy(), m.body_mass.copy()
    A["body_ipos"] = m.body_ipos.copy()
    inertia = np.zeros((m.nbody, 6))
    for b in range(m.nbody):
        R = quat2mat(m.body_iquat[b])
        I = R @ np.diag(m.body_inertia[b]) @ R.T
        inertia[b] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
    A["body_inertia"] = inertia
    A["body_invweight0"] = m.body_invweight0.copy()
    jt = m.jnt_type.astype(np.int32)
    A["jnt_type"] = np.where(jt == int(mujoco.mjtJoint.mjJNT_FREE), 0, jt).astype(np.int32)
    A["jnt_body"], A["jnt_qposadr"], A["jnt_dofadr"] = (m.jnt_bodyid.astype(np.int32), m.jnt_qposadr.astype(np.int32),
                                                        m.jnt_dofadr.astype(np.int32))
    A["jnt_axis"], A["jnt_pos"], A["jnt_range"] = m.jnt_axis.copy(), m.jnt_pos.copy(), m.jnt_range.copy()
    A["jnt_limited"] = m.jnt_limited.astype(np.int32)
    hinge = [j for j in range(m.njnt) if jt[j] == int(mujoco.mjtJoint.mjJNT_HINGE)]
    A["jnt_solref"], A["jnt_solimp"] = m.jnt_solref[hinge[0]].copy(), m.jnt_solimp[hinge[0]].copy()
    A["qpos0"] = m.qpos0.copy()
    A["dof_damping"], A["dof_armature"], A["dof_invweight0"] = m.dof_damping.copy(), m.dof_armature.copy(), m.dof_invweight0.copy()
    A["dof_body"] = m.dof_bodyid.astype(np.int32)
    A["act_dof"] = np.array([m.jnt_dofadr[m.actuator_trnid[i, 0]] for i in range(m.nu)], np.int32)
    A["act_gear"] = m.actuator_gear[:, 0].copy()
    A["act_gain"] = m.actuator_gainprm[:, 0].copy()
    A["act_bias"] = m.actuator_biasprm[:, :3].copy()
    A["act_tau"] = m.actuator_dynprm[:, 0].copy()
    A["act_ctrlrange"], A["act_ctrllimited"] = m.actuator_ctrlrange.copy(), m.actuator_ctrllimited.astype(np.int32)
    A["act_frcrange"], A["act_frclimited"] = m.actuator_forcerange.copy(), m.actuator_forcelimited.astype(np.int32)
    g = np.array(robot_geoms)
    A["geom_body"], A["geom_pos"], A["geom_quat"] = m.geom_bodyid[g].astype(np.int32), m.geom_pos[g].copy(), m.geom_quat[g].copy()
    A["geom_mesh"], A["geom_rbound"] = m.geom_dataid[g].astype(np.int32), m.geom_rbound[g].copy()
    A["geom_margin"] = np.maximum(m.geom_margin[g], m.geom_margin[pl])
    A["geom_mu"] = np.maximum(m.geom_friction[g, 0], m.geom_friction[pl, 0])
    A["geom_solref"] = 0.5 * (m.geom_solref[g] + m.geom_solref[pl])
    A["geom_solimp"] = 0.5 * (m.geom_solimp[g] + m.geom_solimp[pl])
    # hull graphs: mesh_graph = [numvert, numface, vert_edgeadr[numvert], vert_globalid[numvert], edge_localid[...]]
    vadr, vnum, eadr, verts, vedge, edges = [], [], [], [], [], []
    for me in range(m.nmesh):
        gadr = m.mesh_graphadr[me]
        if gadr < 0:
            raise ValueError("mesh without convex-hull graph")
        gr = m.mesh_graph[gadr:]
        nvert, nface = int(gr[0]), int(gr[1])
        v_edgeadr, v_global = gr[2:2 + nvert], gr[2 + nvert:2 + 2 * nvert]
        e_local = gr[2 + 2 * nvert:2 + 2 * nvert + nvert + 3 * nface]
        V = m.mesh_vert[m.mesh_vertadr[me]:m.mesh_vertadr[me] + m.mesh_vertnum[me]]
        vadr.append(sum(vnum)); vnum.append(nvert); eadr.append(len(edges))
        verts.append(V[v_global])
        vedge.append(np.asarray(v_edgeadr, np.int32))
        edges.extend(int(x) for x in e_local)
    A["mesh_vertadr"], A["mesh_vertnum"], A["mesh_edgeadr"] = (np.array(vadr, np.int32), np.array(vnum, np.int32), np.array(eadr, np.int32))
    A["mesh_vert"] = np.concatenate(verts, 0)
    A["mesh_vert_edge"] = np.concatenate(vedge, 0).astype(np.int32)
    A["mesh_edge"] = np.array(edges, np.int32)
    return CompiledModel(arrays=A, source="mujoco.MjModel")
