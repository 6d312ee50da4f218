"""MJCF + OBJ -> compiled model blob (own parser; no MuJoCo needed).

This replaces ``mujoco.MjModel.from_xml_path`` (/root/reference/src/envs/quadruped.py:59) for the
model class of the reference robot (/root/reference/src/models/quadruped/quadruped.xml:1-218,
scene.xml:1-22): one free-floating root body, a tree of hinge joints, mesh geoms colliding with one
horizontal floor plane, ``position`` actuators with first-order activation filter, and the 19-sensor
block of the MJCF.  Every constant the kernels and the oracle need is computed here in float64.

MuJoCo's compiler semantics are restated from its public XML reference (recalled; MuJoCo is not
installable in this environment - see DESIGN.md "parity unpinned"):

* defaults classes / ``childclass`` inheritance, ``compiler angle="degree"``, euler sequence xyz
  (intrinsic), ``ref`` -> ``qpos0``, ``range`` in qpos units;
* mesh inertia: selectable ``mesh_inertia`` in {"legacy", "exact", "convex"}; the mesh is re-centred
  at its CoM and re-oriented to its principal axes, geom pose is composed with that frame;
  geom ``mass`` is explicit so only shape (CoM, I/m) depends on the mode;
* collision uses the convex hull of each mesh (qhull via scipy) with a vertex adjacency graph;
* ``body_invweight0`` / ``dof_invweight0`` / ``meaninertia`` from M(qpos0) as in ``mj_setConst``.
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import blob as _blob

# ----------------------------------------------------------------------------- math helpers


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat2mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def mat2quat(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def euler2quat(e_rad, seq="xyz"):
    """Intrinsic rotations in the order of ``seq`` (lower case = rotating axes), MuJoCo ``eulerseq``."""
    q = np.array([1.0, 0, 0, 0])
    for ang, ax in zip(e_rad, seq):
        h = 0.5 * ang
        r = np.array([math.cos(h), 0.0, 0.0, 0.0])
        r["xyz".index(ax.lower()) + 1] = math.sin(h)
        q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
    return q


def _floats(s: Optional[str], n: Optional[int] = None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=float)
    v = np.array([float(t) for t in s.split()], dtype=float)
    if n is not None and v.size != n:
        if default is not None and v.size < n:  # MuJoCo pads with defaults (friction, solimp ...)
            d = np.array(default, dtype=float)
            d[: v.size] = v
            return d
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


# ----------------------------------------------------------------------------- meshes


@dataclass
class Mesh:
    name: str
    vert: np.ndarray  # original frame [n,3]
    face: np.ndarray  # [m,3]
    volume: float = 0.0
    pos: np.ndarray = None  # CoM in original frame
    quat: np.ndarray = None  # principal axes in original frame
    inertia_unit: np.ndarray = None  # principal moments per unit mass (I/m)
    hull_vert: np.ndarray = None  # hull vertices in the re-centred, re-oriented mesh frame [h,3]
    hull_edgeadr: np.ndarray = None  # [h] start of each vertex' neighbour list in hull_edge
    hull_edge: np.ndarray = None  # neighbour lists, each terminated by -1
    hull_cedgeadr: np.ndarray = None  # same layout, polytope edges only (no diagonals of coplanar facets):
    hull_cedge: np.ndarray = None     # the graph the kernels hill-climb on for the support search
    aabb_absmax: np.ndarray = None  # per-axis max |coord| in the mesh frame


def load_obj(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """OBJ reader: ``v x y z [r g b]`` lines, ``f a/b/c ...`` polygons (fan-triangulated)."""
    verts: List[List[float]] = []
    faces: List[List[int]] = []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                t = line.split()
                verts.append([float(t[1]), float(t[2]), float(t[3])])
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.array(verts, dtype=float), np.array(faces, dtype=np.int64)


def _tet_covariance(a, b, c, vol):
    """sum over tetrahedra (apex at origin) of  int x x^T dV,  ``vol`` per-tetra volumes [m]."""
    s = a + b + c
    C = np.einsum("m,mi,mj->ij", vol / 20.0, s, s)
    for v in (a, b, c):
        C += np.einsum("m,mi,mj->ij", vol / 20.0, v, v)
    return C


def _mesh_mass_props(vert, face, mode):
    """-> (volume, com, inertia about com for unit density) in the original mesh frame."""
    a, b, c = vert[face[:, 0]], vert[face[:, 1]], vert[face[:, 2]]
    nrm = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    cen = (a + b + c) / 3.0
    if mode == "legacy":
        # tetrahedra between each face and the area-weighted surface centroid, |volume| per face
        facecen = (area[:, None] * cen).sum(0) / area.sum()
        vol = np.abs(np.einsum("mi,mi->m", nrm, cen - facecen)) / 6.0
        volume = vol.sum()
        com = (vol[:, None] * (0.75 * cen + 0.25 * facecen)).sum(0) / volume
        a0, b0, c0 = a - com, b - com, c - com
        vol_i = np.abs(np.einsum("mi,mi->m", a0, np.cross(b0, c0))) / 6.0
        C = _tet_covariance(a0, b0, c0, vol_i)
    elif mode == "exact":
        vol = np.einsum("mi,mi->m", a, np.cross(b, c)) / 6.0
        volume = vol.sum()
        com = (vol[:, None] * (a + b + c) / 4.0).sum(0) / volume
        a0, b0, c0 = a - com, b - com, c - com
        vol_i = np.einsum("mi,mi->m", a0, np.cross(b0, c0)) / 6.0
        C = _tet_covariance(a0, b0, c0, vol_i)
        if volume < 0:
            volume, C = -volume, -C
    else:
        raise ValueError(mode)
    inertia = np.trace(C) * np.eye(3) - C
    return float(volume), com, inertia


def _hull(vert):
    from scipy.spatial import ConvexHull

    h = ConvexHull(vert, qhull_options="Qt")
    ids = np.array(sorted(set(h.simplices.ravel().tolist())), dtype=np.int64)
    local = {int(g): i for i, g in enumerate(ids)}
    nbr: List[List[int]] = [[] for _ in ids]
    # neighbour order = order of appearance while walking the hull facets (qhull facet order)
    for tri in h.simplices:
        for u in tri:
            lu = local[int(u)]
            for w in tri:
                lw = local[int(w)]
                if lw != lu and lw not in nbr[lu]:
                    nbr[lu].append(lw)
    # polytope edges only: drop the diagonals qhull's triangulation (Qt) puts inside merged coplanar facets.
    # A linear function over a convex polytope still has a strictly improving polytope-edge neighbour at every
    # non-optimal vertex, and the fan centres of flat faces lose their ~100 spokes.
    edge_faces = {}
    for fi, tri in enumerate(h.simplices):
        for a, b in ((0, 1), (1, 2), (0, 2)):
            key = (min(int(tri[a]), int(tri[b])), max(int(tri[a]), int(tri[b])))
            edge_faces.setdefault(key, []).append(fi)
    cnbr: List[List[int]] = [[] for _ in ids]
    for lu, lst in enumerate(nbr):
        gu = int(ids[lu])
        for lw in lst:
            gw = int(ids[lw])
            fs = edge_faces[(min(gu, gw), max(gu, gw))]
            diagonal = len(fs) == 2 and np.allclose(h.equations[fs[0]], h.equations[fs[1]], rtol=0, atol=1e-10)
            if not diagonal:
                cnbr[lu].append(lw)
    # orient hull faces outward for the "convex" inertia mode
    tris = h.simplices.copy()
    cen = vert[ids].mean(0)
    for k, tri in enumerate(tris):
        a, b, c = vert[tri]
        if np.dot(np.cross(b - a, c - a), a - cen) < 0:
            tris[k] = tri[[0, 2, 1]]
    return ids, nbr, tris, cnbr


def process_mesh(name: str, path: str, mode: str = "legacy") -> Mesh:
    vert, face = load_obj(path)
    m = Mesh(name=name, vert=vert, face=face)
    ids, nbr, hull_tris, cnbr = _hull(vert)
    if mode == "convex":
        volume, com, inertia = _mesh_mass_props(vert, hull_tris, "exact")
    else:
        volume, com, inertia = _mesh_mass_props(vert, face, mode)
    # principal axes, eigenvalues descending, right-handed
    w, V = np.linalg.eigh(inertia)
    order = np.argsort(-w)
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    m.volume, m.pos, m.quat = volume, com, mat2quat(V)
    R = quat2mat(m.quat)
    m.inertia_unit = w / volume
    local = (vert - com) @ R  # rows: R^T (v - com)
    m.hull_vert = local[ids]
    adr, edges = [], []
    for lst in nbr:
        adr.append(len(edges))
        edges.extend(lst)
        edges.append(-1)
    m.hull_edgeadr = np.array(adr, dtype=np.int32)
    m.hull_edge = np.array(edges, dtype=np.int32)
    adr, edges = [], []
    for lst in cnbr:
        adr.append(len(edges))
        edges.extend(lst)
        edges.append(-1)
    m.hull_cedgeadr = np.array(adr, dtype=np.int32)
    m.hull_cedge = np.array(edges, dtype=np.int32)
    m.aabb_absmax = np.abs(local).max(0)
    return m


# ----------------------------------------------------------------------------- MJCF parsing

_MAIN_DEFAULTS = {
    "geom": {"type": "sphere", "friction": "1 0.005 0.0001", "margin": "0", "gap": "0", "condim": "3",
             "contype": "1", "conaffinity": "1", "solref": "0.02 1", "solimp": "0.9 0.95 0.001 0.5 2",
             "solmix": "1", "priority": "0", "pos": "0 0 0"},
    "joint": {"type": "hinge", "axis": "0 0 1", "pos": "0 0 0", "damping": "0", "armature": "0",
              "ref": "0", "stiffness": "0", "margin": "0", "frictionloss": "0",
              "solreflimit": "0.02 1", "solimplimit": "0.9 0.95 0.001 0.5 2"},
    "position": {"kp": "1", "kv": "0", "gear": "1", "ctrllimited": "auto", "forcelimited": "auto"},
}


class _Defaults:
    def __init__(self):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {"main": {k: dict(v) for k, v in _MAIN_DEFAULTS.items()}}

    def add(self, node: ET.Element, parent: str):
        name = node.get("class", "main")
        cls = {k: dict(v) for k, v in self.classes[parent].items()} if name != parent else self.classes[parent]
        for child in node:
            if child.tag == "default":
                continue
            cls.setdefault(child.tag, {}).update(child.attrib)
        self.classes[name] = cls
        for child in node:
            if child.tag == "default":
                self.add(child, name)

    def resolve(self, elem: ET.Element, childclass: str) -> Dict[str, str]:
        cname = elem.get("class", childclass or "main")
        if cname not in self.classes:
            raise ValueError(f"unknown default class {cname!r}")
        out = dict(self.classes[cname].get(elem.tag, {}))
        out.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return out


def _load_xml(path: str) -> ET.Element:
    root = ET.parse(path).getroot()
    base = os.path.dirname(os.path.abspath(path))

    def expand(node):
        for i, ch in enumerate(list(node)):
            if ch.tag == "include":
                inc = _load_xml(os.path.join(base, ch.get("file")))
                node.remove(ch)
                for j, sub in enumerate(list(inc)):
                    node.insert(i + j, sub)
            else:
                expand(ch)

    expand(root)
    return root


@dataclass
class CompiledModel:
    """Numeric model (float64 / int32 arrays).  ``sections()`` is what goes into the blob."""
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    sensor_names: List[str] = field(default_factory=list)
    sensor_adr: List[int] = field(default_factory=list)
    sensor_dim: List[int] = field(default_factory=list)
    joint_names: List[str] = field(default_factory=list)
    body_names: List[str] = field(default_factory=list)
    source: str = ""

    def __getitem__(self, k):
        return self.arrays[k]

    def to_blob(self) -> bytes:
        return _blob.pack(self.arrays)

    @staticmethod
    def from_blob(buf: bytes) -> "CompiledModel":
        return CompiledModel(arrays=_blob.unpack(buf))


# sensor block the fused epilogue implements (quadruped.xml:174-217): (type, dim)
_SENSOR_LAYOUT = [("jointpos", 1)] * 12 + [("accelerometer", 3), ("gyro", 3), ("framepos", 3),
                                            ("framelinvel", 3), ("framexaxis", 3), ("framezaxis", 3),
                                            ("velocimeter", 3)]


def compile_mjcf(path: str, mesh_inertia: str = "legacy") -> CompiledModel:
    if not os.path.exists(path):
        raise FileNotFoundError(f"Model file not found: {path}")
    root = _load_xml(path)
    base = os.path.dirname(os.path.abspath(path))

    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    deg = comp.get("angle", "degree") == "degree"
    ang = (math.pi / 180.0) if deg else 1.0
    eulerseq = comp.get("eulerseq", "xyz")
    meshdir = os.path.join(base, comp.get("meshdir", ""))

    opt = {}
    for o in root.findall("option"):
        opt.update(o.attrib)
    integrator = opt.get("integrator", "Euler")
    cone = opt.get("cone", "pyramidal")

    dfl = _Defaults()
    for d in root.findall("default"):
        dfl.add(d, "main")

    meshes: Dict[str, Mesh] = {}
    for a in root.findall("asset"):
        for me in a.findall("mesh"):
            f = me.get("file")
            name = me.get("name", os.path.splitext(os.path.basename(f))[0])
            meshes[name] = process_mesh(name, os.path.join(meshdir, f), me.get("inertia", mesh_inertia))
    mesh_names = list(meshes.keys())

    def orient(attr: Dict[str, str]) -> np.ndarray:
        if "quat" in attr:
            q = _floats(attr["quat"], 4)
            return q / np.linalg.norm(q)
        if "euler" in attr:
            return euler2quat(_floats(attr["euler"], 3) * ang, eulerseq)
        return np.array([1.0, 0, 0, 0])

    # ---- walk the body tree (depth-first declaration order = MuJoCo ids)
    bodies, joints, geoms, sites = [], [], [], {}
    plane = None

    def walk(node: ET.Element, parent: int, childclass: str):
        nonlocal plane
        for ch in node:
            if ch.tag == "geom":
                g = dfl.resolve(ch, childclass)
                if g.get("type") == "plane":
                    if parent != 0:
                        raise ValueError("plane geoms must be attached to the world body")
                    plane = g
                elif g.get("type") == "mesh":
                    geoms.append((parent, g))
                else:
                    raise ValueError(f"unsupported geom type {g.get('type')!r} (mesh and plane only)")
            elif ch.tag == "joint" or ch.tag == "freejoint":
                if ch.tag == "freejoint":
                    j = {"type": "free", "damping": "0", "armature": "0", "name": ch.get("name", "")}
                else:
                    j = dfl.resolve(ch, childclass)
                joints.append((parent, j))
            elif ch.tag == "site":
                sites[ch.get("name")] = (parent, _floats(ch.get("pos"), 3, [0, 0, 0]), orient(ch.attrib))
            elif ch.tag == "body":
                bid = len(bodies)
                bodies.append({"name": ch.get("name", f"body{bid}"), "parent": parent,
                               "pos": _floats(ch.get("pos"), 3, [0, 0, 0]), "quat": orient(ch.attrib)})
                walk(ch, bid, ch.get("childclass", childclass))

    bodies.append({"name": "world", "parent": -1, "pos": np.zeros(3), "quat": np.array([1.0, 0, 0, 0])})
    for wb in root.findall("worldbody"):
        walk(wb, 0, None)
    nbody = len(bodies)
    if plane is None:
        raise ValueError("model has no floor plane")
    plane_pos = _floats(plane.get("pos"), 3, [0, 0, 0])
    plane_R = quat2mat(orient(plane))
    if abs(plane_R[2, 2] - 1.0) > 1e-12:
        raise ValueError("only a horizontal floor plane (normal +z) is supported")

    # ---- joints / dofs
    nq = nv = 0
    jt, jbody, jqadr, jdadr, jaxis, jpos, jrange, jlim, qpos0 = [], [], [], [], [], [], [], [], []
    dof_damping, dof_armature, dof_body, joint_names = [], [], [], []
    for b, j in joints:
        typ = j.get("type", "hinge")
        joint_names.append(j.get("name", ""))
        jbody.append(b)
        jqadr.append(nq)
        jdadr.append(nv)
        jpos.append(_floats(j.get("pos"), 3, [0, 0, 0]))
        ax = _floats(j.get("axis"), 3, [0, 0, 1])
        jaxis.append(ax / np.linalg.norm(ax))
        damping, arm = float(j.get("damping", 0)), float(j.get("armature", 0))
        if typ == "free":
            if bodies[b]["parent"] != 0:
                raise ValueError("free joint must be on a child of the world body")
            jt.append(0)
            jrange.append([0.0, 0.0])
            jlim.append(0)
            qpos0.extend(list(bodies[b]["pos"]) + list(bodies[b]["quat"]))
            nq += 7
            for _ in range(6):
                dof_damping.append(damping)
                dof_armature.append(arm)
                dof_body.append(b)
            nv += 6
        elif typ == "hinge":
            jt.append(3)
            rng = _floats(j.get("range"), 2, [0, 0]) * ang
            limited = j.get("limited", "auto")
            lim = (limited == "true") or (limited == "auto" and "range" in j)
            jrange.append(list(rng))
            jlim.append(int(lim))
            qpos0.append(float(j.get("ref", 0)) * ang)
            nq += 1
            dof_damping.append(damping)
            dof_armature.append(arm)
            dof_body.append(b)
            nv += 1
            if float(j.get("stiffness", 0)) != 0 or float(j.get("frictionloss", 0)) != 0:
                raise ValueError("joint stiffness / frictionloss are not supported")
            if float(j.get("margin", 0)) != 0:
                raise ValueError("joint margin is not supported")
        else:
            raise ValueError(f"unsupported joint type {typ!r}")
    jname2id = {n: i for i, n in enumerate(joint_names)}

    # ---- geoms -> body inertial properties + collision tables
    plane_fr = _floats(plane.get("friction"), 3, [1, 0.005, 0.0001])
    g_body, g_pos, g_quat, g_mesh, g_rbound, g_margin, g_mu = [], [], [], [], [], [], []
    g_solref, g_solimp = [], []
    body_mass = np.zeros(nbody)
    body_mc = np.zeros((nbody, 3))
    per_body: List[List[Tuple[float, np.ndarray, np.ndarray]]] = [[] for _ in range(nbody)]
    for b, g in geoms:
        me = meshes[g["mesh"]]
        pos = _floats(g.get("pos"), 3, [0, 0, 0])
        q = orient(g)
        Rg = quat2mat(q)
        gpos = pos + Rg @ me.pos
        gquat = quat_mul(q, me.quat)
        if "mass" in g:
            mass = float(g["mass"])
        else:
            mass = float(g.get("density", 1000.0)) * me.volume
        Rf = quat2mat(gquat)
        I_body = Rf @ np.diag(me.inertia_unit * mass) @ Rf.T  # about geom CoM, body axes
        per_body[b].append((mass, gpos, I_body))
        body_mass[b] += mass
        body_mc[b] += mass * gpos
        g_body.append(b)
        g_pos.append(gpos)
        g_quat.append(gquat)
        g_mesh.append(mesh_names.index(g["mesh"]))
        g_rbound.append(float(np.linalg.norm(me.aabb_absmax)))
        margin = max(float(g.get("margin", 0)), float(plane.get("margin", 0)))
        gap = max(float(g.get("gap", 0)), float(plane.get("gap", 0)))
        if gap != 0:
            raise ValueError("geom gap is not supported")
        g_margin.append(margin)
        fr = _floats(g.get("friction"), 3, [1, 0.005, 0.0001])
        g_mu.append(max(fr[0], plane_fr[0]))
        if int(g.get("condim", 3)) != 3 or int(plane.get("condim", 3)) != 3:
            raise ValueError("only condim=3 contacts are supported")
        # equal priority, equal solmix -> average of the two geoms' solref / solimp
        sr = 0.5 * (_floats(g.get("solref"), 2, [0.02, 1]) + _floats(plane.get("solref"), 2, [0.02, 1]))
        si = 0.5 * (_floats(g.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2]) +
                    _floats(plane.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2]))
        g_solref.append(sr)
        g_solimp.append(si)

    body_ipos = np.zeros((nbody, 3))
    body_inertia = np.zeros((nbody, 6))  # xx yy zz xy xz yz about the CoM, body axes
    for b in range(1, nbody):
        if body_mass[b] <= 0:
            raise ValueError(f"body {bodies[b]['name']} has no mass")
        com = body_mc[b] / body_mass[b]
        I = np.zeros((3, 3))
        for mass, gpos, Ig in per_body[b]:
            d = gpos - com
            I += Ig + mass * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
        body_ipos[b] = com
        body_inertia[b] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]

    # ---- actuators (position servos on joints)
    a_dof, a_gear, a_kp, a_bias, a_tau, a_ctrlr, a_ctrll, a_frcr, a_frcl = [], [], [], [], [], [], [], [], []
    for act in root.findall("actuator"):
        for el in act:
            if el.tag != "position":
                raise ValueError(f"unsupported actuator <{el.tag}> (position only)")
            a = dfl.resolve(el, None)
            jid = jname2id[a["joint"]]
            if jt[jid] != 3:
                raise ValueError("actuators must act on hinge joints")
            kp, kv = float(a.get("kp", 1)), float(a.get("kv", 0))
            a_dof.append(jdadr[jid])
            a_gear.append(_floats(a.get("gear"), None, [1])[0])
            a_kp.append(kp)
            a_bias.append([0.0, -kp, -kv])
            a_tau.append(float(a.get("timeconst", 0)))
            cr = _floats(a.get("ctrlrange"), 2, [0, 0])
            fr = _floats(a.get("forcerange"), 2, [0, 0])
            cl, fl = a.get("ctrllimited", "auto"), a.get("forcelimited", "auto")
            a_ctrlr.append(cr)
            a_ctrll.append(int(cl == "true" or (cl == "auto" and "ctrlrange" in a)))
            a_frcr.append(fr)
            a_frcl.append(int(fl == "true" or (fl == "auto" and "forcerange" in a)))
    nu = len(a_dof)

    # ---- sensors: the fused epilogue implements exactly the reference's block
    sensor_names, sensor_adr, sensor_dim = [], [], []
    adr = 0
    sens = [s for blk in root.findall("sensor") for s in blk]
    if len(sens) != len(_SENSOR_LAYOUT):
        raise ValueError("sensor block does not match the supported layout (quadruped.xml:174-217)")
    site_body = None
    for k, (s, (typ, dim)) in enumerate(zip(sens, _SENSOR_LAYOUT)):
        if s.tag != typ:
            raise ValueError(f"sensor {k}: expected <{typ}>, found <{s.tag}>")
        if typ == "jointpos":
            if jqadr[jname2id[s.get("joint")]] != 7 + k:
                raise ValueError("jointpos sensors must follow joint order")
        else:
            sname = s.get("site") or s.get("objname")
            sb, spos, squat = sites[sname]
            if np.abs(spos).max() > 0 or abs(squat[0] - 1) > 1e-12 or jt[jbody.index(sb)] != 0:
                raise ValueError("frame sensors must sit on a site at the free body's origin")
            site_body = sb
        sensor_names.append(s.get("name", ""))
        sensor_adr.append(adr)
        sensor_dim.append(dim)
        adr += dim

    A: Dict[str, np.ndarray] = {}
    A["sizes"] = np.array([nq, nv, nu, nbody, len(joints), len(g_body), len(mesh_names), adr], dtype=np.int32)
    timestep = float(opt.get("timestep", 0.002))
    grav = _floats(opt.get("gravity"), 3, [0, 0, -9.81])
    # [timestep, gx, gy, gz, tolerance, ls_tolerance, impratio, plane_z, meaninertia(filled below)]
    A["opt_f"] = np.array([timestep, grav[0], grav[1], grav[2], float(opt.get("tolerance", 1e-8)),
                           float(opt.get("ls_tolerance", 0.01)), float(opt.get("impratio", 1.0)),
                           plane_pos[2], 0.0])
    # [integrator(0 Euler,1 implicitfast), cone(0 pyramidal,1 elliptic), iterations, ls_iterations,
    #  plane-mesh extra-vertex rule (0 = far from all previous contacts, 1 = far from the first)]
    integ = {"Euler": 0, "implicitfast": 1}.get(integrator)
    if integ is None:
        raise ValueError(f"unsupported integrator {integrator!r}")
    A["opt_i"] = np.array([integ, {"pyramidal": 0, "elliptic": 1}[cone], int(opt.get("iterations", 100)),
                           int(opt.get("ls_iterations", 50)), 0], dtype=np.int32)
    A["body_parent"] = np.array([b["parent"] for b in bodies], dtype=np.int32)
    A["body_pos"] = np.array([b["pos"] for b in bodies])
    A["body_quat"] = np.array([b["quat"] for b in bodies])
    A["body_mass"] = body_mass
    A["body_ipos"] = body_ipos
    A["body_inertia"] = body_inertia
    A["jnt_type"] = np.array(jt, dtype=np.int32)
    A["jnt_body"] = np.array(jbody, dtype=np.int32)
    A["jnt_qposadr"] = np.array(jqadr, dtype=np.int32)
    A["jnt_dofadr"] = np.array(jdadr, dtype=np.int32)
    A["jnt_axis"] = np.array(jaxis)
    A["jnt_pos"] = np.array(jpos)
    A["jnt_range"] = np.array(jrange)
    A["jnt_limited"] = np.array(jlim, dtype=np.int32)
    A["jnt_solref"] = np.array([0.02, 1.0])
    A["jnt_solimp"] = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
    A["qpos0"] = np.array(qpos0)
    A["dof_damping"] = np.array(dof_damping)
    A["dof_armature"] = np.array(dof_armature)
    A["dof_body"] = np.array(dof_body, dtype=np.int32)
    A["act_dof"] = np.array(a_dof, dtype=np.int32)
    A["act_gear"] = np.array(a_gear)
    A["act_gain"] = np.array(a_kp)
    A["act_bias"] = np.array(a_bias)
    A["act_tau"] = np.array(a_tau)
    A["act_ctrlrange"] = np.array(a_ctrlr)
    A["act_ctrllimited"] = np.array(a_ctrll, dtype=np.int32)
    A["act_frcrange"] = np.array(a_frcr)
    A["act_frclimited"] = np.array(a_frcl, dtype=np.int32)
    A["geom_body"] = np.array(g_body, dtype=np.int32)
    A["geom_pos"] = np.array(g_pos)
    A["geom_quat"] = np.array(g_quat)
    A["geom_mesh"] = np.array(g_mesh, dtype=np.int32)
    A["geom_rbound"] = np.array(g_rbound)
    A["geom_margin"] = np.array(g_margin)
    A["geom_mu"] = np.array(g_mu)
    A["geom_solref"] = np.array(g_solref)
    A["geom_solimp"] = np.array(g_solimp)
    vadr, vnum, eadr, verts, vedge, edges = [], [], [], [], [], []
    ceadr, cvedge, cedges = [], [], []
    for name in mesh_names:
        me = meshes[name]
        vadr.append(sum(vnum))
        vnum.append(len(me.hull_vert))
        eadr.append(len(edges))
        verts.append(me.hull_vert)
        vedge.append(me.hull_edgeadr)
        edges.extend(me.hull_edge.tolist())
        ceadr.append(len(cedges))
        cvedge.append(me.hull_cedgeadr)
        cedges.extend(me.hull_cedge.tolist())
    A["mesh_vertadr"] = np.array(vadr, dtype=np.int32)
    A["mesh_vertnum"] = np.array(vnum, dtype=np.int32)
    A["mesh_edgeadr"] = np.array(eadr, dtype=np.int32)  # start of the mesh' edge list in mesh_edge
    A["mesh_vert"] = np.concatenate(verts, 0)
    A["mesh_vert_edge"] = np.concatenate(vedge, 0).astype(np.int32)  # per vertex, relative to mesh_edgeadr
    A["mesh_edge"] = np.array(edges, dtype=np.int32)  # local vertex ids, -1 terminates a list
    # polytope-edge graph (support-search hill climbing); same layout as the three sections above
    A["mesh_cedgeadr"] = np.array(ceadr, dtype=np.int32)
    A["mesh_vert_cedge"] = np.concatenate(cvedge, 0).astype(np.int32)
    A["mesh_cedge"] = np.array(cedges, dtype=np.int32)

    cm = CompiledModel(arrays=A, sensor_names=sensor_names, sensor_adr=sensor_adr, sensor_dim=sensor_dim,
                       joint_names=joint_names, body_names=[b["name"] for b in bodies], source=path)
    set_const(cm)
    return cm


# ----------------------------------------------------------------------------- mj_setConst restatement


def kinematics(A: Dict[str, np.ndarray], qpos: np.ndarray):
    """World poses of all bodies -> (xpos[nbody,3], xmat[nbody,3,3])."""
    nbody = len(A["body_parent"])
    xpos = np.zeros((nbody, 3))
    xmat = np.tile(np.eye(3), (nbody, 1, 1))
    jb = A["jnt_body"]
    for b in range(1, nbody):
        p = A["body_parent"][b]
        jids = np.nonzero(jb == b)[0]
        if len(jids) and A["jnt_type"][jids[0]] == 0:
            adr = A["jnt_qposadr"][jids[0]]
            xpos[b] = qpos[adr:adr + 3]
            xmat[b] = quat2mat(qpos[adr + 3:adr + 7])
            continue
        xpos[b] = xpos[p] + xmat[p] @ A["body_pos"][b]
        R = xmat[p] @ quat2mat(A["body_quat"][b])
        for j in jids:
            th = qpos[A["jnt_qposadr"][j]] - A["qpos0"][A["jnt_qposadr"][j]]
            ax = A["jnt_axis"][j]
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            Rj = np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)
            anchor = A["jnt_pos"][j]
            xpos[b] = xpos[b] + R @ anchor - R @ Rj @ anchor
            R = R @ Rj
        xmat[b] = R
    return xpos, xmat


def jacobian(A, xpos, xmat, body: int, point: np.ndarray):
    """Translational and rotational Jacobians (3 x nv each) of ``point`` fixed to ``body``."""
    nv = len(A["dof_body"])
    jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
    b = body
    while b > 0:
        for j in np.nonzero(A["jnt_body"] == b)[0]:
            d = A["jnt_dofadr"][j]
            if A["jnt_type"][j] == 0:
                jp[:, d:d + 3] = np.eye(3)
                for k in range(3):
                    ax = xmat[b][:, k]
                    jr[:, d + 3 + k] = ax
                    jp[:, d + 3 + k] = np.cross(ax, point - xpos[b])
            else:
                ax = xmat[b] @ A["jnt_axis"][j]
                anchor = xpos[b] + xmat[b] @ A["jnt_pos"][j]
                jr[:, d] = ax
                jp[:, d] = np.cross(ax, point - anchor)
        b = A["body_parent"][b]
    return jp, jr


def _sym6(v):
    return np.array([[v[0], v[3], v[4]], [v[3], v[1], v[5]], [v[4], v[5], v[2]]])


def mass_matrix(A, qpos):
    xpos, xmat = kinematics(A, qpos)
    nv = len(A["dof_body"])
    M = np.diag(A["dof_armature"].astype(float))
    for b in range(1, len(A["body_parent"])):
        com = xpos[b] + xmat[b] @ A["body_ipos"][b]
        jp, jr = jacobian(A, xpos, xmat, b, com)
        Iw = xmat[b] @ _sym6(A["body_inertia"][b]) @ xmat[b].T
        M += A["body_mass"][b] * jp.T @ jp + jr.T @ Iw @ jr
    return M, xpos, xmat


def set_const(cm: CompiledModel) -> None:
    """``body_invweight0`` / ``dof_invweight0`` / ``meaninertia`` at ``qpos0`` (mj_setConst)."""
    A = cm.arrays
    M, xpos, xmat = mass_matrix(A, A["qpos0"])
    Minv = np.linalg.inv(M)
    nbody, nv = len(A["body_parent"]), M.shape[0]
    biw = np.zeros((nbody, 2))
    for b in range(1, nbody):
        com = xpos[b] + xmat[b] @ A["body_ipos"][b]
        jp, jr = jacobian(A, xpos, xmat, b, com)
        biw[b, 0] = np.trace(jp @ Minv @ jp.T) / 3.0
        biw[b, 1] = np.trace(jr @ Minv @ jr.T) / 3.0
    diw = np.zeros(nv)
    for j in range(len(A["jnt_type"])):
        d = A["jnt_dofadr"][j]
        if A["jnt_type"][j] == 0:
            diw[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            diw[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        else:
            diw[d] = Minv[d, d]
    A["body_invweight0"] = biw
    A["dof_invweight0"] = diw
    A["opt_f"][8] = np.trace(M) / nv
