"""What does leaving out SELF-collision cost?  (MuJoCo also collides the robot's own geoms, 226
pairs after the parent/child filter; only the adjacent-leg {shin, foot, ankle servo} hulls can touch, and only past the
joint ranges.)  This tool MEASURES it on the workloads of BASELINE configs 3 and 5: the CPU oracle rolls the robots out
under random actions, and at every env.step() the 36 candidate hull pairs are tested with an exact convex-convex distance
(GJK on the hull vertices, after a bounding-sphere cull).  A pair closer than the contact margin (1 mm) is a contact that
MuJoCo would have generated and this simulator does not.

    python tools/selfcollision_exposure.py [n_envs] [n_steps] > profiles/r2_selfcollision_exposure.txt

CPU only (test infrastructure: uses the oracle).
"""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.oracle import OracleData, OracleModel
from quadruped_gym_b200.model import DEFAULT_BLOB, blob as qblob

N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
T = int(sys.argv[2]) if len(sys.argv) > 2 else 250
MARGIN = 1e-3


def closest_on_simplex(P):
    """Closest point to the origin on conv(P) (P: k x 3, k <= 4) -> (point, indices of the supporting sub-simplex)."""
    k = len(P)
    best, bi = None, None
    for r in range(1, k + 1):
        for idx in itertools.combinations(range(k), r):
            Q = P[list(idx)]
            if r == 1:
                lam = np.ones(1)
            else:
                D = (Q[1:] - Q[0]).T                      # 3 x (r-1)
                try:
                    t = np.linalg.solve(D.T @ D, -D.T @ Q[0])
                except np.linalg.LinAlgError:
                    continue
                lam = np.r_[1 - t.sum(), t]
                if (lam < -1e-12).any():
                    continue
            x = lam @ Q
            d = x @ x
            if best is None or d < best[0] - 1e-18:
                best, bi = (d, x), idx
    return best[1], list(bi)


def gjk_distance(A, B, iters=64):
    """Distance between conv(A) and conv(B) (vertex arrays, world frame); 0 when they intersect."""
    d = A.mean(0) - B.mean(0)
    if not d.any():
        d = np.array([1.0, 0, 0])
    sup = lambda d: A[np.argmin(A @ d)] - B[np.argmax(B @ d)]     # support of A - B in direction -d
    S = [sup(d)]
    x = S[0]
    for _ in range(iters):
        n2 = x @ x
        if n2 < 1e-16:
            return 0.0
        w = sup(x)
        if n2 - x @ w <= 1e-10 * max(n2, 1e-12):                  # no progress along -x: x is the closest point
            return float(np.sqrt(n2))
        S.append(w)
        x, keep = closest_on_simplex(np.array(S))
        S = [S[i] for i in keep]
        if len(S) == 4:
            return 0.0                                            # origin enclosed by a tetrahedron
    return float(np.sqrt(x @ x))


def quat2mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def setup(A):
    nb = len(A["body_parent"])
    leg_of, level_of = {}, {}
    for b in range(2, nb):
        p = A["body_parent"][b]
        leg_of[b], level_of[b] = (len([1 for bb in range(2, b) if A["body_parent"][bb] == 1]), 1) if p == 1 else (leg_of[p], level_of[p] + 1)
    geoms = []
    for g in range(len(A["geom_body"])):
        b = int(A["geom_body"][g])
        if b >= 2 and level_of[b] >= 2:                           # shin (level 2) and foot + ankle servo (level 3)
            me = int(A["geom_mesh"][g])
            v0, vn = int(A["mesh_vertadr"][me]), int(A["mesh_vertnum"][me])
            V = A["mesh_vert"].reshape(-1, 3)[v0:v0 + vn]
            geoms.append(dict(g=g, body=b, leg=leg_of[b], R=quat2mat(A["geom_quat"].reshape(-1, 4)[g]), p=A["geom_pos"].reshape(-1, 3)[g],
                              V=V, rb=float(A["geom_rbound"][g])))
    pairs = [(i, j) for i in range(len(geoms)) for j in range(i + 1, len(geoms))
             if (geoms[i]["leg"] - geoms[j]["leg"]) % 4 in (1, 3)]                  # adjacent legs only (App. D)
    return geoms, pairs


def run(name, blob_bytes, pose_fn=None, seed=0):
    A = qblob.unpack(blob_bytes)
    geoms, pairs = setup(A)
    om = OracleModel(blob_bytes)
    rng = np.random.default_rng(seed)
    hits = tested = culled = 0
    env_hit = np.zeros(N, bool)
    mind = np.inf
    limit_steps = 0
    hit_with_limit = 0
    for e in range(N):
        d = OracleData(om)
        d.ctrl[:] = [0, 0, -0.5] * 4
        if pose_fn is not None:
            d.set_state(pose_fn(rng), np.zeros(18), np.zeros(12), np.zeros(18), 0.0, np.array([0, 0, -0.5] * 4, float))
        a = rng.uniform(-1, 1, 12)
        for t in range(T):
            if t % 5 == 0:
                a = rng.uniform(-1, 1, 12)
            d.env_step(a, 4)
            limit_steps += int(d.nlimit > 0)
            xpos, xmat = d.xpos.reshape(-1, 3), d.xmat.reshape(-1, 3, 3)
            W = {}
            step_hit = False
            for i, j in pairs:
                gi, gj = geoms[i], geoms[j]
                ci = xpos[gi["body"]] + xmat[gi["body"]] @ gi["p"]
                cj = xpos[gj["body"]] + xmat[gj["body"]] @ gj["p"]
                if np.linalg.norm(ci - cj) > gi["rb"] + gj["rb"] + MARGIN:
                    culled += 1
                    continue
                for k, gk, ck in ((i, gi, ci), (j, gj, cj)):
                    if k not in W:
                        W[k] = ck + gk["V"] @ (xmat[gk["body"]] @ gk["R"]).T
                tested += 1
                dist = gjk_distance(W[i], W[j])
                mind = min(mind, dist)
                if dist < MARGIN:
                    hits += 1
                    step_hit = True
            if step_hit:
                env_hit[e] = True
                hit_with_limit += int(d.nlimit > 0)
    steps = N * T
    print(f"{name}: {N} envs x {T} env.step() (frame_skip 4) = {steps} states; candidate pairs {len(pairs)} per state")
    print(f"   bounding-sphere culled {culled}, GJK-tested {tested}; pairs within the 1 mm margin: {hits} "
          f"({hits / steps:.4f} per state, {100 * hits / max(1, steps * len(pairs)):.4f} % of the pair tests); smallest hull-hull distance seen {mind * 1e3:.2f} mm")
    print(f"   environments that ever had a would-be self-contact: {int(env_hit.sum())}/{N}; states with an active joint-limit row: {limit_steps} "
          f"({100 * limit_steps / steps:.2f} %); self-contact states that also had an active limit: {hit_with_limit}")


def c5_pose(rng):
    yaw, tilt, tdir = rng.uniform(0, 2 * np.pi), rng.uniform(0, np.pi / 6), rng.uniform(0, 2 * np.pi)
    qt = np.array([np.cos(tilt / 2), np.cos(tdir) * np.sin(tilt / 2), np.sin(tdir) * np.sin(tilt / 2), 0.0])
    qy = np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)])
    w1, x1, y1, z1 = qy; w2, x2, y2, z2 = qt
    quat = np.array([w1*w2 - x1*x2 - y1*y2 - z1*z2, w1*x2 + x1*w2 + y1*z2 - z1*y2, w1*y2 - x1*z2 + y1*w2 + z1*x2, w1*z2 + x1*y2 - y1*x2 + z1*w2])
    lo, hi = np.tile(np.deg2rad([-45, -45, -90]), 4), np.tile(np.deg2rad([45, 120, 90]), 4)
    return np.r_[0, 0, rng.uniform(0.05, 0.2), quat, lo + (hi - lo) * rng.random(12)]


base = open(DEFAULT_BLOB, "rb").read()
print("self-collision exposure (what MuJoCo's robot-robot contacts would add; this simulator collides every geom with the floor only)\n")
run("C3 workload (stock model, reset pose, random actions U(-1,1) held 40 ms)", base)
A = qblob.unpack(base)
A["opt_i"][1] = 1
A["act_gain"] = A["act_gain"] * 5.0
A["act_bias"] = A["act_bias"].reshape(-1, 3) * np.array([1.0, 5.0, 1.0])
run("C5 workload (elliptic cone, servo gains x5, randomised initial poses)", qblob.pack(A), c5_pose, seed=1)
