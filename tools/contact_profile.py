"""Where do the contacts of the steady-state C3 population sit?  (sizing of the solver's per-contact loops: lane l of a
quad owns leg l, so a leg with 4 contacts makes its lane run 4 trips while the other three idle.)  Device rollout, then
the CPU oracle's forward pass on the device states gives contact -> geom -> body -> leg."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.oracle import OracleData, OracleModel
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R
from quadruped_gym_b200.model import DEFAULT_BLOB, blob as qblob
n = 4096
blob = open(DEFAULT_BLOB, "rb").read()
A = qblob.unpack(blob)
om = OracleModel(blob)
env = VecQuadrupedEnv(n, "cuda:0", termination_fns={"flip": R.flip_termination()}, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for i in range(150):
    env.step(torch.rand((n, 12), device="cuda", generator=g) * 2 - 1)
st = {k: getattr(env.data, k).cpu().numpy() for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "ctrl")}
parent = A["body_parent"]
def leg_of(b):
    if b == 1: return -1
    while parent[b] != 1: b = parent[b]
    return [bb for bb in range(2, len(parent)) if parent[bb] == 1].index(b)
per_leg = np.zeros((n, 5), int)
for e in range(n):
    d = OracleData(om)
    d.set_state(st["qpos"][e].astype(float), st["qvel"][e].astype(float), st["act"][e].astype(float), st["qacc_warmstart"][e].astype(float), float(st["time"][e]), st["ctrl"][e].astype(float))
    d.forward()
    for c in range(d.ncon):
        per_leg[e, leg_of(int(A["geom_body"][int(d.con_geom[c])]))] += 1
tot = per_leg.sum(1); mx = per_leg[:, :4].max(1); base = per_leg[:, 4]
print(f"{n} envs after 150 steps: contacts per env mean {tot.mean():.2f}; envs with contacts {np.mean(tot > 0):.3f}")
print("hist of total contacts per env  :", np.bincount(tot, minlength=17)[:17].tolist())
print("hist of MAX contacts on one leg :", np.bincount(mx, minlength=9)[:9].tolist())
print("hist of base-body contacts      :", np.bincount(base, minlength=9)[:9].tolist())
c = tot > 0
print(f"among envs in contact: serial trips now = max-leg count, mean {mx[c].mean():.2f}; balanced over 4 lanes = ceil(total/4), mean {np.ceil(tot[c] / 4).mean():.2f}; "
      f"legs in contact mean {(per_leg[c][:, :4] > 0).sum(1).mean():.2f}")
