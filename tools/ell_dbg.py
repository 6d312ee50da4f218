import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle.oracle import OracleModel, OracleData
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.model import blob as qblob, DEFAULT_BLOB
from tests.conftest import rollout_states
A = qblob.unpack(open(DEFAULT_BLOB,'rb').read()); A['opt_i'][1]=1; A['opt_f'][6]=4.0
eb = qblob.pack(A); om = OracleModel(eb)
n=256
st = rollout_states(om, n, 150, seed=17)
st32 = {k: v.astype(np.float32) for k, v in st.items() if k != 'time'}
env = VecQuadrupedEnv(n, 'cuda:0', auto_reset=False, model_blob=eb)
env.set_state(qpos=st32['qpos'], qvel=st32['qvel'], act=st32['act'], qacc_warmstart=st32['warm'], time=st['time'], ctrl=st32['ctrl'])
ctrl = np.random.default_rng(5).uniform(-1, 1, (n, 12)).astype(np.float32)
out = {k: v.cpu().numpy() for k, v in env.debug_step(ctrl).items()}
rows=[]
for e in range(n):
    d = OracleData(om)
    d.set_state(st32['qpos'][e].astype(float), st32['qvel'][e].astype(float), st32['act'][e].astype(float), st32['warm'][e].astype(float), st['time'][e], ctrl[e].astype(float))
    d.forward()
    err = np.abs(out['qacc'][e]-d.qacc).max()/max(1,np.abs(d.qacc).max())
    rows.append((err, e, d.ncon, out['counts'][e,0], d.solver_niter, out['counts'][e,2], d.ls_evals, out['counts'][e,3]))
rows.sort(reverse=True)
print('err env ncon_o ncon_g it_o it_g ls_o ls_g')
for r in rows[:12]: print('%.2e'%r[0], r[1:])
print('median err', np.median([r[0] for r in rows]), 'iters o/g', np.mean([r[4] for r in rows]), np.mean([r[5] for r in rows]))
