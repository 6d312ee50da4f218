"""BASELINE config 2: 4,096 envs, random actions, frame_skip 4 -- every env.step() is ALSO teacher-forced through the
CPU oracle from the device's own pre-step state and compared (qpos, qvel, sensordata), then a soak run at 65,536 envs.
Writes a text report (profiles/r1_validation.txt / r2_validation.txt are copies of one run; the driver-visible form of the
first half is tests/test_parity_configs_gpu.py::test_config2_every_step_teacher_forced)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.oracle import OracleBatch, OracleModel
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R
from quadruped_gym_b200.model import DEFAULT_BLOB

N, T, FS = 4096, int(os.environ.get("T", 60)), 4
om = OracleModel(open(DEFAULT_BLOB, "rb").read())
env = VecQuadrupedEnv(N, "cuda:0", frame_skip=FS, auto_reset=False)
env.reset()
ob = OracleBatch(om, N)
rng = np.random.default_rng(0)
print(f"C2 teacher-forced parity: {N} envs x {T} env.step() (frame_skip {FS}), oracle restarted from the device state every step")
print("step  ncon/env  same-contact-count%   qpos max|err| (p50 / p99 / max)      qvel rel err (p50 / p99 / max)     sensordata(no accel) max")
tot_same = tot = 0
worst = dict(qpos=0.0, qvel=0.0)
for t in range(T):
    a = rng.uniform(-1, 1, (N, 12)).astype(np.float32)
    st = {k: getattr(env.data, k).cpu().numpy() for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "ctrl")}
    for e in range(N):
        d = ob.env(e)
        d.set_state(st["qpos"][e].astype(np.float64), st["qvel"][e].astype(np.float64), st["act"][e].astype(np.float64),
                    st["qacc_warmstart"][e].astype(np.float64), float(st["time"][e]), st["ctrl"][e].astype(np.float64))
    c0 = env.counters(reset=True)
    obs, *_ = env.step(torch.from_numpy(a).cuda())
    cg = env.counters(reset=True)
    oo = ob.rollout(a[None].astype(np.float64), FS, 1e9, False, want_obs=True)[0]
    qg, vg = env.data.qpos.cpu().numpy(), env.data.qvel.cpu().numpy()
    qo = np.array([ob.env(e).qpos.copy() for e in range(N)]); vo = np.array([ob.env(e).qvel.copy() for e in range(N)])
    nco = np.array([ob.env(e).ncon for e in range(N)])
    eq = np.abs(qg - qo).max(1); ev = np.abs(vg - vo).max(1) / np.maximum(1.0, np.abs(vo).max(1))
    es = np.abs(obs.cpu().numpy() - oo); es[:, 12:15] = 0; es = es.max(1)
    good = eq < 1e-4          # envs whose contact set flipped differ by O(1) in qacc
    tot_same += int(good.sum()); tot += N
    if t % 5 == 0 or t == T - 1:
        print(f"{t:4d}  {nco.mean():6.2f}   {100*good.mean():8.3f}        {np.median(eq):.2e} / {np.percentile(eq,99):.2e} / {eq.max():.2e}      "
              f"{np.median(ev):.2e} / {np.percentile(ev,99):.2e} / {ev.max():.2e}    {np.median(es):.2e}")
    worst["qpos"] = max(worst["qpos"], float(np.percentile(eq, 99))); worst["qvel"] = max(worst["qvel"], float(np.percentile(ev, 99)))
print(f"fraction of (env, step) pairs within 1e-4 abs on qpos after one env.step(): {tot_same/tot:.5f}; worst p99 qpos {worst['qpos']:.2e}, qvel rel {worst['qvel']:.2e}")
env.close()

# soak: 65,536 envs x 1500 env.step() with auto-reset (flip / time limit): no divergence, no table overflow
n = 65536
env = VecQuadrupedEnv(n, "cuda:0", frame_skip=4, auto_reset=True, termination_fns={"flip": R.flip_termination()})
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
t0 = time.time()
for i in range(1500):
    if i % 5 == 0:
        a = torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
    obs, rew, term, _, _ = env.step(a)
torch.cuda.synchronize()
c = env.counters()
print(f"soak: {n} envs x 1500 steps in {time.time()-t0:.1f}s wall; finite obs {bool(torch.isfinite(obs).all())}; counters {c}")
ps = c['physics_steps']
print("per physics step: contacts %.3f rows %.3f newton %.3f ls %.3f verts %.2f; episodes %d (flip or 10 s limit)" % (
    c['contacts']/ps, c['efc_rows']/ps, c['newton_iters']/ps, c['ls_evals']/ps, c['verts_tested']/ps, c['episodes']))
