"""Exploratory GPU check: teacher-forced single-step parity of the CUDA path against the CPU oracle,
stage by stage, on states sampled from oracle rollouts.  Prints error statistics (development aid;
the assertions live in tests/)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.oracle import OracleData, OracleModel  # noqa: E402
from quadruped_gym_b200 import VecQuadrupedEnv  # noqa: E402
from quadruped_gym_b200.model import DEFAULT_BLOB  # noqa: E402
from tests.conftest import rollout_states  # noqa: E402

np.set_printoptions(precision=4, suppress=False, linewidth=200)


def to_bform(M, R):
    T = np.eye(18)
    T[:3, :3] = R
    return T.T @ M @ T


def main():
    n = int(os.environ.get("N", 256))
    blob = open(DEFAULT_BLOB, "rb").read()
    om = OracleModel(blob)
    t0 = time.time()
    st = rollout_states(om, n, 120, seed=1)
    print(f"sampled {n} states in {time.time() - t0:.1f}s")
    f32 = {k: v.astype(np.float32) for k, v in st.items() if k != "time"}
    env = VecQuadrupedEnv(n, "cuda:0", auto_reset=False)
    env.set_state(qpos=f32["qpos"], qvel=f32["qvel"], act=f32["act"], qacc_warmstart=f32["warm"], time=st["time"], ctrl=f32["ctrl"])
    rng = np.random.default_rng(7)
    ctrl = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
    out = env.debug_step(ctrl)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in out.items()}
    gq, gv = env.data.qpos.cpu().numpy(), env.data.qvel.cpu().numpy()
    ga, gw = env.data.act.cpu().numpy(), env.data.qacc_warmstart.cpu().numpy()

    err = {k: [] for k in ("M", "bias", "qacc_smooth", "qacc", "qacc_abs", "sens", "qpos", "qvel", "act", "acc_sensor")}
    ncon_o, ncon_g, nit_o, nit_g = [], [], [], []
    mism = 0
    for e in range(n):
        d = OracleData(om)
        d.set_state(f32["qpos"][e].astype(np.float64), f32["qvel"][e].astype(np.float64), f32["act"][e].astype(np.float64),
                    f32["warm"][e].astype(np.float64), st["time"][e], ctrl[e].astype(np.float64))
        d.forward()
        R = d.xmat[9:18].reshape(3, 3)
        MB = to_bform(d.M.copy(), R)
        err["M"].append(np.abs(g["M"][e] - MB).max() / np.abs(MB).max())
        bias = d.qfrc_bias.copy()
        bias[:3] = R.T @ bias[:3]
        err["bias"].append(np.abs(g["qfrc_bias"][e] - bias).max() / (np.abs(bias).max() + 1e-9))
        err["qacc_smooth"].append(np.abs(g["qacc_smooth"][e] - d.qacc_smooth).max() / (np.abs(d.qacc_smooth).max() + 1e-9))
        err["qacc"].append(np.abs(g["qacc"][e] - d.qacc).max() / (np.abs(d.qacc).max() + 1e-9))
        err["qacc_abs"].append(np.abs(g["qacc"][e] - d.qacc).max())
        sens = d.sensordata.copy()
        ds = np.abs(g["sensordata"][e] - sens)
        err["acc_sensor"].append(ds[12:15].max())
        ds[12:15] = 0
        err["sens"].append(ds.max())
        ncon_o.append(d.ncon); ncon_g.append(g["counts"][e, 0]); nit_o.append(d.solver_niter); nit_g.append(g["counts"][e, 2])
        if d.ncon != g["counts"][e, 0] or d.nefc != g["counts"][e, 1]:
            mism += 1
        # finish the step in the oracle
        d2 = OracleData(om)
        d2.set_state(f32["qpos"][e].astype(np.float64), f32["qvel"][e].astype(np.float64), f32["act"][e].astype(np.float64),
                     f32["warm"][e].astype(np.float64), st["time"][e], ctrl[e].astype(np.float64))
        d2.step()
        err["qpos"].append(np.abs(gq[e] - d2.qpos).max())
        err["qvel"].append(np.abs(gv[e] - d2.qvel).max() / (np.abs(d2.qvel).max() + 1e-3))
        err["act"].append(np.abs(ga[e] - d2.act).max())
    ncon_o, ncon_g = np.array(ncon_o), np.array(ncon_g)
    print("contacts oracle mean %.2f  gpu mean %.2f  mismatching envs %d / %d" % (ncon_o.mean(), ncon_g.mean(), mism, n))
    print("newton iters oracle mean %.2f max %d | gpu mean %.2f max %d" % (np.mean(nit_o), np.max(nit_o), np.mean(nit_g), np.max(nit_g)))
    same = ncon_o == ncon_g
    for k, v in err.items():
        v = np.array(v)
        vs = v[same]
        print(f"{k:12s} all: median {np.median(v):.3e} p99 {np.percentile(v, 99):.3e} max {v.max():.3e} | same-contact-set: max {vs.max():.3e}"
              f" | no-contact max {v[ncon_o == 0].max() if (ncon_o == 0).any() else float('nan'):.3e}")
    worst = int(np.argmax(np.array(err["qacc"])))
    print("worst qacc env", worst, "ncon", ncon_o[worst], ncon_g[worst], "iters", nit_o[worst], nit_g[worst])
    print(env.counters())

    # short-horizon open-loop rollout parity from reset
    n2, T = 64, 60
    env2 = VecQuadrupedEnv(n2, "cuda:0", auto_reset=False, frame_skip=4)
    env2.reset()
    acts = rng.uniform(-1, 1, (T, n2, 12)).astype(np.float32)
    ds = [OracleData(om) for _ in range(n2)]
    for d in ds:
        d.ctrl[:] = [0, 0, -0.5] * 4
    for t in range(T):
        obs, rew, term, trunc, info = env2.step(torch.from_numpy(acts[t]).cuda())
        o = obs.cpu().numpy()
        oo = np.zeros((n2, 33))
        for e, d in enumerate(ds):
            d.env_step(acts[t, e].astype(np.float64), 4)
            oo[e] = d.sensordata
        dd = np.abs(o - oo)
        dd[:, 12:15] = 0
        if t % 5 == 0 or t == T - 1:
            qg = env2.data.qpos.cpu().numpy()
            qo = np.array([d.qpos.copy() for d in ds])
            print(f"t={t:3d} obs err (no accel) max {dd.max():.3e} median-env {np.median(dd.max(1)):.3e} | qpos err max {np.abs(qg - qo).max():.3e} median {np.median(np.abs(qg - qo).max(1)):.3e}  ncon {np.mean([d.ncon for d in ds]):.2f}")

    # throughput quick look
    for nn in (4096, 65536):
        envb = VecQuadrupedEnv(nn, "cuda:0", auto_reset=True, termination_fns={})
        from quadruped_gym_b200.envs import rewards as R
        envb.termination_fns["default"] = R.time_limit()
        envb.termination_fns["flip"] = R.flip_termination()
        envb.reset()
        a = torch.rand((nn, 12), device="cuda") * 2 - 1
        for _ in range(50):
            envb.step(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 100
        for i in range(K):
            if i % 5 == 0:
                a = torch.rand((nn, 12), device="cuda") * 2 - 1
            envb.step(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"N={nn}: {ms / K:.3f} ms/step  -> {nn * 4 * K / (ms * 1e-3):.3e} physics env-steps/s", envb.counters(reset=True))


if __name__ == "__main__":
    main()
