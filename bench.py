#!/usr/bin/env python
"""bench.py -- physics env-steps/s of the fused QuadrupedEnv.step() path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" = one ``env.step()`` over the whole batch = ONE launch of qg_step_kernel
(frame_skip physics steps + sensor / reward / termination / auto-reset epilogue).
Default workload = BASELINE.json configs[2]: 65,536 envs per GPU, random actions, default
frame_skip = 4, auto-reset on fall (zaxis_z < 0) or time limit -- the configuration the metric and
the 1e8 target are quoted on.  `value` = N_envs * frame_skip * K / (sum of the K CUDA-event step
times, max over ranks); L2 is flushed between timed steps.  `e2e` = the same metric through
qg_step_host (HOST action buffer in, HOST obs/reward/terminated out, copies inside the timed
region).  `--impl reference` times the reference's CPU implementation of the path -- here the CPU
oracle port (MuJoCo itself is not installable in this image) -- on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NPOOL = int(os.environ.get("QG_BENCH_NPOOL", "64"))   # independent random action tensors cycled through (period 64 steps = 0.5 s of simulated time)
PREROLL = 100   # untimed env steps before the warm-up (robots have landed and stumble under random actions)
METRIC = "env-steps/sec (physics steps incl. frame_skip)"
UNIT = "physics env-steps/s"

WORKLOADS = {
    # name: (envs per GPU, frame_skip, max_time, description)
    "c2": (4096, 4, 10.0, "C2: batched QuadrupedEnv 4,096 envs, random actions, frame_skip 4"),
    "c3": (65536, 4, 10.0, "C3: 65,536 envs/GPU random-action rollout, frame_skip 4, auto-reset on fall/time-limit"),
    "c4": (8192, 10, 20.0, "C4: PPO rollout collection 8,192 envs x 24-step horizon, frame_skip 10, forward+control+alive rewards"),
    "c5": (262144, 4, 10.0, "C5: contact-heavy stress 262,144 envs/GPU, randomised initial poses, actuator kp x5, elliptic friction cone"),
}


def workload_config(name: str, envs: int, fs: int, max_time: float, desc: str) -> dict:
    """The part of `config` that names the WORKLOAD: identical for our arm and the reference arm."""
    return {"workload": desc, "envs_per_gpu": envs, "frame_skip": fs, "max_time_s": max_time,
            "actions": "U(-1,1)^12 per env and env.step()", "auto_reset": "on fall (zaxis_z < 0) or time limit",
            "rewards": "forward(qvel_x) - 0.1*sum(ctrl^2) + alive"}


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One worker = one env stepped in a tight C loop (oracle port), like one SubprocVecEnv process
    of the reference (train_quadruped.py:50) without the pickling."""
    seed, n_env_steps, frame_skip, max_time = args
    from oracle.oracle import OracleBatch, OracleModel
    from quadruped_gym_b200.model import DEFAULT_BLOB
    om = OracleModel(open(DEFAULT_BLOB, "rb").read())
    ob = OracleBatch(om, 1)
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, (n_env_steps, 1, 12))
    ob.rollout(acts[:50], frame_skip, max_time)  # warm-up (drop + landing)
    t0 = time.perf_counter()
    ob.rollout(acts, frame_skip, max_time)
    return time.perf_counter() - t0


_POOL = {}


def _close_pools():
    for p in _POOL.values():
        p.terminate()
        p.join()
    _POOL.clear()


def _pool(procs):
    import multiprocessing as mp
    if procs not in _POOL:
        _POOL[procs] = mp.get_context("fork").Pool(procs)
    return _POOL[procs]


def cpu_rate(procs: int, n_env_steps: int, frame_skip: int, max_time: float):
    """physics env-steps/s of `procs` independent single-env workers (aggregate / max worker wall)."""
    args = [(1000 + i, n_env_steps, frame_skip, max_time) for i in range(procs)]
    walls = [_cpu_worker(args[0])] if procs == 1 else _pool(procs).map(_cpu_worker, args)
    return procs * n_env_steps * frame_skip / max(walls), max(walls)


def run_reference(a):
    """The reference arm: CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    envs, fs, max_time, desc = WORKLOADS[a.workload]
    cores = os.cpu_count() or 1
    rate, _ = cpu_rate(cores, 300, fs, max_time)  # calibrate the bounded sample (also warms the pool)
    for _ in range(a.warmup):
        cpu_rate(cores, 100, fs, max_time)
    wall_per_step = min(0.5, 90.0 / max(a.steps, 1))
    per_step = max(50, int(rate / (cores * fs) * wall_per_step))  # env.step() calls per worker per bench step
    t_tot, units = 0.0, 0
    for _ in range(a.steps):
        _, wall = cpu_rate(cores, per_step, fs, max_time)
        t_tot += wall
        units += cores * per_step * fs
    value = units / t_tot
    sample = f"{cores} processes x 1 env x {per_step} env.step() (frame_skip {fs}) per bench step, random actions, auto-reset"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t_tot / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(a.workload, envs, fs, max_time, desc),
                       note="same workload definition, BOUNDED SAMPLE: each of the %d host processes steps 1 env (the reference's "
                            "SubprocVecEnv shape, train_quadruped.py:50) instead of envs_per_gpu envs; CPU oracle port of mj_step for "
                            "this model class in float64 (the MuJoCo wheel is not installable in this image), so the ratio against "
                            "this arm is 'vs our own CPU port', not 'vs MuJoCo'" % cores),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                smax = float(p[1])
                if t0 <= t <= t1 + 0.1:
                    sm.append(float(p[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ GPU arm
def flops_per_physics_step(c):
    """SURVEY.md App. E work model evaluated with the kernel's own counters (per physics env-step)."""
    n = max(c["physics_steps"], 1)
    nc, nefc, nit, v = c["contacts"] / n, c["efc_rows"] / n, c["newton_iters"] / n, c["verts_tested"] / n
    n_act = c.get("active_rows", 0.5 * c["efc_rows"]) / n
    return 13000 + 5 * v + 470 * nc + nit * (68 * nefc + 342 * n_act + 2600), dict(contacts=nc, efc_rows=nefc, active_rows=n_act, newton_iters=nit, verts_tested=v, ls_evals=c["ls_evals"] / n)


def run_ours(a):
    import torch
    import torch.distributed as dist

    from quadruped_gym_b200 import VecQuadrupedEnv, _lib
    from quadruped_gym_b200.envs import rewards as R

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG stays whatever the launcher set (the driver counts ranks from NCCL's INFO lines); NCCL prints to
        # stdout, which main() has pointed at stderr so that the real stdout stays the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    envs, fs, max_time, desc = WORKLOADS[a.workload]
    if a.envs:
        envs = a.envs
    rewards = {"forward": R.forward_velocity(1.0), "control_cost": R.ctrl_sq(-0.1), "alive_bonus": R.alive_bonus(1.0)}
    model_blob, random_init = None, False
    if a.workload == "c5":   # stress variant of the model: elliptic cone, position-servo gain x5
        from quadruped_gym_b200.model import DEFAULT_BLOB, blob as qblob
        A = qblob.unpack(open(DEFAULT_BLOB, "rb").read())
        A["opt_i"][1] = 1
        A["act_gain"] = A["act_gain"] * 5.0
        A["act_bias"] = A["act_bias"].reshape(-1, 3) * np.array([1.0, 5.0, 1.0])
        model_blob, random_init = qblob.pack(A), True
    env = VecQuadrupedEnv(envs, dev, max_time=max_time, frame_skip=fs, reward_fns=rewards,
                          termination_fns={"flip": R.flip_termination()}, use_default_termination=True,
                          auto_reset=True, seed=0, env_offset=rank * envs, model_blob=model_blob, random_init=random_init)
    env.reset()
    if a.workload == "c5":   # randomised initial poses: yaw U(0,2pi), tilt <= 30 deg, height U(0.05,0.2), joints U(range)
        g0 = torch.Generator(device=dev)
        g0.manual_seed(99 + rank)
        u = lambda *shape: torch.rand(shape, device=dev, generator=g0)
        q = env.data.qpos
        yaw, tilt, tdir = u(envs) * 2 * np.pi, u(envs) * np.pi / 6, u(envs) * 2 * np.pi
        ax = torch.stack([torch.cos(tdir), torch.sin(tdir), torch.zeros_like(tdir)], 1) * torch.sin(tilt / 2)[:, None]
        qt = torch.cat([torch.cos(tilt / 2)[:, None], ax], 1)
        qy = torch.stack([torch.cos(yaw / 2), torch.zeros_like(yaw), torch.zeros_like(yaw), torch.sin(yaw / 2)], 1)
        w1, x1, y1, z1 = qy.unbind(1); w2, x2, y2, z2 = qt.unbind(1)
        q[:, 3:7] = torch.stack([w1*w2 - x1*x2 - y1*y2 - z1*z2, w1*x2 + x1*w2 + y1*z2 - z1*y2,
                                 w1*y2 - x1*z2 + y1*w2 + z1*x2, w1*z2 + x1*y2 - y1*x2 + z1*w2], 1)
        q[:, 2] = 0.05 + 0.15 * u(envs)
        lo = torch.tensor(np.tile(np.deg2rad([-45, -45, -90]), 4), device=dev, dtype=torch.float32)
        hi = torch.tensor(np.tile(np.deg2rad([45, 120, 90]), 4), device=dev, dtype=torch.float32)
        q[:, 7:] = lo + (hi - lo) * u(envs, 12)
        env.set_state(qpos=q)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    pool = [torch.rand((envs, 12), device=dev, generator=gen) * 2 - 1 for _ in range(NPOOL)]  # resident in HBM
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed pre-roll into the steady state: every episode starts with a ~0.1 m drop (28 contact-free env steps), so
    # without it a short run would time free-fall instead of the contact-rich rollout the metric is about
    for i in range(PREROLL):
        env.step(pool[i % NPOOL])
    for i in range(max(a.warmup, 3)):
        env.step(pool[i % NPOOL])
    barrier()
    env.counters(reset=True)
    launches0 = _lib.lib().qg_launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t0 = time.perf_counter()
    rollout = None
    if a.workload == "c4":   # PPO rollout collection: obs/action/reward/done of a 24-step horizon stay in HBM as [T,N,.]
        from quadruped_gym_b200.rollout import RolloutBuffer
        rollout = RolloutBuffer(env, 24)
    for i in range(a.steps):
        flush.fill_(float(i))  # evict the state planes from L2 between timed steps
        ev[i][0].record()
        o, r, te, _, _ = env.step(pool[i % NPOOL])
        if rollout is not None:
            rollout.store(i % 24, pool[i % NPOOL], o, r, te)
        ev[i][1].record()
    barrier()
    t1 = time.perf_counter()
    launches = _lib.lib().qg_launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = float(sum(step_ms))
    ctr = env.counters(reset=True)

    # ---- end to end through the host-buffer C-ABI call (H2D + kernel + D2H inside the timed region)
    e2e_s = float("nan")
    if not a.profile:
        hpool_t = [p.cpu().pin_memory() for p in pool]   # this step's inputs live in pinned host memory
        hpool = [t.numpy() for t in hpool_t]
        for i in range(3):
            env.step_host(hpool[i % NPOOL])
        barrier()
        te0 = time.perf_counter()
        for i in range(a.steps):
            env.step_host(hpool[i % NPOOL])
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - te0
        barrier()

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    cvec = torch.tensor([float(ctr[k]) for k in sorted(ctr)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # timing: max over ranks
        dist.all_reduce(cvec, op=dist.ReduceOp.SUM)    # rollout statistics: the only collective on this path
    total_ms, e2e_ms = float(t[0]), float(t[1])
    ctr_all = {k: float(v) for k, v in zip(sorted(ctr), cvec.tolist())}
    units = envs * fs * a.steps * world
    value = units / (total_ms * 1e-3)
    if rank == 0:
        clocks = sampler.stop(t0, t1)
        fl, means = flops_per_physics_step(ctr_all)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        import ctypes
        tf = ctypes.c_double(0.0)
        if not a.profile:
            _lib.lib().qg_fp32_peak(local, 4096, ctypes.byref(tf))
        launch_ms = float(np.mean(step_ms))
        # dram bytes per launch: cannot be measured outside a profiler, so it is READ from the committed ncu capture of the
        # same kernel (and scaled to this batch size) and labelled as such; null when no capture is committed
        traffic, traffic_from = None, None
        for cand in ("r2_step_kernel_metrics.json", "r1_step_kernel_metrics.json"):
            try:
                pm = json.load(open(os.path.join(ROOT, "profiles", cand)))
                traffic = (pm["dram__bytes_read.sum"] + pm["dram__bytes_write.sum"]) / pm["envs"] * envs
                traffic_from = "profiles/" + cand + " (ncu --set full of qg_step_kernel, per launch, scaled to envs_per_gpu)"
                break
            except Exception:
                continue
        algo_bytes = 737.0 * envs  # 737 B per env.step() per env
        hbm_ach = algo_bytes / (launch_ms * 1e-3) / 1e9
        fp32_ach = fl * envs * fs / (launch_ms * 1e-3) / 1e12
        cores = os.cpu_count() or 1
        r1, rp = (float("nan"), float("nan")) if a.profile else (cpu_rate(1, 3000, fs, max_time)[0], cpu_rate(cores, 3000, fs, max_time)[0])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(a.workload, envs, fs, max_time, desc),
                           inputs="%d action tensors resident in HBM, cycled" % NPOOL,
                           l2="flushed (256 MiB write) between timed steps", preroll_steps=PREROLL,
                           parallelism=f"env-sharded x{world}, no data-path collective"),
            "clocks": clocks,
            "step_ms": {"mean": launch_ms, "median": float(np.median(step_ms)), "min": float(np.min(step_ms)), "max": float(np.max(step_ms)),
                        "std": float(np.std(step_ms)), "note": "per-step CUDA-event times of this rank over the timed steps"},
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": envs * 12 * 4,
                    "d2h_bytes_per_step": envs * (33 * 4 + 4 + 1)},
            "gpu_launches": int(launches),
            # the bounding roof of this path is FP32 CUDA-core issue (tiny per-env matrices, no dense contraction,
            # 0.5 % of HBM): `roofline` is that object, the HBM figure rides along as `roofline_hbm`
            "roofline": {"bound": "fp32-cuda-core", "achieved": fp32_ach, "peak": tf.value, "unit": "TFLOP/s",
                         "frac": fp32_ach / tf.value if tf.value else None, "traffic": traffic, "traffic_from": traffic_from,
                         "flops_per_physics_step": fl,
                         "peak_source": "qg_fp32_peak: FFMA dependent chains on every SM, measured in this run (nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4)",
                         "means_per_physics_step": means,
                         "note": "achieved = SURVEY App. E flop model evaluated with the kernel's own counters x env-substeps per launch / mean launch time (CUDA events)"},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                             "traffic": traffic, "traffic_from": traffic_from, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                             "algorithmic_bytes_per_env_step": 737},
            "cpu_baseline": {"value": rp, "unit": UNIT, "cores": cores, "kind": "port",
                             "single_core_value": r1,
                             "sample": f"{cores} processes x 1 env x 3000 env.step() (frame_skip {fs}), random actions; oracle port, not MuJoCo"},
            "counters": ctr_all,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL writes its version banner and its
    NCCL_DEBUG lines to stdout from C), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
    private duplicate of the original stdout.  NCCL_DEBUG is left as the launcher set it: its lines are in stderr."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def emit(line: dict):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--profile", action="store_true", help="ncu runs: skip the e2e, FP32-peak and CPU-baseline legs")
    a = ap.parse_args()
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
    finally:
        _close_pools()


if __name__ == "__main__":
    main()
