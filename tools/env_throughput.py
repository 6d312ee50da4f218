"""Throughput of the task environments (walking reward stack: 3 launches per step; PO observation: 4 launches)."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200.envs.walking_quad import VecWalkingQuadrupedEnv
from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv
from quadruped_gym_b200.envs.quadruped import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R
from quadruped_gym_b200.envs.sb3 import SB3VecEnv
opts = {"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3}
def run(cls, n, **kw):
    if cls is VecQuadrupedEnv:   # the physics launch alone, same termination rules as the walking env
        env = cls(n, "cuda:0", max_time=20, frame_skip=10, termination_fns={"flip": R.flip_termination()}, **kw)
    else:
        env = cls(n, "cuda:0", max_time=20, frame_skip=10, random_controls=True, reset_options=opts, **kw)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = [torch.rand((n, 12), device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    for i in range(60): env.step(acts[i % 8])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 100
    for i in range(K): env.step(acts[i % 8])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{cls.__name__:28s} N={n:6d} frame_skip=10 {kw}: {ms:.3f} ms/step -> {n*10/(ms*1e-3):.3e} physics env-steps/s")
    env.close()
for n in (4096, 65536):
    run(VecQuadrupedEnv, n)
    run(VecWalkingQuadrupedEnv, n)
    run(VecPOWalkingQuadrupedEnv, n, obs_window=10)
# the training configuration of the reference (train_quadruped.py:15-22,49) through the SB3 VecEnv, numpy in/out, with the
# script's RewardCallback access pattern (np.mean([info[key] for info in infos]) for the 11 keys, train_quadruped.py:86-92)
for n, cb, cp in ((10, True, True), (1024, True, True), (8192, True, True), (8192, False, True), (65536, False, True), (65536, False, False)):
    sb3 = SB3VecEnv(VecPOWalkingQuadrupedEnv(n, "cuda:0", max_time=20, frame_skip=10, obs_window=10, random_controls=True, reset_options=opts), copy=cp)
    sb3.reset(); rng = np.random.default_rng(0)
    acts = [rng.uniform(-1, 1, (n, 12)).astype(np.float32) for _ in range(4)]
    for i in range(10): sb3.step(acts[i % 4])
    t0 = time.perf_counter(); K = 50
    for i in range(K):
        o, r, d, infos = sb3.step(acts[i % 4])
        if cb:
            comps = {key: np.mean([info[key] for info in infos]) for key in sb3.reward_keys}
    dt = (time.perf_counter() - t0) / K
    print(f"SB3VecEnv(PO, obs_window 10) N={n:5d} {'with' if cb else 'without'} the RewardCallback loop, copy={cp}: {dt*1e3:.3f} ms/step wall (numpy in/out) -> {n*10/dt:.3e} physics env-steps/s")
    sb3.close()
