"""Throughput of the task environments (walking reward stack: 3 launches per step; PO observation: 4 launches)."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200.envs.walking_quad import VecWalkingQuadrupedEnv
from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv, SB3VecEnvAdapter
opts = {"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3}
def run(cls, n, **kw):
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
    run(VecWalkingQuadrupedEnv, n)
    run(VecPOWalkingQuadrupedEnv, n, obs_window=10)
# the training configuration of the reference (train_quadruped.py:15-22,49) through the SB3 adapter, numpy in/out
for n in (10, 1024):
    sb3 = SB3VecEnvAdapter(VecPOWalkingQuadrupedEnv(n, "cuda:0", max_time=20, frame_skip=10, obs_window=10, random_controls=True, reset_options=opts))
    sb3.reset(); rng = np.random.default_rng(0)
    for _ in range(20): sb3.step(rng.uniform(-1, 1, (n, 12)))
    t0 = time.perf_counter(); K = 200
    for _ in range(K): sb3.step(rng.uniform(-1, 1, (n, 12)))
    dt = (time.perf_counter() - t0) / K
    print(f"SB3VecEnvAdapter(PO, obs_window 10) N={n:5d}: {dt*1e3:.3f} ms/step wall (numpy in/out, info dicts) -> {n*10/dt:.3e} physics env-steps/s")
    sb3.close()
