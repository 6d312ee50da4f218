"""Random-action rollout of N environments on one B200 (BASELINE configs 2/3), everything on the device.

    python examples/random_rollout.py [N]
"""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200 import VecQuadrupedEnv, rewards as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = VecQuadrupedEnv(n, "cuda:0", frame_skip=4, max_time=10.0,
                      reward_fns={"forward": R.forward_velocity(1.0), "control_cost": R.ctrl_sq(-0.1), "alive_bonus": R.alive_bonus(1.0)},
                      termination_fns={"flip": R.flip_termination()})          # + default time limit, SB3-style auto-reset
obs, info = env.reset()
torch.cuda.synchronize()
t0, ret = time.perf_counter(), torch.zeros(n, device="cuda")
for t in range(500):
    obs, reward, terminated, truncated, info = env.step(torch.rand(n, 12, device="cuda") * 2 - 1)
    ret += reward
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{n} envs x 500 steps x frame_skip 4: {n * 500 * 4 / dt:.3e} physics env-steps/s wall; mean return {ret.mean().item():.2f}; counters {env.counters()}")
