"""Workload for the ncu capture of the NON-physics kernels (walking reward stack, PO observation, masked reset, binning):
VecPOWalkingQuadrupedEnv, 65,536 envs, frame_skip 10, obs_window 10 (the reference's training configuration,
train_quadruped.py:15-22), random actions.  Run under
    ncu --set full --clock-control none -k regex:'qg_(walk|po|reset|bin)' -s 100 -c 10 -o gpurun_out/r2_env_kernels python tools/profile_env_kernels.py
"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv
n = int(os.environ.get("N", 65536))
env = VecPOWalkingQuadrupedEnv(n, "cuda:0", max_time=20, frame_skip=10, obs_window=10, random_controls=True,
                               reset_options={"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3})
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [torch.rand((n, 12), device="cuda", generator=g) * 2 - 1 for _ in range(4)]
for i in range(int(os.environ.get("STEPS", 30))):
    env.step(acts[i % 4])
torch.cuda.synchronize()
env.close()
