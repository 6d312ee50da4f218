"""Experiment: how much longer than the average environment does the slowest environment of a warp (8 envs) / block
(64 envs) iterate?  The solver loop is block-synchronised, so a block runs max-over-64 Newton iterations."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R
n = 65536
env = VecQuadrupedEnv(n, "cuda:0", termination_fns={"flip": R.flip_termination()}, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for i in range(150):
    env.step(torch.rand((n, 12), device="cuda", generator=g) * 2 - 1)
cnt = env.debug_step(torch.rand((n, 12), device="cuda", generator=g) * 2 - 1)["counts"].cpu().float()
for name, col in (("ncon", 0), ("nefc", 1), ("niter", 2), ("nls", 3)):
    v = cnt[:, col]
    print("%-6s mean %.2f  mean of warp-max %.2f  mean of block-max %.2f  max %d  hist %s" % (
        name, v.mean(), v.view(-1, 8).max(1).values.mean(), v.view(-1, 64).max(1).values.mean(), int(v.max()),
        torch.bincount(v.long().clamp(max=24))[:25].tolist()))
