"""Experiment: step time when every env gets the same actions (no inter-env divergence) vs random actions."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R
n = 65536
for mode in os.environ.get("MODES", "same,random,quad_same").split(","):
    env = VecQuadrupedEnv(n, "cuda:0", termination_fns={"flip": R.flip_termination()}, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    def act():
        if mode == "same":
            return (torch.rand((1, 12), device="cuda", generator=g) * 2 - 1).expand(n, 12).contiguous()
        if mode == "quad_same":   # all 8 envs of a warp identical, warps differ
            return (torch.rand((n // 8, 1, 12), device="cuda", generator=g) * 2 - 1).expand(n // 8, 8, 12).reshape(n, 12).contiguous()
        return torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
    for i in range(80):
        if i % 5 == 0: a = act()
        env.step(a)
    env.counters(reset=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        if i % 5 == 0: a = act()
        env.step(a)
    e1.record(); torch.cuda.synchronize()
    c = env.counters()
    ps = c["physics_steps"]
    print(mode, "%.3f ms/step" % (e0.elapsed_time(e1) / 40), {k: round(v / ps, 2) for k, v in c.items() if k in ("contacts", "newton_iters", "ls_evals", "verts_tested")})
    env.close()
