"""Experiment: step time vs solver iteration caps; iteration-count histogram."""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200 import VecQuadrupedEnv, _lib
from quadruped_gym_b200.envs import rewards as R
n = 65536
def run(mi, li):
    env = VecQuadrupedEnv(n, "cuda:0", termination_fns={"flip": R.flip_termination()}, auto_reset=True)
    env.reset(); env._sync_tables()
    _lib.check(_lib.lib().qg_set_options(env._batch, 10.0, 1, 1, mi, li))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for i in range(80):
        if i % 5 == 0: a = torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
        env.step(a)
    env.counters(reset=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        if i % 5 == 0: a = torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
        env.step(a)
    e1.record(); torch.cuda.synchronize()
    c = env.counters(); ps = c["physics_steps"]
    print(f"max_iter {mi:3d} ls_iter {li:3d}: {e0.elapsed_time(e1)/40:.3f} ms/step", {k: round(v / ps, 3) for k, v in c.items() if k in ("contacts", "newton_iters", "ls_evals")})
    return env
for mi, li in ((20, 12), (8, 6), (4, 4), (2, 2), (1, 1)):
    env = run(mi, li)
    if (mi, li) != (1, 1): env.close()
# histogram of iterations / evals per env for one debug step
_lib.check(_lib.lib().qg_set_options(env._batch, 10.0, 1, 1, 20, 12))
out = env.debug_step(torch.rand((n, 12), device="cuda") * 2 - 1)
cnt = out["counts"].cpu()
for name, col in (("ncon", 0), ("nefc", 1), ("niter", 2), ("nls", 3)):
    v = cnt[:, col]
    print(name, "mean %.2f" % v.float().mean(), "hist", torch.bincount(v.clamp(max=30))[:31].tolist())
