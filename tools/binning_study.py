"""How well does last-step solver effort predict this step's?  The solver loop is block-synchronised: a block of 64
environments runs max-over-64 Newton iterations, so the slot -> env permutation should put environments with equal
iteration counts together.  Two consecutive physics steps (new random control for the second); environments are sorted by
a key built from the FIRST step's counters and the mean block-max of the SECOND step's counters is reported, against no
sorting and against the unattainable sort by the second step's own count."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R
n = 65536
os.environ["QG_BINNING"] = "0"
env = VecQuadrupedEnv(n, "cuda:0", termination_fns={"flip": R.flip_termination()}, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for i in range(150):
    env.step(torch.rand((n, 12), device="cuda", generator=g) * 2 - 1)
a1 = torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
c1 = env.debug_step(a1)["counts"].cpu().long()
same = env.debug_step(a1)["counts"].cpu().long()          # same control held (substeps 2..4 of an env.step)
c3 = env.debug_step(torch.rand((n, 12), device="cuda", generator=g) * 2 - 1)["counts"].cpu().long()   # new control (substep 1)
def report(tag, prev, cur):
    nit, nls = cur[:, 2].float(), cur[:, 3].float()
    keys = {"none (env order)": torch.arange(n), "ncon": prev[:, 0], "niter": prev[:, 2], "nls": prev[:, 3],
            "min(ncon,3)*16+min(nls,15) (~current)": prev[:, 0].clamp(max=3) * 16 + prev[:, 3].clamp(max=15),
            "niter*32+min(nls,31)": prev[:, 2] * 32 + prev[:, 3].clamp(max=31), "nls*32+ncon": prev[:, 3] * 32 + prev[:, 0].clamp(max=31),
            "oracle: this step's niter*32+nls": cur[:, 2] * 32 + cur[:, 3].clamp(max=31)}
    print(f"{tag}: niter mean {nit.mean():.2f}, nls mean {nls.mean():.2f}; corr(prev niter, niter) {torch.corrcoef(torch.stack([prev[:,2].float(), nit]))[0,1]:.2f}")
    for k, v in keys.items():
        order = torch.argsort(v, stable=True)
        bn, bl = nit[order].view(-1, 64).max(1).values.mean(), nls[order].view(-1, 64).max(1).values.mean()
        wn = nit[order].view(-1, 8).max(1).values.mean()
        print(f"   sorted by {k:42s}: mean block-max niter {bn:.2f}  nls {bl:.2f}   (warp-max niter {wn:.2f})")
report("next substep, SAME control", c1, same)
report("next substep, NEW control ", same, c3)
