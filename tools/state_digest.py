"""Bitwise regression aid: sha256 over obs/reward/terminated of every step and the final state of a fixed seeded
rollout.  Two builds of libquadgym.so (QG_LIB selects the library) that are meant to compute the same thing must
print the same digest.  Usage: QG_LIB=... python tools/state_digest.py [n_envs] [steps] [cone]"""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
env = VecQuadrupedEnv(n, "cuda:0", termination_fns={"flip": R.flip_termination()}, auto_reset=True, max_time=1.0,
                      reward_fns={"alive": R.alive_bonus(), "ctrl": R.control_cost()})
env.reset(seed=3)
g = torch.Generator(device="cuda")
g.manual_seed(1)
h = hashlib.sha256()
for i in range(steps):
    a = torch.rand((n, 12), device="cuda", generator=g) * 2.4 - 1.2
    obs, rew, term, trunc, info = env.step(a)
    if i % 10 == 9 or i == steps - 1:
        h.update(obs.cpu().numpy().tobytes())
        h.update(rew.cpu().numpy().tobytes())
        h.update(term.cpu().numpy().tobytes())
d = env.data
for t in (d.qpos, d.qvel, d.act, d.ctrl, d.time, d.qacc_warmstart):
    h.update(t.cpu().numpy().tobytes())
if os.environ.get("QG_DUMP"):   # numerical comparison of two builds (fp-level differences are expected, others are not)
    import numpy as np
    np.savez(os.environ["QG_DUMP"], qpos=d.qpos.cpu().numpy(), qvel=d.qvel.cpu().numpy(), obs=obs.cpu().numpy())
c = env.counters()
print("digest", h.hexdigest()[:32], "contacts/step %.3f verts %.2f" % (c["contacts"] / c["physics_steps"], c["verts_tested"] / c["physics_steps"]))
