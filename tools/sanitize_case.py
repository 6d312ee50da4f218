"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the library on ragged batch
sizes -- base env with auto-reset and binning, elliptic model, host-buffer path with 3 segments, walking and PO envs.
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("QG_BINNING", "1")
os.environ.setdefault("QG_HOST_SEGMENTS", "3")
from quadruped_gym_b200 import VecQuadrupedEnv
from quadruped_gym_b200.envs import rewards as R
from quadruped_gym_b200.envs.po_walking_quad import VecPOWalkingQuadrupedEnv
from quadruped_gym_b200.model import DEFAULT_BLOB, blob as qblob

steps = int(os.environ.get("STEPS", 40))
rng = np.random.default_rng(0)
for n, ell in ((333, False), (200, True)):
    mb = None
    if ell:
        A = qblob.unpack(open(DEFAULT_BLOB, "rb").read()); A["opt_i"][1] = 1; mb = qblob.pack(A)
    env = VecQuadrupedEnv(n, "cuda:0", auto_reset=True, max_time=0.2, model_blob=mb, random_init=True, termination_fns={"flip": R.flip_termination()},
                          reward_fns={"forward": R.forward_velocity(1.0), "cc": R.control_cost(-2.0, 0.8), "alive": R.alive_bonus(1.0)})
    env.reset()
    for t in range(steps):
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        if t % 3 == 2:
            env.step_host(a, want_terms=True, want_terminal_obs=True)
        else:
            env.step(torch.from_numpy(a).cuda())
    env.debug_step(np.zeros((n, 12), np.float32))
    print("base env", n, "elliptic" if ell else "pyramidal", env.counters())
    env.close()
env = VecPOWalkingQuadrupedEnv(77, "cuda:0", obs_window=4, max_time=0.3, frame_skip=10, random_controls=True, random_init=True)
env.reset()
for t in range(steps):
    env.step(torch.from_numpy(rng.uniform(-1, 1, (77, 12)).astype(np.float32)).cuda())
env.reset(mask=torch.arange(77) % 2 == 0)
torch.cuda.synchronize()
print("po env ok")
env.close()
