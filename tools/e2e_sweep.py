"""End-to-end (host buffers in / out) step time of qg_step_host against the device-timed step, for several segment
counts of the pipelined host path (QG_HOST_SEGMENTS).  One process per setting (the variable is read at first use)."""
import os, subprocess, sys, json, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from quadruped_gym_b200 import VecQuadrupedEnv
    from quadruped_gym_b200.envs import rewards as R
    n = int(os.environ.get("N", 65536))
    env = VecQuadrupedEnv(n, "cuda:0", auto_reset=True, termination_fns={"flip": R.flip_termination()},
                          reward_fns={"forward": R.forward_velocity(1.0), "control_cost": R.ctrl_sq(-0.1), "alive_bonus": R.alive_bonus(1.0)})
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    pool = [torch.rand((n, 12), device="cuda", generator=g) * 2 - 1 for _ in range(16)]
    hpool = [p.cpu().pin_memory().numpy() for p in pool]
    for i in range(130):
        env.step(pool[i % 16])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60):
        env.step(pool[i % 16])
    e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 60
    for i in range(10):
        env.step_host(hpool[i % 16])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(60):
        env.step_host(hpool[i % 16])
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3 / 60
    print(f"segments {os.environ.get('QG_HOST_SEGMENTS', 'default'):>7}: device {dev_ms:.3f} ms/step (no L2 flush), host-buffer path {host_ms:.3f} ms/step, gap {100 * (host_ms / dev_ms - 1):.1f} %")
else:
    for seg in sys.argv[1:] or ["1", "2", "4", "6", "8", "12"]:
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, QG_HOST_SEGMENTS=seg))
