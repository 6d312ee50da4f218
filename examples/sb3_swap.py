"""How the reference's training script swaps its SubprocVecEnv for the device-resident environments.

/root/reference/src/train_quadruped.py:49-50 builds
    env = SubprocVecEnv([lambda: make_env(options) for _ in range(num_envs)])
with make_env = POWalkingQuadrupedEnv(max_time=20, frame_skip=10, obs_window=10, random_controls=True, reset_options=options).
The drop-in replacement (numpy in / numpy out, same-step auto-reset, the 11 reward keys in every info dict):
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quadruped_gym_b200 import SB3VecEnv, VecPOWalkingQuadrupedEnv

options = {"fixed_heading_angle": 0.0, "fixed_velocity_angle": 0.0, "fixed_speed": 0.3}      # train_quadruped.py:40-46
num_envs = 1024                                                                                # 10 in the reference
env = SB3VecEnv(VecPOWalkingQuadrupedEnv(num_envs, "cuda:0", max_time=20, frame_skip=10, obs_window=10,
                                        random_controls=True, reset_options=options))
# model = PPO("MlpPolicy", env, policy_kwargs={"net_arch": [256, 256, 128], "activation_fn": torch.nn.Tanh})   # as :52-58
obs = env.reset()
for _ in range(100):
    obs, rewards, dones, infos = env.step(np.random.uniform(-1, 1, (num_envs, 12)))
print(obs.shape, rewards.mean(), dones.sum(), sorted(infos[0])[:4], "...")
env.close()
