"""Device-resident rollout collection (BASELINE config 4: PPO rollout collection, 8,192 envs x 24-step horizon).

The reference collects rollouts through SB3's ``collect_rollouts`` over ``SubprocVecEnv`` pipes
(/root/reference/src/train_quadruped.py:49-58,132-134).  Here observations, actions, rewards and done flags of a
whole horizon stay in HBM as ``[T, N, .]`` tensors; nothing is copied to the host.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class RolloutBuffer:
    def __init__(self, env, horizon: int):
        n, dev = env.num_envs, env.device
        d = env.observation_space.shape[0]
        self.env, self.horizon = env, int(horizon)
        self.obs = torch.zeros((horizon + 1, n, d), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((horizon, n, 12), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((horizon, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((horizon, n), dtype=torch.bool, device=dev)
        self._last_obs: Optional[torch.Tensor] = None

    def store(self, t: int, action, obs, reward, done):
        """Slot t of the horizon: the transition produced by one ``env.step(action)`` (device-to-device copies only)."""
        self.actions[t].copy_(action)
        self.rewards[t].copy_(reward)
        self.dones[t].copy_(done)
        self.obs[t + 1].copy_(obs)

    def collect(self, policy: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, generator=None):
        """Run ``horizon`` env steps; ``policy(obs) -> action [N,12]`` (default: U(-1,1) random actions)."""
        env = self.env
        if self._last_obs is None:
            self._last_obs, _ = env.reset()
        self.obs[0].copy_(self._last_obs)
        for t in range(self.horizon):
            if policy is None:
                a = torch.rand((env.num_envs, 12), device=env.device, generator=generator) * 2 - 1
            else:
                a = policy(self.obs[t])
            obs, rew, term, trunc, _ = env.step(a)
            self.store(t, a, obs, rew, term)
        self._last_obs = self.obs[self.horizon]
        return self

    def stats(self) -> dict:
        """Small rollout-statistics vector (what `sharding.reduce_rollout_stats` all-reduces across GPUs)."""
        return {"env_steps": float(self.rewards.numel()), "reward_sum": float(self.rewards.sum()),
                "episodes": float(self.dones.sum())}
