"""CPU oracle for the rollout path: GAE.  TEST INFRASTRUCTURE ONLY.

Restates ``stable_baselines3.common.buffers.RolloutBuffer.compute_returns_and_advantage``
(SB3 2.3.2 -- third-party, not in the reference tree; the reference reaches it through
``PPO(...).learn`` at ``scripts/train.py:119-144`` with ``gamma`` / ``gae_lambda`` from
``config.yaml:100-104``) in numpy float64.  Parity status: unpinned (no SB3 here)."""
from __future__ import annotations

import numpy as np


def compute_returns_and_advantage(rewards, values, episode_starts, last_values, dones, gamma, gae_lambda):
    rewards = np.asarray(rewards, np.float64)
    values = np.asarray(values, np.float64)
    T = rewards.shape[0]
    advantages = np.zeros_like(rewards)
    last_gae_lam = np.zeros(rewards.shape[1])
    for step in reversed(range(T)):
        if step == T - 1:
            next_non_terminal = 1.0 - np.asarray(dones, np.float64)
            next_values = np.asarray(last_values, np.float64)
        else:
            next_non_terminal = 1.0 - np.asarray(episode_starts[step + 1], np.float64)
            next_values = values[step + 1]
        delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
        last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
        advantages[step] = last_gae_lam
    return advantages, advantages + values
