"""CPU oracle for VecNormalize.  TEST INFRASTRUCTURE ONLY.

The reference wraps its vectorised environment in ``stable_baselines3.common.vec_env.
VecNormalize(norm_obs=True, norm_reward=True, clip_obs=10.0, gamma=GAMMA)``
(``scripts/train.py:86-88``; re-loaded for inference at ``backend/rl/agent.py:73-80``).
stable-baselines3 is a third-party dependency that is NOT in the reference tree and not
installed here (``requirements.txt:7`` leaves it unpinned; the shipped checkpoints record
2.3.2 in ``assets/model/v3/best_model.zip:system_info.txt``).  This file restates the
published algorithm of that version -- ``common/running_mean_std.py`` (Chan et al. parallel
moments) and ``common/vec_env/vec_normalize.py`` (``step_wait``, ``reset``, ``normalize_obs``,
``normalize_reward``, ``_update_reward``) -- in numpy float64.

Parity status: **unpinned** (no copy of SB3 to run against, and the reference has no test that
touches it); the anchor is the reference's call site and SB3's
documented behaviour.
"""
from __future__ import annotations

import numpy as np


class RunningMeanStd:
    """SB3 ``RunningMeanStd(epsilon=1e-4, shape)``: mean 0, var 1, count epsilon; ``update``
    folds a batch in with the parallel-variance formula."""

    def __init__(self, epsilon: float = 1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = float(epsilon)

    def update(self, arr: np.ndarray) -> None:
        batch_mean = np.mean(arr, axis=0)
        batch_var = np.var(arr, axis=0)
        self.update_from_moments(batch_mean, batch_var, arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count) -> None:
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / tot_count
        self.mean, self.var, self.count = new_mean, m_2 / tot_count, tot_count


class VecNormalizeOracle:
    """``VecNormalize`` over a stream of (obs, reward, done) batches supplied by the caller."""

    def __init__(self, num_envs: int, obs_dim: int = 8, training: bool = True, norm_obs: bool = True,
                 norm_reward: bool = True, clip_obs: float = 10.0, clip_reward: float = 10.0,
                 gamma: float = 0.99, epsilon: float = 1e-8):
        self.obs_rms = RunningMeanStd(shape=(obs_dim,))
        self.ret_rms = RunningMeanStd(shape=())
        self.returns = np.zeros(num_envs, np.float64)
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon

    def normalize_obs(self, obs: np.ndarray) -> np.ndarray:
        if not self.norm_obs:
            return obs.astype(np.float32)
        out = np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon),
                      -self.clip_obs, self.clip_obs)
        return out.astype(np.float32)

    def normalize_reward(self, reward: np.ndarray) -> np.ndarray:
        if not self.norm_reward:
            return reward.astype(np.float32)
        out = np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        return out.astype(np.float32)

    def reset(self, obs: np.ndarray) -> np.ndarray:
        self.returns = np.zeros_like(self.returns)
        if self.training and self.norm_obs:
            self.obs_rms.update(obs.astype(np.float64))
        return self.normalize_obs(obs)

    def step(self, obs, reward, done, terminal_obs=None):
        """``step_wait``: update obs statistics, normalise obs, update the discounted return and
        its statistics, normalise the reward, normalise terminal observations of done envs,
        zero the return of done envs.  Returns (obs_n, reward_n, terminal_obs_n or None)."""
        obs64 = obs.astype(np.float64)
        if self.training and self.norm_obs:
            self.obs_rms.update(obs64)
        obs_n = self.normalize_obs(obs)
        if self.training:
            self.returns = self.returns * self.gamma + reward.astype(np.float64)
            self.ret_rms.update(self.returns)
        reward_n = self.normalize_reward(reward)
        tobs_n = None
        if terminal_obs is not None:
            tobs_n = terminal_obs.astype(np.float32).copy()
            tobs_n[done] = self.normalize_obs(terminal_obs[done])
        self.returns[done] = 0.0
        return obs_n, reward_n, tobs_n
