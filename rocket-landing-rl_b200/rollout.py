"""Device-resident PPO rollout collection (BASELINE.json configs[4]).

The reference collects rollouts with stable-baselines3: ``PPO("MlpPolicy", ...,
policy_kwargs=dict(net_arch=dict(pi=[256, 256], vf=[256, 256])))`` (``scripts/train.py:117-134``)
steps ``SubprocVecEnv`` workers over pipes and fills a host ``RolloutBuffer``.  Here the
environments, the (optional) ``VecNormalize`` statistics, the policy and the buffer all live on one
GPU: nothing crosses PCIe during collection.

* ``ActorCriticMLP`` is the SB3 ``MlpPolicy`` the shipped checkpoints use: separate tanh MLPs
  8 -> 256 -> 256 for the policy and the value function, a linear action-mean head, a
  state-independent log-std, a linear value head -- 136 965 parameters, as
  ``assets/model/v3/best_model.zip``.  It is plain PyTorch (cuBLAS
  GEMMs): the MLP is a neighbour of the hot path, not the path.
* ``FusedActorCritic`` evaluates the same two nets with the hand-written tcgen05 / TMEM kernel of
  ``csrc/rl_policy.cu``: one launch per net, the 256-wide activations never leave the SM, bf16
  weights with fp32 accumulation, and SB3's action sampling (Gaussian noise from Philox, log-prob,
  Box clipping) fused into the policy net's epilogue.  Same ``forward`` / ``predict_values``
  surface, so ``collect_rollout`` takes either.
* ``collect_rollout`` mirrors ``OnPolicyAlgorithm.collect_rollouts``: sample, clip the action to
  the Box (SB3 clips before ``env.step``), step, bootstrap time-limit truncations with the value
  of the terminal observation, store, and finally compute GAE with the ``rl_gae`` CUDA kernel.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import _lib


class ActorCriticMLP(nn.Module):
    def __init__(self, obs_dim: int = 8, act_dim: int = 2, hidden=(256, 256), log_std_init: float = 0.0):
        super().__init__()

        def mlp():
            layers, d = [], obs_dim
            for h in hidden:
                layers += [nn.Linear(d, h), nn.Tanh()]
                d = h
            return nn.Sequential(*layers)

        self.policy_net, self.value_net_body = mlp(), mlp()
        self.action_net = nn.Linear(hidden[-1], act_dim)
        self.value_net = nn.Linear(hidden[-1], 1)
        self.log_std = nn.Parameter(torch.full((act_dim,), float(log_std_init)))

    def forward(self, obs: torch.Tensor, generator: Optional[torch.Generator] = None):
        """actions ~ N(mean, exp(log_std)), values, log-probabilities (SB3 ``policy.forward``)."""
        mean = self.action_net(self.policy_net(obs))
        values = self.value_net(self.value_net_body(obs)).squeeze(-1)
        std = self.log_std.exp()
        noise = torch.randn(mean.shape, device=mean.device, dtype=mean.dtype, generator=generator)
        actions = mean + std * noise
        log_prob = (-0.5 * noise.pow(2) - self.log_std - 0.5 * math.log(2.0 * math.pi)).sum(-1)
        return actions, values, log_prob

    def predict_values(self, obs: torch.Tensor) -> torch.Tensor:
        return self.value_net(self.value_net_body(obs)).squeeze(-1)

    # ---- the reference's shipped checkpoints (assets/model/v*/best_model.zip) ------------------
    _SB3_KEYS = {"mlp_extractor.policy_net.0": "policy_net.0", "mlp_extractor.policy_net.2": "policy_net.2",
                 "mlp_extractor.value_net.0": "value_net_body.0", "mlp_extractor.value_net.2": "value_net_body.2",
                 "action_net": "action_net", "value_net": "value_net"}

    @classmethod
    def from_sb3_state_dict(cls, sd) -> "ActorCriticMLP":
        """Build the module from a stable-baselines3 ``MlpPolicy`` state dict (tensors or arrays):
        ``mlp_extractor.{policy,value}_net.{0,2}``, ``action_net``, ``value_net``, ``log_std`` -- the
        contents of ``policy.pth`` inside the zip ``PPO.save`` writes (``scripts/train.py:140-144``)."""
        w0, w2 = sd["mlp_extractor.policy_net.0.weight"], sd["mlp_extractor.policy_net.2.weight"]
        net = cls(obs_dim=int(w0.shape[1]), act_dim=int(sd["action_net.weight"].shape[0]),
                  hidden=(int(w0.shape[0]), int(w2.shape[0])))        # v2: 128-128, v3: 256-256
        own = net.state_dict()
        for src, dst in cls._SB3_KEYS.items():
            for leaf in ("weight", "bias"):
                t = torch.as_tensor(sd[f"{src}.{leaf}"], dtype=torch.float32)
                if t.shape != own[f"{dst}.{leaf}"].shape:
                    raise ValueError(f"{src}.{leaf}: shape {tuple(t.shape)}, expected {tuple(own[dst + '.' + leaf].shape)}")
                own[f"{dst}.{leaf}"] = t
        own["log_std"] = torch.as_tensor(sd["log_std"], dtype=torch.float32)
        net.load_state_dict(own)
        return net

    @classmethod
    def from_sb3_zip(cls, path: str) -> "ActorCriticMLP":
        """``PPO.load``'s policy without stable-baselines3: read ``policy.pth`` out of the zip."""
        import io
        import zipfile
        with zipfile.ZipFile(path) as z:
            sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        return cls.from_sb3_state_dict(sd)


MLP_BLOB_BYTES = 143376          # include/rocket_landing_b200.h: RL_MLP_BLOB_BYTES


def pack_mlp(body: nn.Sequential, head: nn.Linear, device) -> torch.Tensor:
    """One tanh MLP 8 -> 256 -> 256 -> head (1 or 2 outputs) in the layout ``rl_mlp_forward`` reads
    (``include/rocket_landing_b200.h``): bf16 W1 stacked twice, f32 b1, bf16 W2 (tiled), f32 b2, f32 W3 [2][256],
    f32 b3 [4]."""
    l1, l2 = body[0], body[2]
    if (l1.weight.shape, l2.weight.shape, head.weight.shape[1]) != ((256, 8), (256, 256), 256) or head.weight.shape[0] not in (1, 2):
        raise ValueError("rl_mlp_forward evaluates 8 -> 256 -> 256 -> {1, 2} nets only")
    blob = torch.zeros(MLP_BLOB_BYTES, dtype=torch.uint8, device=device)
    off = 0

    def put(t: torch.Tensor) -> None:
        nonlocal off
        t = t.detach().to(device).contiguous()
        nbytes = t.numel() * t.element_size()
        blob[off:off + nbytes].view(t.dtype).copy_(t.reshape(-1))
        off += nbytes

    put(torch.cat([l1.weight, l1.weight], dim=1).to(torch.bfloat16))           # [256][16]
    put(l1.bias.float())
    # W2 in the tensor core's K-major tile order (the image of the kernel's shared-memory operand, so that one
    # bulk copy stages it): element (n, k) of nn.Linear.weight at [k // 8][n][k % 8]
    put(l2.weight.detach().to(torch.bfloat16).view(256, 32, 8).permute(1, 0, 2))   # [32][256][8]
    put(l2.bias.float())
    w3 = torch.zeros((2, 256), dtype=torch.float32, device=head.weight.device)
    w3[:head.weight.shape[0]] = head.weight.detach().float()
    put(w3)
    b3 = torch.zeros(4, dtype=torch.float32, device=head.bias.device)
    b3[:head.bias.shape[0]] = head.bias.detach().float()
    put(b3)
    assert off == MLP_BLOB_BYTES
    return blob


class FusedActorCritic:
    """``ActorCriticMLP`` evaluated by the fused tensor-core kernel (inference only: rollout
    collection does not need gradients).  Call ``refresh()`` after the wrapped module's parameters
    change (an optimiser step) to re-pack the weights."""

    def __init__(self, net: ActorCriticMLP, device="cuda", seed: int = 0, low=(0.0, -1.0), high=(1.0, 1.0),
                 row_offset: int = 0):
        import ctypes as C
        self._L = _lib.load()
        self.net = net
        self.device = torch.device(device)
        self.seed = int(seed) & (2**64 - 1)
        self.draw = 0                                   # Philox counter word: one per forward() call
        self.row_offset = int(row_offset)               # global id of row 0 (the shard's env_id_offset)
        self._low = (C.c_float * 2)(*low)               # the env's Box (lander.py:65-69)
        self._high = (C.c_float * 2)(*high)
        assert int(self._L.rl_mlp_blob_bytes()) == MLP_BLOB_BYTES
        self.refresh()

    def refresh(self) -> None:
        self._pi = pack_mlp(self.net.policy_net, self.net.action_net, self.device)
        self._vf = pack_mlp(self.net.value_net_body, self.net.value_net, self.device)
        self._log_std = self.net.log_std.detach().to(self.device, torch.float32).clone()

    def _run(self, blob: torch.Tensor, obs: torch.Tensor, out_dim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if obs.dtype != torch.float32 or not obs.is_contiguous():
            obs = obs.to(torch.float32).contiguous()
        n = obs.shape[0]
        if out is None:
            out = torch.empty((n, out_dim), dtype=torch.float32, device=obs.device)
        assert out.is_contiguous() and out.numel() == n * out_dim and out.dtype == torch.float32
        if n == 0:
            return out
        rc = self._L.rl_mlp_forward(blob.data_ptr(), obs.data_ptr(), out.data_ptr(), n, out_dim,
                                    torch.cuda.current_stream(obs.device).cuda_stream)
        if rc != 0:
            raise _lib.RocketLibraryError(f"rl_mlp_forward failed ({rc}): {self._L.rl_policy_last_error().decode()}")
        return out

    def action_mean(self, obs: torch.Tensor) -> torch.Tensor:
        return self._run(self._pi, obs, 2)

    def predict_values(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Values ``[n]``; ``out`` (a contiguous float32 ``[n]`` tensor, e.g. a slice of the rollout
        buffer) is written in place when given."""
        return self._run(self._vf, obs, 1, out).reshape(-1)

    def act(self, obs: torch.Tensor, actions: Optional[torch.Tensor] = None,
            log_prob: Optional[torch.Tensor] = None, want_mean: bool = False):
        """Sample with the fused epilogue (``rl_policy_sample``).  Returns ``(actions, clipped,
        log_prob, mean or None)``; ``actions`` / ``log_prob`` may be given (slices of the rollout
        buffer) and are then written in place."""
        if obs.dtype != torch.float32 or not obs.is_contiguous():
            obs = obs.to(torch.float32).contiguous()
        n = obs.shape[0]
        f32 = dict(dtype=torch.float32, device=obs.device)
        actions = torch.empty((n, 2), **f32) if actions is None else actions
        log_prob = torch.empty((n,), **f32) if log_prob is None else log_prob
        clipped = torch.empty((n, 2), **f32)
        mean = torch.empty((n, 2), **f32) if want_mean else None
        assert actions.is_contiguous() and log_prob.is_contiguous()
        rc = self._L.rl_policy_sample(self._pi.data_ptr(), obs.data_ptr(), self._log_std.data_ptr(), self.seed, self.draw,
                                      self.row_offset, self._low, self._high, actions.data_ptr(), clipped.data_ptr(), log_prob.data_ptr(),
                                      mean.data_ptr() if mean is not None else None, n,
                                      torch.cuda.current_stream(obs.device).cuda_stream)
        if rc != 0:
            raise _lib.RocketLibraryError(f"rl_policy_sample failed ({rc}): {self._L.rl_policy_last_error().decode()}")
        self.draw += 1
        return actions, clipped, log_prob, mean

    def act_and_value(self, obs: torch.Tensor, actions: Optional[torch.Tensor] = None,
                      log_prob: Optional[torch.Tensor] = None, values: Optional[torch.Tensor] = None,
                      want_mean: bool = False):
        """``act`` and ``predict_values`` of the same observations in ONE launch
        (``rl_policy_value_sample``: some CTAs run the policy net, the others the value net).  Returns
        ``(actions, clipped, log_prob, mean or None, values)``, bit-identical to the two separate calls."""
        if obs.dtype != torch.float32 or not obs.is_contiguous():
            obs = obs.to(torch.float32).contiguous()
        n = obs.shape[0]
        f32 = dict(dtype=torch.float32, device=obs.device)
        actions = torch.empty((n, 2), **f32) if actions is None else actions
        log_prob = torch.empty((n,), **f32) if log_prob is None else log_prob
        values = torch.empty((n,), **f32) if values is None else values
        clipped = torch.empty((n, 2), **f32)
        mean = torch.empty((n, 2), **f32) if want_mean else None
        assert actions.is_contiguous() and log_prob.is_contiguous() and values.is_contiguous() and values.numel() == n
        if n == 0:
            return actions, clipped, log_prob, mean, values
        rc = self._L.rl_policy_value_sample(self._pi.data_ptr(), self._vf.data_ptr(), obs.data_ptr(), self._log_std.data_ptr(),
                                            self.seed, self.draw, self.row_offset, self._low, self._high, actions.data_ptr(),
                                            clipped.data_ptr(), log_prob.data_ptr(),
                                            mean.data_ptr() if mean is not None else None, values.data_ptr(), n,
                                            torch.cuda.current_stream(obs.device).cuda_stream)
        if rc != 0:
            raise _lib.RocketLibraryError(f"rl_policy_value_sample failed ({rc}): {self._L.rl_policy_last_error().decode()}")
        self.draw += 1
        return actions, clipped, log_prob, mean, values

    def forward(self, obs: torch.Tensor, generator: Optional[torch.Generator] = None):
        """``(actions, values, log_prob)`` as ``ActorCriticMLP.forward``.  ``generator`` is ignored:
        the noise comes from the kernel's own Philox stream (``seed``, ``draw``)."""
        actions, _, log_prob, _, values = self.act_and_value(obs)
        return actions, values, log_prob

    __call__ = forward


class DeviceRolloutBuffer:
    """Pre-allocated ``[T, N, ...]`` device tensors (SB3 ``RolloutBuffer`` fields)."""

    def __init__(self, n_steps: int, num_envs: int, device, obs_dim: int = 8, act_dim: int = 2):
        f32 = dict(dtype=torch.float32, device=device)
        self.n_steps, self.num_envs = n_steps, num_envs
        self.observations = torch.zeros((n_steps, num_envs, obs_dim), **f32)
        self.actions = torch.zeros((n_steps, num_envs, act_dim), **f32)
        self.rewards = torch.zeros((n_steps, num_envs), **f32)
        self.values = torch.zeros((n_steps, num_envs), **f32)
        self.log_probs = torch.zeros((n_steps, num_envs), **f32)
        self.episode_starts = torch.zeros((n_steps, num_envs), dtype=torch.bool, device=device)
        self.advantages = torch.zeros((n_steps, num_envs), **f32)
        self.returns = torch.zeros((n_steps, num_envs), **f32)

    def compute_returns_and_advantage(self, last_values: torch.Tensor, dones: torch.Tensor, gamma: float,
                                      gae_lambda: float) -> None:
        L = _lib.load()
        last_values = last_values.to(torch.float32).contiguous()
        dones = dones.to(torch.bool).contiguous()
        rc = L.rl_gae(self.rewards.data_ptr(), self.values.data_ptr(), self.episode_starts.data_ptr(),
                      last_values.data_ptr(), dones.data_ptr(), self.advantages.data_ptr(), self.returns.data_ptr(),
                      self.n_steps, self.num_envs, float(gamma), float(gae_lambda),
                      torch.cuda.current_stream(self.rewards.device).cuda_stream)
        if rc != 0:
            raise _lib.RocketLibraryError(f"rl_gae failed ({rc}): {L.rl_rollout_last_error().decode()}")


@torch.no_grad()
def collect_rollout(env, policy, buffer: DeviceRolloutBuffer, last_obs: torch.Tensor,
                    last_episode_starts: torch.Tensor, gamma: float = 0.99, gae_lambda: float = 0.95,
                    bootstrap_timeouts: bool = True, generator: Optional[torch.Generator] = None):
    """One ``collect_rollouts`` pass of ``buffer.n_steps`` steps.  ``env`` is a
    ``RocketLandingVecEnv`` or a ``DeviceVecNormalize`` around one (auto-reset on).  Returns
    ``(last_obs, last_episode_starts)`` for the next pass."""
    low = torch.tensor([0.0, -1.0], device=last_obs.device)
    high = torch.tensor([1.0, 1.0], device=last_obs.device)
    T = buffer.n_steps
    fused = hasattr(policy, "act_and_value")
    direct = hasattr(env, "venv")          # DeviceVecNormalize can write its outputs straight into the buffer
    buffer.observations[0].copy_(last_obs)
    buffer.episode_starts[0].copy_(last_episode_starts)
    obs = buffer.observations[0]
    starts = last_episode_starts
    pending = None                         # (step, time-limit mask) whose bootstrap is still to be decided
    L = _lib.load()
    n_envs = buffer.num_envs
    flag = _pinned_flags(T, last_obs.device)
    scratch = _scratch(T, n_envs, last_obs.device)
    if bootstrap_timeouts:
        flag["host"].zero_()               # (every event of the previous pass has been synchronised: no kernel writes it now)
    for t in range(T):
        if fused:     # sample, clip and log-prob in the policy net's epilogue, the value net in the same launch, straight into the buffer
            _, clipped, _, _, _ = policy.act_and_value(obs, buffer.actions[t], buffer.log_probs[t], buffer.values[t])
        else:
            actions, values, log_probs = policy(obs, generator)
            buffer.actions[t].copy_(actions)
            buffer.values[t].copy_(values)
            buffer.log_probs[t].copy_(log_probs)
            clipped = torch.clamp(actions, low, high).contiguous()   # SB3 clips Box actions before env.step
        nxt = buffer.observations[t + 1] if t + 1 < T else None      # the next step's observation slot
        if direct:
            obs, rewards, terminated, truncated, info = env.step(clipped, out_obs=nxt, out_reward=buffer.rewards[t])
        else:
            obs, rewards, terminated, truncated, info = env.step(clipped)
            buffer.rewards[t].copy_(rewards)
            if nxt is not None:
                nxt.copy_(obs)
                obs = nxt
        starts_out = buffer.episode_starts[t + 1] if t + 1 < T else scratch["last_starts"]
        if bootstrap_timeouts:
            # SB3 adds gamma V(terminal observation) to the reward of a step that ended by the time limit
            # (infos[i]["TimeLimit.truncated"]).  Asking the device "did any rocket time out?" every
            # step would stall the host behind the GPU, so the question is asked asynchronously and
            # answered one step late: a rocket that timed out at step t was reset at step t and cannot
            # finish another episode at step t + 1, so its row of the terminal-observation buffer is
            # still intact then.  One kernel (rl_episode_flags) derives the next episode_start, the
            # time-limit mask and the "any" flag from the step's two flag arrays.
            timeout = scratch["timeout"][t & 1]
            # (the "any" word is written straight into pinned host memory -- under unified addressing the device
            # reaches it by the host pointer -- so no copy sits between this kernel and the next step's)
            rc = L.rl_episode_flags(terminated.data_ptr(), truncated.data_ptr(), starts_out.data_ptr(), timeout.data_ptr(),
                                    flag["host"][t:t + 1].data_ptr(), n_envs, torch.cuda.current_stream(last_obs.device).cuda_stream)
            if rc != 0:
                raise _lib.RocketLibraryError(f"rl_episode_flags failed ({rc}): {L.rl_rollout_last_error().decode()}")
            flag["events"][t].record()
            if pending is not None:
                _bootstrap(policy, buffer, info, pending, flag, gamma, starts_out)
            pending = (t, timeout)
            starts = starts_out
        else:
            starts = torch.logical_or(terminated, truncated, out=starts_out)
    if pending is not None:
        _bootstrap(policy, buffer, info, pending, _pinned_flags(T, last_obs.device), gamma, None)
    last_values = policy.predict_values(obs)
    buffer.compute_returns_and_advantage(last_values, starts, gamma, gae_lambda)
    return obs, starts.clone()             # (the last step's dones live in per-shape scratch: hand out a copy)


_FLAGS: dict = {}
_SCRATCH: dict = {}


def _pinned_flags(n_steps: int, device) -> dict:
    """Per-device pinned host flags + events for the asynchronous "any time-limit truncation?" test."""
    key = (str(device), n_steps)
    if key not in _FLAGS:
        _FLAGS[key] = {"host": torch.zeros(n_steps, dtype=torch.int32).pin_memory(),
                       "events": [torch.cuda.Event() for _ in range(n_steps)]}
    return _FLAGS[key]


def _scratch(n_steps: int, n_envs: int, device) -> dict:
    """Per-(device, shape) device scratch of ``collect_rollout``: two alternating time-limit masks (a mask is
    read one step after it was written) and the last step's dones."""
    key = (str(device), n_steps, n_envs)
    if key not in _SCRATCH:
        _SCRATCH[key] = {"timeout": torch.zeros((2, n_envs), dtype=torch.bool, device=device),
                         "last_starts": torch.zeros(n_envs, dtype=torch.bool, device=device)}
    return _SCRATCH[key]


def _bootstrap(policy, buffer: DeviceRolloutBuffer, info, pending, flag, gamma: float, done_now) -> None:
    """Resolve the time-limit bootstrap of an earlier step (see ``collect_rollout``)."""
    t, mask = pending
    flag["events"][t].synchronize()        # recorded a whole env step ago: already complete in steady state
    if not int(flag["host"][t]):
        return
    if done_now is not None:               # rows overwritten since (cannot happen with episodes longer than one step)
        mask = torch.logical_and(mask, torch.logical_not(done_now))
    tv = policy.predict_values(info["terminal_observation"][mask].contiguous())
    buffer.rewards[t][mask] += gamma * tv
