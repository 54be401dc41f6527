"""Rollout path (SURVEY.md section 8f row 2 / BASELINE.json configs[4]): policy shape, GAE kernel
against the restated SB3 recurrence, and an end-to-end device-resident collection."""
import numpy as np
import pytest

from oracle.ppo_rollout_oracle import compute_returns_and_advantage


def test_policy_has_the_parameter_count_of_the_shipped_checkpoints():
    from rocket_landing_rl_b200 import ActorCriticMLP
    net = ActorCriticMLP()
    assert sum(p.numel() for p in net.parameters()) == 136_965        # assets/model/v3
    import torch
    a, v, lp = net(torch.zeros(5, 8))
    assert a.shape == (5, 2) and v.shape == (5,) and lp.shape == (5,)


def test_gae_oracle_closed_form():
    # single env, no terminations, lambda = 1: advantage = discounted return + bootstrap - value
    r = np.array([[1.0], [2.0], [3.0]])
    v = np.array([[0.5], [0.25], [0.125]])
    adv, ret = compute_returns_and_advantage(r, v, np.zeros((3, 1), bool), np.array([4.0]), np.array([False]), 0.9, 1.0)
    want0 = 1 + 0.9 * 2 + 0.81 * 3 + 0.729 * 4 - 0.5
    assert adv[0, 0] == pytest.approx(want0) and ret[0, 0] == pytest.approx(want0 + 0.5)
    # an episode boundary cuts the bootstrap
    adv, _ = compute_returns_and_advantage(r, v, np.array([[0], [0], [1]], bool), np.array([4.0]), np.array([False]), 0.9, 1.0)
    assert adv[1, 0] == pytest.approx(2.0 - 0.25)


@pytest.mark.gpu
def test_gae_kernel_matches_the_oracle():
    import torch
    from rocket_landing_rl_b200 import DeviceRolloutBuffer
    T, n = 37, 5003
    g = torch.Generator(device="cuda").manual_seed(0)
    buf = DeviceRolloutBuffer(T, n, "cuda")
    buf.rewards.copy_(torch.randn((T, n), device="cuda", generator=g) * 10)
    buf.values.copy_(torch.randn((T, n), device="cuda", generator=g))
    buf.episode_starts.copy_(torch.rand((T, n), device="cuda", generator=g) < 0.02)
    last_v = torch.randn(n, device="cuda", generator=g)
    dones = torch.rand(n, device="cuda", generator=g) < 0.05
    buf.compute_returns_and_advantage(last_v, dones, 0.99, 0.95)
    adv, ret = compute_returns_and_advantage(buf.rewards.cpu().numpy(), buf.values.cpu().numpy(),
                                             buf.episode_starts.cpu().numpy(), last_v.cpu().numpy(),
                                             dones.cpu().numpy(), 0.99, 0.95)
    np.testing.assert_allclose(buf.advantages.cpu().numpy(), adv, rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(buf.returns.cpu().numpy(), ret, rtol=2e-5, atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("n,offset", [(1, 0), (15, 0), (16, 0), (4099, 0), (1 << 20, 0), (4099, 3)])
def test_episode_flags_kernel(n, offset):
    """rl_episode_flags against the three torch expressions it replaces in collect_rollout (SB3 derives the
    same from dones / infos on the host); ragged sizes and a misaligned view take the scalar path."""
    import torch
    from rocket_landing_rl_b200 import _lib
    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(n + offset)
    base = torch.rand((2, n + offset), device="cuda", generator=g)
    terminated = (base[0] < 0.3)[offset:]
    truncated = (base[1] < (0.3 if n > 16 else 0.9))[offset:]
    if offset:
        assert terminated.data_ptr() % 16 != 0
    else:
        terminated, truncated = terminated.contiguous(), truncated.contiguous()
    starts = torch.zeros(n + 32, dtype=torch.bool, device="cuda")
    limit = torch.zeros(n + 32, dtype=torch.bool, device="cuda")
    any_flag = torch.zeros(2, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    assert L.rl_episode_flags(terminated.data_ptr(), truncated.data_ptr(), starts.data_ptr(), limit.data_ptr(),
                              any_flag.data_ptr(), n, stream) == 0
    want_limit = truncated & ~terminated
    assert torch.equal(starts[:n], terminated | truncated) and not starts[n:].any()      # nothing written past n
    assert torch.equal(limit[:n], want_limit) and not limit[n:].any()
    assert any_flag.tolist() == [int(want_limit.any()), 0]
    # a null episode_start output is allowed; no time-limit rocket leaves the word alone
    none = torch.zeros_like(truncated)
    assert L.rl_episode_flags(terminated.data_ptr(), none.data_ptr(), None, limit.data_ptr(), any_flag[1:].data_ptr(), n, stream) == 0
    assert not limit.any() and any_flag.tolist()[1] == 0
    assert L.rl_episode_flags(None, none.data_ptr(), None, limit.data_ptr(), any_flag.data_ptr(), n, stream) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_device_resident_rollout_collection(fused):
    """configs[4] in miniature: envs -> VecNormalize -> policy -> buffer -> GAE, all on the GPU,
    with the policy evaluated by PyTorch or by the fused tensor-core kernel."""
    import torch
    from rocket_landing_rl_b200 import (ActorCriticMLP, DeviceRolloutBuffer, DeviceVecNormalize,
                                        FusedActorCritic, RocketLandingVecEnv, collect_rollout)
    torch.manual_seed(0)
    n, T = 8192, 160
    env = DeviceVecNormalize(RocketLandingVecEnv(n, seed=0), gamma=0.99)
    policy = ActorCriticMLP().cuda()
    if fused:
        policy = FusedActorCritic(policy)
    buf = DeviceRolloutBuffer(T, n, "cuda")
    obs = env.reset()
    starts = torch.ones(n, dtype=torch.bool, device="cuda")
    obs, starts = collect_rollout(env, policy, buf, obs, starts, gamma=0.99, gae_lambda=0.95)
    assert bool(torch.isfinite(buf.advantages).all()) and bool(torch.isfinite(buf.returns).all())
    assert bool(buf.episode_starts[0].all()) and 0 < int(buf.episode_starts[1:].sum()) < n * 3
    assert float(buf.observations.abs().max()) <= 10.0 + 1e-5           # clip_obs
    assert float(buf.rewards.abs().max()) <= 10.0 + 1e-4               # clip_reward (no time-limit bootstraps here)
    # the stored actions are the UNCLIPPED samples, as in SB3
    assert float(buf.actions[..., 0].min()) < 0.0 and float(buf.actions[..., 0].max()) > 1.0
    adv, ret = compute_returns_and_advantage(buf.rewards.cpu().numpy(), buf.values.cpu().numpy(),
                                             buf.episode_starts.cpu().numpy(),
                                             policy.predict_values(obs).detach().cpu().numpy(), starts.cpu().numpy(),
                                             0.99, 0.95)
    np.testing.assert_allclose(buf.advantages.cpu().numpy(), adv, rtol=1e-4, atol=1e-3)
    stats = env.venv.stats()
    assert stats.env_steps == n * T and stats.episodes > 0.8 * n
    env.close()
    env.venv.close()


@pytest.mark.gpu
def test_fused_mlp_kernel_matches_the_pytorch_module():
    """rl_mlp_forward (tcgen05 tensor cores, bf16 weights, fp32 accumulation, activations on chip)
    against the fp32 PyTorch module it packs.  Tolerance: bf16 weights carry 2^-9 relative error
    each and the hidden activations are rounded to bf16 once; over 256-term dot products of O(1)
    activations with O(1/16) weights that is ~1e-3 absolute on outputs of O(0.3); 1e-2 is asserted."""
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic
    torch.manual_seed(3)
    net = ActorCriticMLP().cuda()
    with torch.no_grad():                       # biases off zero so that every term of the kernel is exercised
        for p in net.parameters():
            if p.ndim == 1:
                p.uniform_(-0.3, 0.3)
    fused = FusedActorCritic(net)
    for n in (1, 15, 16, 17, 127, 128, 129, 5000, 262144 + 77):
        obs = (torch.randn((n, 8), device="cuda") * 3.0).clamp_(-10, 10)      # normalised observations, clip_obs = 10
        with torch.no_grad():
            want_mean = net.action_net(net.policy_net(obs))
            want_val = net.predict_values(obs)
        got_mean = fused.action_mean(obs)
        got_val = fused.predict_values(obs)
        assert got_mean.shape == (n, 2) and got_val.shape == (n,)
        assert float((got_mean - want_mean).abs().max()) < 1e-2, n
        assert float((got_val - want_val).abs().max()) < 1e-2, n
    # every row is computed from its own observation only: permuting rows permutes outputs exactly
    obs = torch.randn((1000, 8), device="cuda")
    perm = torch.randperm(1000, device="cuda")
    assert torch.equal(fused.action_mean(obs)[perm], fused.action_mean(obs[perm].contiguous()))
    a, v, lp = fused(obs)
    assert a.shape == (1000, 2) and v.shape == (1000,) and lp.shape == (1000,)


def test_fma_pipe_tanh_polynomial_of_the_mlp_kernel():
    """The fused MLP kernel evaluates part of its tanh on the FMA pipe (csrc/rl_policy.cu: tanh_poly_core).
    The coefficients are read from the source, so that they cannot drift from tools/tanh_poly_fit.py's
    claim: both clamped forms (epilogue 1 clamps u = x^2 and the result, epilogue 2 clamps x) stay within
    1.4e-3 of tanh for EVERY fp32 input, are odd and never leave [-1, 1] by more than that."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "rocket-landing-rl_b200", "csrc", "rl_policy.cu")).read()
    body = re.search(r"tanh_poly_core\(uint64_t x, uint64_t u\) \{\s*const float c\[8\] = \{([^}]*)\}", src).group(1)
    coef = np.array([np.float32(v.strip().rstrip("f")) for v in body.split(",")], np.float32)
    clamp = np.float32(re.search(r"kTanhClamp = ([0-9.]+)f", src).group(1))
    assert coef.shape == (8,)

    def core(x, u):                                              # Horner in fp32, as the packed FFMA2 chain
        r = (u * coef[0] + coef[1]).astype(np.float32)
        for k in range(2, 8):
            r = (r * u + coef[k]).astype(np.float32)
        return (x * r).astype(np.float32)

    x = np.concatenate([np.linspace(-20, 20, 2_000_001), [0.0, 1e-30, -1e-30, 1e4, -1e4, clamp, -clamp]]).astype(np.float32)
    ref = np.tanh(x.astype(np.float64))
    e1 = np.clip(core(x, np.minimum(x * x, clamp * clamp)), -1.0, 1.0)          # epilogue 1 (before the bf16 rounding)
    xc = np.clip(x, -clamp, clamp)
    e2 = core(xc, xc * xc)                                                        # epilogue 2
    for got in (e1, e2):
        assert np.abs(got - ref).max() < 1.4e-3
        assert np.abs(got).max() <= 1.0 + 1.4e-3
    assert np.array_equal(core(x, np.minimum(x * x, clamp * clamp)), -core(-x, np.minimum(x * x, clamp * clamp)))


@pytest.mark.gpu
def test_fused_mlp_kernel_with_saturating_units():
    """The same comparison with layer-1 weights scaled by 4: more than half of the hidden units then sit in
    the clamped part of the FMA-pipe tanh (|x| > 3.727) or far out on the MUFU one -- the regime a trained
    policy's large normalised observations reach (measured: 6e-3, PyTorch's bf16 autocast 8e-3)."""
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic
    torch.manual_seed(11)
    net = ActorCriticMLP().cuda()
    with torch.no_grad():
        for mod in (net.policy_net, net.value_net_body):
            first = [m for m in mod.modules() if isinstance(m, torch.nn.Linear)][0]
            first.weight.mul_(4.0)
            first.bias.uniform_(-1.0, 1.0)
    fused = FusedActorCritic(net)
    obs = (torch.randn((65536 + 3, 8), device="cuda") * 3.0).clamp_(-10, 10)
    with torch.no_grad():
        first = [m for m in net.policy_net.modules() if isinstance(m, torch.nn.Linear)][0]
        pre = torch.nn.functional.linear(obs, first.weight, first.bias)
        assert float((pre.abs() > 3.727).float().mean()) > 0.4               # the regime is reached
        want_mean = net.action_net(net.policy_net(obs))
        want_val = net.predict_values(obs)
    assert float((fused.action_mean(obs) - want_mean).abs().max()) < 1e-2
    assert float((fused.predict_values(obs) - want_val).abs().max()) < 1e-2


@pytest.mark.gpu
def test_fused_action_sampling():
    """rl_policy_sample: the policy net with SB3's DiagGaussian sampling in its epilogue.  The noise
    is Philox + Box-Muller inside the kernel, so the check is distributional (standardised samples
    have zero mean, unit variance, no correlation between the two components, Gaussian tails) plus the
    exact relations between the outputs: log-prob formula, Box clipping, reproducibility."""
    import math
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic
    torch.manual_seed(1)
    net = ActorCriticMLP(log_std_init=-0.7).cuda()
    with torch.no_grad():
        net.log_std.copy_(torch.tensor([-0.7, 0.3]))
    fused = FusedActorCritic(net, seed=123)
    n = 1 << 20
    obs = torch.randn((n, 8), device="cuda").clamp_(-10, 10)
    actions, clipped, logp, mean = fused.act(obs, want_mean=True)
    assert torch.allclose(mean, fused.action_mean(obs), atol=0, rtol=0)          # same kernel, same numbers
    std = net.log_std.detach().exp()
    z = (actions - mean) / std
    assert abs(float(z.mean())) < 4e-3 and abs(float(z.var()) - 1.0) < 6e-3
    assert abs(float((z[:, 0] * z[:, 1]).mean())) < 4e-3
    assert 0.0024 < float((z.abs() > 3.0).float().mean()) < 0.0030             # P(|N| > 3) = 0.0027
    want_logp = (-0.5 * z.pow(2) - net.log_std.detach() - 0.5 * math.log(2 * math.pi)).sum(-1)
    assert float((logp - want_logp).abs().max()) < 2e-3
    low, high = torch.tensor([0.0, -1.0], device="cuda"), torch.tensor([1.0, 1.0], device="cuda")
    assert torch.equal(clipped, torch.clamp(actions, low, high))
    # a new draw gives new noise; the same (seed, draw) gives the same bits
    a2, _, _, _ = fused.act(obs)
    assert not torch.equal(a2, actions)
    again = FusedActorCritic(net, seed=123)
    a3, _, _, _ = again.act(obs)
    assert torch.equal(a3, actions)
    # in-place into (slices of) a rollout buffer
    buf_a = torch.zeros((2, n, 2), device="cuda")
    buf_l = torch.zeros((2, n), device="cuda")
    again.act(obs, buf_a[1], buf_l[1])
    assert torch.equal(buf_a[1], a2) and float(buf_a[0].abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 127, 129, 300, 5000, 20000, 262144 + 77])
def test_policy_and_value_net_in_one_launch_are_bit_identical_to_two(n):
    """rl_policy_value_sample runs the policy net on some CTAs and the value net on the others: every
    output word equals what rl_policy_sample + rl_mlp_forward produce (same seed and draw), for batches
    smaller than one row block, smaller than the grid, and far larger; rows outside [0, n) stay untouched."""
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic
    torch.manual_seed(5)
    net = ActorCriticMLP().cuda()
    with torch.no_grad():
        net.log_std.fill_(-0.7)
    obs = (torch.randn((n, 8), device="cuda") * 3.0).clamp_(-10, 10)
    two = FusedActorCritic(net, seed=9, row_offset=1000)
    one = FusedActorCritic(net, seed=9, row_offset=1000)
    for _ in range(2):                                           # two draws: the counter advances alike
        a2, c2, l2, m2 = two.act(obs, want_mean=True)
        v2 = two.predict_values(obs)
        pad = 64
        acts = torch.full((n + pad, 2), 7.0, device="cuda")
        logp = torch.full((n + pad,), 7.0, device="cuda")
        vals = torch.full((n + pad,), 7.0, device="cuda")
        a1, c1, l1, m1, v1 = one.act_and_value(obs, acts[:n], logp[:n], vals[:n], want_mean=True)
        for got, want in ((a1, a2), (c1, c2), (l1, l2), (m1, m2), (v1, v2)):
            assert torch.equal(got, want)
        assert bool((acts[n:] == 7.0).all()) and bool((logp[n:] == 7.0).all()) and bool((vals[n:] == 7.0).all())
    a, v, lp = one(obs)
    assert a.shape == (n, 2) and v.shape == (n,) and lp.shape == (n,)


@pytest.mark.gpu
def test_policy_noise_is_independent_of_the_reset_stream_for_equal_seeds():
    """Both streams are Philox4x32-10 and both default to seed 0.  With the SAME key the policy's
    counter (row, draw, 0) would coincide with the reset's counter (global id, epoch, block 0)
    whenever draw == epoch -- which is exactly the first action of every episode in collect_rollout
    -- making the exploration noise a function of the initial x, y the policy observes.  The
    policy's key carries a domain tag: its uniforms must be uncorrelated with the reset's, and
    shards that share a seed must draw different noise (row_offset)."""
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic, RocketLandingVecEnv
    n = 1 << 17
    net = ActorCriticMLP(log_std_init=0.0).cuda()
    with torch.no_grad():
        for p_ in net.parameters():
            p_.zero_()                                    # mean 0, std 1: actions == the noise itself
    obs = torch.zeros((n, 8), device="cuda")
    for epoch in (0, 3):
        env = RocketLandingVecEnv(n, seed=0, auto_reset=False)
        for _ in range(epoch + 1):
            reset_obs = env.reset()[0].clone()            # the reset at launch epoch `epoch`
        env.close()
        ux = (reset_obs[:, 0] + 1500.0) / 3000.0          # u01(w[0]) of Philox block 0 (config.yaml:53)
        uy = (reset_obs[:, 1] - 1800.0) / 400.0           # u01(w[1])
        fused = FusedActorCritic(net, seed=0)
        fused.draw = epoch
        z, _, _, _ = fused.act(obs)
        u1 = torch.exp(-0.5 * (z * z).sum(-1))            # Box-Muller radius -> its uniform
        u2 = (torch.atan2(z[:, 1], z[:, 0]) / (2 * torch.pi)) % 1.0
        for a_, b_ in ((u1, ux), (u2, uy), (u1, uy), (u2, ux)):
            c = float(torch.corrcoef(torch.stack([a_, b_]))[0, 1])
            assert abs(c) < 0.02, (epoch, c)              # same key: |c| = 1
    shard = FusedActorCritic(net, seed=0, row_offset=n // 2)
    whole = FusedActorCritic(net, seed=0)
    zw, _, _, _ = whole.act(obs)
    zs, _, _, _ = shard.act(obs[: n // 2].contiguous())
    assert torch.equal(zs, zw[n // 2:]) and not torch.equal(zs, zw[: n // 2])


@pytest.mark.gpu
def test_time_limit_bootstrap_is_resolved_one_step_late_but_exactly():
    """collect_rollout adds gamma V(terminal observation) to the reward of steps that end by the time
    limit (SB3's TimeLimit.truncated handling) without a host synchronisation per step: the "any
    timeout?" flag is read back one step late.  With max_episode_steps = 5 every rocket times out at
    steps 4, 9, ...: the rewards with and without bootstrapping must differ there, and only there,
    by gamma V(observation an env WITHOUT auto-reset shows after the same five actions)."""
    import torch
    from rocket_landing_rl_b200 import (ActorCriticMLP, DeviceRolloutBuffer, FusedActorCritic,
                                        RocketLandingVecEnv, collect_rollout, load_params)
    p = load_params()
    p.max_episode_steps = 5
    n, T, gamma = 3000, 12, 0.9
    torch.manual_seed(0)
    net = ActorCriticMLP().cuda()

    def run(bootstrap):
        env = RocketLandingVecEnv(n, seed=4, params=p)
        pol = FusedActorCritic(net, seed=9)
        buf = DeviceRolloutBuffer(T, n, "cuda")
        obs, _ = env.reset()
        collect_rollout(env, pol, buf, obs, torch.ones(n, dtype=torch.bool, device="cuda"), gamma=gamma,
                        bootstrap_timeouts=bootstrap)
        env.close()
        return buf

    a, b = run(True), run(False)
    assert torch.equal(a.actions, b.actions) and torch.equal(a.observations, b.observations)
    diff = a.rewards - b.rewards
    for t in range(T):
        if t % 5 == 4:
            assert bool(a.episode_starts[t + 1].all()) if t + 1 < T else True
            assert float(diff[t].abs().min()) > 0.0
        else:
            assert float(diff[t].abs().max()) == 0.0
    # the first time limit: terminal observation = what an env without auto-reset returns at step 4
    env = RocketLandingVecEnv(n, seed=4, params=p, auto_reset=False)
    env.reset()
    low, high = torch.tensor([0.0, -1.0], device="cuda"), torch.tensor([1.0, 1.0], device="cuda")
    for t in range(5):
        obs, _, term, trunc, _ = env.step(torch.clamp(a.actions[t], low, high).contiguous())
    assert bool(trunc.all()) and not bool(term.any())
    want = gamma * FusedActorCritic(net).predict_values(obs)
    # (diff is a difference of fp32 rewards of magnitude up to ~1e2: 1e-6 relative of those, not of itself)
    assert torch.allclose(diff[4], want, rtol=1e-4, atol=3e-4)
    env.close()


def test_pack_mlp_layout_matches_the_header():
    """Host logic of the fused MLP path (no GPU): the packed blob has the layout
    include/rocket_landing_b200.h documents for rl_mlp_forward, byte for byte."""
    import re
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, pack_mlp
    from rocket_landing_rl_b200.rollout import MLP_BLOB_BYTES
    import os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                            "rocket_landing_b200.h")).read()
    assert int(re.search(r"#define RL_MLP_BLOB_BYTES (\d+)", hdr).group(1)) == MLP_BLOB_BYTES
    torch.manual_seed(0)
    net = ActorCriticMLP()
    for body, head in ((net.policy_net, net.action_net), (net.value_net_body, net.value_net)):
        blob = pack_mlp(body, head, "cpu")
        assert blob.dtype == torch.uint8 and blob.numel() == MLP_BLOB_BYTES
        off = 0

        def take(nbytes, dtype):
            nonlocal off
            t = blob[off:off + nbytes].view(dtype)
            off += nbytes
            return t

        w1 = take(256 * 16 * 2, torch.bfloat16).view(256, 16)
        assert torch.equal(w1[:, :8], body[0].weight.detach().to(torch.bfloat16)) and torch.equal(w1[:, 8:], w1[:, :8])
        assert torch.equal(take(256 * 4, torch.float32), body[0].bias.detach())
        w2 = take(256 * 256 * 2, torch.bfloat16).view(32, 256, 8)                 # [k / 8][n][k % 8]
        assert torch.equal(w2.permute(1, 0, 2).reshape(256, 256), body[2].weight.detach().to(torch.bfloat16))
        assert torch.equal(take(256 * 4, torch.float32), body[2].bias.detach())
        w3 = take(2 * 256 * 4, torch.float32).view(2, 256)
        k = head.weight.shape[0]
        assert torch.equal(w3[:k], head.weight.detach()) and float(w3[k:].abs().sum()) == 0.0
        b3 = take(16, torch.float32)
        assert torch.equal(b3[:k], head.bias.detach()) and float(b3[k:].abs().sum()) == 0.0
        assert off == MLP_BLOB_BYTES
    with pytest.raises(ValueError):
        pack_mlp(torch.nn.Sequential(torch.nn.Linear(8, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh()),
                 torch.nn.Linear(64, 2), "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 127, 129, 1000, 148 * 128 + 5])
def test_fused_mlp_kernels_stay_inside_their_output_rows(n):
    """compute-sanitizer is not available on the GPU pool: ragged batches write into views of larger
    canary-filled tensors, and nothing beyond row n may change (row blocks are 128 rockets wide)."""
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, FusedActorCritic
    torch.manual_seed(2)
    fused = FusedActorCritic(ActorCriticMLP().cuda())
    obs = torch.randn((n, 8), device="cuda")
    pad = 300

    def canary(*shape):
        return torch.full(shape, -7.0, device="cuda")

    vals = canary(n + pad)
    fused.predict_values(obs, out=vals[:n])
    acts, logp = canary(n + pad, 2), canary(n + pad)
    _, clipped, _, mean = fused.act(obs, acts[:n], logp[:n], want_mean=True)
    torch.cuda.synchronize()
    for t in (vals, acts, logp):
        assert bool((t[n:] == -7.0).all()) and bool(torch.isfinite(t[:n]).all())
    assert clipped.shape == (n, 2) and mean.shape == (n, 2)
    assert torch.equal(fused.predict_values(obs), vals[:n])


_PDL_DIGEST = r"""
import hashlib, sys, torch
sys.path.insert(0, sys.argv[1])
from rocket_landing_rl_b200 import (ActorCriticMLP, DeviceRolloutBuffer, DeviceVecNormalize, FusedActorCritic,
                                    RocketLandingVecEnv, collect_rollout)
torch.manual_seed(2)
n, T = 20000, 12
env = DeviceVecNormalize(RocketLandingVecEnv(n, device="cuda:0", seed=4), gamma=0.99)
policy = FusedActorCritic(ActorCriticMLP().cuda(), seed=6)
buf = DeviceRolloutBuffer(T, n, "cuda:0")
obs = env.reset()
starts = torch.ones(n, dtype=torch.bool, device="cuda:0")
h = hashlib.sha256()
for _ in range(3):
    obs, starts = collect_rollout(env, policy, buf, obs, starts)
    for t in (buf.observations, buf.actions, buf.rewards, buf.values, buf.log_probs, buf.advantages, buf.returns):
        h.update(t.cpu().numpy().tobytes())
    h.update(buf.episode_starts.cpu().numpy().tobytes())
print("DIGEST", h.hexdigest())
"""


@pytest.mark.gpu
def test_programmatic_dependent_launch_changes_no_bit():
    """The kernels of a rollout step are chained by programmatic dependent launch (csrc/rl_launch.cuh): each
    may be set up while its predecessor drains and waits for it before it touches memory.  Three collection
    passes (actor-critic, env step with auto-reset, VecNormalize, flags, GAE) with and without it
    (RL_B200_PDL=0 launches everything the ordinary way) must produce the same bytes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    for pdl in ("1", "0"):
        env = dict(os.environ, RL_B200_PDL=pdl)
        out = subprocess.run([sys.executable, "-c", _PDL_DIGEST, root], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([line for line in out.stdout.splitlines() if line.startswith("DIGEST")][0])
    assert digests[0] == digests[1]
