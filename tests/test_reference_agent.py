"""The reference's shipped agent (assets/model/v3: SB3 PPO MlpPolicy + VecNormalize statistics)
running on this repo's stack -- B200 environment step, device VecNormalize, fused tcgen05 policy
kernel -- against what the same agent does in the reference's own environment
(tests/golden/ref_policy_v3.npz, produced by tests/golden/make_policy_fixture.py from the live
reference).  Closing the loop this way checks observation order and clipping, normalisation, action
semantics and the physics together: a trained policy is a sensitive detector of any of them."""
import os

import numpy as np
import pytest

from oracle.ref_loader import reference_available, reference_root


@pytest.fixture(scope="module")
def fixture(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "ref_policy_v3.npz")))


@pytest.fixture(scope="module")
def fixture_v2(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "ref_policy_v2.npz")))


def _state_dict(fx):
    return {k[2:].replace("__", "."): v for k, v in fx.items() if k.startswith("w_")}


def test_checkpoint_readers_without_sb3(fixture):
    """policy.pth inside the SB3 zip and vecnormalize.pkl are read without stable-baselines3."""
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, read_sb3_vecnormalize
    net = ActorCriticMLP.from_sb3_state_dict(_state_dict(fixture))
    assert sum(p.numel() for p in net.parameters()) == 136_965
    with torch.no_grad():
        mean = net.action_net(net.policy_net(torch.from_numpy(fixture["trip_norm"]))).numpy()
    assert np.abs(mean - fixture["trip_mean"]).max() < 2e-5                  # fp32 torch vs the fixture's fp64 numpy
    if reference_available():
        base = os.path.join(reference_root(), "assets", "model", "v3")
        net2 = ActorCriticMLP.from_sb3_zip(os.path.join(base, "best_model.zip"))
        for a, b in zip(net.parameters(), net2.parameters()):
            assert torch.equal(a, b)
        st = read_sb3_vecnormalize(os.path.join(base, "vecnormalize.pkl"))
        assert np.array_equal(st["obs_mean"], fixture["obs_mean"]) and st["clip_obs"] == float(fixture["clip_obs"])


@pytest.mark.gpu
def test_open_loop_normalisation_and_policy_match_the_reference_agent(fixture):
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, DeviceVecNormalize, FusedActorCritic, RocketLandingVecEnv
    env = DeviceVecNormalize(RocketLandingVecEnv(64, seed=0), training=False, norm_reward=False,
                             clip_obs=float(fixture["clip_obs"]), epsilon=float(fixture["epsilon"]))
    env.load_stats({k: fixture[k] for k in ("obs_mean", "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count")})
    norm = env.normalize_obs(torch.from_numpy(fixture["trip_obs"]).cuda())
    assert float((norm.cpu() - torch.from_numpy(fixture["trip_norm"])).abs().max()) < 1e-5
    fused = FusedActorCritic(ActorCriticMLP.from_sb3_state_dict(_state_dict(fixture)).cuda())
    mean = fused.action_mean(norm)
    # bf16 weights: tolerance as in tests/test_rollout.py, relative to outputs of magnitude ~1
    assert float((mean.cpu() - torch.from_numpy(fixture["trip_mean"])).abs().max()) < 2e-2
    env.close()
    env.venv.close()


@pytest.mark.gpu
@pytest.mark.parametrize("version", ["v3", "v2"])
def test_closed_loop_outcomes_match_the_reference_environment(fixture, fixture_v2, version):
    """v3 (8-256-256, evaluated by the fused tcgen05 kernel) and v2 (8-128-128, 950 000 training steps,
    evaluated by PyTorch: the fused kernel is specialised for the 256-wide nets the reference trains now).
    16 384 rockets, each flown by the agent through its first episode (deterministic actions, no
    auto-reset).  The reference's 1 500 episodes give the expected frequencies of the outcomes (time
    limit / out of bounds, safe, good, ok, unsafe) with binomial standard errors of 0.5-1.3 %; the
    B200 stack must agree within three of them (+ 0.5 %), and in mean episode length within 5 %.
    Measured: B200 0.607 / 0.049 / 0.036 / 0.247 / 0.061, reference 0.601 / 0.053 / 0.039 / 0.249 / 0.057."""
    import torch
    from rocket_landing_rl_b200 import ActorCriticMLP, DeviceVecNormalize, FusedActorCritic, RocketLandingVecEnv
    fixture = fixture if version == "v3" else fixture_v2
    n = 16384
    raw = RocketLandingVecEnv(n, seed=7, auto_reset=False)
    env = DeviceVecNormalize(raw, training=False, norm_reward=False, clip_obs=float(fixture["clip_obs"]),
                             epsilon=float(fixture["epsilon"]))
    env.load_stats({k: fixture[k] for k in ("obs_mean", "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count")})
    net = ActorCriticMLP.from_sb3_state_dict(_state_dict(fixture)).cuda()
    if version == "v3":
        action_mean = FusedActorCritic(net).action_mean
    else:
        torch.backends.cuda.matmul.allow_tf32 = False

        def action_mean(o):
            with torch.no_grad():
                return net.action_net(net.policy_net(o))
    low, high = torch.tensor([0.0, -1.0], device="cuda"), torch.tensor([1.0, 1.0], device="cuda")
    obs = env.reset()
    first_done = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    outcome = torch.zeros(n, dtype=torch.int8, device="cuda")
    for t in range(1000):
        action = torch.clamp(action_mean(obs), low, high).contiguous()            # PPO.predict(deterministic=True)
        obs, _, term, trunc, info = env.step(action)
        new = (term | trunc) & (first_done < 0)
        first_done[new] = t + 1
        outcome[new] = torch.where(term, info["landing"], torch.zeros_like(info["landing"]))[new]
    assert bool((first_done > 0).all())                                           # the time limit ends everything
    got = torch.bincount(outcome.long(), minlength=5).cpu().numpy() / n
    ref_counts = np.bincount(fixture["outcomes"].astype(np.int64), minlength=5)
    m = ref_counts.sum()
    want = ref_counts / m
    sigma = np.sqrt(np.maximum(want * (1 - want), 1.0 / m) / m)
    print(version, "outcome frequencies (none, safe, good, ok, unsafe): B200", np.round(got, 3), "reference", np.round(want, 3))
    assert np.all(np.abs(got - want) <= 3 * sigma + 0.005), (got, want, sigma)
    mean_len = float(first_done.float().mean())
    print("mean episode length: B200 %.1f, reference %.1f" % (mean_len, fixture["lengths"].mean()))
    assert abs(mean_len - fixture["lengths"].mean()) < 0.05 * fixture["lengths"].mean(), (mean_len, fixture["lengths"].mean())
    env.close()
    raw.close()
