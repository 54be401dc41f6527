"""VecNormalize: the oracle restates stable-baselines3 2.3.2; the
CUDA implementation is checked against it on the real environment's outputs."""
import numpy as np
import pytest

from oracle.vecnormalize_oracle import RunningMeanStd, VecNormalizeOracle


def test_running_mean_std_matches_batch_statistics():
    rng = np.random.default_rng(0)
    data = rng.normal(3.0, 2.0, size=(5000, 4))
    rms = RunningMeanStd(shape=(4,))
    for chunk in np.array_split(data, 7):
        rms.update(chunk)
    # with the 1e-4 pseudo-count the running moments equal the pooled ones to ~1e-7
    np.testing.assert_allclose(rms.mean, data.mean(0), rtol=0, atol=1e-6)
    np.testing.assert_allclose(rms.var, data.var(0), rtol=1e-6)
    assert rms.count == pytest.approx(5000 + 1e-4)


def test_oracle_step_semantics():
    vn = VecNormalizeOracle(3, obs_dim=2, gamma=0.5)
    obs = np.array([[0.0, 10.0], [2.0, 10.0], [4.0, 10.0]], np.float32)
    out = vn.reset(obs)
    assert out.dtype == np.float32 and abs(out[:, 0].mean()) < 1e-3      # centred by its own batch
    rew = np.array([1.0, 2.0, 3.0], np.float32)
    done = np.array([False, True, False])
    o, r, t = vn.step(obs, rew, done, terminal_obs=obs.copy())
    assert vn.returns.tolist() == [1.0, 0.0, 3.0]                        # done -> return zeroed AFTER the update
    assert np.all(np.abs(o) <= 10.0) and np.all(np.abs(r) <= 10.0)
    np.testing.assert_allclose(t[1], o[1])                               # terminal obs normalised with the same stats
    vn.training = False
    before = (vn.obs_rms.mean.copy(), vn.ret_rms.var)
    vn.step(obs * 5, rew, done)
    assert np.array_equal(before[0], vn.obs_rms.mean) and before[1] == vn.ret_rms.var


@pytest.mark.gpu
def test_device_vecnormalize_matches_the_oracle_on_real_rollouts():
    import torch
    from rocket_landing_rl_b200 import DeviceVecNormalize, RocketLandingVecEnv
    n, T = 20_000, 260
    env = RocketLandingVecEnv(n, seed=4)
    vn = DeviceVecNormalize(env, gamma=0.99)
    oracle = VecNormalizeOracle(n, gamma=0.99)
    obs_n = vn.reset()
    want = oracle.reset(vn.get_original_obs().cpu().numpy())
    np.testing.assert_allclose(obs_n.cpu().numpy(), want, rtol=0, atol=2e-5)
    gen = torch.Generator(device="cuda").manual_seed(0)
    worst_obs = worst_rew = 0.0
    for t in range(T):
        a = torch.rand((n, 2), device="cuda", generator=gen)
        a[:, 1] = a[:, 1] * 2 - 1
        if t == 200:
            vn.training = oracle.training = False                        # evaluation mode: statistics frozen
        raw_tobs = None
        o, r, term, trunc, info = vn.step(a)
        done = (term | trunc).cpu().numpy()
        raw_obs = vn.get_original_obs().cpu().numpy()
        raw_rew = vn.get_original_reward().cpu().numpy()
        wo, wr, _ = oracle.step(raw_obs, raw_rew, done)
        go, gr = o.cpu().numpy(), r.cpu().numpy()
        worst_obs = max(worst_obs, float(np.abs(go - wo).max()))
        worst_rew = max(worst_rew, float(np.abs(gr - wr).max()))
        np.testing.assert_allclose(go, wo, rtol=0, atol=3e-5, err_msg=f"obs t={t}")
        np.testing.assert_allclose(gr, wr, rtol=1e-5, atol=3e-5, err_msg=f"reward t={t}")
        if done.any():                                                    # terminal observations are normalised too
            tn = info["terminal_observation"].cpu().numpy()[done]
            assert np.all(np.abs(tn) <= 10.0 + 1e-6)
    st = vn.stats()
    np.testing.assert_allclose(st["obs_mean"], oracle.obs_rms.mean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(st["obs_var"], oracle.obs_rms.var, rtol=1e-8)
    assert st["obs_count"] == pytest.approx(oracle.obs_rms.count, rel=1e-12)
    assert st["ret_var"] == pytest.approx(oracle.ret_rms.var, rel=1e-5)   # returns are kept in fp32 on the device
    assert st["ret_count"] == pytest.approx(oracle.ret_rms.count, rel=1e-12)
    print(f"VecNormalize: worst |obs diff| {worst_obs:.2e}, worst |reward diff| {worst_rew:.2e}")
    # save / load of the statistics (vecnormalize.pkl equivalent) and the inference path of agent.py:181
    vn2 = DeviceVecNormalize(RocketLandingVecEnv(64, seed=1), training=False)
    vn2.load_stats(st)
    batch = vn.get_original_obs()[:64].clone()
    np.testing.assert_allclose(vn2.normalize_obs(batch).cpu().numpy(), oracle.normalize_obs(batch.cpu().numpy()),
                               rtol=0, atol=3e-5)
    assert vn.launch_count == 1 + 2 + 3 * 200 + 60
    vn.close()
    vn2.close()
    env.close()
