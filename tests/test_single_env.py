"""``RocketLandingEnv`` (one rocket, the reference's own class surface: backend/envs/lander.py:18)
driven the way the reference's callers drive it -- reset, step until ``terminated or truncated``,
reset again -- beside the pinned oracle started from the same initial state."""
import numpy as np
import pytest

from oracle.rocket_oracle import OracleEnv

pytestmark = pytest.mark.gpu


def test_surface_matches_the_reference_class():
    from rocket_landing_rl_b200 import RocketLandingEnv
    env = RocketLandingEnv(seed=5)
    assert env.action_space.shape == (2,) and env.observation_space.shape == (8,)
    assert env.action_space.low.tolist() == [0.0, -1.0] and env.action_space.high.tolist() == [1.0, 1.0]
    assert env.observation_space.low.tolist() == [-1500.0, 0.0, -25.0, -200.0, -5.0, -5.0, -15.0, -10.0]   # lander.py:73-101
    assert env.observation_space.high.tolist() == [1500.0, 2200.0, 25.0, -150.0, 5.0, 5.0, 15.0, 10.0]
    assert env.max_episode_steps == 1000
    obs, info = env.reset(seed=123)
    assert obs.dtype == np.float32 and obs.shape == (8,) and env.observation_space.contains(obs)
    assert set(info) == {"raw_state", "speed", "altitude", "steps"} and info["steps"] == 0
    assert {"x", "y", "vx", "vy", "ax", "ay", "angle", "angularVelocity", "mass", "fuelMass", "speed",
            "relativeAngle", "totalMass"} <= set(info["raw_state"])
    out = env.step(np.array([0.5, 0.1], np.float32))
    assert len(out) == 5
    obs, reward, terminated, truncated, info = out
    assert isinstance(reward, float) and isinstance(terminated, bool) and isinstance(truncated, bool)
    assert info["steps"] == env.current_step == 1 and info["altitude"] == info["raw_state"]["y"]
    assert env.render() is None
    env.close()


def test_episodes_beside_the_oracle():
    """Three episodes with random actions.  Each starts from whatever the B200 env sampled; the oracle
    is injected with that state and both are stepped with the same actions."""
    import parity
    from rocket_landing_rl_b200 import RocketLandingEnv
    env = RocketLandingEnv(seed=11)
    ref = OracleEnv()
    rng = np.random.default_rng(0)
    scale = parity.SCALE[:8]                                                     # tolerances of tests/parity.py
    for episode in range(3):
        obs, info = env.reset()
        s = info["raw_state"]
        want = ref.inject(s["x"], s["y"], s["vx"], s["vy"], s["angle"], s["angularVelocity"], s["mass"], s["fuelMass"],
                          ax=s["ax"], ay=s["ay"])
        assert np.allclose(obs, want, rtol=0, atol=parity.RTOL * scale)
        for t in range(1000):
            a = rng.uniform([0, -1], [1, 1]).astype(np.float32)
            obs, rew, term, trunc, info = env.step(a)
            wobs, wrew, wterm, wtrunc, winfo = ref.step(a)
            assert (term, trunc) == (wterm, wtrunc), (episode, t)
            assert np.all(np.abs(obs - wobs) <= parity.RTOL * np.maximum(np.abs(wobs), scale)), (episode, t)
            assert abs(rew - wrew) <= max(parity.FREE_REWARD_ATOL, parity.FREE_REWARD_RTOL * abs(wrew))
            assert info["steps"] == winfo["steps"]
            # info["raw_state"] carries the UNCLIPPED accelerations (the engine fires: ay reaches +15 m/s^2)
            for k_ in ("ax", "ay"):
                want_a = getattr(ref.state, k_)
                assert abs(info["raw_state"][k_] - want_a) <= parity.RTOL * max(abs(want_a), 10.0), (episode, t, k_)
            if term or trunc:
                break
        assert term and 60 < t < 200          # random actions: every episode ends on the ground
    env.close()
