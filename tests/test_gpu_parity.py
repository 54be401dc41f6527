"""GPU tier: the CUDA library, called through the C ABI, against the reference's committed
outputs (tests/golden) and the C oracle -- tolerances and guard bands are stated in
tests/parity.py -- plus the contracts around the step: auto-reset, fused substeps, sharding
invariance, host-buffer (VecEnv) API, statistics, checkpoint, CUDA-graph replay, and
size-independent properties at BASELINE.json's full size."""
import os

import numpy as np
import pytest

import parity
from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available()
    return t


def _rand_actions(torch, n, gen):
    a = torch.rand((n, 2), device="cuda", generator=gen)
    a[:, 1] = a[:, 1] * 2.0 - 1.0
    return a


# ---------------------------------------------------------------------------------------------
# parity with the reference
# ---------------------------------------------------------------------------------------------
def test_library_loaded_is_the_in_tree_cuda_build(torch):
    from rocket_landing_rl_b200 import _lib
    _lib.load()
    maps = open("/proc/self/maps").read()
    assert _lib.LIB_PATH in maps, "the in-tree CUDA library must be the one that is loaded"


def test_reference_single_env_trace(torch, golden_dir):
    """BASELINE.json configs[0]: single rocket, 1000-step fixed-seed random-action rollout."""
    g = dict(np.load(os.path.join(golden_dir, "ref_rollout_single.npz")))
    g2 = {k: (v[:, None] if v.ndim >= 1 and k not in ("init_state", "init_obs") else v[None])
          for k, v in g.items()}
    rep = parity.replay_golden_rollout(parity.GpuStepper(1), g2)
    print(rep.summary())
    rep.assert_ok(max_guard_events=0)
    assert rep.episodes == 8


def test_reference_batch_rollouts(torch, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "ref_rollout_batch8.npz")))
    rep = parity.replay_golden_rollout(parity.GpuStepper(8), g)
    print(rep.summary())
    rep.assert_ok(max_guard_events=1)
    assert rep.steps_compared >= 7990


def test_reference_edge_cases(torch, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "ref_edge_cases.npz")))
    rep = parity.replay_edge_cases(parity.GpuStepper(g["init_state"].shape[0]), g)
    print(rep.summary())
    rep.assert_ok(max_guard_events=2)
    assert (rep.landing_counts[1:] > 0).all()


def test_4096_rockets_1000_steps_against_oracle(torch):
    """BASELINE.json configs[1]: 4096 batched rockets, random actions, trajectory parity."""
    n = 4096
    rep = parity.lockstep_vs_oracle(parity.GpuStepper(n), c_oracle.BatchOracle(n), 1000, seed=2024)
    print(rep.summary())
    # ~34 000 episodes; SURVEY.md H2 expects ~1 guard-band touchdown disagreement in fp32
    rep.assert_ok(max_guard_events=12)
    assert rep.episodes > 30_000 and rep.steps_compared > 0.9995 * n * 1000


def test_controlled_descents_reach_every_landing_branch(torch):
    n = 2048
    rep = parity.lockstep_vs_oracle(parity.GpuStepper(n), c_oracle.BatchOracle(n), 700, seed=7,
                                    action_fn=parity.braking_actions)
    print(rep.summary())
    rep.assert_ok(max_guard_events=12)
    assert rep.landing_counts[1] > 100 and rep.landing_counts[2] > 10


# ---- long horizon: episodes flown to the 1000-step limit (lander.py:237-241), word by word -------
@pytest.mark.parametrize("action_fn", [parity.hover_actions, parity.tilted_cruise_actions])
def test_thousand_step_episodes_against_oracle(torch, action_fn):
    """north_star: trajectories "within a stated fp32-vs-fp64 tolerance (e.g. rel 1e-5 per step over
    1000 steps)".  512 rockets kept aloft for the full 1000 steps, free-running against the fp64 C
    oracle, RTOL = 1e-5 on every state / observation word of every step."""
    n = 512
    rep = parity.lockstep_vs_oracle(parity.GpuStepper(n), c_oracle.BatchOracle(n), 1000, seed=5,
                                    action_fn=action_fn)
    print(rep.summary())
    rep.assert_ok(max_guard_events=0)
    assert rep.steps_compared == n * 1000 and rep.episodes == n
    assert rep.max_state_err.max() < 0.5 and rep.max_obs_err < 0.5


@pytest.mark.parametrize("name", ["ref_long_agent_v3.npz", "ref_long_hover.npz"])
def test_thousand_step_reference_trajectories(torch, golden_dir, name):
    """Trajectories of the live reference flown by its own shipped v3 agent and by a hover
    controller (tests/golden/make_long_horizon.py), replayed through the CUDA path."""
    g = dict(np.load(os.path.join(golden_dir, name)))
    rep = parity.replay_golden_rollout(parity.GpuStepper(8), g)
    print(rep.summary())
    rep.assert_ok(max_guard_events=1)
    assert rep.steps_compared >= 8 * 1100 - 2
    assert rep.max_state_err.max() < 0.5 and rep.max_obs_err < 0.5


def test_another_configuration_flown_to_its_time_limit(torch):
    """The run-time-constant kernels on a configuration that shares no derived constant with
    config.yaml (dt 0.05, another fixed-point fuel unit, 1500-step limit): every rocket kept aloft to
    step 1500 against the oracle given the same constants."""
    p = parity.other_config()
    n = 256
    st = parity.GpuStepper(n, params=p)
    assert not st.env.uses_folded_constants and st.env.fuel_unit == 2.0 ** -11
    rep = parity.lockstep_vs_oracle(st, c_oracle.BatchOracle(n, parity.constants_from_params(p)), 1500, seed=9,
                                    action_fn=parity.hover_actions)
    print(rep.summary())
    rep.assert_ok(max_guard_events=0)
    assert rep.steps_compared == n * 1500 and rep.episodes == n
    assert rep.max_state_err.max() < 0.5 and rep.max_obs_err < 0.5


def test_gpu_and_host_build_of_the_device_math_agree(torch):
    """Same source (rl_device.cuh), two compilers: the Philox reset must be bit-identical,
    a step must agree to rounding."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n = 4096
    env = RocketLandingVecEnv(n, seed=99, env_id_offset=1000, auto_reset=False)
    obs, _ = env.reset()
    host = parity.HostEmuStepper(n)
    hobs = host.philox_reset(gid0=1000, seed=99, epoch=0)
    np.testing.assert_array_equal(env.get_state().cpu().numpy(), host.soa)
    np.testing.assert_array_equal(obs.cpu().numpy(), hobs)
    a = np.random.default_rng(0).uniform([0, -1], [1, 1], (n, 2)).astype(np.float32)
    o, r, te, tr, info = env.step(torch.from_numpy(a).cuda())
    ho, hr, hte, htr, hl = host.step(a)
    np.testing.assert_allclose(o.cpu().numpy(), ho, rtol=2e-6, atol=2e-5)
    np.testing.assert_allclose(r.cpu().numpy(), hr, rtol=1e-5, atol=1e-3)
    env.close()


# ---------------------------------------------------------------------------------------------
# contracts around the step
# ---------------------------------------------------------------------------------------------
def test_auto_reset_follows_the_vecenv_contract(torch):
    """a16: on done the returned observation is the RESET observation, the terminal one goes
    to info["terminal_observation"]; episode return/length are those SB3's Monitor reports."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n, T = 2048, 260
    auto = RocketLandingVecEnv(n, seed=5, auto_reset=True)
    twin = RocketLandingVecEnv(n, seed=5, auto_reset=False)       # same stream, no reset
    o0, _ = auto.reset()
    twin.reset()
    assert torch.equal(auto.get_state(), twin.get_state())
    gen = torch.Generator(device="cuda").manual_seed(1)
    ret = torch.zeros(n, device="cuda", dtype=torch.float64)
    length = torch.zeros(n, device="cuda", dtype=torch.int64)
    n_done = 0
    ret_sum = 0.0
    for t in range(T):
        a = _rand_actions(torch, n, gen)
        obs, rew, term, trunc, info = auto.step(a)
        tobs, trew, tterm, ttrunc, _ = twin.step(a)
        done = term | trunc
        assert torch.equal(term, tterm) and torch.equal(trunc, ttrunc) and torch.equal(rew, trew)
        ret += rew.double()
        length += 1
        # not done: same observation; done: terminal obs stashed, reset obs returned
        assert torch.equal(obs[~done], tobs[~done])
        if done.any():
            assert torch.equal(info["terminal_observation"][done], tobs[done])
            st = auto.get_reference_state()
            d = done.cpu().numpy()
            assert st["fresh"][d].all() and (st["steps"][d] == 0).all() and (st["ep_return"][d] == 0).all()
            o = obs[done].cpu().numpy()
            np.testing.assert_array_equal(o[:, 0], st["x"][d].astype(np.float32))
            np.testing.assert_array_equal(o[:, 1], np.clip(st["y"][d], 0, 2200).astype(np.float32))
            np.testing.assert_array_equal(o[:, 6], st["angle"][d].astype(np.float32))
            assert (np.abs(o[:, 4:6]) <= 5.0).all()
            np.testing.assert_allclose(info["episode_return"][done].double().cpu().numpy(),
                                       ret[done].cpu().numpy(), rtol=2e-5, atol=1e-2)
            assert torch.equal(info["episode_length"][done].long(), length[done])
            n_done += int(done.sum())
            ret_sum += float(info["episode_return"][done].double().sum())
            ret[done] = 0
            length[done] = 0
            # re-synchronise the twin with the freshly reset rockets
            twin.set_state(auto.get_state(), mask=done)
    stats = auto.stats()
    assert stats.episodes == n_done and stats.env_steps == n * T
    assert stats.safe + stats.good + stats.ok + stats.unsafe + stats.time_limit + stats.out_of_bounds == n_done
    assert stats.return_sum == pytest.approx(ret_sum, rel=1e-6)
    assert stats.nonfinite == 0 and 100 < stats.mean_length < 140
    auto.close()
    twin.close()


def test_fused_substeps_equal_repeated_single_steps(torch):
    """configs[2] semantics: K fused substeps == K launches of one step with the same action."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n, K = 8192, 8
    fused = RocketLandingVecEnv(n, seed=3, auto_reset=False, substeps=K)
    single = RocketLandingVecEnv(n, seed=3, auto_reset=False, substeps=1)
    fused.reset()
    single.reset()
    gen = torch.Generator(device="cuda").manual_seed(2)
    any_done = False
    for it in range(20):
        a = _rand_actions(torch, n, gen)
        fo, fr, ft, ftr, _ = fused.step(a)
        rew = torch.zeros(n, device="cuda")
        done = torch.zeros(n, device="cuda", dtype=torch.bool)
        obs = torch.zeros((n, 8), device="cuda")
        term = torch.zeros(n, device="cuda", dtype=torch.bool)
        for k in range(K):
            keep = single.get_state()
            o, r, te, tr, _ = single.step(a)
            if done.any():
                single.set_state(keep, mask=done)      # finished rockets must not advance
            live = ~done
            rew[live] += r[live]
            obs[live] = o[live]
            term[live] = te[live]
            done |= live & (te | tr)
        assert torch.equal(ft, term) and torch.equal(ft | ftr, done)
        assert torch.equal(fused.get_state(), single.get_state())
        assert torch.equal(fo, obs)
        assert torch.allclose(fr, rew, rtol=1e-6, atol=1e-5)
        any_done |= bool(done.any())
        if done.any():
            fused.reset(mask=done)
            single.set_state(fused.get_state(), mask=done)
    assert any_done
    fused.close()
    single.close()


def test_sharding_does_not_change_trajectories(torch):
    """configs[3] semantics: Philox is keyed by GLOBAL rocket id, so two shards with
    env_id_offset reproduce one unsharded batch bit for bit, resets included."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv, shard_range
    total = 6000
    whole = RocketLandingVecEnv(total, seed=77)
    (s0, c0), (s1, c1) = shard_range(total, 0, 2), shard_range(total, 1, 2)
    a_ = RocketLandingVecEnv(c0, seed=77, env_id_offset=s0)
    b_ = RocketLandingVecEnv(c1, seed=77, env_id_offset=s1)
    ow, _ = whole.reset()
    oa, _ = a_.reset()
    ob, _ = b_.reset()
    assert torch.equal(ow, torch.cat([oa, ob]))
    gen = torch.Generator(device="cuda").manual_seed(3)
    for t in range(200):
        act = _rand_actions(torch, total, gen)
        ow, rw, tw, _, _ = whole.step(act)
        oa, ra, ta, _, _ = a_.step(act[:c0].contiguous())
        ob, rb, tb, _, _ = b_.step(act[c0:].contiguous())
        assert torch.equal(ow, torch.cat([oa, ob])) and torch.equal(rw, torch.cat([ra, rb]))
        assert torch.equal(tw, torch.cat([ta, tb]))
    sw, sa, sb = whole.stats_tensor().clone(), a_.stats_tensor().clone(), b_.stats_tensor().clone()
    assert sw[0] > 0
    idx = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]               # integer-valued slots: exact
    assert torch.equal(sw[idx], (sa + sb)[idx])
    assert torch.allclose(sw, sa + sb, rtol=1e-9)
    for e in (whole, a_, b_):
        e.close()


@pytest.mark.parametrize("small_host", ["16384", "0"])
def test_host_buffer_vecenv_api_matches_device_api(torch, monkeypatch, small_host):
    """The reference-facing plugin call (numpy in, numpy out, SB3 VecEnv surface), through both
    host paths of rl_step_host: zero-copy pinned staging (small batches, the default at this size)
    and the chunked copy pipeline (RL_B200_SMALL_HOST=0 forces it)."""
    from rocket_landing_rl_b200 import RocketLandingSB3VecEnv, RocketLandingVecEnv
    n = 1024
    monkeypatch.setenv("RL_B200_SMALL_HOST", small_host)
    host = RocketLandingSB3VecEnv(n, seed=11)
    dev = RocketLandingVecEnv(n, seed=11)
    assert host.observation_space.shape == (8,) and host.action_space.shape == (2,)
    oh = host.reset()
    od, _ = dev.reset()
    assert oh.dtype == np.float32 and np.array_equal(oh, od.cpu().numpy())
    rng = np.random.default_rng(0)
    seen_done = 0
    for t in range(160):
        a = rng.uniform([0, -1], [1, 1], (n, 2)).astype(np.float32)
        oh, rh, dh, infos = host.step(a)
        od, rd, td, trd, info = dev.step(torch.from_numpy(a).cuda())
        assert np.array_equal(oh, od.cpu().numpy()) and np.array_equal(rh, rd.cpu().numpy())
        assert np.array_equal(dh, (td | trd).cpu().numpy())
        assert len(infos) == n
        for i in np.flatnonzero(dh):
            assert set(infos[i]) >= {"terminal_observation", "TimeLimit.truncated", "episode"}
            assert np.array_equal(infos[i]["terminal_observation"], info["terminal_observation"][i].cpu().numpy())
            assert infos[i]["episode"]["l"] == int(info["episode_length"][i])
            assert infos[i]["TimeLimit.truncated"] is False
            assert infos[i]["landing"] in ("safe", "good", "ok", "unsafe")
            seen_done += 1
        assert all(infos[i] == {} for i in np.flatnonzero(~dh)[:5])
    assert seen_done > 0.9 * n
    host.close()
    dev.close()


def test_device_and_host_api_interleave_on_one_handle(torch):
    """The device API enqueues on the caller's stream, the host API runs on private streams of the
    handle (include/rocket_landing_b200.h, "Conventions"): a host-API call must see everything the
    device API enqueued before it, with no synchronisation by the caller.  A long dependent chain
    is enqueued on a side stream (a slow kernel first, so that the work is still pending when the
    host call arrives), then the state is read through rl_get_state_host / stepped through
    rl_step_host and compared with a twin driven through one API only."""
    import ctypes as C
    from rocket_landing_rl_b200 import RocketLandingVecEnv, _lib
    L = _lib.load()
    n = 1 << 16
    mixed = RocketLandingVecEnv(n, seed=41)
    twin = RocketLandingVecEnv(n, seed=41)
    side = torch.cuda.Stream()
    gen = torch.Generator(device="cuda").manual_seed(6)
    acts = [_rand_actions(torch, n, gen) for _ in range(24)]
    big = torch.randn((8192, 8192), device="cuda")
    twin.reset()
    for a in acts:
        twin.step(a)
    want_state = twin.get_state().cpu().numpy()
    host_a = acts[0].cpu().numpy().copy()
    want = [x.cpu().numpy().copy() for x in twin.step(acts[0])[:4]]
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(6):
            big = big @ big * 1e-4                      # ~50 ms of work ahead of the env kernels on `side`
        mixed.reset()
        for a in acts:
            mixed.step(a)
    # no synchronisation here: the host API must order itself after `side`
    got_state = np.zeros((10, n), np.int32)
    _lib.check(L.rl_get_state_host(mixed._h, got_state.ctypes.data), "rl_get_state_host")
    np.testing.assert_array_equal(got_state, want_state)
    obs, rew = np.zeros((n, 8), np.float32), np.zeros(n, np.float32)
    te, tr = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    _lib.check(L.rl_step_host(mixed._h, host_a.ctypes.data, obs.ctypes.data, rew.ctypes.data, te.ctypes.data,
                              tr.ctypes.data, None, None, None, None, None, 1, 1), "rl_step_host")
    np.testing.assert_array_equal(obs, want[0])
    np.testing.assert_array_equal(rew, want[1])
    np.testing.assert_array_equal(te.astype(bool), want[2])
    # ... and back: a device-API call after the (synchronous) host call continues the same trajectory
    with torch.cuda.stream(side):
        o2 = mixed.step(acts[1])[0].clone()
    side.synchronize()
    assert torch.equal(o2, twin.step(acts[1])[0])
    assert mixed._L.rl_get_epoch(mixed._h) == twin._L.rl_get_epoch(twin._h) == 27      # 1 reset + 26 steps
    mixed.close()
    twin.close()


def test_reset_distribution_on_the_gpu(torch):
    """a1, get_initial_state (simulation/config.py:20-53, config.yaml:53-67): ten independent
    uniforms.  Through rl_reset AND through the in-step auto-reset, at 1 Mi rockets: per-word
    Kolmogorov-Smirnov distance against the uniform on the config range, pairwise independence
    of the ten words, and no draw repeated across (global id, epoch)."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv, layout
    n = 1 << 20
    p = RocketLandingVecEnv(8).params
    lo = np.array(list(p.reset_low))
    hi = np.array(list(p.reset_high))

    def words(env, obs):
        v = env.get_reference_state()
        o = obs.cpu().numpy().astype(np.float64)
        assert v["fresh"].all() and (v["steps"] == 0).all()
        # x y vx vy (state) ax ay (only ever in the reset observation, Q10) angle omega dry fuel
        return np.stack([v["x"], v["y"], v["vx"], v["vy"], o[:, 4], o[:, 5], v["angle"], v["ang_vel"],
                         v["dry_mass"], v["fuel"]], axis=1)

    def check(w, tag):
        u = (w - lo) / (hi - lo)
        assert u.min() >= -1e-6 and u.max() <= 1.0 + 1e-6, tag        # (v = fixed-point delta / dt)
        m = u.shape[0]
        grid = (np.arange(m) + 0.5) / m
        for k in range(10):
            ks = np.abs(np.sort(u[:, k]) - grid).max()
            assert ks < 1.95 / np.sqrt(m), (tag, k, ks)            # P(KS > 1.95 / sqrt(m)) = 0.1 % per word
        c = np.corrcoef(u.T)
        assert np.abs(c - np.eye(10)).max() < 5.0 / np.sqrt(m), (tag, np.abs(c - np.eye(10)).max())
        return u

    env = RocketLandingVecEnv(n, seed=2024, env_id_offset=5_000_000)
    obs, _ = env.reset()
    w0 = words(env, obs)
    u0 = check(w0, "rl_reset epoch 0")
    obs, _ = env.reset()
    w1 = words(env, obs)
    check(w1, "rl_reset epoch 1")
    # uniqueness across (id, epoch): the 24-bit x draws of two epochs coincide for ~n / 2^24 rockets by chance
    assert (w0[:, 0] == w1[:, 0]).mean() < 3.0 / (1 << 24) + 1e-5
    assert np.unique(np.round(u0[:, :4] * (1 << 24)).astype(np.int64), axis=0).shape[0] == n   # no two rockets alike
    # auto-reset inside the step: collect the FRESH rockets of many steps
    gen = torch.Generator(device="cuda").manual_seed(8)
    got = []
    a = _rand_actions(torch, n, gen)
    for t in range(140):
        obs, _, term, trunc, _ = env.step(a)
        if t >= 80:
            done = (term | trunc).cpu().numpy()
            if done.any():
                v = env.get_reference_state()
                o = obs.cpu().numpy().astype(np.float64)
                got.append(np.stack([v["x"], v["y"], v["vx"], v["vy"], o[:, 4], o[:, 5], v["angle"], v["ang_vel"],
                                     v["dry_mass"], v["fuel"]], axis=1)[done])
    wa = np.concatenate(got)
    assert wa.shape[0] > 200_000
    check(wa, "auto-reset")
    # a shard with the same seed and ids reproduces the words exactly; another seed does not
    same = RocketLandingVecEnv(4096, seed=2024, env_id_offset=5_000_000)
    np.testing.assert_array_equal(words(same, same.reset()[0]), w0[:4096])
    other = RocketLandingVecEnv(4096, seed=2025, env_id_offset=5_000_000)
    assert (words(other, other.reset()[0])[:, 0] != w0[:4096, 0]).mean() > 0.99
    for e in (env, same, other):
        e.close()


def test_batched_raw_state_matches_the_oracle(torch):
    """lander.py:157-166: info["raw_state"] of EVERY rocket after a step, the unclipped ax / ay
    included (the observation clips them to +-5 while ay reaches +15 under thrust)."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n = 4096
    env = RocketLandingVecEnv(n, seed=9, auto_reset=False, raw_state=True)
    orc = c_oracle.BatchOracle(n)
    rng = np.random.default_rng(1)
    init = parity.sample_reset_states(rng, env.params, n)
    env.set_state(torch.from_numpy(parity._fresh_soa(init, 0, env.fuel_unit, env.dt, env.pos_unit)))
    orc.inject(init)
    big = 0
    for t in range(40):
        a = rng.uniform([0, -1], [1, 1], (n, 2)).astype(np.float32)
        _, _, _, _, info = env.step(torch.from_numpy(a).cuda())
        orc.step(a)
        rs = env.raw_state()
        ref = orc.state()
        for k_, col, sc in (("x", 0, 1500.0), ("y", 1, 2000.0), ("vx", 2, 25.0), ("vy", 3, 200.0), ("ax", 4, 10.0),
                            ("ay", 5, 10.0), ("angle", 6, 15.0), ("angularVelocity", 7, 10.0), ("mass", 8, 36000.0),
                            ("fuelMass", 9, 390000.0)):
            tol = parity.RTOL * np.maximum(np.abs(ref[:, col]), sc)
            assert (np.abs(rs[k_] - ref[:, col]) <= tol).all(), (t, k_)
        big += int((np.abs(rs["ay"]) > 5.0).sum())
        np.testing.assert_allclose(rs["totalMass"], ref[:, 8] + ref[:, 9], rtol=1e-6)
    assert big > 1000            # the unclipped values really are beyond the observation's +-5
    env.close()


def test_error_paths_through_the_abi(torch):
    from rocket_landing_rl_b200 import RocketLandingVecEnv, _lib
    env = RocketLandingVecEnv(64)
    with pytest.raises(ValueError):
        env.step(torch.zeros((63, 2), device="cuda"))
    env.substeps = 0
    with pytest.raises(_lib.RocketLibraryError, match="substeps"):
        env.step(torch.zeros((64, 2), device="cuda"))
    env.substeps = 1
    with pytest.raises(TypeError):
        env.set_state(torch.zeros((10, 64), dtype=torch.float64))
    env.close()
    env.close()          # idempotent


def test_checkpoint_resume_is_bit_identical(torch):
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n = 4096
    gen = torch.Generator(device="cuda").manual_seed(9)
    acts = [_rand_actions(torch, n, gen) for _ in range(300)]
    a_ = RocketLandingVecEnv(n, seed=21)
    a_.reset()
    for t in range(150):
        a_.step(acts[t])
    ckpt = a_.state_dict()
    tail_a = [tuple(x.clone() for x in a_.step(acts[t])[:4]) for t in range(150, 300)]
    b_ = RocketLandingVecEnv(n, seed=21)
    b_.load_state_dict(ckpt)
    for t in range(150, 300):
        got = b_.step(acts[t])[:4]
        for x, y in zip(got, tail_a[t - 150]):
            assert torch.equal(x, y)
    a_.close()
    b_.close()


def test_step_is_cuda_graph_capturable_and_replay_advances_the_reset_stream(torch):
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n = 4096
    eager = RocketLandingVecEnv(n, seed=31)
    graphed = RocketLandingVecEnv(n, seed=31)
    eager.reset()
    graphed.reset()
    gen = torch.Generator(device="cuda").manual_seed(4)
    static_a = torch.zeros((n, 2), device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        static_a.copy_(_rand_actions(torch, n, gen))
        with torch.cuda.graph(g, stream=s):
            graphed.step(static_a)
    torch.cuda.current_stream().wait_stream(s)
    gen = torch.Generator(device="cuda").manual_seed(4)
    resets = 0
    for t in range(180):
        a = _rand_actions(torch, n, gen)
        static_a.copy_(a)
        g.replay()                     # (the capture itself did not execute the step)
        o, r, te, tr, _ = eager.step(a)
        assert torch.equal(graphed.obs, o) and torch.equal(graphed.reward, r)
        assert torch.equal(graphed.terminated, te)
        resets += int((te | tr).sum())
    assert resets > 0.9 * n  # resets happened and (since obs match) used fresh Philox epochs
    eager.close()
    graphed.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json full size: size-independent properties (configs[2]: 16 Mi rockets)
# ---------------------------------------------------------------------------------------------
def test_full_size_properties_16m_rockets(torch, monkeypatch):
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n = 16 * 1024 * 1024
    env = RocketLandingVecEnv(n, seed=0)
    obs, _ = env.reset()
    lo = torch.tensor(list(env.params.obs_low), device="cuda")
    hi = torch.tensor(list(env.params.obs_high), device="cuda")
    assert bool(((obs >= lo) & (obs <= hi)).all())
    gen = torch.Generator(device="cuda").manual_seed(0)
    a = _rand_actions(torch, n, gen)
    fuel_prev = env.get_state()[7].view(torch.int32).clone()       # fixed-point fuel word
    T = 130
    checksum = torch.zeros((), device="cuda", dtype=torch.int64)
    done_total = 0
    for t in range(T):
        obs, rew, term, trunc, info = env.step(a)
        done = term | trunc
        done_total += int(done.sum())
        if t % 16 == 0:
            assert bool(((obs >= lo) & (obs <= hi)).all())           # observation clipping (lander.py:149)
            st = env.get_state()
            # angle normalisation (engine.py:230): the int32 angle word spans exactly [-180, 180), so
            # what remains to check is the value the observation reports
            assert bool(((obs[:, 6] >= -15.0) & (obs[:, 6] <= 15.0)).all())
            fuel = st[7].view(torch.int32)
            assert bool(((fuel <= fuel_prev) | done).all()) or t == 0   # fuel only burns within an episode
            assert bool((fuel >= 0).all())
            assert bool(torch.isfinite(rew).all())
        fuel_prev = env.get_state()[7].view(torch.int32).clone() if (t + 1) % 16 == 0 else fuel_prev
        checksum += obs.view(torch.int32).sum(dtype=torch.int64) + rew.view(torch.int32).sum(dtype=torch.int64)
    stats = env.stats()
    assert stats.env_steps == n * T and stats.episodes == done_total
    assert stats.nonfinite == 0
    # constant action for everybody: episode lengths stay near the reference's 77..154 range
    assert 0.5 * n < stats.episodes < 2.0 * n and 60 < stats.mean_length < 160
    # determinism: same seed, same actions -> identical checksum of every output
    env2 = RocketLandingVecEnv(n, seed=0)
    env2.reset()
    checksum2 = torch.zeros((), device="cuda", dtype=torch.int64)
    for t in range(T):
        obs, rew, _, _, _ = env2.step(a)
        checksum2 += obs.view(torch.int32).sum(dtype=torch.int64) + rew.view(torch.int32).sum(dtype=torch.int64)
    assert int(checksum) == int(checksum2)
    env.close()
    env2.close()
    # ... and the same bits from the other grid shape (this size runs 8 consecutive tiles per CTA
    # by default; force the persistent strided grid)
    monkeypatch.setenv("RL_B200_TILES_PER_CTA", "0")
    env3 = RocketLandingVecEnv(n, seed=0)
    env3.reset()
    checksum3 = torch.zeros((), device="cuda", dtype=torch.int64)
    for t in range(T):
        obs, rew, _, _, _ = env3.step(a)
        checksum3 += obs.view(torch.int32).sum(dtype=torch.int64) + rew.view(torch.int32).sum(dtype=torch.int64)
    assert int(checksum) == int(checksum3)
    env3.close()


# ---------------------------------------------------------------------------------------------
# the two instantiations of the step kernel (constants folded in / read at run time)
# ---------------------------------------------------------------------------------------------
def _generic_env(cls, *args, **kwargs):
    os.environ["RL_B200_GENERIC_KERNEL"] = "1"
    try:
        return cls(*args, **kwargs)
    finally:
        del os.environ["RL_B200_GENERIC_KERNEL"]


def test_default_config_uses_folded_constants_and_other_configs_do_not(torch):
    from rocket_landing_rl_b200 import RocketLandingVecEnv, load_params
    env = RocketLandingVecEnv(64)
    assert env.uses_folded_constants
    env.close()
    p = load_params()
    p.max_episode_steps = 50
    env = RocketLandingVecEnv(64, params=p)
    assert not env.uses_folded_constants
    env.reset()
    a = torch.zeros((64, 2), device="cuda")
    for t in range(50):
        obs, rew, term, trunc, info = env.step(a)
        assert bool(trunc.all()) == (t == 49) and not bool(term.any())      # lander.py:237-241
    assert bool((info["episode_length"] == 50).all())
    env.close()
    env = _generic_env(RocketLandingVecEnv, 64)
    assert not env.uses_folded_constants
    env.close()


@pytest.mark.parametrize("substeps", [1, 4])
def test_generic_and_folded_kernels_are_bitwise_identical(torch, substeps):
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n = 20_000                                     # ragged: not a multiple of 32
    folded = RocketLandingVecEnv(n, seed=13, substeps=substeps)
    generic = _generic_env(RocketLandingVecEnv, n, seed=13, substeps=substeps)
    assert folded.uses_folded_constants and not generic.uses_folded_constants
    of, _ = folded.reset()
    og, _ = generic.reset()
    assert torch.equal(of, og)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(300 // substeps):
        a = _rand_actions(torch, n, gen)
        rf = folded.step(a)
        rg = generic.step(a)
        for x, y in zip(rf[:4], rg[:4]):
            assert torch.equal(x, y), t
        for key in ("terminal_observation", "episode_return", "episode_length", "landing"):
            assert torch.equal(rf[4][key], rg[4][key]), (t, key)
    assert torch.equal(folded.get_state(), generic.get_state())
    sf, sg = folded.stats_tensor().clone(), generic.stats_tensor().clone()
    exact = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]            # integer-valued slots
    assert torch.equal(sf[exact], sg[exact]) and sf[0] > 0
    assert torch.allclose(sf, sg, rtol=1e-9)               # float sums: atomics order differs
    folded.close()
    generic.close()


def test_generic_kernel_parity_with_the_reference(torch, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "ref_rollout_batch8.npz")))
    st = _generic_env(parity.GpuStepper, 8)
    assert not st.env.uses_folded_constants
    rep = parity.replay_golden_rollout(st, g)
    rep.assert_ok(max_guard_events=1)
    n = 2048
    st = _generic_env(parity.GpuStepper, n)
    rep = parity.lockstep_vs_oracle(st, c_oracle.BatchOracle(n), 400, seed=5)
    print(rep.summary())
    rep.assert_ok(max_guard_events=8)


@pytest.mark.parametrize("tiles_per_cta", [1, 3, 8])
def test_grid_shapes_are_bitwise_identical(torch, monkeypatch, tiles_per_cta):
    """The step kernel runs either a persistent strided grid (small launches) or `chunk`
    consecutive tiles per CTA in address order (large launches).  Both
    instantiations must produce the same bits: 70 001 rockets (ragged last tile), 260 steps with
    auto-reset, so that every rocket finishes at least one episode."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    n = 70001
    g = torch.Generator(device="cuda").manual_seed(17)
    acts = [torch.rand((n, 2), device="cuda", generator=g) for _ in range(4)]
    for a in acts:
        a[:, 1].mul_(2).sub_(1)

    def run(env):
        env.reset()
        ret = torch.zeros(n, device="cuda", dtype=torch.float64)
        flags = torch.zeros(n, device="cuda", dtype=torch.int64)
        for t in range(260):
            obs, rew, term, trunc, info = env.step(acts[t % 4])
            ret += rew.double()
            flags += term.long() + 2 * trunc.long() + 4 * info["landing"].long()
        torch.cuda.synchronize()
        return env.get_state().clone(), obs.clone(), ret, flags, info["terminal_observation"].clone(), env.stats()

    monkeypatch.setenv("RL_B200_TILES_PER_CTA", "0")          # persistent strided grid
    ref = run(RocketLandingVecEnv(n, device="cuda", seed=5))
    monkeypatch.setenv("RL_B200_TILES_PER_CTA", str(tiles_per_cta))
    got = run(RocketLandingVecEnv(n, device="cuda", seed=5))
    for a, b in zip(ref[:5], got[:5]):
        assert torch.equal(a, b)
    assert ref[5].episodes == got[5].episodes > n
    assert ref[5].env_steps == got[5].env_steps == 260 * n
    assert ref[5].return_sum == pytest.approx(got[5].return_sum, rel=1e-6)    # fp32 partial sums grouped differently


def test_ragged_batch_sizes(torch):
    """Sizes around the warp-tile / CTA / padding boundaries, including a single rocket."""
    from rocket_landing_rl_b200 import RocketLandingVecEnv
    big = RocketLandingVecEnv(1000, seed=3)
    ob = big.reset()[0].clone()
    acts = torch.rand((1000, 2), device="cuda")
    want = [tuple(x.clone() for x in big.step(acts)[:4]) for _ in range(3)]
    big.close()
    for n in (1, 31, 32, 33, 255, 256, 257, 999):
        env = RocketLandingVecEnv(n, seed=3)
        o, _ = env.reset()
        assert torch.equal(o, ob[:n]), n
        canary = torch.full((n + 64, 8), -7.0, device="cuda")     # the kernel must not write past row n
        env.obs = canary[:n]
        for k in range(3):
            got = env.step(acts[:n].contiguous())[:4]
            for x, y in zip(got, want[k]):
                assert torch.equal(x, y[:n]), (n, k)
        assert bool((canary[n:] == -7.0).all()), n
        env.close()


def test_pipelined_host_path_equals_device_path_across_chunks(torch):
    """rl_step_host cuts large batches into chunks on two streams; every chunk reads the same
    epoch snapshot, so resets -- and everything else -- equal one rl_step over the whole batch."""
    from rocket_landing_rl_b200 import RocketLandingSB3VecEnv, RocketLandingVecEnv
    n = 3 * 4096 * 256 + 77                      # three chunks, ragged tail
    host = RocketLandingSB3VecEnv(n, seed=17, build_infos=False)
    dev = RocketLandingVecEnv(n, seed=17)
    assert np.array_equal(host.reset(), dev.reset()[0].cpu().numpy())
    a = torch.rand((n, 2), generator=torch.Generator().manual_seed(0))
    a[:, 1] = a[:, 1] * 2 - 1
    a_np = a.numpy()
    a_dev = a.cuda()
    resets = 0
    for t in range(112):
        oh, rh, th, trh = host.step_raw(a_np)
        od, rd, td, trd, _ = dev.step(a_dev)
        if t % 8 == 0 or t > 95:
            assert np.array_equal(oh, od.cpu().numpy()), t
            assert np.array_equal(rh, rd.cpu().numpy()), t
            assert np.array_equal(th.astype(bool), td.cpu().numpy()), t
        resets += int(th.sum()) + int(trh.sum())
    assert resets > 0.5 * n
    host.close()
    dev.close()
    torch.cuda.synchronize()
    sh = RocketLandingSB3VecEnv(64, seed=1)
    assert sh.launch_count == 0
    sh.close()
