"""Parity harness: steps a "stepper" (the CUDA library on the GPU, or the host build of the
device math in the CPU tier) in lock step with reference outputs (committed goldens, or
the C oracle) and applies the stated tolerances.

Two legs run on every step.

FREE-RUNNING (the stepper's trajectory against the reference's from the same injected
reset state; errors accumulate over the whole episode -- ~120 steps under random actions, the
full 1000 (``lander.py:237-241``) under the hover / tilted-cruise controllers and in the
reference's own agent-flown trajectories of ``tests/golden/ref_long_*.npz``):

* state / observation words: ``|got - ref| <= RTOL * max(|ref|, SCALE[word])``, ``RTOL = 1e-5``
  (the bar of BASELINE.json ``north_star``: "rel 1e-5 per step", scale-relative as in
  SURVEY.md H3).
* flags (terminated / truncated) and landing class: EXACT, except where the reference's own
  decision variable sits inside a declared guard band (``GUARD``) of the threshold it is
  compared with: there a +-1-step disagreement is tolerated, counted and reported (SURVEY.md
  H2: the touchdown test ``y <= 0.1`` is crossed at ~20 m/step, so an altitude error of d
  flips the step with probability d / 20 m).  After such a disagreement -- or an angle wrap
  decided differently inside its guard band -- the rocket is excluded until the next
  injected reset re-synchronises it.
* reward: ``|got - ref| <= max(FREE_REWARD_ATOL, FREE_REWARD_RTOL * |ref|)``.  Loose on purpose:
  the reward amplifies state error (``8 / max(y, 1)`` near the ground, reward.py:192;
  ``effect**2``, reward.py:160), so this leg bounds the PROPAGATED error only.  The shaping
  reward is DISCONTINUOUS in the state at a handful of thresholds (``REWARD_STEPS`` below); where
  the reference's own decision variable sits inside the guard band of one of them the step is
  counted (``reward_guard_events``) instead of compared -- a stabilising controller parks the
  angle at 0 +- 1e-3 deg, where ``correct_direction`` (reward.py:130-146) flips the reward by
  +-0.7 with the sign of an angle the two sides agree on to 1e-6 deg.  The TEACHER-FORCED leg
  checks those very branches with no guard band.

TEACHER-FORCED (the reference's reward / termination / landing functions evaluated in
fp64 by the oracle on the stepper's OWN before/after state, widened from fp32):

* flags and landing class: EXACT, no guard band -- the thresholds in the kernel are rounded
  in the direction that makes every fp32 comparison agree with the fp64 one.
* reward: ``|got - ref| <= max(TF_REWARD_ATOL, TF_REWARD_RTOL * |ref|)`` with
  ``TF_REWARD_RTOL = 1e-5``: the fp32 evaluation error of the reward formula itself.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from dataclasses import dataclass, field

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from rocket_landing_rl_b200 import layout            # noqa: E402
from rocket_landing_rl_b200.config import load_params, RlParams   # noqa: E402
from oracle import c_oracle                          # noqa: E402

RTOL = 1e-5
# reference word order of info["raw_state"]: x y vx vy ax ay angle angularVelocity mass fuelMass
SCALE = np.array([1500.0, 2000.0, 25.0, 200.0, 10.0, 10.0, 15.0, 10.0, 36000.0, 390000.0])
FREE_REWARD_ATOL, FREE_REWARD_RTOL = 2e-2, 2e-3
TF_REWARD_ATOL, TF_REWARD_RTOL = 2e-4, 1e-5
GUARD = {"y": 1e-2, "v": 1e-3, "angle": 1e-3, "wrap": 1e-2, "bounds": 0.5, "angle_sign": 5e-5}
# Thresholds at which calculate_reward jumps (reward.py line -> decision variable -> threshold):
#   :130      |angle_before| > 0.1, |omega_before| > 0.1   (dead band of the cold-gas term)
#   :130-146  sign(angle_before)                            (correct_direction bonus vs penalty)
#   :171-227  y_after > 0.1                                 (airborne terms; also the touchdown test)
#   :205-219  |angle_after| > 10                            (tilt penalty 0.5 th -> th |a| / 90)
#   :278      |angle_after| > 90                            (tip-over, -300)
# (vy crossing 0 and |angle| crossing 5 are continuous: the terms they gate vanish there.)
REWARD_STEPS = {"angle_before": (0.0, 0.1), "omega_before": (0.1,), "angle_after": (10.0, 90.0)}
LAND_V = (30.0, 40.0, 100.0)
LAND_A = (5.0, 8.0, 10.0)


def _fresh_soa(states: np.ndarray, steps=0, fuel_unit=layout.DEFAULT_FUEL_UNIT, dt=0.1,
               pos_unit=layout.DEFAULT_POS_UNIT) -> np.ndarray:
    s = np.asarray(states, np.float64)
    return layout.fresh_state_soa(s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 6], s[:, 7], s[:, 8],
                                  s[:, 9], steps=steps, fuel_unit=fuel_unit, dt=dt, pos_unit=pos_unit)


# ----------------------------------------------------------------------------------------
# steppers
# ----------------------------------------------------------------------------------------
class HostEmuStepper:
    """Host build of rl_device.cuh (tests/hostemu) -- CPU logic check only, never shipped."""

    def __init__(self, n: int, params: RlParams | None = None, folded: bool = False):
        import subprocess
        self.folded = folded
        here = os.path.join(_ROOT, "tests", "hostemu")
        so, src = os.path.join(here, "libhostemu.so"), os.path.join(here, "hostemu.cpp")
        deps = [src] + [os.path.join(_ROOT, "rocket-landing-rl_b200", "csrc", f)
                        for f in ("rl_device.cuh", "rl_params.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared",
                                   "-ffp-contract=off", "-o", so, src])
        self.L = C.CDLL(so)
        vp = C.c_void_p
        self.L.emu_step.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        self.L.emu_matches_default.argtypes = [vp]
        self.L.emu_reset.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, vp]
        self.L.emu_philox.argtypes = [vp, vp, vp]
        self.n = n
        self.params = params or load_params()
        self.dt = self.params.time_step
        self.fuel_unit = layout.fuel_unit_for(self.params.reset_high[9])
        self.pos_unit = layout.pos_unit_for(self.params)
        self.soa = np.zeros((10, n), np.int32)

    def inject_fresh(self, states: np.ndarray, mask=None, steps=0) -> None:
        new = _fresh_soa(states, steps, self.fuel_unit, self.dt, self.pos_unit)
        if mask is None:
            self.soa[:] = new
        else:
            self.soa[:, mask] = new[:, mask]

    def inject_soa(self, soa: np.ndarray) -> None:
        self.soa[:] = layout.as_words(soa)

    def step(self, actions: np.ndarray, substeps: int = 1):
        a = np.ascontiguousarray(actions, np.float32)
        obs = np.zeros((self.n, 8), np.float32)
        rew = np.zeros(self.n, np.float32)
        term = np.zeros(self.n, np.uint8)
        trunc = np.zeros(self.n, np.uint8)
        land = np.zeros(self.n, np.int8)
        rc = self.L.emu_step(C.addressof(self.params), self.soa.ctypes.data, self.n,
                             a.ctypes.data, obs.ctypes.data, rew.ctypes.data, term.ctypes.data,
                             trunc.ctypes.data, land.ctypes.data, substeps, int(self.folded))
        assert rc == 0, rc
        return obs, rew, term.astype(bool), trunc.astype(bool), land

    def view(self) -> dict:
        return layout.reference_view(self.soa, self.dt, self.fuel_unit, self.pos_unit)

    def philox_reset(self, gid0: int, seed: int, epoch: int):
        obs = np.zeros((self.n, 8), np.float32)
        rc = self.L.emu_reset(C.addressof(self.params), self.soa.ctypes.data, self.n, gid0, seed,
                              epoch, obs.ctypes.data)
        assert rc == 0
        return obs

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        out = (C.c_uint32 * 4)()
        self.L.emu_philox(c, k, out)
        return list(out)


class GpuStepper:
    """The product: RocketLandingVecEnv (CUDA through the C ABI), auto-reset off so that the
    harness injects the reference's own reset states."""

    def __init__(self, n: int, params: RlParams | None = None, device="cuda:0"):
        import torch
        from rocket_landing_rl_b200 import RocketLandingVecEnv
        self.torch = torch
        self.env = RocketLandingVecEnv(n, device=device, auto_reset=False, params=params)
        self.n = n
        self.params = self.env.params
        self.dt = self.env.dt

    def inject_fresh(self, states: np.ndarray, mask=None, steps=0) -> None:
        self.env.set_state(self.torch.from_numpy(_fresh_soa(states, steps, self.env.fuel_unit, self.dt, self.env.pos_unit)),
                           None if mask is None else self.torch.from_numpy(np.asarray(mask, np.uint8)))

    def inject_soa(self, soa: np.ndarray) -> None:
        self.env.set_state(self.torch.from_numpy(layout.as_words(soa)))

    def step(self, actions: np.ndarray, substeps: int = 1):
        self.env.substeps = substeps
        a = self.torch.from_numpy(np.ascontiguousarray(actions, np.float32)).to(self.env.device)
        obs, rew, term, trunc, info = self.env.step(a)
        return (obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy().astype(bool),
                trunc.cpu().numpy().astype(bool), info["landing"].cpu().numpy())

    def view(self) -> dict:
        return self.env.get_reference_state()


# ----------------------------------------------------------------------------------------
# comparison
# ----------------------------------------------------------------------------------------
@dataclass
class ParityReport:
    steps_compared: int = 0
    max_state_err: np.ndarray = field(default_factory=lambda: np.zeros(10))   # units of tolerance
    max_obs_err: float = 0.0
    max_free_reward_err: float = 0.0     # units of the free-running reward tolerance
    max_tf_reward_err: float = 0.0       # units of the teacher-forced reward tolerance
    tf_steps: int = 0
    flag_guard_events: int = 0           # free-running disagreements inside a guard band
    wrap_guard_events: int = 0
    reward_guard_events: int = 0         # steps whose free-running reward sits on a reward discontinuity
    flag_violations: list = field(default_factory=list)     # outside any guard band
    value_violations: list = field(default_factory=list)
    tf_violations: list = field(default_factory=list)       # teacher-forced: any disagreement
    episodes: int = 0
    landing_counts: np.ndarray = field(default_factory=lambda: np.zeros(5, np.int64))

    def ok(self) -> bool:
        return not (self.flag_violations or self.value_violations or self.tf_violations)

    def summary(self) -> str:
        names = "x y vx vy ax ay angle omega dry fuel".split()
        errs = ", ".join(f"{n}={e:.3f}" for n, e in zip(names, self.max_state_err) if e > 0)
        return (f"{self.steps_compared} rocket-steps free-running, {self.tf_steps} teacher-forced, "
                f"{self.episodes} episodes, landings(none,safe,good,ok,unsafe)={self.landing_counts.tolist()}; "
                f"max err / tol: state[{errs}] obs={self.max_obs_err:.3f} free-reward="
                f"{self.max_free_reward_err:.3f} tf-reward={self.max_tf_reward_err:.3f}; guard-band "
                f"events: flags {self.flag_guard_events}, wraps {self.wrap_guard_events}, reward steps "
                f"{self.reward_guard_events}; VIOLATIONS: "
                f"flags {len(self.flag_violations)}, values {len(self.value_violations)}, "
                f"teacher-forced {len(self.tf_violations)}")

    def assert_ok(self, max_guard_events: int | None = None) -> None:
        msg = self.summary() + "\n" + "\n".join(
            str(v) for v in (self.flag_violations + self.value_violations + self.tf_violations)[:12])
        assert self.ok(), msg
        if max_guard_events is not None:
            assert self.flag_guard_events + self.wrap_guard_events <= max_guard_events, msg


def _near_landing_threshold(vx, vy, ang) -> np.ndarray:
    near = np.zeros(vx.shape, bool)
    for t in LAND_V:
        near |= (np.abs(np.abs(vx) - t) < GUARD["v"]) | (np.abs(np.abs(vy) - t) < GUARD["v"])
    for t in LAND_A:
        near |= np.abs(np.abs(ang) - t) < GUARD["angle"]
    return near


class LockstepChecker:
    """Feed one step of (got, ref) at a time; keeps the per-rocket "in sync" mask."""

    def __init__(self, n: int, constants=None):
        self.n = n
        self.synced = np.ones(n, bool)
        self.rep = ParityReport()
        self.constants = constants          # oracle EnvConstants of a non-default configuration (None: config.yaml)

    def resync(self, mask) -> None:
        self.synced[mask] = True

    # ---- teacher-forced leg: oracle functions on the stepper's own states -------------------
    def teacher_forced(self, t, before_view, after_view, actions, got_rew, got_term, got_trunc,
                       got_land) -> None:
        rep = self.rep
        rew, term, trunc, land = c_oracle.reward_on_states(before_view, after_view, actions, self.constants)
        bad = (term != got_term) | (trunc != got_trunc) | (land != got_land)
        for i in np.flatnonzero(bad):
            rep.tf_violations.append((t, int(i), "tf-flags", (bool(got_term[i]), bool(got_trunc[i]), int(got_land[i])),
                                      (bool(term[i]), bool(trunc[i]), int(land[i])), float(after_view["y"][i])))
        tol = np.maximum(TF_REWARD_ATOL, TF_REWARD_RTOL * np.abs(rew))
        err = np.abs(got_rew.astype(np.float64) - rew) / tol
        err[bad] = 0.0
        finite = np.isfinite(rew)
        err = np.where(finite, err, 0.0)
        rep.max_tf_reward_err = max(rep.max_tf_reward_err, float(err.max()))
        for i in np.flatnonzero(err > 1.0):
            rep.tf_violations.append((t, int(i), "tf-reward", float(got_rew[i]), float(rew[i])))
        rep.tf_steps += self.n

    # ---- free-running leg --------------------------------------------------------------------
    def compare(self, t, got_obs, got_rew, got_term, got_trunc, got_land, got_view,
                ref_state, ref_obs, ref_rew, ref_term, ref_trunc, ref_landing, ref_before) -> None:
        """``ref_before`` [n,10]: the reference's state BEFORE this step (reference word order)."""
        rep = self.rep
        ref_state = np.asarray(ref_state, np.float64)
        ref_before = np.asarray(ref_before, np.float64)
        ref_y_before = ref_before[:, 1]
        x, y, vx, vy, ang = (ref_state[:, 0], ref_state[:, 1], ref_state[:, 2], ref_state[:, 3],
                             ref_state[:, 6])
        live = self.synced
        term_bad = got_term != ref_term
        trunc_bad = got_trunc != ref_trunc
        near_ground = (np.abs(y - 0.1) < GUARD["y"]) | (np.abs(ref_y_before - 0.1) < GUARD["y"])
        near_bounds = ((np.abs(np.abs(x) - 50000.0) < GUARD["bounds"])
                       | (np.abs(y - 50000.0) < GUARD["bounds"]))
        for i in np.flatnonzero(live & term_bad):
            if near_ground[i]:
                rep.flag_guard_events += 1
            else:
                rep.flag_violations.append((t, int(i), "terminated", bool(got_term[i]), bool(ref_term[i]), float(y[i])))
        for i in np.flatnonzero(live & trunc_bad & ~term_bad):
            if near_bounds[i]:
                rep.flag_guard_events += 1
            else:
                rep.flag_violations.append((t, int(i), "truncated", bool(got_trunc[i]), bool(ref_trunc[i]), float(x[i]), float(y[i])))
        flag_bad = term_bad | trunc_bad
        # landing class (only meaningful where both agree the rocket touched down)
        class_bad = ref_term & got_term & (np.asarray(got_land) != np.asarray(ref_landing))
        class_guard = _near_landing_threshold(vx, vy, ang)
        for i in np.flatnonzero(live & class_bad):
            if class_guard[i]:
                rep.flag_guard_events += 1
            else:
                rep.flag_violations.append((t, int(i), "landing", int(got_land[i]), int(ref_landing[i]),
                                            float(vx[i]), float(vy[i]), float(ang[i])))
        # angle wrap decided differently inside its guard band
        got_ang = got_view["angle"]
        wrap_bad = np.abs(got_ang - ang) > 180.0
        for i in np.flatnonzero(live & wrap_bad & ~flag_bad):
            if 180.0 - abs(ang[i]) < GUARD["wrap"]:
                rep.wrap_guard_events += 1
            else:
                rep.value_violations.append((t, int(i), "angle-wrap", float(got_ang[i]), float(ang[i])))

        cmp = live & ~flag_bad & ~wrap_bad
        idx = np.flatnonzero(cmp)
        if idx.size:
            got_vec = np.stack([got_view["x"], got_view["y"], got_view["vx"], got_view["vy"],
                                np.zeros(self.n), np.zeros(self.n), got_view["angle"],
                                got_view["ang_vel"], got_view["dry_mass"], got_view["fuel"]], axis=1)
            tol = RTOL * np.maximum(np.abs(ref_state), SCALE)
            err = np.abs(got_vec - ref_state) / tol
            err[:, 4:6] = 0.0                               # ax, ay are step outputs: checked via obs
            e = err[idx]
            rep.max_state_err = np.maximum(rep.max_state_err, e.max(axis=0))
            for j, w in zip(*np.nonzero(e > 1.0)):
                rep.value_violations.append((t, int(idx[j]), f"state[{w}]", float(got_vec[idx[j], w]),
                                             float(ref_state[idx[j], w])))
            ref_obs64 = ref_obs.astype(np.float64)
            otol = RTOL * np.maximum(np.abs(ref_obs64), SCALE[:8])
            oe = (np.abs(got_obs.astype(np.float64) - ref_obs64) / otol)[idx]
            rep.max_obs_err = max(rep.max_obs_err, float(oe.max()))
            for j, w in zip(*np.nonzero(oe > 1.0)):
                rep.value_violations.append((t, int(idx[j]), f"obs[{w}]", float(got_obs[idx[j], w]),
                                             float(ref_obs[idx[j], w])))
            rtol = np.maximum(FREE_REWARD_ATOL, FREE_REWARD_RTOL * np.abs(ref_rew))
            rerr = np.abs(got_rew.astype(np.float64) - ref_rew) / rtol
            rerr = np.where(class_bad | ~np.isfinite(ref_rew), 0.0, rerr)
            # reward discontinuities: the reference's decision variable inside a guard band
            on_step = near_ground.copy()
            ab, wb = np.abs(ref_before[:, 6]), np.abs(ref_before[:, 7])
            for thr in REWARD_STEPS["angle_before"]:
                on_step |= np.abs(ab - thr) < (GUARD["angle_sign"] if thr == 0.0 else GUARD["angle"])
            for thr in REWARD_STEPS["omega_before"]:
                on_step |= np.abs(wb - thr) < GUARD["angle"]
            for thr in REWARD_STEPS["angle_after"]:
                on_step |= np.abs(np.abs(ang) - thr) < GUARD["angle"]
            rep.reward_guard_events += int((on_step[idx] & ~(ref_term[idx])).sum())
            rerr = np.where(on_step & ~ref_term, 0.0, rerr)
            re_ = rerr[idx]
            rep.max_free_reward_err = max(rep.max_free_reward_err, float(re_.max()))
            for j in np.flatnonzero(re_ > 1.0):
                rep.value_violations.append((t, int(idx[j]), "reward", float(got_rew[idx[j]]),
                                             float(ref_rew[idx[j]])))
            rep.steps_compared += int(idx.size)

        self.synced &= ~(flag_bad | wrap_bad)
        rep.episodes += int((ref_term | ref_trunc).sum())
        rep.landing_counts += np.bincount(np.asarray(ref_landing)[ref_term].astype(np.int64),
                                          minlength=5)[:5]


# ----------------------------------------------------------------------------------------
# drivers
# ----------------------------------------------------------------------------------------
def replay_golden_rollout(stepper, g: dict) -> ParityReport:
    """Replay a committed reference rollout ([T, E, ...] arrays of make_golden.py): inject the
    reference's initial and reset states, feed its actions, compare every step."""
    actions = g["actions"]
    T, E = actions.shape[0], actions.shape[1]
    assert stepper.n == E
    stepper.inject_fresh(g["init_state"])
    chk = LockstepChecker(E)
    ref_before = np.asarray(g["init_state"], np.float64).copy()
    before = stepper.view()
    for t in range(T):
        obs, rew, term, trunc, land = stepper.step(actions[t])
        after = stepper.view()
        chk.teacher_forced(t, before, after, actions[t], rew, term, trunc, land)
        done = g["terminated"][t] | g["truncated"][t]
        ref_obs = np.where(done[:, None], g["terminal_obs"][t], g["obs"][t])
        chk.compare(t, obs, rew, term, trunc, land, after, g["state"][t], ref_obs, g["reward"][t],
                    g["terminated"][t], g["truncated"][t], g["landing"][t], ref_before)
        ref_before = np.asarray(g["state"][t], np.float64).copy()
        before = after
        if done.any():
            stepper.inject_fresh(g["reset_state"][t], mask=done)
            chk.resync(done)
            ref_before[done] = g["reset_state"][t][done]
            before = stepper.view()
    return chk.rep


def replay_edge_cases(stepper, g: dict) -> ParityReport:
    """Teacher-forced wild states of ``ref_edge_cases.npz`` (outputs of the reference itself):
    inject, step twice with the same action, compare both steps."""
    n = g["init_state"].shape[0]
    assert stepper.n == n
    stepper.inject_fresh(g["init_state"], steps=g["init_step"])
    chk = LockstepChecker(n)
    ref_before = np.asarray(g["init_state"], np.float64).copy()
    before = stepper.view()
    for k in range(2):
        obs, rew, term, trunc, land = stepper.step(g["actions"])
        after = stepper.view()
        chk.teacher_forced(k, before, after, g["actions"], rew, term, trunc, land)
        chk.compare(k, obs, rew, term, trunc, land, after, g["state"][k], g["obs"][k], g["reward"][k],
                    g["terminated"][k], g["truncated"][k], g["landing"][k], ref_before)
        ref_before = np.asarray(g["state"][k], np.float64).copy()
        before = after
    return chk.rep


def sample_reset_states(rng: np.random.Generator, params: RlParams, n: int) -> np.ndarray:
    lo = np.array(list(params.reset_low))
    hi = np.array(list(params.reset_high))
    return rng.uniform(lo, hi, size=(n, 10))


def random_actions(t, oracle, rng):
    return rng.uniform([0.0, -1.0], [1.0, 1.0], size=(oracle.n, 2)).astype(np.float32)


def braking_actions(t, oracle, rng):
    """A crude suicide-burn controller + attitude damper, so that episodes END in every
    landing class instead of the 'unsafe' that random actions always produce."""
    r = oracle.rockets
    y, vy, ang, om = r["y"], r["vy"], r["angle"], r["ang_vel"]
    m = r["dry_mass"] + r["fuel"]
    a_max = 1.0e7 / m - 9.81
    burn = (vy * vy) / (2.0 * np.maximum(y - 2.0, 0.5)) > a_max * rng.uniform(0.55, 1.05, size=oracle.n)
    hover = (9.81 * m / 1.0e7) * (1.0 + np.clip(-(vy + 8.0) * 0.15, -0.9, 3.0))
    th = np.where(burn, 1.0, np.where(y < 60.0, hover, 0.0))
    cg = np.clip(-(ang * 0.9 + om * 1.6) * rng.uniform(0.6, 1.4, size=oracle.n), -1.0, 1.0)
    a = np.stack([th, cg], axis=1) + rng.normal(0.0, 0.02, size=(oracle.n, 2))
    return a.astype(np.float32)


def constants_from_params(p: RlParams):
    """The oracle's EnvConstants for an arbitrary RlParams (non-default configurations)."""
    from oracle.rocket_oracle import EnvConstants
    return EnvConstants(
        dt=p.time_step, gravity=p.gravity, air_density=p.air_density, thrust_power=p.thrust_power,
        cold_gas_thrust_power=p.cold_gas_thrust_power, fuel_rate=p.fuel_consumption_rate, radius=p.radius,
        moment_arm=p.cold_gas_moment_arm, reference_area=p.reference_area, drag_coefficient=p.drag_coefficient,
        angular_damping=p.angular_damping,
        reset_ranges=tuple((float(lo), float(hi)) for lo, hi in zip(p.reset_low, p.reset_high)),
        land_perfect=tuple(p.land_perfect), land_good=tuple(p.land_good), land_ok=tuple(p.land_ok),
        max_episode_steps=int(p.max_episode_steps), max_horizontal_position=p.max_horizontal_position,
        max_altitude=p.max_altitude, tip_over_angle=p.tip_over_angle, r_landing_perfect=p.landing_perfect,
        r_landing_good=p.landing_good, r_landing_ok=p.landing_ok, r_crash_ground=p.crash_ground,
        r_out_of_bounds=p.out_of_bounds, r_tipped_over=p.tipped_over,
        throttle_descent_reward_scale=p.throttle_descent_reward_scale, cold_gas_reward_scale=p.cold_gas_reward_scale,
        free_fall_penalty_scale=p.free_fall_penalty_scale, angle_aware_throttle_scale=p.angle_aware_throttle_scale,
        correct_direction_bonus=p.correct_direction_bonus, gamma=p.gamma,
        obs_low=tuple(p.obs_low), obs_high=tuple(p.obs_high))


def other_config() -> RlParams:
    """A configuration that shares NO derived constant with the reference's config.yaml: half the
    time step, a heavier rocket with a bigger engine and tank (another fixed-point fuel unit),
    stronger cold gas and damping, a 1500-step time limit.  Exercises the run-time-constant kernels."""
    p = load_params()
    p.time_step = 0.05
    p.thrust_power = 2.4e7
    p.cold_gas_thrust_power = 1.1e5
    p.fuel_consumption_rate = 2900.0
    p.angular_damping = 0.08
    p.drag_coefficient = 0.7
    p.reset_low[8], p.reset_high[8] = 52000.0, 61000.0
    p.reset_low[9], p.reset_high[9] = 600000.0, 900000.0
    p.max_episode_steps = 1500
    return p


def hover_actions(t, oracle, rng):
    """Keeps every rocket aloft to the 1000-step limit (lander.py:237-241): descend to ~300 m, hold
    it with a throttle loop, PD attitude loop, 2 % action noise.  The worst case for ACCUMULATED
    error: the episode is ten times longer than a random-action one."""
    r = oracle.rockets
    y, vy, a, w = r["y"], r["vy"], r["angle"], r["ang_vel"]
    m = r["dry_mass"] + r["fuel"]
    tv = np.where(y > 400, -40.0, np.clip(-(y - 300) * 0.2, -20, 20))
    th = (9.81 * m / oracle.c.thrust_power) * (1 + np.clip((tv - vy) * 0.3, -1, 6))
    cg = np.clip(-(a * 0.9 + w * 1.6), -1, 1)
    return (np.stack([th, cg], 1) + rng.normal(0, 0.02, (oracle.n, 2))).astype(np.float32)


def tilted_cruise_actions(t, oracle, rng):
    """Hover while HOLDING a per-rocket tilt of up to +-12 deg (thrust has a large horizontal
    component for the whole episode, so vx and x integrate every error of the angle and of T/m),
    with a slow roll reversal; rockets drift sideways at tens of m/s."""
    r = oracle.rockets
    y, vy, a, w = r["y"], r["vy"], r["angle"], r["ang_vel"]
    m = r["dry_mass"] + r["fuel"]
    ids = np.arange(oracle.n)
    target = 12.0 * np.sin(0.37 * ids + 1.0) * np.cos(2.0 * np.pi * t / 400.0)
    tv = np.where(y > 600, -30.0, np.clip(-(y - 500) * 0.2, -20, 20))
    th = (9.81 * m / 1e7) * (1 + np.clip((tv - vy) * 0.3, -1, 6)) / np.maximum(np.cos(np.deg2rad(a)), 0.5)
    cg = np.clip(-((a - target) * 0.9 + w * 1.6), -1, 1)
    return (np.stack([th, cg], 1) + rng.normal(0, 0.02, (oracle.n, 2))).astype(np.float32)


def lockstep_vs_oracle(stepper, oracle, steps: int, seed: int = 0, action_fn=random_actions,
                       view_every: int = 1) -> ParityReport:
    """Free-running comparison against the C oracle (oracle.c_oracle.BatchOracle): both start
    from the same injected states; whenever the ORACLE finishes an episode a fresh state is
    drawn on the host and injected into both."""
    n = stepper.n
    rng = np.random.default_rng(seed)
    init = sample_reset_states(rng, stepper.params, n)
    stepper.inject_fresh(init)
    oracle.inject(init)
    chk = LockstepChecker(n, oracle.c)
    before = stepper.view()
    for t in range(steps):
        a = action_fn(t, oracle, rng)
        ref_before = oracle.state()
        obs, rew, term, trunc, land = stepper.step(a)
        oobs, orew, oterm, otrunc, oland = oracle.step(a)
        after = stepper.view()
        chk.teacher_forced(t, before, after, a, rew, term, trunc, land)
        chk.compare(t, obs, rew, term, trunc, land, after, oracle.state(), oobs, orew, oterm,
                    otrunc, oland, ref_before)
        before = after
        done = oterm | otrunc
        if done.any():
            fresh = sample_reset_states(rng, stepper.params, n)
            stepper.inject_fresh(fresh, mask=done)
            oracle.inject(fresh, mask=done)
            chk.resync(done)
            before = stepper.view()
    return chk.rep
