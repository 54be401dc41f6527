"""CPU oracle: a scalar fp64 restatement of the reference's environment step.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Allowed users:
``tests/``, ``__graft_entry__.smoke()``, ``bench.py``'s CPU-baseline legs and the
golden-vector generator.

Parity status: **pinned against the live reference.**  The reference's own tests hold
no golden numbers for this path (``tests/test_physics.py`` recomputes
every expectation from the same formula), so the pin is
``oracle/pin_against_reference.py`` / ``tests/test_oracle_vs_reference.py``: this
restatement is stepped beside the unmodified reference (imported through
``oracle/ref_loader.py``) on identical injected states and actions and must agree
bit-for-bit on every state word, reward and flag; and the committed fixtures under
``tests/golden/`` are outputs of the reference itself (``tests/golden/make_golden.py``).

Every function cites the reference ``file:line`` it follows (paths are relative to
the reference root).  Operation order is kept exactly as in the reference because the
pin is bit-exact in fp64; all arithmetic is Python ``float`` (IEEE double) except
where the reference itself calls numpy (``np.sin``, ``np.cos``, ``np.dot``,
``np.sqrt``, ``np.deg2rad``, ``np.rad2deg``), which is kept so the two agree to the
last bit on the same numpy build.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

# landing classes, numbered as the reference's wire format numbers them
# (backend/protocol.py:31-41): 0 none, 1 safe, 2 good, 3 ok, 4 unsafe
LANDING_NONE, LANDING_SAFE, LANDING_GOOD, LANDING_OK, LANDING_UNSAFE = 0, 1, 2, 3, 4
LANDING_NAMES = ("none", "safe", "good", "ok", "unsafe")


# ----------------------------------------------------------------------------------
# constants (config.yaml; backend/config.py:21-37 strict dotted lookup)
# ----------------------------------------------------------------------------------
def _lookup(tree, dotted):
    node = tree
    for part in dotted.split("."):
        if not isinstance(node, dict) or part not in node:
            raise KeyError(f"Missing config key: '{dotted}'")
        node = node[part]
    if node is None:
        raise ValueError(f"Config key '{dotted}' exists but has no value (None).")
    return node


@dataclass
class EnvConstants:
    """All numbers the step reads.  Field comments give the reference call site."""
    dt: float = 0.1                       # simulation.time_step (engine.py:11)
    gravity: float = -9.81                # environment.gravity (engine.py:10)
    air_density: float = 1.225            # engine.py:23
    thrust_power: float = 1.0e7           # engine.py:14
    cold_gas_thrust_power: float = 7.0e4  # engine.py:15-17
    fuel_rate: float = 1700.0             # engine.py:18-20
    radius: float = 1.85                  # engine.py:26
    moment_arm: float = 1.85              # engine.py:27-29
    reference_area: float = 10.8          # engine.py:25
    drag_coefficient: float = 0.8         # engine.py:24
    angular_damping: float = 0.05         # engine.py:32
    # reset ranges in draw order (simulation/config.py:26-40)
    reset_ranges: tuple = ((-1500.0, 1500.0), (1800.0, 2200.0), (-25.0, 25.0),
                           (-200.0, -150.0), (-5.0, 5.0), (-5.0, 5.0), (-15.0, 15.0),
                           (-10.0, 10.0), (34000.0, 38000.0), (370000.0, 410000.0))
    # landing thresholds (utils.py:7-17): (vx, vy, angle) for perfect / good / ok
    land_perfect: tuple = (30.0, 30.0, 5.0)
    land_good: tuple = (40.0, 40.0, 8.0)
    land_ok: tuple = (100.0, 100.0, 10.0)
    # rl.* (reward.py:13-24, lander.py:57)
    max_episode_steps: int = 1000
    max_horizontal_position: float = 50000.0
    max_altitude: float = 50000.0
    tip_over_angle: float = 90.0
    r_landing_perfect: float = 4000.0
    r_landing_good: float = 3000.0
    r_landing_ok: float = 2000.0
    r_crash_ground: float = -800.0
    r_out_of_bounds: float = -300.0
    r_tipped_over: float = -300.0
    throttle_descent_reward_scale: float = 0.3
    cold_gas_reward_scale: float = 0.7
    free_fall_penalty_scale: float = 0.2
    angle_aware_throttle_scale: float = 0.3
    correct_direction_bonus: float = 1.5
    gamma: float = 0.99
    # observation Box bounds (lander.py:73-101): low / high for
    # [x, y, vx, vy, ax, ay, angle, angular velocity]; y low is the literal 0.0
    obs_low: tuple = field(default=None)
    obs_high: tuple = field(default=None)

    def __post_init__(self):
        rr = self.reset_ranges
        if self.obs_low is None:
            self.obs_low = (rr[0][0], 0.0, rr[2][0], rr[3][0], rr[4][0], rr[5][0],
                            rr[6][0], rr[7][0])
        if self.obs_high is None:
            self.obs_high = (rr[0][1], rr[1][1], rr[2][1], rr[3][1], rr[4][1], rr[5][1],
                             rr[6][1], rr[7][1])

    @classmethod
    def from_yaml(cls, path: str | None = None) -> "EnvConstants":
        import yaml
        if path is None:
            path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "config.yaml")
        with open(path, "r") as fh:
            tree = yaml.safe_load(fh) or {}
        g = lambda k: _lookup(tree, k)            # noqa: E731
        pair = lambda k: (float(g(k)[0]), float(g(k)[1]))   # noqa: E731
        thr = lambda lvl: tuple(float(g(f"landing.thresholds.{lvl}.{n}"))   # noqa: E731
                                for n in ("speed_vx", "speed_vy", "angle"))
        rw = lambda k: float(g(f"rl.rewards.{k}"))  # noqa: E731
        return cls(
            dt=float(g("simulation.time_step")), gravity=float(g("environment.gravity")),
            air_density=float(g("environment.air_density")),
            thrust_power=float(g("rocket.thrust_power")),
            cold_gas_thrust_power=float(g("rocket.cold_gas_thrust_power")),
            fuel_rate=float(g("rocket.fuel_consumption_rate")),
            radius=float(g("rocket.radius")), moment_arm=float(g("rocket.cold_gas_moment_arm")),
            reference_area=float(g("rocket.reference_area")),
            drag_coefficient=float(g("rocket.drag_coefficient")),
            angular_damping=float(g("rocket.angular_damping")),
            reset_ranges=(pair("rocket.position_limits.x"), pair("rocket.position_limits.y"),
                          pair("rocket.velocity_limits.vx"), pair("rocket.velocity_limits.vy"),
                          pair("rocket.acceleration_limits.ax"),
                          pair("rocket.acceleration_limits.ay"),
                          pair("rocket.attitude_limits.angle"),
                          pair("rocket.attitude_limits.angular_velocity"),
                          pair("rocket.mass_limits.dry_mass"),
                          pair("rocket.mass_limits.fuel_mass")),
            land_perfect=thr("perfect"), land_good=thr("good"), land_ok=thr("ok"),
            max_episode_steps=int(g("rl.max_episode_steps")),
            max_horizontal_position=float(g("rl.max_horizontal_position")),
            max_altitude=float(g("rl.max_altitude")),
            tip_over_angle=float(g("rl.tip_over_angle")),
            r_landing_perfect=rw("landing_perfect"), r_landing_good=rw("landing_good"),
            r_landing_ok=rw("landing_ok"), r_crash_ground=rw("crash_ground"),
            r_out_of_bounds=rw("out_of_bounds"), r_tipped_over=rw("tipped_over"),
            throttle_descent_reward_scale=rw("throttle_descent_reward_scale"),
            cold_gas_reward_scale=rw("cold_gas_reward_scale"),
            free_fall_penalty_scale=rw("free_fall_penalty_scale"),
            angle_aware_throttle_scale=rw("angle_aware_throttle_scale"),
            correct_direction_bonus=rw("correct_direction_bonus"), gamma=rw("gamma"))


# ----------------------------------------------------------------------------------
# state
# ----------------------------------------------------------------------------------
@dataclass
class RocketState:
    """The ten words of ``Rocket.state`` that matter plus the three words of
    ``Rocket.previous_state`` the integrator reads (engine.py:181-216 reads only
    previous x, y, angle)."""
    x: float = 0.0
    y: float = 0.0
    vx: float = 0.0
    vy: float = 0.0
    ax: float = 0.0
    ay: float = 0.0
    angle: float = 0.0
    ang_vel: float = 0.0
    ang_acc: float = 0.0
    dry_mass: float = 0.0
    fuel: float = 0.0
    prev_x: float = 0.0
    prev_y: float = 0.0
    prev_angle: float = 0.0

    def copy(self) -> "RocketState":
        return RocketState(**{k: getattr(self, k) for k in self.__dataclass_fields__})


STATE_WORDS = ("x", "y", "vx", "vy", "ax", "ay", "angle", "ang_vel", "dry_mass", "fuel")


# ----------------------------------------------------------------------------------
# physics  (backend/physics/engine.py)
# ----------------------------------------------------------------------------------
def normalize_angle_180(angle_deg: float) -> float:
    """engine.py:230-237 -- Python ``%`` 360 (result in [0, 360)), minus 360 when
    >= 180.  Output in [-180, 180).  The ``elif < -180`` arm there is unreachable."""
    a = angle_deg % 360.0
    if a >= 180.0:
        a -= 360.0
    return a


def net_force(c: EnvConstants, total_mass: float, throttle: float, angle_deg: float,
              vx: float, vy: float):
    """engine.py:100-115 = gravity (:48-52) + thrust (:54-76) + drag (:78-98), summed
    in that order.  ``throttle`` arrives already clipped by the caller but is clipped
    again at :61.  Returns (Fx, Fy)."""
    if total_mass <= 0:
        return 0.0, 0.0
    gx, gy = 0.0, (0.0 if total_mass < 0 else total_mass) * c.gravity
    th = min(max(throttle, 0.0), 1.0)
    if th > 1e-6:
        rad = float(np.deg2rad(angle_deg))
        mag = th * c.thrust_power
        tx, ty = mag * float(np.sin(rad)), mag * float(np.cos(rad))
    else:
        tx, ty = 0.0, 0.0
    v = np.array([vx, vy])
    speed_sq = float(np.dot(v, v))           # engine.py:83 (BLAS dot, keeps the pin bit-exact)
    if speed_sq > 1e-9:
        speed = float(np.sqrt(speed_sq))
        drag_mag = 0.5 * c.air_density * c.drag_coefficient * c.reference_area * speed_sq
        dx, dy = drag_mag * (-vx / speed), drag_mag * (-vy / speed)
    else:
        dx, dy = 0.0, 0.0
    return (gx + tx) + dx, (gy + ty) + dy


def linear_acceleration(fx: float, fy: float, total_mass: float):
    """engine.py:117-123."""
    if total_mass <= 1e-6:
        return 0.0, 0.0
    return fx / total_mass, fy / total_mass


def angular_acceleration(c: EnvConstants, cold_gas: float, total_mass: float) -> float:
    """engine.py:125-155 -- deg/s^2 from cold-gas torque over a solid-cylinder inertia."""
    cg = min(max(cold_gas, -1.0), 1.0)
    if total_mass <= 1e-6 or c.radius <= 1e-6:
        return 0.0
    inertia = 0.5 * total_mass * (c.radius ** 2)
    if inertia < 1e-6:
        return 0.0
    torque = (c.cold_gas_thrust_power * cg) * c.moment_arm
    return float(np.rad2deg(torque / inertia))


def bootstrap_previous(c: EnvConstants, s: RocketState) -> None:
    """base.py:65-130 -- the Verlet bootstrap: zero-control acceleration at the
    current state, then x_prev = x - v dt + a0 dt^2 / 2 (same for y), and
    angle_prev = normalise(angle - omega dt + 0) since alpha0 = 0.  The sampled
    ax/ay are not used."""
    dt = c.dt
    m0 = s.dry_mass + s.fuel
    if m0 <= 1e-6:
        ax0 = ay0 = alpha0 = 0.0
    else:
        fx, fy = net_force(c, m0, 0.0, s.angle, s.vx, s.vy)
        ax0, ay0 = linear_acceleration(fx, fy, m0)
        alpha0 = angular_acceleration(c, 0.0, m0)
    s.prev_x = s.x - (s.vx * dt) + (0.5 * ax0 * dt ** 2)
    s.prev_y = s.y - (s.vy * dt) + (0.5 * ay0 * dt ** 2)
    s.prev_angle = normalize_angle_180(s.angle - (s.ang_vel * dt) + (0.5 * alpha0 * dt ** 2))


def apply_action(c: EnvConstants, s: RocketState, throttle: float, cold_gas: float) -> None:
    """base.py:132-184 (one physics step, in place) with engine.py:157-228 inlined.

    Clip the action (:135-136); fuel gate (:138-141); pre-burn total mass (:143-144);
    forces at the current state; position-Verlet for x, y; damped Verlet for the
    angle with the angular velocity taken BEFORE normalisation; previous <- current
    (whose angle is the normalised one, :171); burn fuel (:174-175)."""
    dt = c.dt
    th = min(max(float(throttle), 0.0), 1.0)
    cg = min(max(float(cold_gas), -1.0), 1.0)
    fuel_now = s.fuel
    if fuel_now <= 0:
        th = 0.0
        s.fuel = 0.0
    m = s.dry_mass + fuel_now
    if m <= 1e-6:
        return                                              # base.py:145-147
    fx, fy = net_force(c, m, th, s.angle, s.vx, s.vy)
    s.ax, s.ay = linear_acceleration(fx, fy, m)
    s.ang_acc = angular_acceleration(c, cg, m)
    # engine.py:181-186
    nx = 2.0 * s.x - s.prev_x + s.ax * dt ** 2
    ny = 2.0 * s.y - s.prev_y + s.ay * dt ** 2
    # engine.py:193-194
    nvx = (nx - s.x) / dt
    nvy = (ny - s.y) / dt
    # engine.py:206-216
    damp = max(0.0, 1.0 - (c.angular_damping * dt))
    change = (s.angle - s.prev_angle) * damp
    nangle = s.angle + change + s.ang_acc * dt ** 2
    # engine.py:220-226
    nomega = (nangle - s.angle) / dt
    nangle = normalize_angle_180(nangle)
    # base.py:171-172
    s.prev_x, s.prev_y, s.prev_angle = s.x, s.y, s.angle
    s.x, s.y, s.vx, s.vy, s.angle, s.ang_vel = nx, ny, nvx, nvy, nangle, nomega
    # engine.py:239-243, base.py:174-175
    burned = max(0.0, min(max(th, 0.0), 1.0) * c.fuel_rate * dt)
    s.fuel = max(0.0, fuel_now - burned)


# ----------------------------------------------------------------------------------
# reward / termination  (backend/rl/reward.py, backend/utils.py)
# ----------------------------------------------------------------------------------
def classify_landing(c: EnvConstants, vx: float, vy: float, angle: float) -> int:
    """utils.py:1-37 -- strict '<' on the three magnitudes."""
    avx, avy, aa = abs(vx), abs(vy), abs(angle)
    if avx < c.land_perfect[0] and avy < c.land_perfect[1] and aa < c.land_perfect[2]:
        return LANDING_SAFE
    if avx < c.land_good[0] and avy < c.land_good[1] and aa < c.land_good[2]:
        return LANDING_GOOD
    if avx < c.land_ok[0] and avy < c.land_ok[1] and aa < c.land_ok[2]:
        return LANDING_OK
    return LANDING_UNSAFE


def _potential(y: float, vx: float, vy: float, angle: float, omega: float) -> float:
    """reward.py:231-264."""
    yp = max(0.0, y)
    yf = min(1.0, yp / 2000.0)
    w_y = 0.005
    w_vy = 0.015 * (2.0 - yf)
    w_vx = 0.005
    w_ang = 0.01 * (2.0 - yf)
    w_stab = 0.05 * (2.0 - yf)
    return (-w_y * yp - w_vy * abs(vy) - w_vx * abs(vx) - w_ang * abs(angle)
            - w_stab * abs(omega))


def reward_and_flags(c: EnvConstants, before: RocketState, throttle: float, cold_gas: float,
                     after: RocketState):
    """reward.py:27-281.  ``throttle`` / ``cold_gas`` are the RAW action widened from
    float32 (:69-70), not the clipped values physics used.  Returns
    (reward, terminated_on_ground, truncated, landing_class)."""
    total = 0.0
    y_b, ang_b, om_b = before.y, before.angle, before.ang_vel
    y_a, vy_a, ang_a, om_a = after.y, after.vy, after.angle, after.ang_vel

    # ground contact (:73-102): returns before any bounds / tip-over check
    if y_a <= 0.1 and y_b > 0.1:
        cls = classify_landing(c, after.vx, after.vy, after.angle)
        angle_bonus = max(0, 1.0 - (abs(ang_a) / 10.0))
        velocity_bonus = max(0, 1.0 - (abs(vy_a) / 5.0))
        quality = 0.6 * angle_bonus + 0.4 * velocity_bonus
        if cls == LANDING_SAFE:
            total += c.r_landing_perfect * (1.0 + 0.5 * quality)
        elif cls == LANDING_GOOD:
            total += c.r_landing_good * (0.8 + 0.2 * quality)
        elif cls == LANDING_OK:
            total += c.r_landing_ok
        else:
            severity = min(1.0, (abs(vy_a) / 20.0 + abs(ang_a) / 45.0) / 2.0)
            total += c.r_crash_ground * (0.7 + 0.3 * severity)
        return float(total), True, False, cls

    # 1. angular control (:108-124)
    ae_b, ae_a = abs(ang_b), abs(ang_a)
    we_b, we_a = abs(om_b), abs(om_a)
    corr = -(ae_a - ae_b) * 0.5
    corr -= (we_a - we_b) * 0.1
    total += corr

    # cold-gas direction term (:130-167)
    if ae_b > 0.1 or we_b > 0.1:
        correct = (ang_b > 0 and cold_gas < 0) or (ang_b < 0 and cold_gas > 0)
        needed = ae_b + we_b
        effect = (ae_b - ae_a) + (we_b - we_a)
        mult = c.correct_direction_bonus if correct else -c.correct_direction_bonus
        cg_reward = abs(cold_gas) * needed * mult * c.cold_gas_reward_scale
        if effect > 0:
            cg_reward *= 1.0 + effect ** 2
        total += cg_reward
    else:
        total -= abs(cold_gas) * 0.3 * c.cold_gas_reward_scale

    # 2. throttle while descending (:171-186)
    if y_a > 0.1 and vy_a < 0:
        angle_factor = max(0, 1.0 - (abs(ang_a) / 45.0))
        descent = min(1.0, abs(vy_a) / 10.0)
        total += (throttle * (-vy_a) * angle_factor * (1.0 + descent)
                  * c.throttle_descent_reward_scale)

    # 3. free fall (:190-201)
    if y_a > 0.1 and vy_a < 0 and throttle < 0.1:
        proximity = 1.0 + (8.0 / max(y_a, 1.0))
        speed_f = min(1.0, abs(vy_a) / 15.0)
        total -= c.free_fall_penalty_scale * (-vy_a) * proximity * (1.0 + speed_f)

    # 4. throttle while tilted (:205-219)
    if throttle > 0.1 and y_a > 0.1:
        if abs(ang_a) > 10.0:
            total -= throttle * (abs(ang_a) / 90.0) * c.angle_aware_throttle_scale
        elif abs(ang_a) > 5.0 and abs(ang_a) <= 10.0:
            total -= (throttle * ((abs(ang_a) - 5.0) / 5.0)
                      * c.angle_aware_throttle_scale * 0.5)

    # 5. climbing (:222-227)
    if y_a > 0.1 and vy_a > 0:
        alt_f = min(1.0, y_a / 1000.0)
        asc_f = min(1.0, vy_a / 5.0)
        total -= vy_a * 0.5 * alt_f * (1.0 + asc_f)

    # 6. potential shaping (:267-270)
    phi_b = _potential(before.y, before.vx, before.vy, before.angle, before.ang_vel)
    phi_a = _potential(after.y, after.vx, after.vy, after.angle, after.ang_vel)
    total += c.gamma * phi_a - phi_b

    truncated = False
    if abs(after.x) > c.max_horizontal_position or y_a > c.max_altitude:   # :274-276
        total += c.r_out_of_bounds
        truncated = True
    if abs(ang_a) > c.tip_over_angle:                                      # :278-279
        total += c.r_tipped_over
    return float(total), False, truncated, LANDING_NONE


# ----------------------------------------------------------------------------------
# environment  (backend/envs/lander.py) + SB3-style auto-reset
# ----------------------------------------------------------------------------------
def clip_observation(c: EnvConstants, s: RocketState) -> np.ndarray:
    """lander.py:124-150 -- float32 cast first, then clip to the Box bounds."""
    raw = np.array([s.x, s.y, s.vx, s.vy, s.ax, s.ay, s.angle, s.ang_vel], dtype=np.float32)
    return np.clip(raw, np.asarray(c.obs_low, np.float32), np.asarray(c.obs_high, np.float32))


def sample_initial_state(c: EnvConstants, uniform=np.random.uniform) -> RocketState:
    """simulation/config.py:26-53 -- ten draws from the GLOBAL numpy generator, in
    this order: x, y, vx, vy, ax, ay, angle, angular velocity, dry mass, fuel."""
    d = [float(uniform(lo, hi)) for (lo, hi) in c.reset_ranges]
    return RocketState(x=d[0], y=d[1], vx=d[2], vy=d[3], ax=d[4], ay=d[5], angle=d[6],
                       ang_vel=d[7], dry_mass=d[8], fuel=d[9])


class OracleEnv:
    """Single-rocket environment with the reference's ``reset`` / ``step`` surface
    (lander.py:168-255) over the restated physics and reward."""

    def __init__(self, constants: EnvConstants | None = None):
        self.c = constants or EnvConstants.from_yaml()
        self.state = RocketState()
        self.current_step = 0
        self.last_landing = LANDING_NONE
        self.reset()

    # lander.py:168-186; ``seed`` has no effect on the state (it only seeds an unused
    # gymnasium generator) -- the draws come from the global numpy generator.
    def reset(self, *, seed=None):
        self.state = sample_initial_state(self.c)
        bootstrap_previous(self.c, self.state)
        self.current_step = 0
        return clip_observation(self.c, self.state), {"steps": 0}

    def inject(self, x, y, vx, vy, angle, ang_vel, dry_mass, fuel, ax=0.0, ay=0.0,
               current_step=0) -> np.ndarray:
        """Deterministic initial state: set ``Rocket.state`` then recompute the
        consistent previous state (base.py:65), as SURVEY.md section 8c prescribes."""
        self.state = RocketState(x=float(x), y=float(y), vx=float(vx), vy=float(vy),
                                 ax=float(ax), ay=float(ay), angle=float(angle),
                                 ang_vel=float(ang_vel), dry_mass=float(dry_mass),
                                 fuel=float(fuel))
        bootstrap_previous(self.c, self.state)
        self.current_step = int(current_step)
        return clip_observation(self.c, self.state)

    def step(self, action):
        """lander.py:188-255."""
        throttle = float(action[0])
        cold_gas = float(action[1])
        before = self.state.copy()
        apply_action(self.c, self.state, throttle, cold_gas)
        self.current_step += 1
        reward, terminated, truncated, cls = reward_and_flags(
            self.c, before, throttle, cold_gas, self.state)
        if self.current_step >= self.c.max_episode_steps:                 # :237-241
            truncated = True
        self.last_landing = cls
        obs = clip_observation(self.c, self.state)
        return obs, float(reward), terminated, truncated, {"steps": self.current_step,
                                                            "landing": cls}


def rollout_with_auto_reset(env: OracleEnv, actions: np.ndarray):
    """The SB3 VecEnv contract the reference trains under (a16 in SURVEY.md section 8a; SB3
    source is not in the reference tree): on ``done`` keep the terminal observation
    aside and return the observation of a fresh ``reset()`` instead.

    Returns a dict of per-step arrays; ``state`` rows are the RAW post-step state
    (before any reset), ``reset_state`` rows the freshly sampled state on done steps."""
    T = len(actions)
    out = {
        "state": np.zeros((T, 10)), "obs": np.zeros((T, 8), np.float32),
        "terminal_obs": np.zeros((T, 8), np.float32), "reward": np.zeros(T),
        "terminated": np.zeros(T, bool), "truncated": np.zeros(T, bool),
        "landing": np.zeros(T, np.int8), "reset_state": np.zeros((T, 10)),
        "steps": np.zeros(T, np.int32),
    }
    for t in range(T):
        obs, r, term, trunc, info = env.step(actions[t])
        s = env.state
        out["state"][t] = [getattr(s, k) for k in STATE_WORDS]
        out["reward"][t], out["terminated"][t], out["truncated"][t] = r, term, trunc
        out["landing"][t], out["steps"][t] = info["landing"], info["steps"]
        if term or trunc:
            out["terminal_obs"][t] = obs
            obs, _ = env.reset()
            s = env.state
            out["reset_state"][t] = [getattr(s, k) for k in STATE_WORDS]
        out["obs"][t] = obs
    return out


def inverse_of_dt_is_exact(c: EnvConstants) -> bool:
    """True when 1/dt is a small integer (dt = 0.1 -> 10), in which case dividing by dt
    and multiplying by 1/dt agree to ~1 ulp -- used by the tolerance notes."""
    return math.isclose(round(1.0 / c.dt), 1.0 / c.dt, rel_tol=1e-15)
