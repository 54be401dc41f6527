"""Constants of the environment: YAML -> ``RlParams`` (the C ABI's POD struct).

Mirrors the reference's strict loader (``backend/config.py:21-37``): dotted-key ``get``
that raises ``KeyError`` on a missing key and ``ValueError`` on a ``None`` value.  The
YAML is read ONCE on the host; the library bakes it into a kernel-parameter struct.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))


def default_config_path() -> str:
    """``$ROCKET_CONFIG`` if set, else the ``config.yaml`` at the repository root (same
    key paths as the reference's file)."""
    env = os.environ.get("ROCKET_CONFIG")
    return env if env else os.path.join(os.path.dirname(_PKG_DIR), "config.yaml")


class Config:
    """Dotted-key view of a ``config.yaml`` with the reference's key paths.  Lookups are strict in
    the same way as the reference's loader: an unknown key or a path through a non-mapping is a
    ``KeyError``, a key that is present but empty is a ``ValueError``."""

    def __init__(self, path: str | None = None):
        import yaml
        self.config_path = path or default_config_path()
        try:
            with open(self.config_path, "r") as fh:
                tree = yaml.safe_load(fh)
        except FileNotFoundError:
            raise FileNotFoundError(f"no configuration file at {self.config_path}") from None
        except yaml.YAMLError as exc:
            raise ValueError(f"{self.config_path} is not valid YAML: {exc}") from exc
        self._config = tree if isinstance(tree, dict) else {}

    def get(self, key_path: str) -> Any:
        node: Any = self._config
        walked = []
        for part in key_path.split("."):
            if not isinstance(node, dict):
                raise KeyError(f"'{'.'.join(walked)}' is a leaf of {self.config_path}; it has no entry '{part}'")
            walked.append(part)
            try:
                node = node[part]
            except KeyError:
                raise KeyError(f"'{'.'.join(walked)}' is not defined in {self.config_path}") from None
        if node is None:
            raise ValueError(f"'{key_path}' is present in {self.config_path} but empty")
        return node


class RlParams(C.Structure):
    """ctypes image of ``struct RlParams`` (include/rocket_landing_b200.h)."""
    _fields_ = ([(n, C.c_double) for n in (
        "time_step", "gravity", "air_density", "thrust_power", "cold_gas_thrust_power",
        "fuel_consumption_rate", "radius", "cold_gas_moment_arm", "reference_area",
        "drag_coefficient", "angular_damping")]
        + [("reset_low", C.c_double * 10), ("reset_high", C.c_double * 10),
           ("land_perfect", C.c_double * 3), ("land_good", C.c_double * 3),
           ("land_ok", C.c_double * 3)]
        + [(n, C.c_double) for n in (
            "max_horizontal_position", "max_altitude", "tip_over_angle", "landing_perfect",
            "landing_good", "landing_ok", "crash_ground", "out_of_bounds", "tipped_over",
            "throttle_descent_reward_scale", "cold_gas_reward_scale", "free_fall_penalty_scale",
            "angle_aware_throttle_scale", "correct_direction_bonus", "gamma")]
        + [("obs_low", C.c_double * 8), ("obs_high", C.c_double * 8),
           ("max_episode_steps", C.c_int32), ("reserved", C.c_int32)])


# reset draw order of the reference (backend/simulation/config.py:26-40)
_RESET_KEYS = ("rocket.position_limits.x", "rocket.position_limits.y",
               "rocket.velocity_limits.vx", "rocket.velocity_limits.vy",
               "rocket.acceleration_limits.ax", "rocket.acceleration_limits.ay",
               "rocket.attitude_limits.angle", "rocket.attitude_limits.angular_velocity",
               "rocket.mass_limits.dry_mass", "rocket.mass_limits.fuel_mass")


def _pair(cfg: Config, key: str):
    """backend/simulation/config.py:8-17: a list of exactly two numbers."""
    val = cfg.get(key)
    if (isinstance(val, list) and len(val) == 2
            and all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in val)):
        return float(val[0]), float(val[1])
    raise ValueError(f"Config key '{key}' must be a list of 2 numbers.")


def load_params(config: Config | str | None = None) -> RlParams:
    cfg = config if isinstance(config, Config) else Config(config)
    p = RlParams()
    p.time_step = float(cfg.get("simulation.time_step"))
    p.gravity = float(cfg.get("environment.gravity"))
    p.air_density = float(cfg.get("environment.air_density"))
    p.thrust_power = float(cfg.get("rocket.thrust_power"))
    p.cold_gas_thrust_power = float(cfg.get("rocket.cold_gas_thrust_power"))
    p.fuel_consumption_rate = float(cfg.get("rocket.fuel_consumption_rate"))
    p.radius = float(cfg.get("rocket.radius"))
    p.cold_gas_moment_arm = float(cfg.get("rocket.cold_gas_moment_arm"))
    p.reference_area = float(cfg.get("rocket.reference_area"))
    p.drag_coefficient = float(cfg.get("rocket.drag_coefficient"))
    p.angular_damping = float(cfg.get("rocket.angular_damping"))
    for i, key in enumerate(_RESET_KEYS):
        p.reset_low[i], p.reset_high[i] = _pair(cfg, key)
    for field, level in (("land_perfect", "perfect"), ("land_good", "good"), ("land_ok", "ok")):
        arr = getattr(p, field)
        for i, name in enumerate(("speed_vx", "speed_vy", "angle")):
            arr[i] = float(cfg.get(f"landing.thresholds.{level}.{name}"))
    p.max_horizontal_position = float(cfg.get("rl.max_horizontal_position"))
    p.max_altitude = float(cfg.get("rl.max_altitude"))
    p.tip_over_angle = float(cfg.get("rl.tip_over_angle"))
    rewards = cfg.get("rl.rewards")
    for field in ("landing_perfect", "landing_good", "landing_ok", "crash_ground",
                  "out_of_bounds", "tipped_over", "throttle_descent_reward_scale",
                  "cold_gas_reward_scale", "free_fall_penalty_scale",
                  "angle_aware_throttle_scale", "correct_direction_bonus", "gamma"):
        setattr(p, field, float(rewards[field]))          # strict, reward.py:13-24
    # observation Box (lander.py:73-101): reset limits of the first eight words, y low = 0.0
    for i in range(8):
        p.obs_low[i], p.obs_high[i] = p.reset_low[i], p.reset_high[i]
    p.obs_low[1] = 0.0
    p.max_episode_steps = int(cfg.get("rl.max_episode_steps"))
    return p
