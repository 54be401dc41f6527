"""CPU oracle for the reference's SECOND caller of the step: the UI simulation loop and its
binary telemetry.  TEST INFRASTRUCTURE ONLY.

Restates, on top of ``oracle/rocket_oracle.py`` (pinned bit-for-bit against the live reference):

* ``RocketControls.step`` (``backend/rocket/controls.py:27-113``): frozen after touchdown; the
  action is clipped BEFORE it reaches both physics and reward (``:72-75`` -- unlike the gym env,
  where the reward sees the raw action); no max-step truncation, no observation; out-of-bounds
  only costs reward;
* ``SimulationController.step`` (``backend/simulation/controller.py:223-269``): rockets that have
  touched down are skipped and report ``(None, None, True)``;
* ``RocketWebSocketHandler.send_state_update`` (``backend/handler/websocket_handler.py:146-171``):
  the landing message is evaluated on the step a rocket touches down and remembered; the action
  echoed to the client is the one just applied, or zeros for a done / inactive rocket;
* ``BinaryProtocol`` (``backend/protocol.py:5-96``): 1 header byte (1 = telemetry) + 16 float32 per
  rocket: x y vx vy ax ay angle angularVelocity angularAcceleration mass fuelMass reward(NaN if
  inactive) throttle coldGas landing_code(0 none 1 safe 2 good 3 ok 4 unsafe) is_active.

Parity status: pinned -- ``tests/test_simulation.py`` steps this beside the reference's own
``RocketControls`` and ``BinaryProtocol`` where the checkout exists, and
``tests/golden/ref_telemetry.npz`` holds a telemetry stream produced by the reference itself.
"""
from __future__ import annotations

import struct

import numpy as np

from . import rocket_oracle as ro

MSG_TELEMETRY = 1
RECORD_FLOATS = 16


class ControlsOracle:
    """One ``RocketControls``."""

    def __init__(self, constants: ro.EnvConstants | None = None):
        self.c = constants or ro.EnvConstants.from_yaml()
        self.state = ro.RocketState()
        self.touchdown = False
        self.steps = 0

    def inject(self, x, y, vx, vy, angle, ang_vel, dry_mass, fuel) -> None:
        self.state = ro.RocketState(x=float(x), y=float(y), vx=float(vx), vy=float(vy), angle=float(angle),
                                    ang_vel=float(ang_vel), dry_mass=float(dry_mass), fuel=float(fuel))
        ro.bootstrap_previous(self.c, self.state)
        self.touchdown = False
        self.steps = 0

    def step(self, action):
        """controls.py:27-113 -> (state, reward, done, landing_class)."""
        if self.touchdown:
            return self.state, 0.0, True, ro.LANDING_NONE
        self.steps += 1
        throttle = min(max(float(action[0]), 0.0), 1.0)
        cold_gas = min(max(float(action[1]), -1.0), 1.0)
        clipped = np.array([throttle, cold_gas], dtype=np.float32)       # controls.py:75
        before = self.state.copy()
        ro.apply_action(self.c, self.state, throttle, cold_gas)
        reward, self.touchdown, _trunc, cls = ro.reward_and_flags(
            self.c, before, float(clipped[0]), float(clipped[1]), self.state)
        return self.state, float(reward), self.touchdown, cls


def encode_record(state, reward, action, landing_code) -> bytes:
    """protocol.py:52-96."""
    if state is None:
        return struct.pack("16f", *([0.0] * 11), float("nan"), 0.0, 0.0, float(landing_code), 0.0)
    return struct.pack("16f", state.x, state.y, state.vx, state.vy, state.ax, state.ay, state.angle,
                       state.ang_vel, state.ang_acc, state.dry_mass, state.fuel,
                       reward if reward is not None else 0.0, float(action[0]), float(action[1]),
                       float(landing_code), 1.0)


class SimulationOracle:
    """``SimulationController`` + the handler's telemetry for N rockets."""

    def __init__(self, n: int, constants: ro.EnvConstants | None = None):
        self.rockets = [ControlsOracle(constants) for _ in range(n)]
        self.touched = [False] * n
        self.final_code = [0] * n                     # websocket_handler.py final_outcomes

    def inject(self, states: np.ndarray) -> None:
        for r, s in zip(self.rockets, states):
            r.inject(s[0], s[1], s[2], s[3], s[6], s[7], s[8], s[9])
        self.touched = [False] * len(self.rockets)
        self.final_code = [0] * len(self.rockets)

    def step(self, actions: np.ndarray) -> bytes:
        """One tick: controller.py:223-269 + websocket_handler.py:120-171 -> the binary message."""
        msg = bytearray(struct.pack("B", MSG_TELEMETRY))
        for i, r in enumerate(self.rockets):
            if self.touched[i]:
                msg += encode_record(None, None, (0.0, 0.0), self.final_code[i])
                continue
            state, reward, done, cls = r.step(actions[i])
            self.touched[i] = done
            code = 0
            if done:
                code = ro.classify_landing(r.c, state.vx, state.vy, state.angle)
                self.final_code[i] = code
            echoed = (0.0, 0.0) if done else (float(actions[i][0]), float(actions[i][1]))
            msg += encode_record(state, reward, echoed, code)
        return bytes(msg)


def decode(message: bytes) -> np.ndarray:
    assert message[0] == MSG_TELEMETRY and (len(message) - 1) % 64 == 0
    return np.frombuffer(message, dtype="<f4", offset=1).reshape(-1, RECORD_FLOATS)
