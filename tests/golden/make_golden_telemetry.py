"""Generate ``ref_telemetry.npz`` from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_telemetry.py

Drives the reference's own ``SimulationController.step`` (backend/simulation/controller.py:223),
``evaluate_landing`` (backend/utils.py:1) and ``BinaryProtocol`` (backend/protocol.py) for 12
rockets x 170 ticks; only the ~15 lines of glue of ``RocketWebSocketHandler.send_state_update`` /
``send_binary_telemetry`` (backend/handler/websocket_handler.py:120-171) are restated here, because
instantiating the handler needs a live Tornado connection.  Half of the rockets fly a braking
controller so that safe / good landings appear, the rest random actions (unsafe)."""
import os
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(_HERE)))
from oracle.ref_loader import load_reference            # noqa: E402
from oracle.pin_against_reference import REF_KEYS       # noqa: E402


def main():
    ref = load_reference()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="rocket_ref_sim_"))      # the controller opens a log file under ./logs
    try:
        from backend.simulation.controller import SimulationController
        from backend.protocol import BinaryProtocol
        n, ticks = 12, 170
        sim = SimulationController(n, rl_agent=None)
        sim.log_state = sim.log_action = sim.log_reward = False
        rng = np.random.default_rng(3)
        lo = np.array([-1500, 1800, -25, -200, 0, 0, -15, -10, 34000, 370000], float)
        hi = np.array([1500, 2200, 25, -150, 0, 0, 15, 10, 38000, 410000], float)
        init = rng.uniform(lo, hi, size=(n, 10))
        init[:6, 1] = rng.uniform(300, 600, 6)                # the braking half starts low and slow enough to land
        init[:6, 3] = rng.uniform(-60, -40, 6)
        init[:6, 6:8] = rng.uniform(-2, 2, (6, 2))
        for rc, s in zip(sim.rockets, init):
            rc.rocket.state = {k: float(v) for k, v in zip(REF_KEYS, s)}
            rc.rocket.previous_state = rc.rocket.calculate_consistent_previous_state(rc.rocket.state, rc.rocket.dt)
        final_outcomes = {}
        messages, actions_log = [], np.zeros((ticks, n, 2), np.float32)
        for t in range(ticks):
            acts = []
            for i, rc in enumerate(sim.rockets):
                st = rc.rocket.state
                if i < 6:
                    m = st["mass"] + st["fuelMass"]
                    a_max = 1.0e7 / m - 9.81
                    burn = st["vy"] ** 2 / (2 * max(st["y"] - 2.0, 0.5)) > 0.8 * a_max
                    hover = (9.81 * m / 1.0e7) * (1.0 + np.clip(-(st["vy"] + 6.0) * 0.15, -0.9, 3.0))
                    th = 1.0 if burn else (hover if st["y"] < 60 else 0.0)
                    cg = float(np.clip(-(st["angle"] * 0.9 + st["angularVelocity"] * 1.6), -1.3, 1.3))
                    a = np.array([th, cg], np.float32)
                else:
                    a = rng.uniform([-0.2, -1.3], [1.2, 1.3]).astype(np.float32)
                actions_log[t, i] = a
                acts.append({"throttle": float(a[0]), "coldGas": float(a[1])})
            states, rewards, dones = sim.step(acts)
            # --- websocket_handler.py:153-171 -----------------------------------------------------
            prev_actions = sim.prev_action_taken.copy()
            payload, landing = [], [None] * n
            for i in range(n):
                payload.append(prev_actions[i] if states[i] and not dones[i] else {"throttle": 0.0, "coldGas": 0.0})
                if dones[i] and states[i] is not None:
                    msg = ref.evaluate_landing(states[i], sim.config)["landing_message"]
                    landing[i] = msg
                    final_outcomes[i] = {"status": msg, "reward": rewards[i]}
            # --- websocket_handler.py:120-145 -----------------------------------------------------
            data = bytearray(BinaryProtocol.encode_telemetry_header())
            for i in range(n):
                ls = landing[i]
                if ls is None and i in final_outcomes:
                    ls = final_outcomes[i]["status"]
                data.extend(BinaryProtocol.encode_rocket_state(state=states[i], reward=rewards[i],
                                                               action=payload[i], landing_status=ls))
            messages.append(np.frombuffer(bytes(data), np.uint8))
    finally:
        os.chdir(cwd)
    out = os.path.join(_HERE, "ref_telemetry.npz")
    np.savez_compressed(out, init_state=init, actions=actions_log, messages=np.stack(messages))
    rec = np.stack(messages)[:, 1:].copy().view("<f4").reshape(ticks, n, 16)
    print(out, os.path.getsize(out), "bytes; final landing codes", rec[-1, :, 14].tolist(),
          "active at end", rec[-1, :, 15].sum())


if __name__ == "__main__":
    main()
