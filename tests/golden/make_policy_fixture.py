"""The reference's shipped agent as a fixture: weights, VecNormalize statistics and what that agent
does in the reference's OWN environment.

Run in the build container only (needs the reference checkout):

    python tests/golden/make_policy_fixture.py

Reads ``assets/model/v3/best_model.zip`` (``policy.pth``: the SB3 ``MlpPolicy`` state dict, 136 965
parameters) and ``assets/model/v3/vecnormalize.pkl`` (the pickled ``VecNormalize``: only its
``obs_rms`` / ``ret_rms`` / ``clip_obs`` / ``epsilon`` are read, through stand-in classes: neither
stable-baselines3 nor gymnasium is installed here), then drives the UNMODIFIED reference
``RocketLandingEnv`` (``oracle/ref_loader.py``) with the agent exactly as ``backend/rl/agent.py:160-
205`` does: normalise the observation with the loaded statistics, take the deterministic action
(the policy mean), clip it to the action space (what ``PPO.predict`` does for a ``Box``).

The same for ``assets/model/v2`` (an 8 -> 128 -> 128 policy trained for 950 000 steps; ``v1`` takes a
6-word observation of an earlier env and cannot fly this one).

Writes ``ref_policy_v3.npz`` / ``ref_policy_v2.npz``: the 13 weight tensors, the statistics, per-episode outcome of 1 500
reference episodes (landing class by ``backend/utils.evaluate_landing``, return, length) and 4 096
(observation, normalised observation, action mean) triples met on the way, for an open-loop check
of the fused kernels.
"""
from __future__ import annotations

import io
import os
import pickle
import sys
import zipfile

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ROOT)

from oracle.ref_loader import load_reference, reference_root      # noqa: E402

LANDING_CODE = {"safe": 1, "good": 2, "ok": 3, "unsafe": 4}


class _Bag:
    """Stand-in for the pickled SB3 / gymnasium classes: keeps whatever state pickle hands it."""
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith(("stable_baselines3", "gymnasium")):
            return _Bag
        if module.startswith("numpy") or module in ("builtins", "collections"):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"unexpected global {module}.{name}")


def load_checkpoint(version: str = "v3"):
    base = os.path.join(reference_root(), "assets", "model", version)
    with zipfile.ZipFile(os.path.join(base, "best_model.zip")) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    with open(os.path.join(base, "vecnormalize.pkl"), "rb") as fh:
        vn = _Unpickler(fh).load()
    stats = {"obs_mean": np.asarray(vn.obs_rms.mean, np.float64), "obs_var": np.asarray(vn.obs_rms.var, np.float64),
             "obs_count": float(vn.obs_rms.count), "ret_mean": float(vn.ret_rms.mean), "ret_var": float(vn.ret_rms.var),
             "ret_count": float(vn.ret_rms.count), "clip_obs": float(vn.clip_obs), "epsilon": float(vn.epsilon),
             "gamma": float(vn.gamma)}
    return sd, stats


def main(version: str = "v3", episodes: int = 1500) -> None:
    ref = load_reference()
    sd, st = load_checkpoint(version)
    w = {k: v.numpy().astype(np.float32) for k, v in sd.items()}

    def policy_mean(obs_norm: np.ndarray) -> np.ndarray:          # fp64 numpy: the fixture's reference numbers
        h = np.tanh(obs_norm @ w["mlp_extractor.policy_net.0.weight"].astype(np.float64).T + w["mlp_extractor.policy_net.0.bias"])
        h = np.tanh(h @ w["mlp_extractor.policy_net.2.weight"].astype(np.float64).T + w["mlp_extractor.policy_net.2.bias"])
        return h @ w["action_net.weight"].astype(np.float64).T + w["action_net.bias"]

    env = ref.RocketLandingEnv()
    cfg = ref.Config()
    np.random.seed(2024)
    outcomes, returns, lengths, trip_obs, trip_norm, trip_mean = [], [], [], [], [], []
    low, high = np.array([0.0, -1.0]), np.array([1.0, 1.0])
    for ep in range(episodes):
        obs, _ = env.reset()
        ret, done, t = 0.0, False, 0
        while not done:
            norm = np.clip((obs.astype(np.float64) - st["obs_mean"]) / np.sqrt(st["obs_var"] + st["epsilon"]),
                           -st["clip_obs"], st["clip_obs"]).astype(np.float32)      # VecNormalize.normalize_obs
            mean = policy_mean(norm.astype(np.float64))
            if len(trip_obs) < 4096 and (t % 7 == ep % 7):
                trip_obs.append(obs.copy()); trip_norm.append(norm.copy()); trip_mean.append(mean.astype(np.float32))
            action = np.clip(mean, low, high).astype(np.float32)                   # PPO.predict(deterministic=True) on a Box
            obs, r, term, trunc, info = env.step(action)
            ret += r
            t += 1
            done = term or trunc
        cls = LANDING_CODE[ref.evaluate_landing(info["raw_state"], cfg)["landing_message"]] if term else 0
        outcomes.append(cls); returns.append(ret); lengths.append(t)
    out = {"w_" + k.replace(".", "__"): v for k, v in w.items()}
    out.update({k: np.asarray(v) for k, v in st.items()})
    out.update(outcomes=np.asarray(outcomes, np.int8), returns=np.asarray(returns, np.float64),
               lengths=np.asarray(lengths, np.int32), trip_obs=np.asarray(trip_obs, np.float32),
               trip_norm=np.asarray(trip_norm, np.float32), trip_mean=np.asarray(trip_mean, np.float32))
    np.savez_compressed(os.path.join(_HERE, f"ref_policy_{version}.npz"), **out)
    oc = np.bincount(np.asarray(outcomes), minlength=5)
    print(f"{episodes} reference episodes with the {version} agent: none/safe/good/ok/unsafe = {oc.tolist()}, "
          f"mean return {np.mean(returns):.1f}, mean length {np.mean(lengths):.1f}, triples {len(trip_obs)}")


if __name__ == "__main__":
    for v in (sys.argv[1:] or ["v3", "v2"]):       # v1 was trained on an earlier, 6-word observation: not this env
        main(v)
