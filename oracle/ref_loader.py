"""Import the UNMODIFIED reference (adimail/rocket-landing-rl) in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()``, ``bench.py``'s CPU legs and
the golden-vector generator use it.

The reference's hot path (``backend/physics``, ``backend/rocket``,
``backend/rl/reward.py``, ``backend/envs/lander.py``) is plain Python + numpy, but a
plain ``import backend`` fails here because three third-party packages that have
nothing to do with the arithmetic are absent:

* ``backend/__init__.py:1`` imports ``tornado.web``,
* ``backend/envs/lander.py:1-2`` imports ``gymnasium``,
* ``backend/rl/agent.py:8-9`` imports ``stable_baselines3``.

``load_reference()`` pre-seeds ``sys.modules`` with inert stand-ins for exactly
those three, puts the reference root on ``sys.path`` and returns the reference's own
modules.  None of the stand-ins touches a number.  This module only works where the
reference checkout exists (``$ROCKET_REF`` or ``/root/reference``); it does not
travel to the GPU box, which is why golden vectors are committed under
``tests/golden/``.
"""
from __future__ import annotations

import logging
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

_REF_CACHE = None


_ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_backend.bin")
_UNPACKED = None


def _unpack_compiled() -> str | None:
    """oracle/_ref/reference_backend.bin (byte-compiled by oracle/make_ref.py) unpacked into a scratch
    directory, if it exists and was built by THIS interpreter version."""
    global _UNPACKED
    if _UNPACKED is not None:
        return _UNPACKED or None
    import importlib.util
    import json
    import tempfile
    import zipfile
    _UNPACKED = ""
    try:
        with zipfile.ZipFile(_ARCHIVE) as z:
            if json.loads(z.read("BUILD.json"))["magic"] != importlib.util.MAGIC_NUMBER.hex():
                return None
            dest = tempfile.mkdtemp(prefix="rocket_ref_compiled_")
            z.extractall(dest)
        _UNPACKED = dest
    except Exception:   # noqa: BLE001
        return None
    return _UNPACKED


def reference_root(allow_compiled: bool = False) -> str | None:
    """The reference checkout ($ROCKET_REF or /root/reference); with ``allow_compiled`` also the
    byte-compiled copy in ``oracle/_ref`` (the only form that exists on the GPU box; used for TIMING
    the reference there, never for parity -- the goldens under tests/golden are the parity anchor)."""
    for cand in (os.environ.get("ROCKET_REF"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "backend", "envs", "lander.py")):
            return cand
    if allow_compiled:
        return _unpack_compiled()
    return None


def reference_available() -> bool:
    return reference_root() is not None


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    # --- tornado: only class/attribute lookups at import time ---------------------
    if "tornado" not in sys.modules:
        class _Any:
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, item):
                return _Any()

            def __call__(self, *a, **k):
                return _Any()

        web = _mod("tornado.web", Application=_Any, RequestHandler=_Any,
                   StaticFileHandler=_Any)
        ws = _mod("tornado.websocket", WebSocketHandler=_Any,
                  WebSocketClosedError=type("WebSocketClosedError", (Exception,), {}))
        ioloop = _mod("tornado.ioloop", IOLoop=_Any, PeriodicCallback=_Any)
        _mod("tornado", web=web, websocket=ws, ioloop=ioloop)

    # --- gymnasium: Env with a no-op reset(seed=) and a Box that stores bounds ----
    if "gymnasium" not in sys.modules:
        class Space:
            pass

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                self.low = np.asarray(low, dtype=self.dtype)
                self.high = np.asarray(high, dtype=self.dtype)
                self.shape = self.low.shape

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

        class Env:
            def reset(self, *, seed=None, options=None):
                return None

        spaces = _mod("gymnasium.spaces", Box=Box, Space=Space)
        _mod("gymnasium", Env=Env, spaces=spaces)

    # --- stable_baselines3: names only (never instantiated on the hot path) -------
    if "stable_baselines3" not in sys.modules:
        class _Dummy:
            def __init__(self, *a, **k):
                raise RuntimeError("stable_baselines3 stand-in: not available here")

        vec_env = _mod("stable_baselines3.common.vec_env", DummyVecEnv=_Dummy,
                       VecNormalize=_Dummy, SubprocVecEnv=_Dummy)
        common = _mod("stable_baselines3.common", vec_env=vec_env)
        _mod("stable_baselines3", PPO=_Dummy, common=common)


def load_reference(allow_compiled: bool = False, prefer_compiled: bool = False) -> SimpleNamespace:
    """Return the reference's own modules (``RocketLandingEnv``, ``Rocket``,
    ``PhysicsEngine``, ``calculate_reward``, ``evaluate_landing``, ``Config``)."""
    global _REF_CACHE
    if _REF_CACHE is not None:
        return _REF_CACHE
    root = reference_root(allow_compiled)
    if prefer_compiled and _unpack_compiled():
        root = _unpack_compiled()
    if root is None:
        raise RuntimeError("reference checkout not found (set $ROCKET_REF)")
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    # backend/rl/agent.py:14-20 creates logs/agents/ under the CWD at import time:
    # contain that side effect in a scratch directory.
    import tempfile
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="rocket_ref_import_")
    os.chdir(scratch)
    try:
        from backend.envs.lander import RocketLandingEnv
        from backend.rocket.base import Rocket
        from backend.physics.engine import PhysicsEngine
        from backend.rl.reward import calculate_reward
        from backend.utils import evaluate_landing
        from backend.config import Config
        import backend.simulation.config as sim_config
    finally:
        os.chdir(cwd)
    logging.disable(logging.CRITICAL)   # lander.py:12 basicConfig(INFO) + per-step debug
    _REF_CACHE = SimpleNamespace(
        root=root, RocketLandingEnv=RocketLandingEnv, Rocket=Rocket,
        PhysicsEngine=PhysicsEngine, calculate_reward=calculate_reward,
        evaluate_landing=evaluate_landing, Config=Config, sim_config=sim_config)
    return _REF_CACHE
