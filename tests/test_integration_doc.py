"""INTEGRATION.md section 2 shows the binding a maintainer of the reference would add
(``backend/envs/b200.py``: plain ctypes, no dependency on this repository's Python package).
This test EXTRACTS that block from the document and runs it, so the documented binding cannot rot:
on CPU it is compiled and its struct is checked against the library's; on the GPU box it is driven
through reset / step against the package's own VecEnv."""
import ctypes as C
import os
import re
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_binding():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = next(b for b in blocks if "class B200VecEnv" in b)
    from rocket_landing_rl_b200 import _lib
    from rocket_landing_rl_b200.config import Config
    # the reference's backend.config.Config has the same get("a.b.c") surface as this repo's Config
    backend = types.ModuleType("backend")
    backend_config = types.ModuleType("backend.config")
    backend_config.Config = Config
    backend.config = backend_config
    saved = {k: sys.modules.get(k) for k in ("backend", "backend.config")}
    sys.modules.update({"backend": backend, "backend.config": backend_config})
    os.environ["ROCKET_B200_LIB"] = _lib.LIB_PATH
    try:
        ns = {"__name__": "backend.envs.b200"}
        exec(compile(code, "INTEGRATION.md:section-2", "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ns


def test_documented_binding_compiles_and_mirrors_the_struct():
    from rocket_landing_rl_b200.config import Config, RlParams, load_params
    ns = _load_binding()
    assert C.sizeof(ns["RlParams"]) == C.sizeof(RlParams)
    assert [f[0] for f in ns["RlParams"]._fields_] == [f[0] for f in RlParams._fields_]
    p = ns["params_from_config"](Config())
    assert bytes(p) == bytes(load_params())           # every constant, bit for bit


@pytest.mark.gpu
def test_documented_binding_runs_on_the_gpu():
    from rocket_landing_rl_b200 import RocketLandingSB3VecEnv
    ns = _load_binding()
    n = 512
    doc = ns["B200VecEnv"](n, seed=3)
    own = RocketLandingSB3VecEnv(n, seed=3)
    np.testing.assert_array_equal(doc.reset(), own.reset())
    rng = np.random.default_rng(0)
    seen = 0
    for t in range(150):
        a = rng.uniform([0, -1], [1, 1], (n, 2)).astype(np.float32)
        o1, r1, d1, i1 = doc.step(a)
        o2, r2, d2, i2 = own.step(a)
        np.testing.assert_array_equal(o1, o2)
        np.testing.assert_array_equal(r1, r2)
        np.testing.assert_array_equal(d1, d2)
        for i in np.flatnonzero(d1):
            assert i1[i]["episode"] == i2[i]["episode"]
            np.testing.assert_array_equal(i1[i]["terminal_observation"], i2[i]["terminal_observation"])
            assert i1[i]["TimeLimit.truncated"] == i2[i]["TimeLimit.truncated"]
            seen += 1
    assert seen > 0.9 * n
    doc.close()
    own.close()
