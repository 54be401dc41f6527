"""Import shim: ``rocket-landing-rl_b200/`` (the package directory the layout names) is not
a valid Python identifier, so ``import rocket_landing_rl_b200`` resolves here and this
module re-exports that directory as its own package body."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "rocket-landing-rl_b200")
__path__.insert(0, _impl)
with open(_os.path.join(_impl, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _os, _fh
