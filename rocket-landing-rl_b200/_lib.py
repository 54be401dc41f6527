"""ctypes loader of ``librocket_landing_b200.so`` (the C ABI in
``include/rocket_landing_b200.h``).  No fallback: if the library is missing or the ABI
version disagrees this raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

from .config import RlParams

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "librocket_landing_b200.so"
# RL_B200_LIBRARY points the loader at another build of the SAME sources (tools/build_variants.py
# makes them for kernel experiments); it is still this library or nothing -- there is no fallback.
LIB_PATH = os.environ.get("RL_B200_LIBRARY") or os.path.join(_PKG_DIR, LIB_NAME)
ABI_VERSION = 2

# every symbol include/rocket_landing_b200.h declares
EXPORTS = ("rl_abi_version", "rl_last_error", "rl_params_default", "rl_state_bytes", "rl_create",
           "rl_destroy", "rl_num_envs", "rl_fuel_unit", "rl_fuel_unit_for", "rl_pos_unit", "rl_pos_unit_for", "rl_uses_folded_constants", "rl_get_epoch", "rl_set_epoch", "rl_reset", "rl_step",
           "rl_set_state", "rl_get_state", "rl_stats", "rl_step_host", "rl_reset_host", "rl_get_state_host",
           "rl_launch_count", "rl_vecnorm_last_error", "rl_vecnorm_create", "rl_vecnorm_destroy",
           "rl_vecnorm_reset", "rl_vecnorm_step", "rl_vecnorm_normalize_obs", "rl_vecnorm_get_stats",
           "rl_vecnorm_set_stats", "rl_vecnorm_launch_count", "rl_rollout_last_error", "rl_gae", "rl_episode_flags", "rl_sim_step",
           "rl_sim_step_host", "rl_policy_last_error", "rl_mlp_blob_bytes", "rl_mlp_forward", "rl_policy_sample",
           "rl_policy_value_sample")

_lib = None


class RocketLibraryError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RocketLibraryError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built.  Run "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C "
            f"rocket-landing-rl_b200/csrc`).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise RocketLibraryError(f"{LIB_PATH} does not export {name}")
    L.rl_abi_version.restype = C.c_int
    if L.rl_abi_version() != ABI_VERSION:
        raise RocketLibraryError(f"ABI mismatch: library {L.rl_abi_version()}, binding {ABI_VERSION}")
    vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
    L.rl_last_error.restype = C.c_char_p
    L.rl_params_default.argtypes = [C.POINTER(RlParams)]
    L.rl_params_default.restype = None
    L.rl_state_bytes.argtypes = [i64]
    L.rl_state_bytes.restype = i64
    L.rl_create.argtypes = [C.POINTER(RlParams), i64, i64, u64, i32, vp, C.POINTER(vp)]
    L.rl_destroy.argtypes = [vp]
    L.rl_num_envs.argtypes = [vp]
    L.rl_num_envs.restype = i64
    L.rl_fuel_unit.argtypes = [vp]
    L.rl_fuel_unit.restype = C.c_double
    L.rl_fuel_unit_for.argtypes = [C.POINTER(RlParams)]
    L.rl_fuel_unit_for.restype = C.c_double
    L.rl_pos_unit.argtypes = [vp]
    L.rl_pos_unit.restype = C.c_double
    L.rl_pos_unit_for.argtypes = [C.POINTER(RlParams)]
    L.rl_pos_unit_for.restype = C.c_double
    L.rl_uses_folded_constants.argtypes = [vp]
    L.rl_get_epoch.argtypes = [vp]
    L.rl_get_epoch.restype = u64
    L.rl_set_epoch.argtypes = [vp, u64]
    L.rl_reset.argtypes = [vp, vp, vp, vp]
    L.rl_step.argtypes = [vp] + [vp] * 10 + [i32, i32, vp]
    L.rl_set_state.argtypes = [vp, vp, vp, vp]
    L.rl_get_state.argtypes = [vp, vp, vp]
    L.rl_stats.argtypes = [vp, vp, i32, vp]
    L.rl_step_host.argtypes = [vp] + [vp] * 10 + [i32, i32]
    L.rl_reset_host.argtypes = [vp, vp]
    L.rl_get_state_host.argtypes = [vp, vp]
    L.rl_launch_count.argtypes = [vp]
    L.rl_launch_count.restype = i64
    f64 = C.c_double
    L.rl_vecnorm_last_error.restype = C.c_char_p
    L.rl_vecnorm_create.argtypes = [i64, f64, f64, f64, f64, i32, C.POINTER(vp)]
    L.rl_vecnorm_destroy.argtypes = [vp]
    L.rl_vecnorm_reset.argtypes = [vp, vp, vp, i32, i32, vp]
    L.rl_vecnorm_step.argtypes = [vp] + [vp] * 7 + [i32, i32, i32, vp]
    L.rl_vecnorm_normalize_obs.argtypes = [vp, vp, vp, i64, vp]
    L.rl_vecnorm_get_stats.argtypes = [vp, vp]
    L.rl_vecnorm_set_stats.argtypes = [vp, vp]
    L.rl_vecnorm_launch_count.argtypes = [vp]
    L.rl_vecnorm_launch_count.restype = i64
    L.rl_rollout_last_error.restype = C.c_char_p
    L.rl_sim_step.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rl_sim_step_host.argtypes = [vp, vp, vp]
    L.rl_gae.argtypes = [vp] * 7 + [i32, i64, f64, f64, vp]
    L.rl_episode_flags.argtypes = [vp] * 5 + [i64, vp]
    L.rl_policy_last_error.restype = C.c_char_p
    L.rl_mlp_blob_bytes.restype = i64
    L.rl_mlp_forward.argtypes = [vp, vp, vp, i64, i32, vp]
    L.rl_policy_sample.argtypes = [vp, vp, vp, u64, u64, i64, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp, vp, vp, i64, vp]
    L.rl_policy_value_sample.argtypes = [vp, vp, vp, vp, u64, u64, i64, C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp, vp,
                                         vp, vp, i64, vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rl_last_error().decode("utf-8", "replace")
        raise RocketLibraryError(f"{what} failed (code {rc}): {msg}")
