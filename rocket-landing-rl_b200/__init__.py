"""B200-native batched rocket-landing environment step.

Drop-in for the environment-step path of adimail/rocket-landing-rl
(``backend/envs/lander.py`` ``RocketLandingEnv.reset/step`` and everything below it),
run for millions of independent rockets per GPU by hand-written sm_100a CUDA behind the
C ABI in ``include/rocket_landing_b200.h``.  There is no CPU fallback: importing the
environment classes without the built library raises.
"""
from .config import Config, load_params, default_config_path          # noqa: F401
from .spaces import Box                                               # noqa: F401
from .layout import (STATE_WORDS, fresh_state_soa, state_soa_from_reference,   # noqa: F401
                     reference_view)
from .stats import EpisodeStats, shard_range, allreduce_stats          # noqa: F401
from .env import RocketLandingVecEnv                                   # noqa: F401
from .single_env import RocketLandingEnv                               # noqa: F401
from .vec_env import RocketLandingSB3VecEnv                            # noqa: F401
from .normalize import DeviceVecNormalize, read_sb3_vecnormalize       # noqa: F401
from .simulation import RocketSimulation                               # noqa: F401
from .rollout import ActorCriticMLP, DeviceRolloutBuffer, FusedActorCritic, collect_rollout, pack_mlp   # noqa: F401

__all__ = ["Config", "load_params", "default_config_path", "Box", "STATE_WORDS",
           "fresh_state_soa", "state_soa_from_reference", "reference_view", "EpisodeStats",
           "shard_range", "allreduce_stats", "RocketLandingEnv", "RocketLandingVecEnv", "RocketLandingSB3VecEnv", "DeviceVecNormalize", "read_sb3_vecnormalize",
           "ActorCriticMLP", "DeviceRolloutBuffer", "FusedActorCritic", "pack_mlp", "collect_rollout", "RocketSimulation"]
