"""B200-native batched implementation of smart-nanogrid-gym's environment step.

Public surface:
  NanogridConfig            constructor kwargs + derived tables (pure Python)
  BatchedSmartNanogridEnv   E envs per CUDA launch, torch tensors, zero-copy (needs a GPU)
  SmartNanogridEnv, make    the reference's single-env gym API (E = 1, float64 build)
  ScheduleRecords, ...      schedule containers / initial_values.json I/O
  SmartNanogridVecEnv       Stable-Baselines3-style VecEnv facade (numpy in / out, terminal_observation in infos)
  EpisodeRecorder           prediction_results.json export for one env (trace.py)
  collect_rollout, ...      on-device PPO rollout collection + GAE kernel (rollout.py)
"""
from .config import NanogridConfig, PENALTY_MODES  # noqa: F401
from .schedule import (ScheduleRecords, records_from_dense, dense_from_records, concat_records,  # noqa: F401
                       load_initial_values_json, save_initial_values_json)

__all__ = ["NanogridConfig", "BatchedSmartNanogridEnv", "SmartNanogridEnv", "make", "ScheduleRecords",
           "records_from_dense", "dense_from_records", "concat_records", "load_initial_values_json",
           "save_initial_values_json"]


def __getattr__(name):  # torch / CUDA are only imported when the env classes are used
    if name == "BatchedSmartNanogridEnv":
        from .batched_env import BatchedSmartNanogridEnv
        return BatchedSmartNanogridEnv
    if name == "SmartNanogridVecEnv":
        from .vec_env import SmartNanogridVecEnv
        return SmartNanogridVecEnv
    if name == "EpisodeRecorder":
        from .trace import EpisodeRecorder
        return EpisodeRecorder
    if name in ("MlpPolicy", "RolloutBuffer", "collect_rollout", "GraphedRollout"):
        from . import rollout
        return getattr(rollout, name)
    if name in ("SmartNanogridEnv", "make", "register_with_gym", "ENV_ID"):
        from . import env
        return getattr(env, name)
    raise AttributeError(name)
