"""Stable-Baselines3-style VecEnv facade over `BatchedSmartNanogridEnv`.

The reference trains with `PPO("MlpPolicy", env)` (solvers/RL/ppo_train.py:89-92); SB3 wraps the single env
in a DummyVecEnv and talks to it through the VecEnv protocol: numpy in / numpy out, `step_wait` returning
`(obs, rewards, dones, infos)`, automatic reset of finished envs with the last observation of the episode
in `infos[i]["terminal_observation"]`.  This class speaks that protocol (duck-typed: SB3 is not in the
build image) for E envs on one GPU, so an SB3 trainer can be pointed at it unchanged.  Every call crosses
PCIe (pinned host buffers, `sng_step_host`); trainers that can consume CUDA tensors should drive
`BatchedSmartNanogridEnv` / `rollout.collect_rollout` directly.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from .batched_env import BatchedSmartNanogridEnv


class SmartNanogridVecEnv:
    metadata = {"render_modes": []}

    def __init__(self, num_envs: int, device="cuda:0", seed: int = 0, **kwargs):
        self.env = BatchedSmartNanogridEnv(num_envs, device=device, seed=seed, precision="float32", auto_reset=True,
                                           want_terminal_obs=True, **kwargs)
        self.num_envs = int(num_envs)
        self.observation_space = self.env.observation_space
        self.action_space = self.env.action_space
        E, A, D = self.num_envs, self.env.cfg.act_dim, self.env.cfg.obs_dim
        self._a = torch.zeros(E, A, dtype=torch.float32).pin_memory()
        self._o = torch.zeros(E, D, dtype=torch.float32).pin_memory()
        self._r = torch.zeros(E, dtype=torch.float32).pin_memory()
        self._d = torch.zeros(E, dtype=torch.uint8).pin_memory()
        self._pending: Optional[np.ndarray] = None
        self.reset_infos: List[dict] = [{} for _ in range(E)]

    # ---- VecEnv protocol ---------------------------------------------------------------
    def reset(self) -> np.ndarray:
        return self.env.reset().cpu().numpy()

    def step_async(self, actions: np.ndarray) -> None:
        self._pending = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, -1)

    def step_wait(self):
        if self._pending is None:
            raise RuntimeError("step_wait() without step_async()")
        self._a.copy_(torch.from_numpy(self._pending))
        self._pending = None
        self.env.step_host(self._a, self._o, self._r, self._d)
        self.env.check_errors()                      # the reference raises inside step()
        obs = self._o.numpy().copy()
        dones = self._d.numpy().astype(bool)
        infos: List[dict] = [{} for _ in range(self.num_envs)]
        if dones.any():
            term = self.env.terminal_obs[torch.from_numpy(np.flatnonzero(dones)).to(self.env.device)].cpu().numpy()
            for k, i in enumerate(np.flatnonzero(dones)):
                # SB3 semantics: the episode ended by termination (the reference never truncates)
                infos[i] = {"terminal_observation": term[k], "TimeLimit.truncated": False}
        return obs, self._r.numpy().copy(), dones, infos

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        self.env.close()

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        self.env.seed(seed)
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    # ---- attribute plumbing SB3 expects from a VecEnv ------------------------------------
    def _indices(self, indices) -> Sequence[int]:
        if indices is None:
            return range(self.num_envs)
        return [indices] if isinstance(indices, int) else indices

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        return [getattr(self.env, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self.env, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        out = getattr(self.env, method_name)(*args, **kwargs)
        return [out for _ in self._indices(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode: str = "human"):
        return None
