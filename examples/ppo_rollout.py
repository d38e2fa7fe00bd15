#!/usr/bin/env python
"""BASELINE config 3: PPO rollout collection over 65,536 envs -- SB3's MlpPolicy shape in plain torch, actions
clipped to the Box, the step kernel writing observations / rewards / dones straight into the rollout buffer, GAE by
the sng_gae kernel, the whole rollout replayed as one CUDA graph.  Needs a B200."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv, GraphedRollout, MlpPolicy, RolloutBuffer  # noqa: E402


def main(n_envs=65536, n_steps=24, rollouts=20):
    env = BatchedSmartNanogridEnv(n_envs, seed=0, number_of_chargers=10, charging_mode="bounded",
                                  vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
    torch.manual_seed(0)
    policy = MlpPolicy(env.cfg.obs_dim, env.cfg.act_dim).to("cuda:0")
    buf = RolloutBuffer(n_steps, n_envs, env.cfg.obs_dim, env.cfg.act_dim, "cuda:0", gamma=0.99, gae_lambda=0.95)
    obs = env.reset()
    starts = torch.ones(n_envs, dtype=torch.uint8, device="cuda:0")
    collect = GraphedRollout(env, policy, buf)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(rollouts):
        obs, starts = collect(obs, starts)
        # a trainer would now run its PPO epochs on buf.observations[:-1], buf.raw_actions, buf.advantages, buf.returns ...
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d rollouts of %d steps x %d envs: %.3g env-steps/s, mean step reward %.3f, mean advantage %.3f"
          % (rollouts, n_steps, n_envs, rollouts * n_steps * n_envs / dt, buf.rewards.mean().item(), buf.advantages.mean().item()))
    env.close()


if __name__ == "__main__":
    main()
