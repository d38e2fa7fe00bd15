#!/usr/bin/env python
"""BASELINE config 2: 4,096 envs of the default nanogrid under the random policy, 24-step episodes -- the action slab of a
whole episode from ONE counter-based draw (sng_sample_actions), the episode from ONE launch (sng_rollout), auto-reset
included.  Prints the mean episode return next to the live reference's (-403.3, sigma 99.5 over 3,000
episodes of the same station) and the rate.  Needs a B200 (there is no CPU fallback)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv  # noqa: E402

KW = dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")


def main(num_envs=4096, episodes=50, seed=0):
    env = BatchedSmartNanogridEnv(num_envs, seed=seed, **KW)
    T = env.cfg.n_steps
    env.reset()
    actions = torch.empty(T, num_envs, env.cfg.act_dim, device=env.device)
    obs = torch.empty(T, num_envs, env.cfg.obs_dim, device=env.device)
    reward = torch.empty(T, num_envs, device=env.device)
    done = torch.empty(T, num_envs, dtype=torch.uint8, device=env.device)
    returns = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(episodes):
        env.random_actions(seed + 1, step0=k * T, n_steps=T, out=actions)     # env.action_space.sample() for every env and step
        env.rollout(actions, obs, reward, done)                                 # 24 x env.step(); the last one ends the day
        returns.append(reward.sum(0))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r = torch.cat(returns)
    assert bool(done[-1].all()) and env.error_flags() == 0
    print("random policy: %d episodes, mean return %.1f, sigma %.1f (live reference: -403.3, 99.5)" % (r.numel(), r.mean().item(), r.std().item()))
    print("%.3g env-steps/s by the wall clock of this Python loop (%d envs, %d launches; the two kernels of an episode take ~40 us on "
          "the GPU -- bench.py's c2 legs time them without the host in the way)" % (num_envs * T * episodes / dt, num_envs, 2 * episodes))
    env.close()


if __name__ == "__main__":
    main()
