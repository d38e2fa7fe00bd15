#!/usr/bin/env python
"""BASELINE config 1: the default nanogrid (10 charging spots, PV + battery, 24 hourly steps) driven by the
reference's rule-based controller (solvers/RBC/rbc.py) for one episode -- through the single-env gym API, then
the same rule on a batch of envs on the device.  Needs a B200 (there is no CPU fallback)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv, EpisodeRecorder, make  # noqa: E402

KW = dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")


def main():
    # --- one env, the reference's API (reset -> (obs, {}), step -> (obs, reward, terminated, truncated, info)) ---
    env = make("SmartNanogridEnv-v0", **KW)
    obs, _ = env.reset(seed=0)
    ret, t = 0.0, 0
    terminated = False
    while not terminated:
        a = env._b.rbc_actions(torch.tensor(obs[None, :], device="cuda:0"))[0].cpu().numpy()
        obs, r, terminated, truncated, info = env.step(a)
        ret += r
        t += 1
    print("single env: %d steps, episode return %.3f" % (t, ret))
    env.close()

    # --- 65,536 envs at once, same rule, with a prediction_results.json trace of env 0 ---
    benv = BatchedSmartNanogridEnv(65536, seed=0, want_diagnostics=True, want_terminal_obs=True, **KW)
    obs = benv.reset()
    rec = EpisodeRecorder(benv, 0)
    total = torch.zeros(benv.num_envs, device="cuda:0")
    for _ in range(24):
        obs, r, done, _, _ = rec.step(benv.rbc_actions(obs))
        total += r
    print("batched: mean episode return %.3f +- %.3f over %d envs" % (total.mean().item(), total.std().item(), benv.num_envs))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prediction_results.json")
    rec.save(out)
    print("trace of env 0 written to", out, "(keys as in the reference's files/prediction_results.json)")
    benv.close()


if __name__ == "__main__":
    main()
