"""Batch-size sweep of the two step kernels on the C4 station (N = 10, default shape): the one-block-per-warp kernel
(set_pipeline 5) against the one-lane-per-spot kernel (set_pipeline 4, sng_lanes.cuh), per-step launches replayed from a
CUDA graph (with programmatic dependent launch) and 24 steps per launch through sng_rollout.  Prints us per step.
Usage: python scripts/lanes_sweep.py [--sizes 4096,16384,...] [--spots 10]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv  # noqa: E402


def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,4096,8192,16384,32768,65536,131072")
    ap.add_argument("--spots", type=int, default=10)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warps", type=int, default=0, help="warps per CTA (0 = default)")
    ap.add_argument("--two", action="store_true", help="also the block kernel with two lanes per env (set_pipeline 3)")
    args = ap.parse_args()
    dev = "cuda:0"
    out = {}
    for E in [int(x) for x in args.sizes.split(",")]:
        row = {}
        for name, variant in (("block_per_warp", 5), ("lane_per_spot", 4), ("two_lanes_per_env", 3))[:3 if args.two else 2]:
            env = BatchedSmartNanogridEnv(E, device=dev, seed=0, precision="float32", auto_reset=True, number_of_chargers=args.spots,
                                          charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
            env.set_pipeline(variant)
            env.set_tuning(warps_per_cta=args.warps)
            env.reset()
            g = torch.Generator(device=dev).manual_seed(5)
            acts = torch.stack([env.sample_actions(g) for _ in range(args.steps)]).contiguous()
            obs = torch.empty(args.steps, E, env.cfg.obs_dim, device=dev)
            rew = torch.empty(args.steps, E, device=dev)
            done = torch.empty(args.steps, E, device=dev, dtype=torch.uint8)
            env.rollout(acts, obs, rew, done)
            ms_roll = timed(lambda: env.rollout(acts, obs, rew, done), 200) / args.steps
            for pdl in (0, 1):
                env.set_launch_mode(pdl)
                s = torch.cuda.Stream()
                a0 = acts[0]
                with torch.cuda.stream(s):
                    for k in range(3):
                        env.step(a0)
                    s.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=s):
                        for k in range(args.steps):
                            env.step(a0)
                graph.replay()
                row["%s_step_pdl%d_us" % (name, pdl)] = 1e3 * timed(graph.replay, 200) / args.steps
            row[name + "_rollout_us"] = 1e3 * ms_roll
            assert env.error_flags() == 0
            env.close()
        out[E] = row
        print(E, json.dumps({k: round(v, 3) for k, v in row.items()}), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
