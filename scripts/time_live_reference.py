#!/usr/bin/env python
"""Times the LIVE reference (the unmodified Python env under the shims of oracle/ref_loader.py) on this machine's
host cores -- only possible where /root/reference exists (the build container, not the GPU box).  SURVEY 8(d):
P worker processes, each stepping its own reference env with uniform random float32 actions for a fixed wall time,
I/O-stubbed (the per-episode JSON dumps replaced by no-ops: the fair "simulation only" figure).
Writes one JSON line; the committed copy is profiles/r1_live_reference_cpu.json."""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTHONBREAKPOINT", "0")


def worker(n_spots, seconds, q):
    import numpy as np
    from oracle import ref_loader as rl
    env = rl.make_ref_env(number_of_chargers=n_spots)
    rl.seed_reference(os.getpid())
    lo, hi = env.action_space.low, env.action_space.high
    rng = np.random.default_rng(os.getpid())
    for _ in range(2):                       # warm-up episodes
        env.reset()
        for _ in range(24):
            env.step(rng.uniform(lo, hi).astype(np.float32))
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        env.reset()
        for _ in range(24):
            env.step(rng.uniform(lo, hi).astype(np.float32))
        steps += 24
    q.put((steps, time.perf_counter() - t0))


def main():
    from oracle import ref_loader as rl
    if not rl.reference_available():
        print(json.dumps({"unavailable": "reference tree not present"}))
        return
    procs = os.cpu_count() or 1
    out = {"what": "live reference Python env, I/O-stubbed, uniform random actions, reset() included", "processes": procs}
    for n_spots in (10, 64):
        q = mp.Queue()
        ps = [mp.Process(target=worker, args=(n_spots, 15.0, q)) for _ in range(procs)]
        for p in ps:
            p.start()
        res = [q.get() for _ in ps]
        for p in ps:
            p.join()
        out["n%d_env_steps_per_s_all_cores" % n_spots] = sum(s / t for s, t in res)
        out["n%d_env_steps_per_s_per_core" % n_spots] = sum(s / t for s, t in res) / procs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
