#!/usr/bin/env python
"""One-off wide parity sweep on a GPU box (not part of the test suite: several minutes): every combination of
{PV, battery, V2X, different capacities, requested SoC} x 4 penalty modes x N in {4, 10, 64} x {1 h, 15 min (N=64 only)},
float32 and float64 builds vs the float64 oracle, sampled schedules, fused auto-reset, branchy actions, 2 episodes + 5 steps.
Prints one line per configuration and a summary; exit code 1 on any mismatch."""
import itertools
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import OracleBatch  # noqa: E402
from parity_utils import assert_close_f32, branchy_actions, penalty_margin_distance  # noqa: E402
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv  # noqa: E402


def ulp(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 5e-324)


def run(kw, precision, E=384, seed=11):
    f64 = precision == "float64"
    env = BatchedSmartNanogridEnv(E, seed=seed, precision=precision, want_terminal_obs=True, charging_mode="bounded", **kw)
    cfg = env.cfg
    ob = OracleBatch(cfg, E, n_threads=8)
    obs = env.reset().cpu().numpy()
    ob.sample(seed, 0, 0)
    o_ref = ob.observe()
    assert np.array_equal(obs, o_ref) if f64 else np.allclose(obs, o_ref, rtol=1e-5, atol=1e-6)
    rng = np.random.default_rng(3)
    lo, hi = cfg.action_bounds()
    episode = np.zeros(E, np.uint32)
    skipped = 0
    for s in range(2 * cfg.n_steps + 5):
        a = branchy_actions(rng, lo, hi, (E,))
        a_dev = torch.tensor(a, device="cuda:0", dtype=env.real)
        a_or = a if f64 else a.astype(np.float32).astype(np.float64)
        near = penalty_margin_distance(ob) < (0 if f64 else 1e-5)
        skipped += int(near.sum())
        o_ref, r_ref, d_ref = ob.step(a_or)
        o, r, d, _, _ = env.step(a_dev)
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        assert np.array_equal(d, d_ref), s
        if d_ref.any():
            tob = env.terminal_obs.cpu().numpy()
            episode += 1
            ob.sample(seed, 0, episode)
            o_term, o_ref = o_ref, ob.observe()
            assert np.array_equal(tob, o_term) if f64 else np.allclose(tob, o_term, rtol=1e-5, atol=1e-6)
        if f64:
            assert np.array_equal(o, o_ref), s
            assert ulp(r, r_ref).max() <= 4, s
        else:
            assert_close_f32("obs", o, o_ref, atol=1e-6)
            assert_close_f32("reward", r, r_ref, atol=1e-5, mask=~near)
    assert env.error_flags() == int(np.bitwise_or.reduce(ob.err))
    env.close()
    return skipped


def main():
    t0 = time.time()
    n_ok = n_bad = 0
    flags = ("pv_system_available_in_model", "battery_system_available_in_model", "vehicle_to_everything",
             "enable_different_vehicle_battery_capacities", "enable_requested_state_of_charge")
    for N, interval in ((4, "1h"), (10, "1h"), (64, "15min")):
        for bits in itertools.product((False, True), repeat=5):
            for pen in ("no_penalty", "on_departure", "sparse", "dense"):
                if N == 64 and pen in ("no_penalty", "on_departure") and bits[2]:
                    continue      # trim the slowest corner (V2X at 64 spots) to two penalty modes
                kw = dict(zip(flags, bits), number_of_chargers=N, time_interval=interval, vehicle_uncharged_penalty_mode=pen)
                for precision in ("float32", "float64"):
                    try:
                        run(kw, precision)
                        n_ok += 1
                    except AssertionError as ex:
                        n_bad += 1
                        print("MISMATCH", N, interval, bits, pen, precision, str(ex)[:200], flush=True)
    print("parity sweep: %d configurations x precisions ok, %d mismatches, %.0f s" % (n_ok, n_bad, time.time() - t0), flush=True)
    sys.exit(1 if n_bad else 0)


if __name__ == "__main__":
    main()
