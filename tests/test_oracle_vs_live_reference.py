"""Differential test: float64 oracle == live reference, bit for bit.

Only runs where the reference tree exists (the build container); skipped on the GPU box.
The committed golden fixtures (tests/golden/) carry the same evidence to places where the
reference cannot travel.
"""
import itertools
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as rl  # noqa: E402

pytestmark = pytest.mark.skipif(not rl.reference_available(), reason="reference tree not present")


def branchy_actions(rng, lo, hi):
    """U(low, high) with exact zeros and saturated bounds sprinkled in."""
    a = rng.uniform(lo, hi).astype(np.float64)
    z = rng.random(a.shape)
    a[z < 0.15] = 0.0
    m = (z >= 0.15) & (z < 0.20)
    a[m] = hi[m]
    m = (z >= 0.20) & (z < 0.25)
    a[m] = lo[m]
    return a


def run_variant(kw, seed, episodes):
    from oracle.oracle import OracleBatch
    from smart_nanogrid_gym_b200.config import NanogridConfig
    full = dict(rl.DEFAULT_KW)
    full.update(kw)
    env = rl.make_ref_env(**kw)
    cfg = NanogridConfig(**full)
    tabs = rl.constant_tables(env)
    assert np.array_equal(tabs["price"], cfg.price[:48]) and tabs["price_max"] == cfg.price_max
    if cfg.pv:
        assert np.array_equal(tabs["pv_power"], cfg.pv_power)
        assert np.array_equal(tabs["irr"] / tabs["irr_max"], cfg.irr_norm)
    ob = OracleBatch(cfg, 1)
    rng = np.random.default_rng(seed)
    rl.seed_reference(seed)
    lo, hi = cfg.action_bounds()
    steps = 0
    for _ in range(episodes):
        obs, _info = env.reset()
        s = rl.export_schedule(env)
        ob.load_dense(0, s["soc"], s["occ"], s["cap"], s["req"], s["arrivals"], s["departures"],
                      s["pv_shift"], s["soc_b"] if cfg.batt else 0.0)
        assert np.array_equal(ob.observe()[0], obs)
        for t in range(cfg.n_steps):
            a = branchy_actions(rng, lo, hi)
            o, r, d, tr, info = env.step(a.copy())
            o2, r2, d2 = ob.step(a[None])
            assert np.array_equal(o2[0], o), (kw, t)
            assert r2[0] == r, (kw, t, r, r2[0])
            assert bool(d2[0]) == bool(d) and tr is False and info == {}
            steps += 1
        assert d
        # battery SoC survives the reset (quirk Q8) and per-spot SoC history matches
        cs = env.central_management_system.charging_station
        assert np.array_equal(cs.get_vehicles_state_of_charge()[:, :ob.W], ob.soc[0])
    return steps


FLAGS = list(itertools.product([True, False], repeat=5))


@pytest.mark.parametrize("penalty_mode", ["no_penalty", "on_departure", "sparse", "dense"])
def test_oracle_bit_exact_all_variants(penalty_mode):
    total = 0
    for idx, (pv, b, v2x, dc, rs) in enumerate(FLAGS):
        for n in (4, 10):
            kw = dict(number_of_chargers=n, pv_system_available_in_model=pv,
                      battery_system_available_in_model=b, vehicle_to_everything=v2x,
                      enable_different_vehicle_battery_capacities=dc, enable_requested_state_of_charge=rs,
                      vehicle_uncharged_penalty_mode=penalty_mode)
            total += run_variant(kw, seed=1000 * n + idx, episodes=2)
    assert total == len(FLAGS) * 2 * 2 * 24


def test_oracle_bit_exact_2h_interval():
    kw = dict(number_of_chargers=6, time_interval="2h")
    assert run_variant(kw, seed=5, episodes=3) == 36


@pytest.mark.parametrize("price_model", [1, 2, 3, 4])
def test_oracle_bit_exact_price_models(price_model):
    kw = dict(number_of_chargers=5, price_model=price_model)
    assert run_variant(kw, seed=price_model, episodes=2) == 48
