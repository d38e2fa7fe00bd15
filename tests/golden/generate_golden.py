"""Generates the committed golden fixtures by RUNNING THE LIVE REFERENCE in the build
container (`/root/reference`, imported under the shims of oracle/ref_loader.py).

    PYTHONBREAKPOINT=0 python tests/golden/generate_golden.py

The reference is pure Python and cannot travel to the GPU box, so its outputs travel
instead.  Everything written here is data produced by the unmodified reference code:

  g1_*.json / g2_*.json       the reference's own recorded episodes (copied data files,
                              solvers/RL/{training_files,single_prediction_files}/)
  ref_tables.npz              PV / price tables for every runnable (dt, price model)
  ref_tables_multiday.npz     PVSystemManager(2, dt): flat 3-day series, per-day rows, max, power (dt = 1 h, 2 h)
  ref_schedules_seeded.npz    schedules from the reference generator under np.random.seed(s)
  ref_variants.npz            32 flag variants x 2 episodes, branchy float64 actions
  ref_c1_rbc_n10.npz          BASELINE config 1: N=10 default env driven by the RBC rule
  ref_c2_n10_e256.npz         BASELINE config 2 (first 256 of the 4,096 envs): uniform actions
  ref_return_stats.json       random-policy episode-return statistics (for the sampler tests)
  sb3_ppo_4ch_policy.npz      the weights of the reference's shipped PPO checkpoint (policy.pth of
                              solvers/RL/models/PPO-b-pv-bounded-sparse-4ch-1h/999600.zip), SB3 key names
                              (`python tests/golden/generate_golden.py policy` regenerates only this one)
"""
import json
import os
import random
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_loader as rl  # noqa: E402
from smart_nanogrid_gym_b200.config import NanogridConfig  # noqa: E402
from smart_nanogrid_gym_b200.schedule import records_from_dense, concat_records  # noqa: E402
from test_oracle_vs_live_reference import branchy_actions  # noqa: E402

REC_FIELDS = ("arr", "dep", "cap", "soc0", "req", "n_veh")


def rec_dict(rec, prefix="sched_"):
    return {prefix + f: getattr(rec, f) for f in REC_FIELDS}


def copy_recorded_episodes():
    base = os.path.join(rl.REFERENCE_ROOT, "solvers", "RL")
    stem = "PPO-b-pv-bounded-sparse-4ch-1h-"
    for tag, d in (("g1", "training_files"), ("g2", "single_prediction_files")):
        for kind in ("initial_values", "prediction_results"):
            shutil.copyfile(os.path.join(base, d, stem + kind + ".json"),
                            os.path.join(HERE, "%s_%s.json" % (tag, kind)))


def gen_tables():
    out = {}
    for ti in ("1h", "2h"):
        for pm in range(5):
            env = rl.make_ref_env(number_of_chargers=4, time_interval=ti, price_model=pm)
            tabs = rl.constant_tables(env)
            for k, v in tabs.items():
                out["%s_pm%d_%s" % (ti, pm, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_tables.npz"), **out)


def gen_seeded_schedules():
    """Schedules straight from the reference generator, one env per (seed, flags)."""
    out = {}
    for dc in (0, 1):
        for rs in (0, 1):
            env = rl.make_ref_env(number_of_chargers=10, enable_different_vehicle_battery_capacities=bool(dc),
                                  enable_requested_state_of_charge=bool(rs))
            recs, shifts = [], []
            for seed in range(24):
                rl.seed_reference(seed)
                env.reset()
                s = rl.export_schedule(env)
                recs.append(records_from_dense(s["soc"], s["occ"], s["cap"], s["req"], s["arrivals"],
                                               s["departures"], 24))
                shifts.append(s["pv_shift"])
            rec = concat_records(recs)
            for f in REC_FIELDS:
                out["dc%d_rs%d_%s" % (dc, rs, f)] = getattr(rec, f)
            out["dc%d_rs%d_pv_shift" % (dc, rs)] = np.array(shifts)
    np.savez_compressed(os.path.join(HERE, "ref_schedules_seeded.npz"), **out)


def run_episode(env, cfg, actions_fn, want_diag=True):
    """reset + one episode; returns schedule record + per-step outputs of the reference."""
    obs0, _ = env.reset()
    s = rl.export_schedule(env)
    rec = records_from_dense(s["soc"], s["occ"], s["cap"], s["req"], s["arrivals"], s["departures"], cfg.n_steps)
    T = cfg.n_steps
    A, D, N = cfg.act_dim, cfg.obs_dim, cfg.n_spots
    acts = np.zeros((T, A))
    obs = np.zeros((T, D), np.float32)
    rew = np.zeros(T)
    done = np.zeros(T, np.uint8)
    power = np.zeros((T, N))
    diag = np.zeros((T, 6))
    o = obs0
    for t in range(T):
        a = actions_fn(t, o)
        acts[t] = a
        o, r, d, tr, info = env.step(np.array(a, dtype=np.float64))
        obs[t], rew[t], done[t] = o, r, d
        power[t] = env.charger_power_values_per_timestep[-1]
        diag[t] = (env.grid_power_per_timestep[-1], env.grid_energy_cost_per_timestep[-1],
                   env.total_vehicle_penalty_per_timestep[-1], env.total_battery_penalty_per_timestep[-1],
                   env.battery_per_timestep[-1], env.battery_power_value_per_timestep[-1])
    final_soc = env.central_management_system.charging_station.get_vehicles_state_of_charge()[:, :T + 1].copy()
    return dict(rec=rec, pv_shift=s["pv_shift"], soc_b0=s["soc_b"], obs0=obs0, actions=acts, obs=obs, reward=rew,
                done=done, power=power, diag=diag, final_soc=final_soc)


def stack_runs(runs):
    out = rec_dict(concat_records([r["rec"] for r in runs]))
    for k in ("pv_shift", "soc_b0", "obs0", "actions", "obs", "reward", "done", "power", "diag", "final_soc"):
        out[k] = np.stack([np.asarray(r[k]) for r in runs], axis=0)
    return out


def gen_variants():
    """32 flag combinations (pv, batt, v2x, diff_cap, req_soc), penalty mode cycling, N in {4,10}."""
    import itertools
    out = {}
    meta = []
    modes = ["sparse", "dense", "on_departure", "no_penalty"]
    for idx, (pv, b, v2x, dc, rs) in enumerate(itertools.product([True, False], repeat=5)):
        n = 10 if idx % 2 == 0 else 4
        kw = dict(number_of_chargers=n, pv_system_available_in_model=pv, battery_system_available_in_model=b,
                  vehicle_to_everything=v2x, enable_different_vehicle_battery_capacities=dc,
                  enable_requested_state_of_charge=rs, vehicle_uncharged_penalty_mode=modes[idx % 4])
        full = dict(rl.DEFAULT_KW)
        full.update(kw)
        cfg = NanogridConfig(**full)
        env = rl.make_ref_env(**kw)
        rng = np.random.default_rng(7000 + idx)
        rl.seed_reference(7000 + idx)
        lo, hi = cfg.action_bounds()
        runs = [run_episode(env, cfg, lambda t, o: branchy_actions(rng, lo, hi)) for _ in range(2)]
        for k, v in stack_runs(runs).items():
            out["v%02d_%s" % (idx, k)] = v
        meta.append(full)
    out["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "ref_variants.npz"), **out)


def rbc_rule(cfg, obs):
    """solvers/RBC/rbc.py:12-26 with generic offsets; battery action 0."""
    off = (8 if cfg.pv else 4) + cfg.n_spots
    a = np.zeros(cfg.act_dim)
    for i in range(cfg.n_spots):
        d = obs[off + i]
        if d == 0:
            a[i] = 0
        elif 0 < d < 0.16667:
            a[i] = 1
        else:
            a[i] = (float(obs[0]) + float(obs[2])) / 2
    return a


def gen_c1_rbc():
    kw = dict(number_of_chargers=10)
    full = dict(rl.DEFAULT_KW)
    full.update(kw)
    cfg = NanogridConfig(**full)
    env = rl.make_ref_env(**kw)
    rl.seed_reference(0)
    runs = [run_episode(env, cfg, lambda t, o: rbc_rule(cfg, o)) for _ in range(3)]
    np.savez_compressed(os.path.join(HERE, "ref_c1_rbc_n10.npz"), **stack_runs(runs))


def gen_c2(n_envs=256):
    """BASELINE config 2 inputs: env e is seeded with s = e; actions U(low, high)
    from default_rng(1234) with shape [24, 4096, 11] -- the first `n_envs` columns are kept."""
    kw = dict(number_of_chargers=10)
    full = dict(rl.DEFAULT_KW)
    full.update(kw)
    cfg = NanogridConfig(**full)
    lo, hi = cfg.action_bounds()
    all_actions = np.random.default_rng(1234).uniform(lo, hi, size=(24, 4096, 11))
    runs = []
    for e in range(n_envs):
        env = rl.make_ref_env(**kw)  # fresh env: battery starts at 0.5
        rl.seed_reference(e)
        runs.append(run_episode(env, cfg, lambda t, o: all_actions[t, e]))
    out = stack_runs(runs)
    out["actions"] = out["actions"].astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_c2_n10_e%d.npz" % n_envs), **out)


def gen_return_stats():
    """Random-policy episode returns of the live reference."""
    stats = {}
    for name, kw, episodes in (("n10_pv_batt", dict(number_of_chargers=10), 1500),
                               ("n4_pv_batt", dict(number_of_chargers=4), 1500)):
        full = dict(rl.DEFAULT_KW)
        full.update(kw)
        cfg = NanogridConfig(**full)
        env = rl.make_ref_env(**kw)
        rl.seed_reference(0)
        rng = np.random.default_rng(0)
        lo, hi = cfg.action_bounds()
        rets, nveh, occ_steps = [], [], []
        t0 = time.time()
        for _ in range(episodes):
            env.reset()
            cs = env.central_management_system.charging_station
            nveh.append(np.mean([len(a) for a in cs.arrivals]))
            occ_steps.append(cs.get_occupancy_for_all_chargers().sum(axis=1).mean())
            ret = 0.0
            for _t in range(24):
                o, r, d, tr, info = env.step(rng.uniform(lo, hi).astype(np.float32))
                ret += r
            rets.append(ret)
        dt = time.time() - t0
        rets = np.array(rets)
        stats[name] = dict(episodes=episodes, mean_return=float(rets.mean()), std_return=float(rets.std()),
                           mean_vehicles_per_spot=float(np.mean(nveh)),
                           mean_occupied_steps_per_spot=float(np.mean(occ_steps)),
                           ref_steps_per_sec_io_stubbed_one_core=episodes * 24 / dt)
    with open(os.path.join(HERE, "ref_return_stats.json"), "w") as fp:
        json.dump(stats, fp, indent=2)


def gen_multiday_tables():
    """PVSystemManager(days, dt) for days > 1 (the env hard-codes NUMBER_OF_DAYS_TO_PREDICT = 1,
    envs/smart_nanogrid_environment.py:51; the manager itself takes any number the 3-day irradiance file covers)."""
    rl.load_reference()
    from smart_nanogrid_gym.utils.pv_system_manager import PVSystemManager
    out = {}
    for ti, dt in (("1h", 1.0), ("2h", 2.0)):
        pvm = PVSystemManager(2, dt)
        k = "%s_d2_" % ti
        out[k + "irr_flat"] = np.array(pvm.solar_irradiance[0], dtype=np.float64)
        out[k + "irr_rows"] = np.array(pvm.solar_irradiance_2, dtype=np.float64)
        out[k + "irr_max"] = np.float64(pvm.max_radiation)
        out[k + "pv_power"] = np.array(pvm.available_solar_power[0], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_tables_multiday.npz"), **out)


SHIPPED_CHECKPOINT = "solvers/RL/models/PPO-b-pv-bounded-sparse-4ch-1h/999600.zip"


def gen_shipped_policy():
    """The reference's trained policy (C3 "for N=4, the shipped checkpoint"): a data file of the
    reference, stored as plain arrays because the zip cannot travel to the GPU box."""
    import io
    import zipfile
    import torch
    with zipfile.ZipFile(os.path.join(rl.REFERENCE_ROOT, SHIPPED_CHECKPOINT)) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    np.savez_compressed(os.path.join(HERE, "sb3_ppo_4ch_policy.npz"), **{k: v.numpy() for k, v in sd.items()})


if __name__ == "__main__":
    assert rl.reference_available(), "needs /root/reference"
    os.environ["PYTHONBREAKPOINT"] = "0"
    if sys.argv[1:] == ["policy"]:
        gen_shipped_policy()
        sys.exit(0)
    if sys.argv[1:] == ["multiday"]:
        gen_multiday_tables()
        sys.exit(0)
    random.seed(0)
    copy_recorded_episodes()
    gen_tables()
    gen_seeded_schedules()
    gen_variants()
    gen_c1_rbc()
    gen_c2()
    gen_return_stats()
    gen_shipped_policy()
    gen_multiday_tables()
    for f in sorted(os.listdir(HERE)):
        print("%9d  %s" % (os.path.getsize(os.path.join(HERE, f)), f))
