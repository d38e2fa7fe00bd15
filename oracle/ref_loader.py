"""TEST INFRASTRUCTURE ONLY -- imports the live reference (`/root/reference`) under shims.

This module is only usable in the build container (the reference tree does not exist
on the GPU box).  It is used by `tests/golden/generate_golden.py` to produce the
committed golden fixtures and by the optional `tests/test_oracle_vs_live_reference.py`
(skipped when the tree is absent).  Nothing in the product path imports it.

Shims:
  1. a stub `gym` package (gym / gymnasium / SB3 are not installed):
     reference imports at smart_nanogrid_gym/__init__.py:1 and
     envs/smart_nanogrid_environment.py:5-7.
  2. `smart_nanogrid_gym.utils.config` replaced by writable POSIX paths
     (utils/config.py:4-5 builds Windows paths; reset() writes initial_values.json,
     utils/charging_station.py:185).
  3. battery variants: the orchestrator calls the penaliser with 8 kwargs
     (utils/central_management_system.py:176-179) that only `PenaliserOld`
     accepts (utils/penaliser_old.py:98-104).
  4. PYTHONBREAKPOINT=0 for the stray breakpoint() (central_management_system.py:165).
"""
import os
import random
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SNG_REFERENCE_ROOT", "/root/reference")
_loaded = {}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "smart_nanogrid_gym"))


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        shp = tuple(shape) if shape else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shp).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shp).copy()
        self.shape, self.dtype = self.low.shape, dtype


def _install_gym_stub():
    names = ("gym", "gym.spaces", "gym.utils", "gym.utils.seeding", "gym.envs", "gym.envs.registration")
    mods = {n: types.ModuleType(n) for n in names}
    gym = mods["gym"]

    class Env:  # noqa: D401 - minimal base class
        pass

    gym.Env = Env
    mods["gym.spaces"].Box = _Box
    gym.spaces = mods["gym.spaces"]
    gym.utils = mods["gym.utils"]
    mods["gym.utils"].seeding = mods["gym.utils.seeding"]
    reg = mods["gym.envs.registration"]
    reg.registry = {}
    reg.register = lambda **k: reg.registry.update({k["id"]: k})
    reg.make = reg.spec = None
    gym.envs = mods["gym.envs"]
    mods["gym.envs"].registration = reg
    for n, m in mods.items():
        sys.modules.setdefault(n, m)
    return reg


def load_reference(stub_io=True):
    """Import the reference package; return the module namespace we need."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    os.environ["PYTHONBREAKPOINT"] = "0"
    reg = _install_gym_stub()
    work = tempfile.mkdtemp(prefix="sng_ref_")
    shutil.copytree(os.path.join(REFERENCE_ROOT, "smart_nanogrid_gym", "files"), work + "/files")
    for d in ("training_files", "evaluation_files", "single_prediction_files", ""):
        os.makedirs("%s/solvers/RL/%s" % (work, d), exist_ok=True)
    cfg = types.ModuleType("smart_nanogrid_gym.utils.config")
    cfg.data_files_directory_path = work + "/files/"
    cfg.solvers_files_directory_path = work + "/solvers/"
    sys.modules["smart_nanogrid_gym.utils.config"] = cfg
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import smart_nanogrid_gym  # noqa: F401  (registers the env id)
    import smart_nanogrid_gym.utils.central_management_system as cms
    from smart_nanogrid_gym.utils.penaliser_old import PenaliserOld
    cms.Penaliser = PenaliserOld
    import smart_nanogrid_gym.envs.smart_nanogrid_environment as envmod
    import smart_nanogrid_gym.utils.charging_station as csmod
    if stub_io:
        # json dumps only (envs/smart_nanogrid_environment.py:239-309, charging_station.py:185-186)
        envmod.SmartNanogridEnv._SmartNanogridEnv__save_prediction_results = lambda self: None

        class _NullJson:
            @staticmethod
            def dump(*a, **k):
                return None
            load = staticmethod(__import__("json").load)
        csmod.json = _NullJson
    _loaded.update(dict(env_cls=envmod.SmartNanogridEnv, envmod=envmod, csmod=csmod, cms=cms,
                        registry=reg.registry, workdir=work))
    return _loaded


DEFAULT_KW = dict(charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")


def make_ref_env(**kw):
    """SmartNanogridEnv(**kw) of the live reference (envs/smart_nanogrid_environment.py:31-34)."""
    ref = load_reference()
    args = dict(DEFAULT_KW)
    args.update(kw)
    return ref["env_cls"](**args)


def seed_reference(seed):
    """The reference has no seeding (envs/...environment.py:362-365); it draws from the
    global numpy legacy generator and Python's `random`."""
    np.random.seed(seed)
    random.seed(seed)


def export_schedule(env):
    """Dense schedule arrays + ragged lists after env.reset()."""
    cs = env.central_management_system.charging_station
    bess = env.central_management_system.battery_system
    return dict(
        soc=cs.get_vehicles_state_of_charge().copy(),
        occ=cs.get_occupancy_for_all_chargers().copy(),
        cap=cs.get_vehicle_capacities_for_all_chargers().copy(),
        req=cs.get_requested_end_state_of_charge_for_all_chargers().copy(),
        arrivals=[list(map(int, a)) for a in cs.arrivals],
        departures=[list(map(int, d)) for d in cs.departures],
        pv_shift=float(env.random_pv_shift_ratio),
        soc_b=float(bess.current_state_of_charge) if bess else 0.0,
    )


def inject_schedule(env, sched):
    """Load a schedule into a live reference env WITHOUT `load_initial_values`
    (which drops Requested_SOC, quirk Q7; charging_station.py:119-136)."""
    cs = env.central_management_system.charging_station
    env.timestep = 0
    cs.arrivals = [list(a) for a in sched["arrivals"]]
    cs.departures = [list(d) for d in sched["departures"]]
    for i, ch in enumerate(cs.chargers):
        ch.vehicle_arrivals = list(sched["arrivals"][i])
        ch.vehicle_state_of_charge = np.array(sched["soc"][i], dtype=np.float64)
        ch.occupancy = np.array(sched["occ"][i], dtype=np.float64)
        ch.vehicle_capacities = np.array(sched["cap"][i], dtype=np.float64)
        ch.requested_end_state_of_charge = np.array(sched["req"][i], dtype=np.float64)
    env.random_pv_shift_ratio = float(sched["pv_shift"])
    bess = env.central_management_system.battery_system
    if bess is not None:
        bess.current_state_of_charge = float(sched["soc_b"])
    # prime the penalty-check set exactly as reset() does (…environment.py:351)
    return env._SmartNanogridEnv__get_observations()


def constant_tables(env):
    """PV / price tables the reference builds at construction."""
    cmsys = env.central_management_system
    out = {}
    acc = cmsys.accountant
    out["price"] = np.array(acc.energy_price[0], dtype=np.float64)
    out["price_max"] = float(acc.energy_price_max)
    pvm = cmsys.pv_system_manager
    if pvm is not None:
        out["pv_power"] = np.array(pvm.available_solar_power[0], dtype=np.float64)
        out["irr"] = np.array(pvm.solar_irradiance_2[0], dtype=np.float64)
        out["irr_max"] = float(pvm.max_radiation)
    return out
