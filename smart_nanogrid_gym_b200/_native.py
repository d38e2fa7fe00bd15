"""ctypes binding of libsng.so (the C ABI declared in include/sng.h).

There is no CPU fallback: if the library is missing or no CUDA device is present, creating an
environment raises.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNG_LIB_PATH") or os.path.join(_PKG, "libsng.so")   # override: kernel experiments only
CSRC = os.path.join(_PKG, "csrc")
HEADER = os.path.join(os.path.dirname(_PKG), "include", "sng.h")

SNG_F32, SNG_F64 = 32, 64
MAX_VEHICLES = 8
FLAG_NEG_DEMAND, FLAG_BATT_SOC_GT1, FLAG_NAN_ACTION = 1, 2, 4
DIAG = ("total_ch", "total_dis", "solar", "batt_power", "grid_power", "grid_cost", "pen_veh", "pen_batt")

EXPORTS = ("sng_abi_version", "sng_sizeof", "sng_last_error", "sng_query_layout", "sng_create", "sng_destroy", "sng_bind",
           "sng_reset", "sng_load_schedule", "sng_step", "sng_rollout", "sng_step_host", "sng_sample_plan", "sng_sample_actions",
           "sng_error_flags", "sng_launch_count", "sng_set_tuning", "sng_set_pipeline", "sng_gae", "sng_policy_forward", "sng_null_launch",
           "sng_policy_packed_bytes", "sng_policy_pack", "sng_policy_forward_packed", "sng_policy_forward_sampled", "sng_debug_arrival_gap", "sng_set_launch_mode", "sng_policy_set_launch_mode", "sng_policy_step", "sng_debug_traffic_skeleton", "sng_debug_stamp")


class SngConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("precision", C.c_int32), ("n_envs", C.c_int64), ("env_gid0", C.c_int64),
                ("n_spots", C.c_int32), ("n_steps", C.c_int32), ("horizon", C.c_int32), ("table_len", C.c_int32),
                ("pv", C.c_int32), ("batt", C.c_int32), ("v2x", C.c_int32), ("penalty_mode", C.c_int32),
                ("diff_cap", C.c_int32), ("req_soc", C.c_int32), ("default_cap", C.c_int32), ("auto_reset", C.c_int32),
                ("pv_days", C.c_int32), ("_reserved", C.c_int32),
                ("dt", C.c_double), ("ev_pmax", C.c_double), ("ev_eff", C.c_double),
                ("b_cap", C.c_double), ("b_pmax", C.c_double), ("b_eff", C.c_double), ("b_dod", C.c_double),
                ("b_soc0", C.c_double),
                ("sell_coeff", C.c_double), ("cost_weight", C.c_double), ("batt_pen_w", C.c_double),
                ("margin", C.c_double), ("dep_norm", C.c_double),
                ("pv_power", C.c_void_p), ("irr_norm", C.c_void_p), ("price", C.c_void_p), ("price_norm", C.c_void_p)]


class SngLayout(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("act_dim", C.c_int32), ("obs_dim", C.c_int32), ("real_bytes", C.c_int32),
                ("plan_rec_bytes", C.c_int32), ("envst_bytes", C.c_int32), ("plan_slots", C.c_int32),
                ("diag_count", C.c_int32), ("env_block", C.c_int32), ("spot_planes", C.c_int32)]


class SngBuffers(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("_pad", C.c_uint32)] + [(n, C.c_void_p) for n in (
        "actions", "obs", "reward", "done", "terminal_obs", "spot", "envst", "plan", "err", "diag",
        "last_return", "spot_power")]


class SngScheduleView(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_slots", C.c_int32)] + [(n, C.c_void_p) for n in (
        "arr", "dep", "cap", "soc0", "req", "n_veh", "pv_shift", "soc_b")]


class SngMlp(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("obs_dim", C.c_int32), ("hidden", C.c_int32), ("act_dim", C.c_int32)] + [
        (n, C.c_void_p) for n in ("w_pi0", "b_pi0", "w_pi1", "b_pi1", "w_act", "b_act", "log_std",
                                  "w_vf0", "b_vf0", "w_vf1", "b_vf1", "w_val", "b_val")]


class NativeError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libsng.so in-tree (nvcc, -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", "Makefile"))] + [HEADER]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale or force:
        cmd = ["make", "-j4", "-C", CSRC] + (["-B"] if force else [])
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout)
        if res.returncode != 0:
            raise NativeError("building libsng.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load libsng.so; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.sng_abi_version.restype = C.c_int
        L.sng_last_error.restype = C.c_char_p
        L.sng_query_layout.argtypes = [C.POINTER(SngConfig), C.POINTER(SngLayout)]
        L.sng_create.argtypes = [C.POINTER(SngConfig), C.c_int, C.POINTER(C.c_void_p)]
        L.sng_destroy.argtypes = [C.c_void_p]
        L.sng_destroy.restype = None
        L.sng_bind.argtypes = [C.c_void_p, C.POINTER(SngBuffers)]
        L.sng_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
        L.sng_load_schedule.argtypes = [C.c_void_p, C.POINTER(SngScheduleView), C.c_void_p]
        L.sng_step.argtypes = [C.c_void_p, C.c_void_p]
        L.sng_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sng_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sng_sample_plan.argtypes = [C.c_void_p, C.c_void_p]
        L.sng_sample_actions.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.sng_error_flags.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p]
        L.sng_launch_count.argtypes = [C.c_void_p]
        L.sng_launch_count.restype = C.c_int64
        L.sng_set_tuning.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.sng_set_pipeline.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.sng_set_launch_mode.argtypes = [C.c_void_p, C.c_int]
        L.sng_policy_set_launch_mode.argtypes = [C.c_int]
        L.sng_null_launch.argtypes = [C.c_void_p]
        L.sng_debug_arrival_gap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.sng_gae.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_int64, C.c_float, C.c_float, C.c_void_p]
        L.sng_policy_forward.argtypes = [C.POINTER(SngMlp)] + [C.c_void_p] * 8 + [C.c_int64, C.c_void_p]
        L.sng_policy_packed_bytes.argtypes = []
        L.sng_policy_packed_bytes.restype = C.c_size_t
        L.sng_policy_pack.argtypes = [C.POINTER(SngMlp), C.c_void_p, C.c_void_p]
        L.sng_policy_forward_packed.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int64, C.c_void_p]
        L.sng_policy_forward_sampled.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                                  C.c_uint64] + [C.c_void_p] * 7 + [C.c_int64, C.c_void_p])
        L.sng_policy_step.argtypes = ([C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_uint64] + [C.c_void_p] * 11)
        L.sng_debug_traffic_skeleton.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sng_debug_stamp.argtypes = [C.c_void_p, C.c_void_p]
        if L.sng_abi_version() != 3:
            raise NativeError("libsng.so ABI version mismatch")
        for which, st in enumerate((SngConfig, SngLayout, SngBuffers, SngScheduleView)):
            if L.sng_sizeof(which) != C.sizeof(st):
                raise NativeError("ctypes mirror of struct %d is out of sync with include/sng.h" % which)
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise NativeError("libsng error %d: %s" % (rc, lib().sng_last_error().decode()))


def make_config(cfg, n_envs: int, env_gid0: int = 0, precision: int = SNG_F32, auto_reset: bool = True):
    """NanogridConfig -> (SngConfig, keep-alive list of the table arrays)."""
    import numpy as np
    tabs = [np.ascontiguousarray(x, dtype=np.float64) for x in (cfg.pv_power, cfg.irr_norm, cfg.price, cfg.price_norm)]
    n = min(t.shape[0] for t in tabs)
    c = SngConfig()
    c.struct_size = C.sizeof(SngConfig)
    c.precision = precision
    c.n_envs, c.env_gid0 = int(n_envs), int(env_gid0)
    c.n_spots, c.n_steps, c.horizon, c.table_len = cfg.n_spots, cfg.n_steps, cfg.hours_ahead, n
    c.pv, c.batt, c.v2x = int(cfg.pv), int(cfg.batt), int(cfg.v2x)
    c.penalty_mode = cfg.penalty_mode_id
    c.diff_cap = int(cfg.enable_different_vehicle_battery_capacities)
    c.req_soc = int(cfg.enable_requested_state_of_charge)
    c.default_cap = int(cfg.default_vehicle_capacity)
    c.auto_reset = int(auto_reset)
    c.pv_days = int(cfg.pv_days) if (cfg.pv and getattr(cfg, "cycle_pv_days", False)) else 1
    c.dt = cfg.dt
    c.ev_pmax, c.ev_eff = cfg.ev_max_power, cfg.ev_efficiency
    c.b_cap, c.b_pmax, c.b_eff, c.b_dod, c.b_soc0 = (cfg.bess_capacity, cfg.bess_max_power, cfg.bess_efficiency,
                                                     cfg.bess_depth_of_discharge, cfg.bess_initial_soc)
    c.sell_coeff, c.cost_weight, c.batt_pen_w = cfg.sell_coefficient, cfg.cost_weight, cfg.battery_penalty_weight
    c.margin, c.dep_norm = cfg.soc_margin_ratio, cfg.departure_normaliser
    c.pv_power, c.irr_norm, c.price, c.price_norm = [t.ctypes.data for t in tabs]
    return c, tabs


def query_layout(c: SngConfig) -> SngLayout:
    out = SngLayout()
    check(lib().sng_query_layout(C.byref(c), C.byref(out)))
    return out
