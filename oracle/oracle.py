"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the float64 CPU oracle
(oracle/nanogrid_oracle.c).  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only; the product package never imports it.

Parity status: PINNED (see nanogrid_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnanogrid_oracle.so")
MAX_VEHICLES = 8
DIAG = ("total_ch", "total_dis", "solar", "batt_power", "batt_soc", "grid_power", "grid_energy",
        "grid_cost", "pen_veh", "pen_batt", "pen_total", "total_cost")
PENALTY_MODES = {"no_penalty": 0, "on_departure": 1, "sparse": 2, "dense": 3}
ERR_NEG_DEMAND, ERR_BATT_SOC_GT1, ERR_NAN = 1, 2, 4


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("nanogrid_oracle.c", "nanogrid_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if stale or force:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libnanogrid_oracle.so"])
    return _LIB_PATH


class _Cfg(C.Structure):
    _fields_ = [("n_spots", C.c_int32), ("n_steps", C.c_int32), ("dt", C.c_double),
                ("pv", C.c_int32), ("batt", C.c_int32), ("v2x", C.c_int32), ("penalty_mode", C.c_int32),
                ("horizon", C.c_int32), ("diff_cap", C.c_int32), ("req_soc", C.c_int32), ("pv_days", C.c_int32),
                ("ev_pmax", C.c_double), ("ev_eff", C.c_double), ("b_cap", C.c_double), ("b_pmax", C.c_double),
                ("b_eff", C.c_double), ("b_dod", C.c_double), ("sell_coeff", C.c_double),
                ("cost_weight", C.c_double), ("batt_pen_w", C.c_double), ("margin", C.c_double),
                ("dep_norm", C.c_double),
                ("pv_power", C.c_void_p), ("irr_norm", C.c_void_p), ("price", C.c_void_p),
                ("price_norm", C.c_void_p)]


class _State(C.Structure):
    _fields_ = [("n_envs", C.c_int64)] + [(n, C.c_void_p) for n in (
        "occ", "soc", "cap", "req", "arr", "dep", "n_veh", "check", "t", "pv_shift", "soc_b", "err", "pv_base")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()          # no-op unless the sources are newer than the library
        L = C.CDLL(_LIB_PATH)
        L.ngo_obs_dim.restype = C.c_int
        L.ngo_act_dim.restype = C.c_int
        L.ngo_numpy_sum.restype = C.c_double
        L.ngo_numpy_sum.argtypes = [C.c_void_p, C.c_int]
        L.ngo_observe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ngo_step.argtypes = [C.c_void_p] * 8 + [C.c_int]
        L.ngo_rbc_actions.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.ngo_philox4x32_10.argtypes = [C.c_void_p] * 3
        L.ngo_sample_episode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                               C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """E independent reference environments in float64, with the reference's dense arrays.

    `cfg` is any object with the attributes of `smart_nanogrid_gym_b200.config.NanogridConfig`
    (n_spots, n_steps, dt, pv, batt, v2x, penalty_mode_id, tables, constants)."""

    def __init__(self, cfg, n_envs, n_threads=1):
        self.cfg = cfg
        self.E, self.N, self.T = int(n_envs), cfg.n_spots, cfg.n_steps
        self.W = self.T + 1
        self.n_threads = int(n_threads)
        self._tabs = [np.ascontiguousarray(x, dtype=np.float64) for x in
                      (cfg.pv_power, cfg.irr_norm, cfg.price, cfg.price_norm)]
        c = _Cfg()
        c.n_spots, c.n_steps, c.dt = self.N, self.T, cfg.dt
        c.pv, c.batt, c.v2x = int(cfg.pv), int(cfg.batt), int(cfg.v2x)
        c.penalty_mode = cfg.penalty_mode_id
        c.horizon = cfg.hours_ahead
        c.diff_cap = int(cfg.enable_different_vehicle_battery_capacities)
        c.req_soc = int(cfg.enable_requested_state_of_charge)
        c.pv_days = int(cfg.pv_days) if (cfg.pv and getattr(cfg, "cycle_pv_days", False)) else 1
        c.ev_pmax, c.ev_eff = cfg.ev_max_power, cfg.ev_efficiency
        c.b_cap, c.b_pmax, c.b_eff, c.b_dod = (cfg.bess_capacity, cfg.bess_max_power, cfg.bess_efficiency,
                                               cfg.bess_depth_of_discharge)
        c.sell_coeff, c.cost_weight = cfg.sell_coefficient, cfg.cost_weight
        c.batt_pen_w, c.margin, c.dep_norm = (cfg.battery_penalty_weight, cfg.soc_margin_ratio,
                                              cfg.departure_normaliser)
        c.pv_power, c.irr_norm, c.price, c.price_norm = [t.ctypes.data for t in self._tabs]
        self._c = c
        E, N, W = self.E, self.N, self.W
        self.occ = np.zeros((E, N, W))
        self.soc = np.zeros((E, N, W))
        self.cap = np.zeros((E, N, W))
        self.req = np.zeros((E, N, W))
        self.arr = np.zeros((E, N, MAX_VEHICLES), np.int32)
        self.dep = np.zeros((E, N, MAX_VEHICLES), np.int32)
        self.n_veh = np.zeros((E, N), np.int32)
        self.check = np.zeros((E, N), np.uint8)
        self.t = np.zeros(E, np.int32)
        self.pv_shift = np.ones(E)
        self.soc_b = np.full(E, cfg.bess_initial_soc if cfg.batt else 0.0)
        self.err = np.zeros(E, np.uint32)
        self.pv_base = np.zeros(E, np.int32)     # offset of the episode's day in the PV tables (multi-day PV)
        s = _State()
        s.n_envs = E
        for n in ("occ", "soc", "cap", "req", "arr", "dep", "n_veh", "check", "t", "pv_shift", "soc_b", "err", "pv_base"):
            setattr(s, n, getattr(self, n).ctypes.data)
        self._s = s
        self.obs_dim = lib().ngo_obs_dim(C.byref(c))
        self.act_dim = lib().ngo_act_dim(C.byref(c))
        assert self.obs_dim == cfg.obs_dim and self.act_dim == cfg.act_dim

    # ---- schedule ingestion -------------------------------------------------------------
    def load_dense(self, e, soc, occ, cap, req, arrivals, departures, pv_shift=None, soc_b=None):
        """Schedule of one env from the reference's dense [N, T+1] arrays and ragged lists."""
        # the reference always allocates 25 slots (charger.py:16-19); slots > n_steps stay zero
        W = self.W
        for src in (soc, occ, cap, req):
            assert not np.asarray(src)[:, W:].any()
        self.soc[e], self.occ[e], self.cap[e], self.req[e] = (np.asarray(x)[:, :W] for x in (soc, occ, cap, req))
        for i in range(self.N):
            n = len(arrivals[i])
            assert n == len(departures[i]) and n <= MAX_VEHICLES
            self.n_veh[e, i] = n
            self.arr[e, i, :n] = arrivals[i]
            self.dep[e, i, :n] = departures[i]
        if pv_shift is not None:
            self.pv_shift[e] = pv_shift
        if soc_b is not None:
            self.soc_b[e] = soc_b
        self.t[e] = 0

    def load_records(self, arr, dep, cap, soc0, req, n_veh, pv_shift=None, soc_b=None):
        """Schedules of all envs from compact per-vehicle records [E, N, V] (+ counts [E, N])."""
        E, N, W, T = self.E, self.N, self.W, self.T
        self.occ[:], self.soc[:], self.cap[:], self.req[:] = 0, 0, 0, 0
        V = arr.shape[2]
        assert V <= MAX_VEHICLES
        self.arr[:, :, :V], self.dep[:, :, :V] = arr, dep
        self.n_veh[:] = n_veh
        tt = np.arange(W)[None, None, :]
        for v in range(V):
            valid = (v < n_veh)[:, :, None]
            a, d = arr[:, :, v, None], dep[:, :, v, None]
            present = valid & (tt >= a) & (tt < d) & (tt < T)
            self.occ[present] = 1.0
            self.cap[:] = np.where(present, cap[:, :, v, None].astype(np.float64), self.cap)
            self.req[:] = np.where(present, req[:, :, v, None].astype(np.float64), self.req)
            at = valid & (tt == a)
            self.soc[:] = np.where(at, soc0[:, :, v, None].astype(np.float64), self.soc)
        if pv_shift is not None:
            self.pv_shift[:] = pv_shift
        if soc_b is not None:
            self.soc_b[:] = soc_b
        self.t[:] = 0

    def sample(self, seed, env_gid0, episode, mask=None):
        """Counter-based (Philox) episode sampling, mirror of the CUDA sampler."""
        ep = np.ascontiguousarray(np.broadcast_to(np.asarray(episode, np.uint32), (self.E,)))
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().ngo_sample_episode_batch(C.byref(self._c), C.byref(self._s), int(seed), int(env_gid0), _p(ep), _p(m),
                                       self.n_threads)

    # ---- stepping -----------------------------------------------------------------------
    def observe(self):
        obs = np.empty((self.E, self.obs_dim), np.float32)
        lib().ngo_observe(C.byref(self._c), C.byref(self._s), _p(obs), self.n_threads)
        return obs

    def step(self, actions, want_diag=False):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.E, self.act_dim)
        obs = np.empty((self.E, self.obs_dim), np.float32)
        rew = np.empty(self.E)
        done = np.empty(self.E, np.uint8)
        power = np.empty((self.E, self.N)) if want_diag else None
        diag = np.empty((self.E, len(DIAG))) if want_diag else None
        lib().ngo_step(C.byref(self._c), C.byref(self._s), _p(a), _p(obs), _p(rew), _p(done), _p(power), _p(diag),
                       self.n_threads)
        if want_diag:
            return obs, rew, done, power, {k: diag[:, j] for j, k in enumerate(DIAG)}
        return obs, rew, done

    def step_noalloc(self, a, obs, rew, done):
        lib().ngo_step(C.byref(self._c), C.byref(self._s), _p(a), _p(obs), _p(rew), _p(done), None, None,
                       self.n_threads)

    def rbc_actions(self, obs):
        o = np.ascontiguousarray(obs, np.float32)
        a = np.empty((self.E, self.act_dim))
        lib().ngo_rbc_actions(C.byref(self._c), self.E, _p(o), _p(a))
        return a

    def current_soc(self):
        """soc[:, :, t] at the current t (what the next observe() would report)."""
        return np.take_along_axis(self.soc, self.t[:, None, None].astype(np.int64), axis=2)[:, :, 0]


def numpy_sum(a):
    a = np.ascontiguousarray(a, np.float64)
    return lib().ngo_numpy_sum(_p(a), a.shape[0])


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.empty(4, np.uint32)
    lib().ngo_philox4x32_10(_p(c), _p(k), _p(out))
    return out
