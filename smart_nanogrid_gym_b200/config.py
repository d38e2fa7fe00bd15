"""Host-side configuration of the batched nanogrid environment.

`NanogridConfig` takes the same keyword arguments as the reference constructor
(`SmartNanogridEnv.__init__`, envs/smart_nanogrid_environment.py:32-34) and derives the
dimensions, physical constants and the shared PV / price tables that the CUDA step kernel
reads.  Pure Python + numpy: no CUDA needed to build it.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Optional

import numpy as np

PENALTY_MODES = {"no_penalty": 0, "on_departure": 1, "sparse": 2, "dense": 3}

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def parse_time_interval(requested: str) -> float:
    """SmartNanogridEnv.set_time_interval, envs/smart_nanogrid_environment.py:125-138."""
    if requested:
        if "h" in requested:
            return float(requested.replace("h", ""))
        if "min" in requested:
            return float(requested.replace("min", "")) / 60.0
        raise ValueError("Wrong time interval was provided")
    return float(1)


def _price_day_hourly(price_model: int) -> np.ndarray:
    """Accountant.get_price_day, utils/accountant.py:58-101 (hourly tables, models 0-4)."""
    # Accountant.set_grid_tariffs, accountant.py:17-24
    high = 0.028 + 0.148933333 + 0.014
    low = 0.013333333 + 0.087613333 + 0.014
    if price_model == 0:
        return np.array([low] * 7 + [high] * 13 + [low] * 4)
    if price_model == 1:
        return np.array([0.05] * 7 + [0.1] * 13 + [0.05] * 4)
    if price_model == 2:
        return np.array([0.05, 0.05, 0.05, 0.05, 0.05, 0.06, 0.07, 0.08, 0.09, 0.1, 0.1, 0.1, 0.08, 0.06,
                         0.05, 0.05, 0.05, 0.06, 0.06, 0.06, 0.06, 0.05, 0.05, 0.05])
    if price_model == 3:
        return np.array([0.071, 0.060, 0.056, 0.056, 0.056, 0.060, 0.060, 0.060, 0.066, 0.066, 0.076, 0.080,
                         0.080, 0.1, 0.1, 0.076, 0.076, 0.1, 0.082, 0.080, 0.085, 0.079, 0.086, 0.070])
    if price_model == 4:
        return np.array([0.1, 0.1, 0.05, 0.05, 0.05, 0.05, 0.05, 0.08, 0.08, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1,
                         0.1, 0.1, 0.06, 0.06, 0.06, 0.1, 0.1, 0.1, 0.1])
    # price_model == 5 is broken in the reference (accountant.py:90-98)
    raise ValueError("price_model must be 0..4 (model 5 raises in the reference as well)")


def build_price_table(price_model: int, dt: float, n_steps: int, days: int = 1):
    """(days + 1)-day price table (the reference's two days for days = 1; the tariff repeats daily) + its max.  For dt >= 1 h this is the reference's hard-coded
    hourly table used verbatim (accountant.py:49-56,69-73: it is indexed by *step*, quirk Q9).
    For dt < 1 h (no runnable reference) model 0 follows the reference's own
    dt-parameterised (dead-code) rule `low if i < 7/dt or i > 19/dt else high`
    (accountant.py:61-66); models 1-4 hold each hourly value for 1/dt steps."""
    day = _price_day_hourly(price_model)
    if dt < 1.0:
        if price_model == 0:
            low, high = day[0], day[7]
            day = np.array([low if (i < 7 / dt or i > 19 / dt) else high for i in range(int(24 / dt))])
        else:
            day = np.repeat(day, int(round(1.0 / dt)))
    two_days = np.concatenate([day, day], axis=0)
    need = (days + 1) * n_steps
    if two_days.shape[0] < need:
        if two_days.shape[0] < 2 * n_steps:
            two_days = np.resize(two_days, 2 * n_steps)     # 2 h steps: the reference's 48-entry table, indexed by step
        else:
            two_days = np.concatenate([two_days] + [day] * (days - 1), axis=0)
    price_max = two_days.max(where=(two_days >= 0), initial=0)  # accountant.py:51
    return two_days.astype(np.float64), float(price_max)


def load_irradiance_1min() -> np.ndarray:
    """Minute-resolution irradiance (W/m^2, 3 days): the `irradiance` variable of the
    reference's files/solar_irradiance.mat (pv_system_manager.py:30-32), stored as .npy."""
    return np.load(os.path.join(_DATA_DIR, "solar_irradiance_1min.npy"))


def build_pv_tables(dt: float, n_steps: int, irradiance_1min: Optional[np.ndarray] = None, days: int = 1):
    """PVSystemManager(days, dt).__init__, utils/pv_system_manager.py:10-22: per-step irradiance means
    over a (days + 1)-day padded window (:11-15, 34-44), max (:20), PV power = irr*A*eta/1000*1.5/dt (:67-73,87-88).
    The flat series returned here is `solar_irradiance[0]` / `available_solar_power[0]`; row d of the reference's
    `solar_irradiance_2` (:46-65, "repeated middle" reshape) is its window [d * n_steps, (d + 2) * n_steps)."""
    irr_1min = load_irradiance_1min() if irradiance_1min is None else np.asarray(irradiance_1min, np.float64)
    step_min = int(60 * dt)
    padded = n_steps * (days + 1)
    if padded * step_min > irr_1min.shape[0]:
        raise ValueError("the irradiance file holds %d minutes: too short for %d day(s) + 1 of padding" % (irr_1min.shape[0], days))
    irr = np.zeros(padded)
    for k in range(padded):
        irr[k] = np.mean(irr_1min[k * step_min:(k + 1) * step_min])
    irr_max = irr.max(where=(irr >= 0), initial=0)
    scaling_pv = (2.279 * 1.134 * 20) * 0.21 / 1000  # pv_system_manager.py:17,72-73
    energy = irr * scaling_pv * 1.5                   # :67-70
    power = energy / dt                               # :87-88
    return irr, float(irr_max), power


@dataclasses.dataclass
class NanogridConfig:
    # reference constructor kwargs (same names, same defaults)
    price_model: int = 0
    number_of_chargers: int = 8
    pv_system_available_in_model: bool = True
    battery_system_available_in_model: bool = True
    vehicle_to_everything: bool = False
    enable_different_vehicle_battery_capacities: bool = True
    enable_requested_state_of_charge: bool = False
    algorithm_used: str = ""
    environment_mode: str = ""
    time_interval: str = ""
    charging_mode: str = ""
    vehicle_uncharged_penalty_mode: str = ""

    # physical constants, hard-coded in the reference
    ev_max_power: float = 22.0          # charger.py:20-23
    ev_efficiency: float = 0.95
    default_vehicle_capacity: int = 40  # charging_station.py:224
    bess_capacity: float = 80.0         # central_management_system.py:35
    bess_initial_soc: float = 0.5
    bess_max_power: float = 44.0
    bess_efficiency: float = 0.95
    bess_depth_of_discharge: float = 0.15
    sell_coefficient: float = 0.8       # accountant.py:6
    cost_weight: float = 0.75           # accountant.py:35
    battery_penalty_weight: float = 0.8 # penaliser.py:181
    soc_margin_ratio: float = 0.05      # penaliser.py:7
    hours_ahead: int = 3                # ...environment.py:52
    number_of_days_to_predict: int = 1  # ...environment.py:51 (NUMBER_OF_DAYS_TO_PREDICT -> PVSystemManager(days, dt))
    cycle_pv_days: bool = False         # extension: episode k sees day k % days of the irradiance series (the reference
                                        # always reads row 0 of solar_irradiance_2, pv_system_manager.py:81-91)
    departure_normaliser: float = 24.0  # ...environment.py:208 (a literal, not 24/dt)

    def __post_init__(self):
        self.dt = parse_time_interval(self.time_interval)
        steps = 24.0 / self.dt
        if steps != int(steps):
            raise ValueError("24 h must be a whole number of time intervals")
        self.n_steps = int(steps)
        self.n_spots = int(self.number_of_chargers)
        if not (1 <= self.n_spots <= 255):
            raise ValueError("number_of_chargers must be in 1..255")
        if self.n_steps + int(4 / self.dt) > 250:
            raise ValueError("time interval too small: departures must fit a byte")
        self.pv = bool(self.pv_system_available_in_model)
        self.batt = bool(self.battery_system_available_in_model)
        self.v2x = bool(self.vehicle_to_everything)
        self.act_dim = self.n_spots + int(self.batt)                      # ...environment.py:101-118
        self.obs_dim = (1 + int(self.pv)) * (1 + self.hours_ahead) + 2 * self.n_spots + int(self.batt)  # :90-96
        self.pv_days = int(self.number_of_days_to_predict)
        if self.pv_days < 1:
            raise ValueError("number_of_days_to_predict must be >= 1")
        self.price, self.price_max = build_price_table(self.price_model, self.dt, self.n_steps, self.pv_days)
        self.price_norm = self.price / self.price_max                      # accountant.py:41-46
        if self.pv:
            self.irr, self.irr_max, self.pv_power = build_pv_tables(self.dt, self.n_steps, days=self.pv_days)
            self.irr_norm = self.irr / self.irr_max                        # pv_system_manager.py:81-85
        else:
            n = self.price.shape[0]
            self.irr = np.zeros(n)
            self.irr_max = 1.0
            self.pv_power = np.zeros(n)
            self.irr_norm = np.zeros(n)
        n = max(self.price.shape[0], self.pv_power.shape[0])
        self.table_len = n

    # the reference only fails on these when they are first used
    def validate_modes(self):
        if self.vehicle_uncharged_penalty_mode not in PENALTY_MODES:
            raise ValueError("Error: Wrong vehicle uncharged - penalty mode provided!")  # charging_station.py:60
        if self.charging_mode != "bounded":
            raise ValueError("Error: Wrong charging mode provided!")                     # charger.py:88

    @property
    def penalty_mode_id(self) -> int:
        # invalid modes are only rejected at reset(), like the reference (validate_modes)
        return PENALTY_MODES.get(self.vehicle_uncharged_penalty_mode, 0)

    def action_bounds(self):
        """Box bounds of ...environment.py:101-118."""
        low = np.full(self.act_dim, -1.0 if self.v2x else 0.0, dtype=np.float32)
        if self.batt:
            low[-1] = -1.0
        high = np.ones(self.act_dim, dtype=np.float32)
        return low, high

    def reference_kwargs(self):
        names = ("price_model", "number_of_chargers", "pv_system_available_in_model",
                 "battery_system_available_in_model", "vehicle_to_everything",
                 "enable_different_vehicle_battery_capacities", "enable_requested_state_of_charge",
                 "algorithm_used", "environment_mode", "time_interval", "charging_mode",
                 "vehicle_uncharged_penalty_mode")
        return {k: getattr(self, k) for k in names}
