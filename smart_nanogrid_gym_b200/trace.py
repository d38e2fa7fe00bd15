"""Episode traces in the reference's `prediction_results.json` layout.

The reference appends 26 diagnostics per step (envs/smart_nanogrid_environment.py:143-171) and dumps
them at every episode end (:239-309) for its plotting notebooks.  The CUDA step keeps only what enters
the reward plus an optional 8-entry diagnostics row and the per-spot powers; `EpisodeRecorder` rebuilds the
reference's series for ONE chosen env from those, the observation and the decoded spot state, and writes
the same JSON keys.  It is a debugging / notebook aid for single envs, not part of the hot path.
Series that the reference computes but never feeds into the reward and that are identically zero in
its current code (needless-charging, overcharging, low-utilisation penalties: penaliser_old.py
:34,53-56,100-104 are commented out) are written as zeros.
"""
from __future__ import annotations

import json

import numpy as np
import torch

from . import _native as nat


class EpisodeRecorder:
    """Wraps `env.step` of a BatchedSmartNanogridEnv built with want_diagnostics=True."""

    def __init__(self, env, index: int = 0):
        if env.diag is None:
            raise ValueError("EpisodeRecorder needs an env created with want_diagnostics=True")
        self.env, self.i = env, int(index)
        self.cfg = env.cfg
        self.reset_series()

    def reset_series(self):
        N, T = self.cfg.n_spots, self.cfg.n_steps
        self.soc = np.zeros((N, max(T + 1, 25)))     # the reference allocates 25 slots (charger.py:16-19)
        self.t = 0
        self.initial_battery_soc = float(self.env.env_state()["soc_b"][self.i]) if self.cfg.batt else 0.0
        self.series = {k: [] for k in (
            "Grid_power", "Grid_energy", "Utilized_solar_energy", "Total_vehicle_penalties", "Total_battery_penalties",
            "Total_penalties", "Total_cost", "Battery_state_of_charge", "Grid_energy_cost", "Battery_action",
            "Charger_actions", "Total_charging_power", "Total_discharging_power", "Charger_power_values",
            "Battery_power_value", "Battery_SOC_below_DoD_penalties", "Insufficiently_charged_vehicle_penalties",
            "Battery_calculated_power_value", "DisCharging_nonexistent_vehicles_penalties")}

    def _nonexistent_penalty(self, a, st):
        """100 per non-zero action sent to an empty spot (Charger.reset_info_values, utils/charger.py:146-156)."""
        t, i = self.t, self.i
        arr, dep = st["arr"][i], st["dep"][i]
        present = (arr != 255) & (arr <= t) & (t < dep)
        return 100.0 * float(np.count_nonzero(~present & (a != 0)))

    def step(self, actions: torch.Tensor):
        """env.step(actions) + one row of every series for env `index`."""
        env, cfg, i = self.env, self.cfg, self.i
        st = env.spot_state()
        a = actions[i].detach().double().cpu().numpy()
        nonexistent = self._nonexistent_penalty(a[:cfg.n_spots], st)
        out = env.step(actions)
        P = env.spot_power[i].double().cpu().numpy()      # written by the step kernel itself
        obs, rew = out[0][i].cpu().numpy(), float(out[1][i])
        d = env.diag[i].double().cpu().numpy()
        D = {name: d[k] for k, name in enumerate(nat.DIAG)}
        nd = (1 + int(cfg.pv)) * (1 + cfg.hours_ahead)
        # after an auto-reset `obs` already shows the next day: the finished step's SoC is in terminal_obs
        row = env.terminal_obs[i].cpu().numpy() if (bool(out[2][i]) and env.auto_reset and env.terminal_obs is not None) else obs
        if self.t < self.soc.shape[1]:
            self.soc[:, self.t] = row[nd:nd + cfg.n_spots]
        S = self.series
        S["Grid_power"].append(float(D["grid_power"]))
        S["Grid_energy"].append(float(D["grid_power"] * cfg.dt))
        S["Utilized_solar_energy"].append(float(D["solar"]))
        S["Total_vehicle_penalties"].append(float(D["pen_veh"]))
        S["Insufficiently_charged_vehicle_penalties"].append(float(D["pen_veh"]))
        S["Total_battery_penalties"].append(float(D["pen_batt"]))
        S["Battery_SOC_below_DoD_penalties"].append(float(D["pen_batt"]))
        S["Total_penalties"].append(float(cfg.battery_penalty_weight * D["pen_batt"] + D["pen_veh"]))
        S["Total_cost"].append(-rew)
        S["Battery_state_of_charge"].append(float(row[nd + 2 * cfg.n_spots]) if cfg.batt else 0.0)
        S["Grid_energy_cost"].append(float(D["grid_cost"]))
        S["Battery_action"].append(float(a[cfg.n_spots]) if cfg.batt else 0.0)
        S["Charger_actions"].append(a[:cfg.n_spots].tolist())
        S["Total_charging_power"].append(float(D["total_ch"]))
        S["Total_discharging_power"].append(float(D["total_dis"]))
        S["Charger_power_values"].append(P.tolist())
        S["Battery_power_value"].append(float(D["batt_power"]))
        # the reference records the power the action asked for, before the over-discharge limit
        # (battery_energy_storage_system.py:49,79; 0 when the action is 0, :31-33)
        a_b = float(a[cfg.n_spots]) if cfg.batt else 0.0
        S["Battery_calculated_power_value"].append(a_b * cfg.bess_max_power * cfg.bess_efficiency if a_b != 0 else 0.0)
        S["DisCharging_nonexistent_vehicles_penalties"].append(nonexistent)
        self.t += 1
        return out

    def results(self) -> dict:
        """The dict the reference dumps (envs/smart_nanogrid_environment.py:246-275), key for key."""
        cfg, S = self.cfg, self.series
        zeros = [0.0] * len(S["Grid_power"])
        avail = (np.asarray(cfg.pv_power) * cfg.dt)[None, :].tolist() if cfg.pv else []
        return {
            "SOC": self.soc.tolist(), "Grid_power": S["Grid_power"], "Grid_energy": S["Grid_energy"],
            "Utilized_solar_energy": S["Utilized_solar_energy"], "Total_vehicle_penalties": S["Total_vehicle_penalties"],
            "Total_battery_penalties": S["Total_battery_penalties"], "Total_penalties": S["Total_penalties"],
            "Available_solar_energy": avail, "Total_cost": S["Total_cost"],
            "Battery_state_of_charge": S["Battery_state_of_charge"],
            "Initial_battery_state_of_charge": self.initial_battery_soc, "Grid_energy_cost": S["Grid_energy_cost"],
            "Battery_action": S["Battery_action"], "Charger_actions": S["Charger_actions"],
            "Total_charging_power": S["Total_charging_power"], "Total_discharging_power": S["Total_discharging_power"],
            "Charger_power_values": S["Charger_power_values"], "Battery_power_value": S["Battery_power_value"],
            "Battery_SOC_below_DoD_penalties": S["Battery_SOC_below_DoD_penalties"],
            "Low_resource_utilisation_penalties": zeros, "Battery_overcharging_penalties": zeros,
            "Battery_over_discharging_penalties": zeros,
            "Insufficiently_charged_vehicle_penalties": S["Insufficiently_charged_vehicle_penalties"],
            "Needlessly_charged_vehicle_penalties": zeros, "Overcharged_vehicle_penalties": zeros,
            "Over_discharged_vehicle_penalties": zeros,
            "Battery_calculated_power_value": S["Battery_calculated_power_value"],
            "DisCharging_nonexistent_vehicles_penalties": S["DisCharging_nonexistent_vehicles_penalties"],
        }

    def save(self, path: str):
        with open(path, "w") as fp:
            json.dump(self.results(), fp, indent=4)
