"""EV schedule containers and converters (host side).

The reference keeps, per charging spot, four dense arrays of 25 slots (occupancy, SoC,
capacity, requested SoC; utils/charger.py:16-19) plus ragged arrival / departure lists
(utils/charging_station.py:21-22).  The CUDA path keeps one compact record per vehicle:
`(arrival, departure, capacity, arrival SoC, requested SoC)`.  This module converts between
the two, validates the invariants the kernels rely on, and reads / writes the reference's
`initial_values.json` format (charging_station.py:119-136,173-191).
"""
from __future__ import annotations

import dataclasses
import json
from typing import Sequence

import numpy as np

MAX_VEHICLES = 8  # per spot and day; the reference generator yields <= 5 at 1 h, <= 6 at 15 min


@dataclasses.dataclass
class ScheduleRecords:
    """Compact schedules of E envs x N spots x V vehicle slots."""
    arr: np.ndarray    # int32 [E, N, V] arrival step
    dep: np.ndarray    # int32 [E, N, V] departure step (first step the spot is free again)
    cap: np.ndarray    # int32 [E, N, V] battery capacity in kWh (integer, as the generator draws it)
    soc0: np.ndarray   # float64 [E, N, V] state of charge on arrival
    req: np.ndarray    # float64 [E, N, V] requested end state of charge
    n_veh: np.ndarray  # int32 [E, N]

    @property
    def shape(self):
        return self.arr.shape

    def validate(self, n_steps: int):
        """Invariants of generated schedules (charging_station.py:200-255) that the kernels use."""
        E, N, V = self.arr.shape
        if V > MAX_VEHICLES:
            raise ValueError("more than %d vehicles per spot" % MAX_VEHICLES)
        k = np.arange(V)[None, None, :]
        valid = k < self.n_veh[:, :, None]
        a, d = self.arr, self.dep
        if (valid & ((a < 0) | (a >= n_steps))).any():
            raise ValueError("arrival outside the episode")
        if (valid & (d <= a)).any():
            raise ValueError("departure must be after arrival")
        if (valid & (d > 250)).any():
            raise ValueError("departure does not fit a byte")
        # the next vehicle arrives strictly after the previous one left (:239-251)
        nxt = valid[:, :, 1:] & (a[:, :, 1:] <= d[:, :, :-1])
        if nxt.any():
            raise ValueError("vehicles overlap or arrive on a departure step")
        c = self.cap
        if (valid & ((c < 1) | (c > 255))).any():
            raise ValueError("capacity must be an integer in 1..255 kWh")
        return self


def empty_records(n_envs: int, n_spots: int, n_slots: int = MAX_VEHICLES) -> ScheduleRecords:
    z = lambda dt: np.zeros((n_envs, n_spots, n_slots), dt)  # noqa: E731
    return ScheduleRecords(z(np.int32), z(np.int32), z(np.int32), z(np.float64), z(np.float64),
                           np.zeros((n_envs, n_spots), np.int32))


def records_from_dense(soc, occ, cap, req, arrivals: Sequence[Sequence[int]],
                       departures: Sequence[Sequence[int]], n_steps: int, n_slots: int = MAX_VEHICLES,
                       check: bool = True) -> ScheduleRecords:
    """One env: reference dense arrays [N, >= n_steps+1] + ragged lists -> records [1, N, V].

    With `check`, verifies that the dense arrays are exactly what the lists imply (occupancy
    on [arr, dep), constant capacity / requested SoC during the stay, SoC seeded only at the
    arrival slot) so that the compact form loses nothing."""
    soc, occ, cap, req = (np.asarray(x, np.float64) for x in (soc, occ, cap, req))
    N = soc.shape[0]
    rec = empty_records(1, N, n_slots)
    for i in range(N):
        n = len(arrivals[i])
        if n != len(departures[i]):
            raise ValueError("arrivals / departures length mismatch")
        rec.n_veh[0, i] = n
        for v in range(n):
            a, d = int(arrivals[i][v]), int(departures[i][v])
            c = cap[i, a]
            if c != int(c):
                raise ValueError("non-integer vehicle capacity")
            rec.arr[0, i, v], rec.dep[0, i, v], rec.cap[0, i, v] = a, d, int(c)
            rec.soc0[0, i, v], rec.req[0, i, v] = soc[i, a], req[i, a]
    rec.validate(n_steps)
    if check:
        d_soc, d_occ, d_cap, d_req = dense_from_records(rec, n_steps, width=soc.shape[1])
        for name, x, y in (("SOC", soc, d_soc[0]), ("Charger_occupancy", occ, d_occ[0]),
                           ("Vehicle_capacities", cap, d_cap[0]), ("Requested_SOC", req, d_req[0])):
            if not np.array_equal(x, y):
                raise ValueError("dense %s array is inconsistent with the arrival/departure lists" % name)
    return rec


def dense_from_records(rec: ScheduleRecords, n_steps: int, width: int | None = None):
    """records -> the reference's dense (soc, occ, cap, req) arrays [E, N, width]."""
    E, N, V = rec.arr.shape
    W = (n_steps + 1) if width is None else width
    soc, occ, cap, req = (np.zeros((E, N, W)) for _ in range(4))
    tt = np.arange(W)[None, None, :]
    for v in range(V):
        valid = (v < rec.n_veh)[:, :, None]
        a, d = rec.arr[:, :, v, None], rec.dep[:, :, v, None]
        present = valid & (tt >= a) & (tt < d) & (tt < n_steps)
        occ[present] = 1.0
        cap = np.where(present, rec.cap[:, :, v, None].astype(np.float64), cap)
        req = np.where(present, rec.req[:, :, v, None], req)
        soc = np.where(valid & (tt == a), rec.soc0[:, :, v, None], soc)
    return soc, occ, cap, req


def concat_records(items: Sequence[ScheduleRecords]) -> ScheduleRecords:
    return ScheduleRecords(*[np.concatenate([getattr(r, f.name) for r in items], axis=0)
                             for f in dataclasses.fields(ScheduleRecords)])


def load_initial_values_json(path: str, n_steps: int = 24, restore_requested_soc: bool = True) -> ScheduleRecords:
    """Read the reference's `initial_values.json` (charging_station.py:173-180).

    The reference's own loader forgets `Requested_SOC` (quirk Q7, charging_station.py:119-136),
    which silently disables the undercharge penalty; `restore_requested_soc=False` reproduces that."""
    with open(path) as fp:
        d = json.load(fp)
    req = np.array(d["Requested_SOC"]) if restore_requested_soc else np.zeros_like(np.array(d["SOC"]))
    return records_from_dense(d["SOC"], d["Charger_occupancy"], d["Vehicle_capacities"], req,
                              d["Arrivals"], d["Departures"], n_steps, check=restore_requested_soc)


def save_initial_values_json(path: str, rec: ScheduleRecords, n_steps: int = 24, env: int = 0):
    """Write env `env` in the reference's `initial_values.json` layout (25-slot arrays at 1 h)."""
    one = ScheduleRecords(*[getattr(rec, f.name)[env:env + 1] for f in dataclasses.fields(ScheduleRecords)])
    soc, occ, cap, req = dense_from_records(one, n_steps)
    n = one.n_veh[0]
    out = {
        "SOC": soc[0].tolist(),
        "Arrivals": [one.arr[0, i, :n[i]].tolist() for i in range(n.shape[0])],
        "Departures": [one.dep[0, i, :n[i]].tolist() for i in range(n.shape[0])],
        "Charger_occupancy": occ[0].tolist(),
        "Vehicle_capacities": cap[0].tolist(),
        "Requested_SOC": req[0].tolist(),
    }
    with open(path, "w") as fp:
        json.dump(out, fp, indent=4)
