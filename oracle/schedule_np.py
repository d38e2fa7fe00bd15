"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's EV schedule generator with the
legacy numpy generator it uses, in the reference's exact draw order.

Follows ChargingStation.generate_initial_vehicle_presence_per_charger
(utils/charging_station.py:200-255) and the four generate_random_* helpers (:257-279).
Pinned: `tests/golden/ref_schedules_seeded.npz` holds schedules the live reference produced
under `np.random.seed(s)`; tests/test_oracle_golden.py replays the seeds through this module.
"""
import numpy as np


def generate_spot(rs, n_steps, dt, diff_cap, req_soc):
    """One spot.  `rs` is a np.random.RandomState (the reference uses the global one).
    Returns (arrivals, departures, soc0s, caps, reqs)."""
    arrivals, departures, soc0s, caps, reqs = [], [], [], [], []
    present = False
    dep_cur = 0
    for t in range(n_steps):
        if not present:
            arrival = round(rs.rand() - 0.1)                     # :214
            if arrival == 1 and t < n_steps:
                present = True
                soc0 = rs.uniform(0.1, 0.9)                      # :218, 257-259
                lo = soc0 + 0.1 if soc0 <= 0.9 else 1.0          # :261-265
                rs.uniform(lo, 1.0)                              # :219 -- drawn and discarded
                cap = rs.randint(15, 120) if diff_cap else 40    # :220-225, 267-269
                req = rs.uniform(lo, 1.0) if req_soc else 1.0    # :227-232
                low = t + int(4 / dt)                            # :271-279
                high = int(min(t + int(10 / dt), n_steps + int(1 / dt)))
                dep_cur = int(low) if low >= high else rs.randint(low, high)
                arrivals.append(t)
                departures.append(dep_cur)
                soc0s.append(soc0)
                caps.append(cap)
                reqs.append(req)
        if present and t < dep_cur:                              # :239-242
            pass
        else:                                                    # :243-251
            present = False
    return arrivals, departures, soc0s, caps, reqs


def generate_station(rs, n_spots, n_steps, dt, diff_cap, req_soc, n_slots=8):
    """All spots of one env, in charger order (:193-198).  Returns record arrays [N, V] + counts."""
    arr = np.zeros((n_spots, n_slots), np.int32)
    dep = np.zeros((n_spots, n_slots), np.int32)
    cap = np.zeros((n_spots, n_slots), np.int32)
    soc0 = np.zeros((n_spots, n_slots))
    req = np.zeros((n_spots, n_slots))
    n_veh = np.zeros(n_spots, np.int32)
    for i in range(n_spots):
        a, d, s, c, r = generate_spot(rs, n_steps, dt, diff_cap, req_soc)
        n = len(a)
        n_veh[i] = n
        arr[i, :n], dep[i, :n], cap[i, :n], soc0[i, :n], req[i, :n] = a, d, c, s, r
    return arr, dep, cap, soc0, req, n_veh
