"""Shared helpers of the parity tests (CUDA path vs the float64 oracle)."""
import numpy as np

from oracle.oracle import OracleBatch
from smart_nanogrid_gym_b200.schedule import ScheduleRecords

DEFAULT = dict(charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")


def records_from_oracle(ob: OracleBatch) -> ScheduleRecords:
    """Compact records of the oracle's current (pristine) dense schedule."""
    E, N, V = ob.arr.shape
    k = np.arange(V)[None, None, :]
    valid = k < ob.n_veh[:, :, None]
    arr = np.where(valid, ob.arr, 0).astype(np.int32)
    dep = np.where(valid, ob.dep, 0).astype(np.int32)
    a64 = arr.astype(np.int64)
    cap = np.where(valid, np.take_along_axis(ob.cap, a64, axis=2), 0).astype(np.int32)
    soc0 = np.where(valid, np.take_along_axis(ob.soc, a64, axis=2), 0.0)
    req = np.where(valid, np.take_along_axis(ob.req, a64, axis=2), 0.0)
    return ScheduleRecords(arr, dep, cap, soc0, req, ob.n_veh.copy())


def branchy_actions(rng, lo, hi, shape):
    """U(low, high) with exact zeros and saturated bounds sprinkled in."""
    a = rng.uniform(lo, hi, size=shape + lo.shape)
    z = rng.random(a.shape)
    a[z < 0.15] = 0.0
    hi_b = np.broadcast_to(hi, a.shape)
    lo_b = np.broadcast_to(lo, a.shape)
    m = (z >= 0.15) & (z < 0.20)
    a[m] = hi_b[m]
    m = (z >= 0.20) & (z < 0.25)
    a[m] = lo_b[m]
    return a


def penalty_margin_distance(ob: OracleBatch):
    """For the step the oracle is ABOUT to take: per env, the smallest |s - (r - 0.05 r)| over the vehicles
    in the penalty-check set (inf when the set is empty).  The undercharge penalty jumps by 0.25 r^2 there
    (penaliser.py:72-79), so float32 and float64 may legitimately land on different sides."""
    E, N, W = ob.soc.shape
    col = np.where(ob.t == 0, W - 1, ob.t - 1).astype(np.int64)[:, None, None]
    s = np.take_along_axis(ob.soc, col, axis=2)[:, :, 0]
    r = np.take_along_axis(ob.req, col, axis=2)[:, :, 0]
    d = np.abs(s - (r - ob.cfg.soc_margin_ratio * r))
    d = np.where(ob.check.astype(bool) & (r > 0), d, np.inf)     # r == 0 (empty column, e.g. at t = 0): no penalty either way
    return d.min(axis=1)


def assert_close_f32(name, got, want, rtol=1e-5, atol=0.0, mask=None):
    """|got - want| <= atol + rtol * |want| (the north-star tolerance: 1e-5 relative in float32)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    bad = np.abs(got - want) > (atol + rtol * np.abs(want))
    if mask is not None:
        bad &= mask
    if bad.any():
        idx = np.argwhere(bad)[0]
        raise AssertionError("%s: %d mismatches, first at %s: got %r want %r" %
                             (name, bad.sum(), tuple(idx), got[tuple(idx)], want[tuple(idx)]))


def penalty_margin_table(ob: OracleBatch):
    """Per vehicle, for the step the oracle is ABOUT to take: distance |s - 0.95 r| to the penalty threshold
    (inf outside the check set) and the size of the jump ((r - s) * 10)^2 the penalty makes there."""
    E, N, W = ob.soc.shape
    col = np.where(ob.t == 0, W - 1, ob.t - 1).astype(np.int64)[:, None, None]
    s = np.take_along_axis(ob.soc, col, axis=2)[:, :, 0]
    r = np.take_along_axis(ob.req, col, axis=2)[:, :, 0]
    d = np.abs(s - (r - ob.cfg.soc_margin_ratio * r))
    live = ob.check.astype(bool) & (r > 0)
    return np.where(live, d, np.inf), np.where(live, ((r - s) * 10.0) ** 2, 0.0)


def assert_only_threshold_side_differs(name, got, want, dist, jump, near_tol=1e-5, rtol=1e-5, atol=1e-4):
    """Env-steps masked as "near the penalty threshold" may differ from the oracle ONLY by the penalty jump of
    some of their near-threshold vehicles (float32 landed on the other side of s = 0.95 r): |got - want| must be
    a subset sum of those jumps."""
    from itertools import combinations
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    for e in np.flatnonzero(dist.min(axis=1) < near_tol):
        jumps = jump[e][dist[e] < near_tol]
        delta = abs(got[e] - want[e])
        tol = atol + rtol * abs(want[e])
        sums = [0.0] + [sum(c) for k in range(1, len(jumps) + 1) for c in combinations(jumps.tolist(), k)]
        if not any(abs(delta - x) <= tol for x in sums):
            raise AssertionError("%s: env %d near the penalty threshold differs by %r, not by a jump of %r" %
                                 (name, e, delta, jumps.tolist()))
