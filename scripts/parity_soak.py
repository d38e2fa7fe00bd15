import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle.oracle import OracleBatch
from parity_utils import assert_close_f32, branchy_actions, penalty_margin_distance
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
def ulp(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 5e-324)
for precision, E, episodes in (("float32", 8192, 40), ("float64", 2048, 40)):
    f64 = precision == "float64"
    kw = dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h",
              enable_requested_state_of_charge=True)
    env = BatchedSmartNanogridEnv(E, seed=123, precision=precision, want_terminal_obs=True, **kw)
    ob = OracleBatch(env.cfg, E, n_threads=16)
    obs = env.reset().cpu().numpy(); ob.sample(123, 0, 0); o_ref = ob.observe()
    rng = np.random.default_rng(0); lo, hi = env.cfg.action_bounds(); episode = np.zeros(E, np.uint32)
    worst_obs = worst_rew = 0.0; skipped = 0; t0 = time.time()
    for s in range(24 * episodes):
        a = branchy_actions(rng, lo, hi, (E,))
        a_or = a if f64 else a.astype(np.float32).astype(np.float64)
        near = penalty_margin_distance(ob) < (0 if f64 else 1e-5); skipped += int(near.sum())
        o_ref, r_ref, d_ref = ob.step(a_or)
        o, r, d, _, _ = env.step(torch.tensor(a, device="cuda:0", dtype=env.real))
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        assert np.array_equal(d, d_ref)
        if d_ref.any():
            episode += 1; ob.sample(123, 0, episode); o_ref = ob.observe()
        if f64:
            assert np.array_equal(o, o_ref), s
            worst_rew = max(worst_rew, float(ulp(r, r_ref).max())); assert worst_rew <= 4
        else:
            assert_close_f32("obs", o, o_ref, atol=1e-6); assert_close_f32("reward", r, r_ref, atol=1e-5, mask=~near)
            worst_obs = max(worst_obs, float(np.abs(o - o_ref).max()))
            worst_rew = max(worst_rew, float((np.abs(r - r_ref) / np.maximum(np.abs(r_ref), 1.0))[~near].max(initial=0)))
    print("%s: %d envs x %d episodes ok, worst |obs diff| %.2e, worst reward diff %.3g (%s), masked near-threshold env-steps %d of %d, %.0f s"
          % (precision, E, episodes, worst_obs, worst_rew, "ulp" if f64 else "relative", skipped, E * 24 * episodes, time.time() - t0), flush=True)
    env.close()
