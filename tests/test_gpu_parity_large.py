"""Lock-step parity of the CUDA step against the float64 oracle at BASELINE.json's full sizes.

The GPU runs the whole batch (1,048,576 envs = 16,384 CTAs, several waves; block indices above 2^15; global env ids
up to 2^18 * 64 spots) and records every step's outputs; the oracle then replays the batch in chunks of envs
(`env_gid0` offsets bound its dense arrays to ~0.5 GB) and every observation, reward and termination flag of every
env-step is compared.  Needs a B200: run with `-m gpu`.
"""
import numpy as np
import pytest

from parity_utils import (DEFAULT, assert_close_f32, assert_only_threshold_side_differs, penalty_margin_table)

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(n_envs, precision, **kw):
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    full = dict(DEFAULT)
    full.update(kw)
    return BatchedSmartNanogridEnv(n_envs, device="cuda:0", precision=precision, want_terminal_obs=True, auto_reset=True,
                                   **full)


def _branchy_actions_device(env, g):
    """U(low, high) with exact zeros and saturated bounds sprinkled in, generated on the device."""
    E, A = env.num_envs, env.cfg.act_dim
    u = torch.rand(E, A, device=env.device, dtype=env.real, generator=g)
    z = torch.rand(E, A, device=env.device, generator=g)
    lo, hi = env.action_low.expand(E, A), env.action_high.expand(E, A)
    a = lo + (hi - lo) * u
    a = torch.where(z < 0.15, torch.zeros_like(a), a)
    a = torch.where((z >= 0.15) & (z < 0.20), hi, a)
    a = torch.where((z >= 0.20) & (z < 0.25), lo, a)
    return a.contiguous()


def _ulp_diff(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 5e-324)


def _chunked_lockstep(n_envs, precision, chunk, seed, extra_steps=5, env_gid0=0, threads=16, **kw):
    from oracle.oracle import OracleBatch
    f64 = precision == "float64"
    env = _env(n_envs, precision, seed=seed, env_gid0=env_gid0, **kw)
    cfg = env.cfg
    E, A, D, T = n_envs, cfg.act_dim, cfg.obs_dim, cfg.n_steps
    n_steps = T + extra_steps
    dev = env.device
    # ---- the GPU run, every step's inputs and outputs kept on the device ----
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    obs0 = env.reset().clone()
    acts = torch.empty(n_steps, E, A, device=dev, dtype=env.real)
    obs = torch.empty(n_steps, E, D, device=dev)
    rew = torch.empty(n_steps, E, device=dev, dtype=env.real)
    done = torch.empty(n_steps, E, device=dev, dtype=torch.uint8)
    tobs = {}
    for s in range(n_steps):
        acts[s] = _branchy_actions_device(env, g)
        env.step(acts[s], out=(obs[s], rew[s], done[s]))
        if (s + 1) % T == 0:
            tobs[s] = env.terminal_obs.clone()
    assert env.error_flags() == 0
    st = env.env_state()
    # ---- the oracle, chunk by chunk ----
    masked = 0
    for c0 in range(0, E, chunk):
        c1 = min(c0 + chunk, E)
        n = c1 - c0
        ob = OracleBatch(cfg, n, n_threads=threads)
        ob.sample(seed, env_gid0 + c0, 0)
        o_ref = ob.observe()
        got0 = obs0[c0:c1].cpu().numpy()
        assert np.array_equal(got0, o_ref) if f64 else np.allclose(got0, o_ref, rtol=1e-5, atol=1e-6), c0
        episode = np.zeros(n, np.uint32)
        for s in range(n_steps):
            a = acts[s, c0:c1].double().cpu().numpy()          # the oracle sees exactly what the kernel saw
            dist, jump = penalty_margin_table(ob)
            near = dist.min(axis=1) < (0 if f64 else 1e-5)
            masked += int(near.sum())
            o_ref, r_ref, d_ref = ob.step(a)
            o, r, d = obs[s, c0:c1].cpu().numpy(), rew[s, c0:c1].cpu().numpy(), done[s, c0:c1].cpu().numpy()
            assert np.array_equal(d, d_ref), (c0, s)
            if d_ref.any():
                assert d_ref.all()
                episode += 1
                ob.sample(seed, env_gid0 + c0, episode)
                o_term, o_ref = o_ref, ob.observe()
                t_got = tobs[s][c0:c1].cpu().numpy()
                if f64:
                    assert np.array_equal(t_got, o_term), (c0, s)
                else:
                    assert_close_f32("terminal_obs", t_got, o_term, atol=1e-6)
            if f64:
                assert np.array_equal(o, o_ref), (c0, s)
                assert _ulp_diff(r, r_ref).max() <= 4, (c0, s)
            else:
                assert_close_f32("obs[chunk %d, step %d]" % (c0, s), o, o_ref, atol=1e-6)
                assert_close_f32("reward[chunk %d, step %d]" % (c0, s), r, r_ref, atol=1e-5, mask=~near)
                assert_only_threshold_side_differs("reward[chunk %d, step %d]" % (c0, s), r, r_ref, dist, jump)
        assert np.array_equal(st["t"][c0:c1], ob.t) and np.array_equal(st["episode"][c0:c1], episode)
        if f64:
            assert np.array_equal(st["soc_b"][c0:c1], ob.soc_b)
        else:
            assert_close_f32("soc_b", st["soc_b"][c0:c1], ob.soc_b, atol=1e-6)
        del ob
    env.close()
    return masked, E * n_steps


@pytest.mark.parametrize("precision", ["float32", "float64"])
def test_config4_slice_131072_envs_lockstep_vs_oracle(precision):
    """BASELINE config 4, the per-GPU slice at 8 GPUs (the LAST slice: global env ids 917,504 .. 1,048,575)."""
    masked, total = _chunked_lockstep(131072, precision, 65536, seed=404, env_gid0=1048576 - 131072, number_of_chargers=10)
    assert masked < total * 1e-4


def test_config4_whole_batch_1048576_envs_lockstep_vs_oracle():
    """BASELINE config 4 on one GPU: 16,384 CTAs over several waves, one episode + 5 steps, every env-step checked."""
    masked, total = _chunked_lockstep(1048576, "float32", 65536, seed=405, number_of_chargers=10)
    assert masked < total * 1e-4


def test_config5_16384_envs_at_the_top_of_the_id_range_lockstep_vs_oracle():
    """BASELINE config 5 (64 spots, 96 steps): the last 16,384 of its 262,144 envs, so the Philox stream ids
    (global env * 64 + spot) reach 2^24; two lanes per env."""
    masked, total = _chunked_lockstep(16384, "float32", 4096, seed=406, env_gid0=262144 - 16384, number_of_chargers=64,
                                      time_interval="15min")
    assert masked < total * 1e-3


@pytest.mark.parametrize("kw", [dict(number_of_chargers=10, hours_ahead=5), dict(number_of_chargers=6, hours_ahead=1),
                                dict(number_of_chargers=10)])
def test_rbc_rule_with_other_forecast_horizons(kw):
    """The rule-based controller reads the departure entries at (1 + pv)(1 + H) + N: checked against the rule applied
    to the decoded spot state (independent of the observation layout) for H = 5, 1 and 3."""
    env = _env(512, "float32", seed=3, **kw)
    cfg = env.cfg
    obs = env.reset()
    g = torch.Generator(device="cuda:0").manual_seed(0)
    st_obs, t_obs = env.spot_state(), 0          # the state and the time the current observation describes
    for t in range(cfg.n_steps - 1):
        a = env.rbc_actions(obs).cpu().numpy()
        present = (st_obs["arr"] != 255) & (st_obs["arr"] <= t_obs) & (t_obs < st_obs["dep"])
        dep_norm = np.where(present, ((st_obs["dep"] - t_obs) / cfg.departure_normaliser).astype(np.float32), 0.0).astype(np.float64)
        o = obs.cpu().numpy().astype(np.float64)
        rad = ((o[:, 0] + o[:, 2]) / 2)[:, None]
        want = np.where(dep_norm == 0, 0.0, np.where(dep_norm < 0.16667, 1.0, np.broadcast_to(rad, dep_norm.shape)))
        assert np.array_equal(a[:, :cfg.n_spots], want.astype(np.float32)), t
        assert (a[:, cfg.n_spots] == 0).all()
        # step() returns the observation taken BEFORE t += 1 (quirk Q5): it describes the pre-step vehicles at time t
        st_obs, t_obs = env.spot_state(), t
        obs = env.step(env.sample_actions(g))[0]
    env.close()


def test_masked_reset_refuses_to_change_handle_wide_settings():
    """a masked reset must not flip the other envs from replay to sampling or re-key them."""
    from smart_nanogrid_gym_b200 import _native as nat
    env = _env(64, "float32", seed=1, number_of_chargers=4)
    env.reset()
    mask = torch.zeros(64, dtype=torch.bool, device="cuda:0")
    mask[3] = True
    env.reset(mask=mask)                                   # same seed, sampling mode: fine
    with pytest.raises(nat.NativeError):
        env.reset(seed=2, mask=mask)                       # a new seed would re-key every env
    env.seed(1)
    rec = env.sample_plan()
    env.load_schedule(rec)
    with pytest.raises(nat.NativeError):
        env.reset(mask=mask)                               # replay mode: the other envs would start sampling
    env.close()
