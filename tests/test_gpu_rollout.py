"""SURVEY 8f row 1 (BASELINE config 3): on-device PPO rollout collection around the step -- the step kernel
writes into rollout-buffer slabs zero-copy, and the sng_gae kernel matches SB3's advantage recursion."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

KW = dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")


def test_gae_kernel_matches_sb3_recursion():
    from smart_nanogrid_gym_b200.rollout import RolloutBuffer, gae_reference
    n, E = 37, 5003
    g = torch.Generator(device="cuda:0").manual_seed(0)
    buf = RolloutBuffer(n, E, 29, 11, "cuda:0", gamma=0.99, gae_lambda=0.95)
    buf.rewards.copy_(-10 * torch.rand(n, E, device="cuda:0", generator=g))
    buf.values.copy_(-50 * torch.rand(n, E, device="cuda:0", generator=g))
    buf.episode_starts.copy_((torch.rand(n, E, device="cuda:0", generator=g) < 0.1).to(torch.uint8))
    last_values = -50 * torch.rand(E, device="cuda:0", generator=g)
    last_dones = (torch.rand(E, device="cuda:0", generator=g) < 0.3).to(torch.uint8)
    adv, ret = buf.compute_returns_and_advantage(last_values, last_dones)
    adv_ref, ret_ref = gae_reference(buf.rewards, buf.values, buf.episode_starts, last_values, last_dones, 0.99, 0.95)
    assert torch.allclose(adv.double(), adv_ref, rtol=1e-5, atol=1e-4)
    assert torch.allclose(ret.double(), ret_ref, rtol=1e-5, atol=1e-4)


def test_collect_rollout_is_zero_copy_and_matches_manual_stepping():
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer, collect_rollout, gae_reference
    E, n = 2048, 30
    env = BatchedSmartNanogridEnv(E, seed=3, **KW)
    twin = BatchedSmartNanogridEnv(E, seed=3, **KW)
    torch.manual_seed(0)
    policy = MlpPolicy(env.cfg.obs_dim, env.cfg.act_dim).to("cuda:0")
    buf = RolloutBuffer(n, E, env.cfg.obs_dim, env.cfg.act_dim, "cuda:0")
    ptrs = (buf.observations.data_ptr(), buf.rewards.data_ptr(), buf.dones.data_ptr())
    obs = env.reset()
    obs_twin = twin.reset().clone()
    assert torch.equal(obs, obs_twin)
    g = torch.Generator(device="cuda:0").manual_seed(5)
    starts = torch.ones(E, dtype=torch.uint8, device="cuda:0")
    last_obs, last_dones = collect_rollout(env, policy, buf, obs, starts, generator=g, fused=False)
    assert ptrs == (buf.observations.data_ptr(), buf.rewards.data_ptr(), buf.dones.data_ptr())   # nothing reallocated
    lo, hi = env.action_low, env.action_high
    assert (buf.actions >= lo).all() and (buf.actions <= hi).all()
    assert torch.equal(buf.actions, torch.minimum(torch.maximum(buf.raw_actions, lo), hi))
    # the same actions through a second env stepped the ordinary way reproduce every slab bit for bit
    assert torch.equal(buf.observations[0], obs_twin)
    for s in range(n):
        o, r, d, _, _ = twin.step(buf.actions[s].clone())
        assert torch.equal(o, buf.observations[s + 1]) and torch.equal(r, buf.rewards[s]) and torch.equal(d, buf.dones[s]), s
        expect_start = starts if s == 0 else buf.dones[s - 1]
        assert torch.equal(buf.episode_starts[s], expect_start)
    assert buf.dones[23].all() and buf.dones.sum().item() == E          # one termination per env in 30 steps
    assert torch.equal(last_obs, buf.observations[n]) and torch.equal(last_dones, buf.dones[n - 1])
    with torch.no_grad():
        assert torch.allclose(buf.values[7], policy.predict_values(buf.observations[7]), atol=1e-5)
    # the fused policy kernel fills the buffer the same way (same noise stream, float32 rounding apart)
    buf2 = RolloutBuffer(n, E, env.cfg.obs_dim, env.cfg.act_dim, "cuda:0")
    env2 = BatchedSmartNanogridEnv(E, seed=3, **KW)
    collect_rollout(env2, policy, buf2, env2.reset(), starts, generator=torch.Generator(device="cuda:0").manual_seed(5), fused=True)
    assert torch.allclose(buf2.raw_actions[0], buf.raw_actions[0], rtol=1e-4, atol=2e-5)
    assert torch.allclose(buf2.values[0], buf.values[0], rtol=1e-4, atol=2e-5)
    assert torch.equal(buf2.episode_starts, buf.episode_starts) and torch.equal(buf2.dones, buf.dones)
    env2.close()
    adv_ref, ret_ref = gae_reference(buf.rewards, buf.values, buf.episode_starts, buf.last_values, last_dones,
                                     buf.gamma, buf.gae_lambda)
    assert torch.allclose(buf.advantages.double(), adv_ref, rtol=1e-5, atol=1e-3)
    assert torch.allclose(buf.returns.double(), ret_ref, rtol=1e-5, atol=1e-3)
    assert env.error_flags() == 0
    env.close()
    twin.close()


def test_graphed_rollout_equals_eager():
    """The whole rollout (policy, clip, step kernel, GAE) captured in one CUDA graph reproduces the eager loop."""
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer, collect_rollout
    E, n = 4096, 24
    torch.manual_seed(1)
    policy = MlpPolicy(29, 11).to("cuda:0")
    env_e = BatchedSmartNanogridEnv(E, seed=9, **KW)
    env_g = BatchedSmartNanogridEnv(E, seed=9, **KW)
    buf_e = RolloutBuffer(n, E, 29, 11, "cuda:0")
    buf_g = RolloutBuffer(n, E, 29, 11, "cuda:0")
    env_g.reset()
    graphed = GraphedRollout(env_g, policy, buf_g, deterministic=True)     # steps env_g during warm-up and capture
    obs_e = env_e.reset()
    obs_g = env_g.reset(seed=9)
    env_g.load_state_dict(env_e.state_dict())
    starts = torch.ones(E, dtype=torch.uint8, device="cuda:0")
    s_e, s_g = starts, starts
    for _ in range(3):
        obs_e, s_e = collect_rollout(env_e, policy, buf_e, obs_e, s_e, deterministic=True)
        obs_g, s_g = graphed(obs_g, s_g)
        for name in ("observations", "actions", "rewards", "dones", "episode_starts", "values", "advantages", "returns"):
            a, b = getattr(buf_e, name), getattr(buf_g, name)
            assert torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-5), name
        obs_e, obs_g = obs_e.clone(), obs_g.clone()
        s_e, s_g = s_e.clone(), s_g.clone()
    env_e.close()
    env_g.close()


def test_sb3_style_vec_env_protocol():
    """numpy in / numpy out, 4-tuple step_wait, auto-reset with infos[i]['terminal_observation'] (SB3 VecEnv)."""
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.vec_env import SmartNanogridVecEnv
    E = 300
    venv = SmartNanogridVecEnv(E, seed=4, **KW)
    twin = BatchedSmartNanogridEnv(E, seed=4, want_terminal_obs=True, **KW)
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (E, 29) and obs.dtype == np.float32
    assert np.array_equal(obs, twin.reset().cpu().numpy())
    assert venv.observation_space.shape == (29,) and venv.action_space.shape == (11,) and venv.num_envs == E
    rng = np.random.default_rng(0)
    for s in range(26):
        a = rng.uniform(venv.action_space.low, venv.action_space.high, size=(E, 11)).astype(np.float32)
        venv.step_async(a)
        o, r, d, infos = venv.step_wait()
        to, tr, td, _, _ = twin.step(torch.tensor(a, device="cuda:0"))
        assert o.dtype == np.float32 and r.dtype == np.float32 and d.dtype == bool and len(infos) == E
        assert np.array_equal(o, to.cpu().numpy()) and np.array_equal(r, tr.cpu().numpy()) and np.array_equal(d, td.cpu().numpy().astype(bool))
        if s == 23:
            assert d.all()
            term = twin.terminal_obs.cpu().numpy()
            for i in (0, 17, E - 1):
                assert np.array_equal(infos[i]["terminal_observation"], term[i]) and infos[i]["TimeLimit.truncated"] is False
        else:
            assert not d.any() and all(info == {} for info in infos)
    assert venv.env_is_wrapped(object) == [False] * E and len(venv.get_attr("num_envs")) == E
    venv.close()
    twin.close()


@pytest.mark.parametrize("n_spots,obs_dim,act_dim", [(10, 29, 11), (4, 17, 5), (8, 25, 9)])
def test_fused_policy_kernel_matches_torch_modules(n_spots, obs_dim, act_dim):
    """sng_policy_forward (one launch) == the torch nn.Module forward (float32): values, sampled and clipped actions,
    log-probabilities; value-only mode; ragged batch."""
    from smart_nanogrid_gym_b200.rollout import MlpPolicy
    E = 5000 + 7
    torch.manual_seed(n_spots)
    policy = MlpPolicy(obs_dim, act_dim).to("cuda:0")
    with torch.no_grad():
        policy.log_std.copy_(torch.linspace(-1.0, 0.5, act_dim))
    assert policy.fused_supported()
    g = torch.Generator(device="cuda:0").manual_seed(1)
    obs = torch.rand(E, obs_dim, device="cuda:0", generator=g) * 1.5
    noise = torch.randn(E, act_dim, device="cuda:0", generator=g)
    low = torch.zeros(act_dim, device="cuda:0")
    low[-1] = -1.0
    high = torch.ones(act_dim, device="cuda:0")
    raw, act = torch.empty(E, act_dim, device="cuda:0"), torch.empty(E, act_dim, device="cuda:0")
    val, lp = torch.empty(E, device="cuda:0"), torch.empty(E, device="cuda:0")
    policy.fused_forward(obs, noise, low, high, raw, act, val, lp)
    with torch.no_grad():
        a_ref, v_ref, lp_ref = policy(obs, noise)
    assert torch.allclose(raw, a_ref, rtol=1e-4, atol=2e-5)
    assert torch.allclose(act, torch.minimum(torch.maximum(a_ref, low), high), rtol=1e-4, atol=2e-5)
    assert torch.allclose(val, v_ref, rtol=1e-4, atol=2e-5) and torch.allclose(lp, lp_ref, rtol=1e-5, atol=1e-4)
    val2 = torch.empty(E, device="cuda:0")
    policy.fused_forward(obs, None, None, None, None, None, val2, None)
    assert torch.equal(val2, val)
    raw_d, act_d, lp_d = torch.empty_like(raw), torch.empty_like(act), torch.empty_like(lp)
    policy.fused_forward(obs, None, low, high, raw_d, act_d, val2, lp_d)          # deterministic: the mean action
    with torch.no_grad():
        mean_ref, _, lp0_ref = policy(obs, None)
    assert torch.allclose(raw_d, mean_ref, rtol=1e-4, atol=2e-5) and torch.allclose(lp_d, lp0_ref, rtol=1e-5, atol=1e-4)


def test_in_kernel_exploration_noise():
    """sng_policy_forward_sampled: the noise drawn inside the policy kernel is Philox4x32-10 (key = seed, step; counter =
    global env, action / 4) through Box-Muller -- checked value for value against a float64 restatement; the sampled
    actions / log-probs are the torch modules' given that noise; the draw depends on (seed, step, global env) only."""
    from oracle.oracle import philox4x32_10
    from smart_nanogrid_gym_b200.rollout import MlpPolicy
    E, D, A = 4096 + 37, 29, 11
    torch.manual_seed(3)
    policy = MlpPolicy(D, A).to("cuda:0")
    g = torch.Generator(device="cuda:0").manual_seed(2)
    obs = torch.rand(E, D, device="cuda:0", generator=g)
    low, high = torch.zeros(A, device="cuda:0"), torch.ones(A, device="cuda:0")
    low[-1] = -1.0
    new = lambda *shape: torch.empty(*shape, device="cuda:0")  # noqa: E731
    raw, act, val, lp, z = new(E, A), new(E, A), new(E), new(E), new(E, A)
    counter = torch.tensor([7], dtype=torch.int64, device="cuda:0")
    seed, offset, gid0 = 0x1234567890ABCDEF, 3, 1000
    policy.fused_forward(obs, None, low, high, raw, act, val, lp, rng=(seed, counter, offset, gid0), noise_out=z)
    step = 7 + offset
    zc = z.cpu().numpy().astype(np.float64)
    for e in list(range(0, 64)) + [E - 1, E - 37, 2048]:
        for cb in range(3):
            x = philox4x32_10([(gid0 + e) & 0xFFFFFFFF, (gid0 + e) >> 32, cb, step & 0xFFFFFFFF],
                              [seed & 0xFFFFFFFF, (seed >> 32) ^ (step >> 32)])
            u = ((x >> 8).astype(np.float64) + 0.5) * 2.0 ** -24
            want = []
            for k in (0, 2):
                r = np.sqrt(-2.0 * np.log(u[k]))
                want += [r * np.cos(2 * np.pi * u[k + 1]), r * np.sin(2 * np.pi * u[k + 1])]
            got = zc[e, 4 * cb:4 * cb + 4]
            assert np.allclose(got, want[:len(got)], rtol=0, atol=5e-5), (e, cb, got, want)
    with torch.no_grad():
        a_ref, v_ref, lp_ref = policy(obs, z)
    assert torch.allclose(raw, a_ref, rtol=1e-4, atol=2e-5) and torch.allclose(val, v_ref, rtol=1e-4, atol=2e-5)
    assert torch.allclose(act, torch.minimum(torch.maximum(a_ref, low), high), rtol=1e-4, atol=2e-5)
    assert torch.allclose(lp, lp_ref, rtol=1e-5, atol=1e-4)
    # moments of 45k standard normals
    assert abs(float(z.mean())) < 0.02 and abs(float(z.var()) - 1.0) < 0.03 and float(z.abs().max()) < 6.5
    # a shard [h, E) with env_gid0 advanced by h draws what the whole batch drew; another step draws something else
    h = 2048
    raw2, act2, val2, lp2, z2 = new(E - h, A), new(E - h, A), new(E - h), new(E - h), new(E - h, A)
    policy.fused_forward(obs[h:], None, low, high, raw2, act2, val2, lp2, repack=False, rng=(seed, counter, offset, gid0 + h), noise_out=z2)
    assert torch.equal(z2, z[h:]) and torch.equal(raw2, raw[h:])
    counter += 1
    policy.fused_forward(obs[h:], None, low, high, raw2, act2, val2, lp2, repack=False, rng=(seed, counter, offset, gid0 + h), noise_out=z2)
    assert not torch.equal(z2, z[h:]) and abs(float((z2 * z[h:]).mean())) < 0.02


def test_graphed_rollout_with_in_kernel_noise_advances_between_replays():
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer
    E, n = 1024, 6
    env = BatchedSmartNanogridEnv(E, device="cuda:0", seed=1, number_of_chargers=10, charging_mode="bounded",
                                  vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
    torch.manual_seed(0)
    policy = MlpPolicy(env.cfg.obs_dim, env.cfg.act_dim).to("cuda:0")
    buf = RolloutBuffer(n, E, env.cfg.obs_dim, env.cfg.act_dim, "cuda:0")
    obs = env.reset()
    collect = GraphedRollout(env, policy, buf, rng_seed=11)
    c0 = int(policy.rng_counter.item())
    starts = torch.ones(E, dtype=torch.uint8, device="cuda:0")
    obs, starts = collect(obs, starts)
    first = buf.raw_actions.clone()
    assert int(policy.rng_counter.item()) == c0 + n
    obs, starts = collect(obs, starts)
    assert int(policy.rng_counter.item()) == c0 + 2 * n and not torch.equal(first, buf.raw_actions)
    # episode_starts[s + 1] is dones[s] (one array), and the GAE consumed them
    assert torch.equal(buf.episode_starts[1:], buf.dones[:-1]) and bool(torch.isfinite(buf.advantages).all())
    # noise of different steps of one rollout is different
    zs = (buf.raw_actions[1] - buf.raw_actions[0]).abs().mean()
    assert float(zs) > 1e-3
    env.close()


def test_programmatic_dependent_launch_changes_nothing_but_timing():
    """sng_set_launch_mode / sng_policy_set_launch_mode: the same rollout with and without programmatic dependent launch
    (step kernel in mode 2: state loads ahead of the wait; policy kernel in mode 1: weight image ahead of the wait; either
    one alone; policy CTAs claiming the whole shared memory) is bit-identical, eager and graphed;
    step-after-step launches in mode 1 equal ordinary launches."""
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer
    E, n = 4096 + 32, 30
    torch.manual_seed(2)
    policy = MlpPolicy(29, 11).to("cuda:0")
    res = []
    for pdl in (False, True, "policy", "step+x"):
        env = BatchedSmartNanogridEnv(E, device="cuda:0", seed=3, **KW)
        buf = RolloutBuffer(n, E, 29, 11, "cuda:0")
        env.reset()
        collect = GraphedRollout(env, policy, buf, rng_seed=4, pdl=pdl)
        obs = env.reset(reset_battery=True)
        policy.rng_counter.zero_()
        starts = torch.ones(E, dtype=torch.uint8, device="cuda:0")
        for _ in range(3):
            obs, starts = collect(obs, starts)
        res.append((buf.raw_actions.clone(), buf.rewards.clone(), buf.observations.clone(), buf.advantages.clone(), env._spot.clone()))
        assert env.error_flags() == 0
        env.close()
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert torch.equal(a, b)
    envs = [BatchedSmartNanogridEnv(E, device="cuda:0", seed=5, **KW) for _ in range(2)]
    envs[1].set_launch_mode(1)
    for e in envs:
        e.reset()
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for s in range(30):
        a = envs[0].sample_actions(g)
        o0, o1 = envs[0].step(a), envs[1].step(a)
        assert torch.equal(o0[0], o1[0]) and torch.equal(o0[1], o1[1]) and torch.equal(o0[2], o1[2])
    assert torch.equal(envs[0]._spot, envs[1]._spot)
    for e in envs:
        e.close()


@pytest.mark.parametrize("E,n_spots", [(1024, 10), (65536, 10), (38272, 10), (4096, 4)])
def test_fused_policy_step_equals_policy_then_step(E, n_spots):
    """sng_policy_step (ONE launch per rollout step: the policy kernel's io warps run the env step of their tile) gives bit
    for bit what the policy kernel followed by the step kernel gives: everything the rollout buffer stores, and the env
    state.  30 steps > one 24-step episode, so the fused auto-reset is crossed; 65,536 envs = four tiles per CTA,
    38,272 envs = CTAs with two and with three tiles."""
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer
    kw = dict(KW, number_of_chargers=n_spots)
    n = 30
    torch.manual_seed(6)
    res = []
    policy = None
    for fuse in (False, True):
        env = BatchedSmartNanogridEnv(E, device="cuda:0", seed=11, **kw)
        if policy is None:
            policy = MlpPolicy(env.cfg.obs_dim, env.cfg.act_dim).to("cuda:0")
        assert env.supports_policy_step()
        buf = RolloutBuffer(n, E, env.cfg.obs_dim, env.cfg.act_dim, "cuda:0")
        env.reset()
        collect = GraphedRollout(env, policy, buf, rng_seed=4, fuse_step=fuse, pdl="policy" if fuse else False)
        obs = env.reset(reset_battery=True)
        policy.rng_counter.zero_()
        starts = torch.ones(E, dtype=torch.uint8, device="cuda:0")
        for _ in range(2):
            obs, starts = collect(obs, starts)
        torch.cuda.synchronize()
        res.append([x.clone() for x in (buf.raw_actions, buf.actions, buf.values, buf.log_probs, buf.rewards, buf.observations,
                                        buf.dones, buf.advantages, env._spot, env._envst, env.last_return)])
        assert env.error_flags() == 0
        env.close()
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    assert int(res[0][6].sum()) > 0          # episodes ended inside the rollouts


def test_fused_policy_step_with_supplied_noise():
    """sng_policy_step with caller-supplied noise rows (per io set: its own noise stage and mbarrier) and with zero noise
    equals sng_policy_forward_packed followed by sng_step, step after step across an auto-reset."""
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import MlpPolicy
    E = 148 * 128 * 2 + 5 * 128            # CTAs with two and with three tiles
    torch.manual_seed(8)
    policy = MlpPolicy(29, 11).to("cuda:0")
    policy.pack_weights()
    envs = [BatchedSmartNanogridEnv(E, device="cuda:0", seed=13, **KW) for _ in range(2)]
    obs = [e.reset().clone() for e in envs]
    low, high = envs[0].action_low.float(), envs[0].action_high.float()
    z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device="cuda:0")  # noqa: E731
    bufs = [dict(raw=z(E, 11), act=z(E, 11), val=z(E), lp=z(E), obs=z(E, 29), rew=z(E), done=z(E, dtype=torch.uint8)) for _ in range(2)]
    g = torch.Generator(device="cuda:0").manual_seed(3)
    for s in range(27):
        noise = torch.randn(E, 11, device="cuda:0", generator=g) if s % 3 else torch.zeros(E, 11, device="cuda:0")
        a, b = bufs
        policy.fused_forward(obs[0], noise, low, high, a["raw"], a["act"], a["val"], a["lp"], repack=False)
        envs[0].step(a["act"], out=(a["obs"], a["rew"], a["done"]))
        envs[1].policy_step(policy._packed, obs[1], low, high, b["raw"], b["act"], b["val"], b["lp"],
                            out=(b["obs"], b["rew"], b["done"]), noise=noise)
        for k in a:
            assert torch.equal(a[k], b[k]), (s, k)
        obs = [a["obs"].clone(), b["obs"].clone()]
    assert torch.equal(envs[0]._spot, envs[1]._spot) and torch.equal(envs[0]._envst, envs[1]._envst)
    for e in envs:
        assert e.error_flags() == 0
        e.close()


def test_sharded_rollout_equals_the_unsharded_one():
    """Two env shards collected side by side (ShardedGraphedRollout: parallel graph branches, in-kernel noise keyed by
    global env id) produce bit for bit what one env of the summed size does."""
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer, ShardedGraphedRollout
    E, n, seed = 1024, 5, 9
    torch.manual_seed(4)
    policy = MlpPolicy(29, 11).to("cuda:0")
    one = BatchedSmartNanogridEnv(E, device="cuda:0", seed=seed, **KW)
    buf1 = RolloutBuffer(n, E, 29, 11, "cuda:0")
    one.reset()
    whole = GraphedRollout(one, policy, buf1, rng_seed=5)              # steps `one` 2 x n times (warm-up, capture)
    halves = [BatchedSmartNanogridEnv(E // 2, device="cuda:0", seed=seed, env_gid0=k * (E // 2), **KW) for k in range(2)]
    bufs = [RolloutBuffer(n, E // 2, 29, 11, "cuda:0") for _ in range(2)]
    for h in halves:
        h.reset()
    sharded = ShardedGraphedRollout(halves, policy, bufs, rng_seed=5)
    # both sides: second reset -> episode 1, battery back to its initial SoC (it survives plain resets, quirk Q8)
    obs1 = one.reset(reset_battery=True)
    obs2 = [h.reset(reset_battery=True) for h in halves]
    assert torch.equal(torch.cat(obs2), obs1)
    ones = torch.ones(E, dtype=torch.uint8, device="cuda:0")
    for rep in range(2):
        policy.rng_counter.fill_(100 * rep)
        o1, s1 = whole(obs1 if rep == 0 else o1, ones if rep == 0 else s1)
        a1, r1, v1, adv1 = buf1.raw_actions.clone(), buf1.rewards.clone(), buf1.values.clone(), buf1.advantages.clone()
        policy.rng_counter.fill_(100 * rep)
        o2, s2 = sharded(obs2 if rep == 0 else o2, [ones[:E // 2], ones[E // 2:]] if rep == 0 else s2)
        assert torch.equal(torch.cat([b.raw_actions for b in bufs], dim=1), a1)
        assert torch.equal(torch.cat([b.rewards for b in bufs], dim=1), r1)
        assert torch.equal(torch.cat([b.values for b in bufs], dim=1), v1)
        assert torch.equal(torch.cat([b.advantages for b in bufs], dim=1), adv1)
        assert torch.equal(torch.cat(o2), o1) and torch.equal(torch.cat(s2), s1)
    one.close()
    for h in halves:
        h.close()


@pytest.mark.timeout(240)
def test_concurrent_policy_and_step_kernels_do_not_stall():
    """Regression: with step-kernel CTAs of another stream sharing its SMs, a delayed compute warp of the policy kernel
    could fall two completions behind the issuer on the mbarrier that signalled both the heads and the next tile's
    layer 0, and wait forever (r2: 50 back-to-back replays of a two-branch rollout graph at 65,536 envs hung every
    time).  Heads now complete on their own barrier; 60 replays must finish."""
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer, ShardedGraphedRollout
    per, n = 32768, 24
    envs = [BatchedSmartNanogridEnv(per, device="cuda:0", seed=0, env_gid0=k * per, **KW) for k in range(2)]
    torch.manual_seed(0)
    policy = MlpPolicy(29, 11).to("cuda:0")
    bufs = [RolloutBuffer(n, per, 29, 11, "cuda:0") for _ in range(2)]
    obs = [e.reset() for e in envs]
    starts = [torch.ones(per, dtype=torch.uint8, device="cuda:0") for _ in range(2)]
    collect = ShardedGraphedRollout(envs, policy, bufs)
    for _ in range(60):
        obs, starts = collect(obs, starts)
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(b.advantages).all()) for b in bufs) and all(e.error_flags() == 0 for e in envs)
    for e in envs:
        e.close()


def test_shipped_sb3_policy_on_the_recorded_episode():
    """the reference's shipped PPO checkpoint (tests/golden/sb3_ppo_4ch_policy.npz) drives the
    N = 4 station: on the observations of the reference's recorded episode G1 the fused kernel's deterministic
    actions and values equal the torch forward of the same weights, and a whole rollout runs with it."""
    import json
    import os
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv, load_initial_values_json
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(gold, "sb3_ppo_4ch_policy.npz"))
    policy = MlpPolicy.from_sb3_state_dict({k: torch.tensor(z[k]) for k in z.files}).to("cuda:0")
    assert policy.fused_supported()
    kw = dict(KW, number_of_chargers=4)
    # G1's observations: replay the recorded schedule and actions
    rec = load_initial_values_json(os.path.join(gold, "g1_initial_values.json"))
    with open(os.path.join(gold, "g1_prediction_results.json")) as fp:
        p = json.load(fp)
    env = BatchedSmartNanogridEnv(1, auto_reset=False, **kw)
    obs = [env.load_schedule(rec, pv_shift=0.02, soc_b=p["Initial_battery_state_of_charge"]).clone()]
    for t in range(23):
        a = torch.tensor([p["Charger_actions"][t] + [p["Battery_action"][t]]], device="cuda:0", dtype=torch.float32)
        obs.append(env.step(a)[0].clone())
    env.close()
    obs = torch.cat(obs)                                                       # [24, 17]
    low = torch.tensor([0., 0, 0, 0, -1], device="cuda:0")
    high = torch.ones(5, device="cuda:0")
    raw, act = torch.empty(24, 5, device="cuda:0"), torch.empty(24, 5, device="cuda:0")
    val, lp = torch.empty(24, device="cuda:0"), torch.empty(24, device="cuda:0")
    policy.fused_forward(obs, None, low, high, raw, act, val, lp)
    with torch.no_grad():
        mean_ref, v_ref, _ = policy(obs, None)
    assert torch.allclose(raw, mean_ref, rtol=1e-4, atol=2e-5) and torch.allclose(val, v_ref, rtol=1e-4, atol=1e-4)
    assert torch.equal(act, torch.minimum(torch.maximum(raw, low), high))
    # the trained policy charges: its clipped actions are not all at the lower bound on this episode
    assert (act[:, :4] > 0).any()
    # ... and collects rollouts over a batch
    E = 4096
    env = BatchedSmartNanogridEnv(E, seed=2, **kw)
    buf = RolloutBuffer(24, E, 17, 5, "cuda:0")
    o = env.reset()
    collect = GraphedRollout(env, policy, buf, deterministic=True)
    collect(o, torch.ones(E, dtype=torch.uint8, device="cuda:0"))
    torch.cuda.synchronize()
    random_policy_return = -173.3                                              # tests/golden/ref_return_stats.json, N = 4
    ret = buf.rewards.sum(0).mean().item()
    # the trained policy beats random actions by far: -54.8 +- 16.5 per episode with the float64 oracle in the loop
    assert -70 < ret < -40 and ret > random_policy_return
    assert env.error_flags() == 0
    env.close()
