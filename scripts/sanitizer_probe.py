"""Small workloads for compute-sanitizer: step kernels (specialised, generic, four lanes, rollout), policy kernels, GAE."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer, collect_rollout
dev = "cuda:0"
KW = dict(charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
g = torch.Generator(device=dev).manual_seed(0)
# (small batches of the default station shapes run the one-lane-per-spot kernel; variant 5 keeps them on the block kernel)
for kw, E, variant in ((dict(number_of_chargers=10), 333, 0), (dict(number_of_chargers=10), 333, 5),
                       (dict(number_of_chargers=4, vehicle_to_everything=True, vehicle_uncharged_penalty_mode="dense"), 77, 4),
                       (dict(number_of_chargers=10, hours_ahead=5), 97, 0),
                       (dict(number_of_chargers=64, time_interval="15min"), 96, 0), (dict(number_of_chargers=7, vehicle_to_everything=True), 65, 0)):
    k = dict(KW); k.update(kw)
    env = BatchedSmartNanogridEnv(E, device=dev, seed=1, want_terminal_obs=True, want_diagnostics=True, **k)
    env.set_pipeline(variant)
    env.reset()
    for s in range(env.cfg.n_steps + 3):
        env.step(env.sample_actions(g))
    acts = torch.stack([env.sample_actions(g) for _ in range(4)]).contiguous()
    env.rollout(acts)
    assert env.error_flags() == 0
    env.close()
    print("step ok", kw, "variant", variant, flush=True)
env = BatchedSmartNanogridEnv(640 + 17, device=dev, seed=2, number_of_chargers=10, **KW)
policy = MlpPolicy(29, 11).to(dev)
buf = RolloutBuffer(6, env.num_envs, 29, 11, dev)
obs = env.reset()
starts = torch.ones(env.num_envs, dtype=torch.uint8, device=dev)
for kwargs in (dict(rng_seed=3), dict(generator=g), dict(deterministic=True), dict(rng_seed=3, pdl=True), dict(rng_seed=3, pdl="policy")):
    obs, starts = collect_rollout(env, policy, buf, obs, starts, **kwargs)
    torch.cuda.synchronize()
    print("rollout ok", kwargs.keys(), flush=True)
E = 128 * 9          # aligned: the copy-engine paths of the policy kernel
o = torch.rand(E, 29, device=dev); raw = torch.empty(E, 11, device=dev); act = torch.empty_like(raw)
val = torch.empty(E, device=dev); lp = torch.empty(E, device=dev)
low = torch.zeros(11, device=dev); high = torch.ones(11, device=dev)
policy.fused_forward(o, torch.randn(E, 11, device=dev), low, high, raw, act, val, lp)
policy.fused_forward(o, None, None, None, None, None, val, None, repack=False)
policy.fused_forward(o, torch.randn(E, 11, device=dev), low, high, raw, act, val, lp, cuda_cores=True)
torch.cuda.synchronize()
print("policy ok", flush=True)
# fused policy + env step (sng_policy_step): in-kernel noise, supplied noise, programmatic launch; two tiles per CTA at most
env = BatchedSmartNanogridEnv(128 * 5, device=dev, seed=4, number_of_chargers=10, **KW)
buf = RolloutBuffer(26, env.num_envs, 29, 11, dev)
obs = env.reset()
starts = torch.ones(env.num_envs, dtype=torch.uint8, device=dev)
for kwargs in (dict(rng_seed=3, fuse_step=True), dict(generator=g, fuse_step=True), dict(rng_seed=3, fuse_step=True, pdl="policy")):
    obs, starts = collect_rollout(env, policy, buf, obs, starts, **kwargs)
    torch.cuda.synchronize()
    print("fused rollout ok", sorted(kwargs.keys()), flush=True)
assert env.error_flags() == 0
import ctypes as C
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for variant in (0, 1, 2, 4):
    assert env._lib.sng_debug_traffic_skeleton(env._h, variant, stream) == 0
torch.cuda.synchronize()
print("skeleton ok", flush=True)
