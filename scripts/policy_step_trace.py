"""In-kernel clock64 trace of one CTA of the policy kernel with in-kernel exploration noise (two io sets), or -- argument
`fused` -- of the fused policy + env step kernel (sng_policy_step): compute warp 0 (slots 0..63), first thread of io set 0
(64..127) and of io set 1 (128..191).  Usage: policy_step_trace.py [forward|fused]"""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv, _native as nat
from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer, collect_rollout
dev = "cuda:0"; E = 65536
env = BatchedSmartNanogridEnv(E, device=dev, seed=1, number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
policy = MlpPolicy(29, 11).to(dev)
buf = RolloutBuffer(4, E, 29, 11, dev)
obs = env.reset(); starts = torch.ones(E, dtype=torch.uint8, device=dev)
for _ in range(2): collect_rollout(env, policy, buf, obs, starts, rng_seed=1, fuse_step=True)
tr = torch.zeros(256, dtype=torch.int64, device=dev)
lib = nat.lib(); lib.sng_policy_debug_trace.argtypes = [C.c_void_p]
lib.sng_policy_debug_trace(C.c_void_p(tr.data_ptr()))
low, high = env.action_low.float(), env.action_high.float()
if len(sys.argv) > 1 and sys.argv[1] == "fused":
    env.policy_step(policy._packed, buf.observations[0], low, high, buf.raw_actions[0], buf.actions[0], buf.values[0], buf.log_probs[0],
                    out=(buf.observations[1], buf.rewards[0], buf.dones[0]), rng=(1, policy.rng_counter, 0))
else:
    policy.fused_forward(buf.observations[0], None, low, high, buf.raw_actions[0], buf.actions[0], buf.values[0], buf.log_probs[0],
                         repack=False, rng=(1, policy.rng_counter, 0, 0))
torch.cuda.synchronize()
lib.sng_policy_debug_trace(None)
t = tr.cpu().tolist(); base = t[255]
print("kernel end at", t[254] - base)
for name, lo in (("compute", 0), ("io set 0", 64), ("io set 1", 128)):
    x = [v - base for v in t[lo:lo + 64] if v]
    print(name, "abs:", x)
