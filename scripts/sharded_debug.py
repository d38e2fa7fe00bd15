"""Debug: ShardedGraphedRollout at bench size, with progress prints."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer, ShardedGraphedRollout, collect_rollout
KW = dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
shards = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 24
per = E // shards
dev = "cuda:0"
def P(*a):
    print(*a, flush=True)
envs = [BatchedSmartNanogridEnv(per, device=dev, seed=0, env_gid0=k * per, **KW) for k in range(shards)]
policy = MlpPolicy(29, 11).to(dev)
bufs = [RolloutBuffer(n, per, 29, 11, dev) for _ in range(shards)]
obs = [e.reset() for e in envs]
starts = [torch.ones(per, dtype=torch.uint8, device=dev) for _ in range(shards)]
torch.cuda.synchronize(); P("envs ready")
# eager, two streams, no graph
policy.pack_weights()
policy.rng_counter = torch.zeros(1, dtype=torch.int64, device=dev)
sts = [torch.cuda.Stream(device=dev) for _ in range(shards)]
for k, st in enumerate(sts):
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        collect_rollout(envs[k], policy, bufs[k], obs[k], starts[k], rng_seed=0, pack=False, advance_counter=False)
for st in sts:
    torch.cuda.current_stream().wait_stream(st)
torch.cuda.synchronize(); P("eager two-stream rollout done")
c = ShardedGraphedRollout(envs, policy, bufs)
torch.cuda.synchronize(); P("graph captured")
t0 = time.time()
for i in range(5):
    obs, starts = c(obs, starts)
    torch.cuda.synchronize(); P("replay", i, "%.3f ms" % ((time.time() - t0) * 1e3)); t0 = time.time()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(50):
    obs, starts = c(obs, starts)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 50
P("per rollout %.3f ms, per step %.2f us, %.3g env-steps/s" % (ms, ms / n * 1e3, E * n / (ms * 1e-3)))
