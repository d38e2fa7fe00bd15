import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer, ShardedGraphedRollout
KW = dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
E, shards, det, nrep, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
per = E // shards; dev = "cuda:0"
def P(*a): print(*a, flush=True)
envs = [BatchedSmartNanogridEnv(per, device=dev, seed=0, env_gid0=k * per, **KW) for k in range(shards)]
policy = MlpPolicy(29, 11).to(dev)
bufs = [RolloutBuffer(n, per, 29, 11, dev) for _ in range(shards)]
obs = [e.reset() for e in envs]
starts = [torch.ones(per, dtype=torch.uint8, device=dev) for _ in range(shards)]
c = ShardedGraphedRollout(envs, policy, bufs, deterministic=bool(det))
torch.cuda.synchronize(); P("captured")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(nrep):
    obs, starts = c(obs, starts)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / nrep
P("E=%d shards=%d det=%d reps=%d n=%d: per step %.2f us, %.3g env-steps/s" % (E, shards, det, nrep, n, ms / n * 1e3, E * n / (ms * 1e-3)))
