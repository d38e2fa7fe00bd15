"""Debug: the tensor-core policy kernel under concurrency (two streams) and at 2 tiles per CTA."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
from smart_nanogrid_gym_b200.rollout import MlpPolicy
dev = "cuda:0"
mode = sys.argv[1]
E = int(sys.argv[2])
def P(*a):
    print(*a, flush=True)
policy = MlpPolicy(29, 11).to(dev)
policy.pack_weights()
ctr = torch.zeros(1, dtype=torch.int64, device=dev)
low = torch.zeros(11, device=dev); high = torch.ones(11, device=dev)
def bufs():
    return (torch.rand(E, 29, device=dev), torch.empty(E, 11, device=dev), torch.empty(E, 11, device=dev),
            torch.empty(E, device=dev), torch.empty(E, device=dev))
A, B = bufs(), bufs()
def pol(b, gid0=0):
    policy.fused_forward(b[0], None, low, high, b[1], b[2], b[3], b[4], repack=False, rng=(0, ctr, 0, gid0))
torch.cuda.synchronize()
if mode == "single":
    for i in range(2000):
        pol(A)
    torch.cuda.synchronize(); P("single ok")
elif mode == "two":
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for i in range(1000):
        with torch.cuda.stream(s1): pol(A)
        with torch.cuda.stream(s2): pol(B, E)
    torch.cuda.synchronize(); P("two streams ok")
elif mode == "mixed":
    KW = dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
    env = BatchedSmartNanogridEnv(E, device=dev, seed=0, **KW); env.reset()
    acts = env.sample_actions(torch.Generator(device=dev).manual_seed(0)).contiguous()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for i in range(1000):
        with torch.cuda.stream(s1): pol(A)
        with torch.cuda.stream(s2): env.step(acts)
    torch.cuda.synchronize(); P("policy + step concurrently ok")
