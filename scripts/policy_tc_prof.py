"""A few launches of the tensor-core policy kernel for ncu (65,536 envs, N = 10 station shapes)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200.rollout import MlpPolicy
dev = "cuda:0"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
policy = MlpPolicy(29, 11).to(dev)
obs = torch.rand(E, 29, device=dev); noise = torch.randn(E, 11, device=dev)
low = torch.zeros(11, device=dev); high = torch.ones(11, device=dev)
raw, act = torch.empty(E, 11, device=dev), torch.empty(E, 11, device=dev)
val, lp = torch.empty(E, device=dev), torch.empty(E, device=dev)
policy.pack_weights()
ctr = torch.zeros(1, dtype=torch.int64, device=dev)
for _ in range(6):      # the rollout's call: exploration noise drawn in the kernel (two io-warp sets)
    policy.fused_forward(obs, None, low, high, raw, act, val, lp, repack=False, rng=(1, ctr, 0, 0))
torch.cuda.synchronize()
print("done")
