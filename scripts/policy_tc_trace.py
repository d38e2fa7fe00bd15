import os, sys, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from smart_nanogrid_gym_b200.rollout import MlpPolicy
from smart_nanogrid_gym_b200 import _native as nat
dev="cuda:0"; E=65536
policy = MlpPolicy(29, 11).to(dev)
obs = torch.rand(E, 29, device=dev); noise = torch.randn(E, 11, device=dev)
low = torch.zeros(11, device=dev); high = torch.ones(11, device=dev)
raw, act = torch.empty(E, 11, device=dev), torch.empty(E, 11, device=dev)
val, lp = torch.empty(E, device=dev), torch.empty(E, device=dev)
policy.pack_weights()
for _ in range(3): policy.fused_forward(obs, noise, low, high, raw, act, val, lp, repack=False)
tr = torch.zeros(256, dtype=torch.int64, device=dev)
lib = nat.lib(); lib.sng_policy_debug_trace.argtypes=[C.c_void_p]; lib.sng_policy_debug_trace(C.c_void_p(tr.data_ptr()))
policy.fused_forward(obs, noise, low, high, raw, act, val, lp, repack=False)
torch.cuda.synchronize()
t = tr.cpu().tolist()
base = t[255]
print('kernel end at', t[254]-base)
x = [v-base for v in t[:64] if v]
print('abs:', x)
print('deltas:', [b-a for a,b in zip(x, x[1:])])
