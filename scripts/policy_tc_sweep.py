import os, sys, torch
sys.path.insert(0, os.getcwd())
from smart_nanogrid_gym_b200.rollout import MlpPolicy
dev="cuda:0"
policy = MlpPolicy(29, 11).to(dev)
for E in (128, 256, 128*148, 128*296, 128*296+128, 128*592, 65536, 131072, 1048576):
    obs = torch.rand(E, 29, device=dev); noise = torch.randn(E, 11, device=dev)
    low = torch.zeros(11, device=dev); high = torch.ones(11, device=dev)
    raw, act = torch.empty(E, 11, device=dev), torch.empty(E, 11, device=dev)
    val, lp = torch.empty(E, device=dev), torch.empty(E, device=dev)
    for vo in (False, True):
        f = (lambda: policy.fused_forward(obs, None, None, None, None, None, val, None, repack=False)) if vo else (lambda: policy.fused_forward(obs, noise, low, high, raw, act, val, lp, repack=False))
        policy.pack_weights()
        for _ in range(5): f()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ev0.record()
        for _ in range(100): f()
        ev1.record(); torch.cuda.synchronize()
        print("E=%d value_only=%d: %.2f us" % (E, vo, 1e3*ev0.elapsed_time(ev1)/100), flush=True)
