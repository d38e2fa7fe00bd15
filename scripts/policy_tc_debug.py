import os, sys, torch
sys.path.insert(0, os.getcwd())
from smart_nanogrid_gym_b200.rollout import MlpPolicy
dev="cuda:0"; E=int(sys.argv[1]) if len(sys.argv)>1 else 65536
torch.manual_seed(0)
policy = MlpPolicy(29, 11).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
obs = torch.rand(E, 29, device=dev, generator=g) * 1.5
noise = torch.randn(E, 11, device=dev, generator=g)
low = torch.zeros(11, device=dev); low[-1] = -1; high = torch.ones(11, device=dev)
raw, act = torch.full((E, 11), float("nan"), device=dev), torch.full((E, 11), float("nan"), device=dev)
val, lp = torch.full((E,), float("nan"), device=dev), torch.full((E,), float("nan"), device=dev)
policy.fused_forward(obs, noise, low, high, raw, act, val, lp)
torch.cuda.synchronize()
with torch.no_grad():
    a_ref, v_ref, lp_ref = policy(obs, noise)
bad_v = ((val - v_ref).abs() > 1e-4) | torch.isnan(val)
bad_a = ((raw - a_ref).abs() > 1e-4).any(1) | torch.isnan(raw).any(1)
bad_l = ((lp - lp_ref).abs() > 1e-3) | torch.isnan(lp)
for name, bad in (("val", bad_v), ("raw", bad_a), ("lp", bad_l)):
    idx = bad.nonzero().flatten()
    tiles = torch.unique(idx // 128)
    print(name, "bad envs", idx.numel(), "bad tiles", tiles.numel(), tiles[:20].tolist(), "first idx", idx[:8].tolist())
if bad_a.any():
    i = bad_a.nonzero().flatten()[0].item()
    print("raw", raw[i].tolist(), "\nref", a_ref[i].tolist())
