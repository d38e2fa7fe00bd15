"""Where a rollout step's time goes: %globaltimer stamps (sng_debug_stamp, a one-thread kernel) between the policy kernel and
the step kernel of a CAPTURED rollout loop (65,536 envs, in-kernel noise).  A span = one kernel + the boundaries around it;
the same loop with two stamps and no kernel in between calibrates what a stamp costs.  Usage: rollout_timeline.py [pdl]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv, _native as nat
from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer
dev = "cuda:0"; E = int(os.environ.get("ENVS", 65536)); n = 24
pdl = len(sys.argv) > 1 and sys.argv[1] == "pdl"
env = BatchedSmartNanogridEnv(E, device=dev, seed=1, number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", time_interval="1h")
policy = MlpPolicy(29, 11).to(dev); policy.pack_weights()
policy.rng_counter = torch.zeros(1, dtype=torch.int64, device=dev)
buf = RolloutBuffer(n, E, 29, 11, dev)
buf.observations[0].copy_(env.reset())
low, high = env.action_low.float(), env.action_high.float()
stamps = torch.zeros(3 * n + 1, dtype=torch.int64, device=dev)
lib = nat.lib()
def stamp(k):
    lib.sng_debug_stamp(C.c_void_p(stamps.data_ptr() + 8 * k), C.c_void_p(torch.cuda.current_stream().cuda_stream))
def rollout():
    stamp(0)
    for s in range(n):
        if pdl: lib.sng_policy_set_launch_mode(1 if s else 0)
        policy.fused_forward(buf.observations[s], None, low, high, buf.raw_actions[s], buf.actions[s], buf.values[s], buf.log_probs[s],
                             repack=False, rng=(1, policy.rng_counter, s, 0))
        lib.sng_policy_set_launch_mode(0)
        stamp(3 * s + 1)
        env.step(buf.actions[s], out=(buf.observations[s + 1], buf.rewards[s], buf.dones[s]))
        stamp(3 * s + 2)
        stamp(3 * s + 3)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    rollout()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    rollout()
pol, stp, cal = [], [], []
for rep in range(30):
    g.replay(); torch.cuda.synchronize()
    t = stamps.cpu().tolist()
    for s in range(2, n):
        pol.append((t[3 * s + 1] - t[3 * s]) / 1e3); stp.append((t[3 * s + 2] - t[3 * s + 1]) / 1e3); cal.append((t[3 * s + 3] - t[3 * s + 2]) / 1e3)
med = lambda x: sorted(x)[len(x) // 2]
print("%s, %d envs: stamp->stamp %.2f us | policy span %.2f us | step span %.2f us | per rollout step (with 3 stamps) %.2f us" % (
    "policy kernel launched programmatically" if pdl else "ordinary launches", E, med(cal), med(pol), med(stp), med(cal) + med(pol) + med(stp)))
