#!/usr/bin/env python
"""Numerics + timing of the tensor-core policy kernel (sng_policy_forward_packed) against the FP32 torch modules
and the CUDA-core kernel (sng_policy_forward).  Run on a B200:  python scripts/policy_tc_check.py [E]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smart_nanogrid_gym_b200.rollout import MlpPolicy  # noqa: E402


def run(obs_dim, act_dim, E, seed=0, time_it=False):
    dev = "cuda:0"
    torch.manual_seed(seed)
    policy = MlpPolicy(obs_dim, act_dim).to(dev)
    with torch.no_grad():
        policy.log_std.copy_(torch.linspace(-1.0, 0.5, act_dim))
    g = torch.Generator(device=dev).manual_seed(1)
    obs = torch.rand(E, obs_dim, device=dev, generator=g) * 1.5
    noise = torch.randn(E, act_dim, device=dev, generator=g)
    low = torch.zeros(act_dim, device=dev)
    low[-1] = -1.0
    high = torch.ones(act_dim, device=dev)
    out = {}
    for kind in ("tc", "cc"):
        if kind == "cc" and not policy.cuda_core_supported():
            continue
        raw, act = torch.full((E, act_dim), float("nan"), device=dev), torch.full((E, act_dim), float("nan"), device=dev)
        val, lp = torch.full((E,), float("nan"), device=dev), torch.full((E,), float("nan"), device=dev)
        policy.fused_forward(obs, noise, low, high, raw, act, val, lp, cuda_cores=kind == "cc")
        torch.cuda.synchronize()
        out[kind] = (raw, act, val, lp)
    with torch.no_grad():
        a_ref, v_ref, lp_ref = policy(obs.double().float(), noise)
        pd = MlpPolicy(obs_dim, act_dim).to(dev).double()
        pd.load_state_dict({k: v.double() for k, v in policy.state_dict().items()})
        a64, v64, _ = pd(obs.double(), noise.double())
    for kind, (raw, act, val, lp) in out.items():
        print("D=%d A=%d E=%d %s: max|raw-fp32|=%.3e max|val-fp32|=%.3e max|lp-fp32|=%.3e | vs fp64: raw %.3e val %.3e (torch fp32 vs fp64: raw %.3e val %.3e)"
              % (obs_dim, act_dim, E, kind, (raw - a_ref).abs().max().item(), (val - v_ref).abs().max().item(),
                 (lp - lp_ref).abs().max().item(), (raw.double() - a64).abs().max().item(), (val.double() - v64).abs().max().item(),
                 (a_ref.double() - a64).abs().max().item(), (v_ref.double() - v64).abs().max().item()), flush=True)
    raw, act, val, lp = out["tc"]
    ok = (torch.allclose(raw, a_ref, rtol=1e-4, atol=2e-5) and torch.allclose(val, v_ref, rtol=1e-4, atol=2e-5) and
          torch.allclose(lp, lp_ref, rtol=1e-5, atol=1e-4) and
          torch.allclose(act, torch.minimum(torch.maximum(a_ref, low), high), rtol=1e-4, atol=2e-5))
    # value-only and deterministic modes
    val2 = torch.empty(E, device=dev)
    policy.fused_forward(obs, None, None, None, None, None, val2, None)
    ok = ok and torch.equal(val2, val)
    print("  -> %s" % ("OK" if ok else "MISMATCH"), flush=True)
    if time_it:
        for kind in ("tc", "cc"):
            if kind == "cc" and not policy.cuda_core_supported():
                continue
            raw, act, val, lp = out[kind]
            f = lambda: policy.fused_forward(obs, noise, low, high, raw, act, val, lp, repack=False, cuda_cores=kind == "cc")  # noqa: E731
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(5):
                    f()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()            # replayed from a graph: the host's launch cost stays out of the number
            with torch.cuda.graph(graph):
                for _ in range(20):
                    f()
            graph.replay()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(10):
                graph.replay()
            ev1.record()
            torch.cuda.synchronize()
            us = 1e3 * ev0.elapsed_time(ev1) / 200
            flops = 2.0 * 2 * (obs_dim * 64 + 64 * 64) + 2.0 * 64 * (act_dim + 1)
            print("  %s: %.2f us per %d-env call (%.1f useful TFLOP/s)" % (kind, us, E, flops * E / us / 1e6), flush=True)
    return ok


if __name__ == "__main__":
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    ok = True
    ok &= run(29, 11, 128 * 3)              # whole tiles only
    ok &= run(29, 11, 5007)                 # ragged
    ok &= run(17, 5, 5007)
    ok &= run(25, 9, 4096 + 3)
    ok &= run(13, 3, 1000)                  # a shape only the tensor-core kernel takes
    ok &= run(29, 11, E, time_it=True)
    ok &= run(17, 5, E, time_it=True)
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
