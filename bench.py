#!/usr/bin/env python
"""bench.py -- batched env-steps/sec of the nanogrid step on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 2400 --warmup 240
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU restatement of the reference on the host cores

Workload (BASELINE config 4 per GPU): N = 10 charging spots, PV + battery, bounded / sparse, 1 h
steps (24 per episode), fused step + auto-reset + in-kernel Philox schedule sampling,
`--envs` environments per GPU (default 1,048,576: the working set of one step, ~0.4 GB, is larger
than the 126 MB L2, so no flush is needed between iterations).  One "step" = one sng_step launch
over all envs of the rank (the launches of one episode are captured in a CUDA graph and replayed,
`--graph-steps 0` launches them one by one).  Envs shard trivially: rank r owns global env ids
[r*E, (r+1)*E) and there is no collective on the step path (scaling = weak).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE config 4 (the configuration the metric is quoted on), per GPU
    "c4": dict(kw=dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse",
                       time_interval="1h"), envs=1048576,
               name="C4: fused step + auto-reset + in-kernel EV schedule sampling, N=10 spots, PV+battery, 24-step episodes"),
    # BASELINE config 5 (an extension: the reference cannot run 15-minute steps)
    "c5": dict(kw=dict(number_of_chargers=64, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse",
                       time_interval="15min"), envs=262144,
               name="C5: scaled station, N=64 spots, 15-min steps (96 per episode), PV+battery, fused step + auto-reset + sampling"),
}
ENV_KW = WORKLOADS["c4"]["kw"]
METRIC = "batched env-steps/sec"
UNIT = "env-steps/s"


def algorithmic_bytes_per_env_step(n_spots, batt, pv):
    """4A + 4D + 4 + 1 + 8N + 8b + 4 + 12N (float32 build)."""
    a = n_spots + batt
    d = 4 * (1 + pv) + 2 * n_spots + batt
    return 4 * a + 4 * d + 4 + 1 + 8 * n_spots + 8 * batt + 4 + 12 * n_spots


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes(n_envs):
    """dram read+write bytes per launch from the committed ncu capture, if it matches this size."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fp:
            t = json.load(fp)
        if int(t.get("n_envs", -1)) == int(n_envs):
            return float(t["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_throughput(n_envs, seconds, threads):
    """The float64 oracle (a C port of the reference step, oracle/) stepped on the host cores on a
    bounded sample of the workload: `n_envs` envs, full episodes with re-sampling at every episode end."""
    import numpy as np
    from oracle.oracle import OracleBatch
    from smart_nanogrid_gym_b200.config import NanogridConfig
    cfg = NanogridConfig(**ENV_KW)
    ob = OracleBatch(cfg, n_envs, n_threads=threads)
    lo, hi = cfg.action_bounds()
    rng = np.random.default_rng(0)
    a = rng.uniform(lo, hi, size=(n_envs, cfg.act_dim))
    obs = np.empty((n_envs, cfg.obs_dim), np.float32)
    rew = np.empty(n_envs)
    done = np.empty(n_envs, np.uint8)
    ob.sample(0, 0, 0)
    ob.observe()
    episode, steps = 0, 0
    for _ in range(cfg.n_steps):  # warm-up episode
        ob.step_noalloc(a, obs, rew, done)
    t0 = time.perf_counter()
    while True:
        episode += 1
        ob.sample(0, 0, episode)
        ob.observe()
        for _ in range(cfg.n_steps):
            ob.step_noalloc(a, obs, rew, done)
        steps += cfg.n_steps
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return n_envs * steps / el, el, steps


def bind_to_gpu_numa_node(gpu_index):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host
    buffers of the end-to-end leg are allocated on the NUMA node the GPU hangs off (first touch)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return len(allowed)
    except Exception:  # noqa: BLE001 -- best effort; the bench is valid without it
        return 0


def rollout_leg(env, n_steps, dev):
    """BASELINE config 3: policy in the loop.  Reported beside the headline, not as it."""
    import torch
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer
    torch.manual_seed(0)
    policy = MlpPolicy(env.cfg.obs_dim, env.cfg.act_dim).to(dev)
    buf = RolloutBuffer(n_steps, env.num_envs, env.cfg.obs_dim, env.cfg.act_dim, dev)
    obs = env.reset()
    starts = torch.ones(env.num_envs, dtype=torch.uint8, device=dev)
    collect = GraphedRollout(env, policy, buf)
    for _ in range(2):
        obs, starts = collect(obs, starts)
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    ev0.record()
    for _ in range(reps):
        obs, starts = collect(obs, starts)
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    return {"value": env.num_envs * n_steps * reps / (ms * 1e-3), "unit": UNIT, "n_steps": n_steps, "envs": env.num_envs,
            "launch": "one CUDA graph per rollout",
            "policy": "tanh MLP %d-64-64-%d actor + critic%s, actions clipped to the Box, GAE by sng_gae" % (
                env.cfg.obs_dim, env.cfg.act_dim, " (fused sng_policy_forward kernel)" if policy.fused_supported() else " (torch ops)"),
            "mean_step_reward": float(buf.rewards.mean())}


def run_reference_arm(args):
    """--impl reference: the reference's own algorithm on the host CPU.  The reference is pure Python and
    cannot be installed on the GPU box (no gym, read-only tree absent), so this times the oracle port
    (oracle/nanogrid_oracle.c, pinned bit-exactly to the live reference) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = os.cpu_count() or 1
    n_envs = args.ref_envs
    # each "step" is one step of a bounded sample (n_envs envs) of the workload
    import numpy as np
    from oracle.oracle import OracleBatch
    from smart_nanogrid_gym_b200.config import NanogridConfig
    cfg = NanogridConfig(**ENV_KW)
    ob = OracleBatch(cfg, n_envs, n_threads=threads)
    lo, hi = cfg.action_bounds()
    a = np.random.default_rng(0).uniform(lo, hi, size=(n_envs, cfg.act_dim))
    obs = np.empty((n_envs, cfg.obs_dim), np.float32)
    rew = np.empty(n_envs)
    done = np.empty(n_envs, np.uint8)
    episode = 0
    ob.sample(0, 0, 0)
    ob.observe()

    def one_step():
        nonlocal episode
        ob.step_noalloc(a, obs, rew, done)
        if done[0]:
            episode += 1
            ob.sample(0, 0, episode)
            ob.observe()

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    el = time.perf_counter() - t0
    value = n_envs * args.steps / el
    sample = "%d envs x %d steps (fused re-sampling at episode ends), float64 C port of the reference step" % (
        n_envs, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 station (N=10 spots, PV+battery, 24-step episodes), CPU sample of %d envs" % n_envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner to stdout) get stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2400)
    ap.add_argument("--warmup", type=int, default=240)
    ap.add_argument("--graph-steps", type=int, default=24, help="steps per captured CUDA graph (0 = plain launches)")
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (0 = the workload's size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-envs", type=int, default=65536, help="sample size of the CPU reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="wall time of the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--warps", type=int, default=0, help="tuning: warps (blocks of 32 envs) per CTA (0 = auto)")
    ap.add_argument("--generic", action="store_true", help="tuning: use the generic runtime-N kernel")
    ap.add_argument("--bulk", type=int, default=1, help="tuning: row staging: -1 scalar, 0 vector loads/stores, 1 copy-engine loads + vector stores, 3 copy engine both ways")
    ap.add_argument("--host-chunks", type=int, default=0, help="tuning: env chunks of the pipelined host path")
    ap.add_argument("--variant", type=int, default=0, help="tuning: 0 default, 1 persistent pipelined kernel, 2 one lane per env even for large stations")
    ap.add_argument("--ctas", type=int, default=0, help="tuning: cap on resident CTAs per SM (pipelined kernel)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--rollout", type=int, default=0, help="extra leg: PPO rollout collection (tanh 64-64 MLP policy in the "
                    "loop, obs/reward/done written into rollout-buffer slabs, GAE kernel) with this many steps per rollout")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world if world > 1 else 1
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    wl = WORKLOADS[args.workload]
    E = args.envs or wl["envs"]
    env = BatchedSmartNanogridEnv(E, device=dev, seed=0, env_gid0=rank * E, precision="float32", auto_reset=True,
                                  **wl["kw"])
    env.set_tuning(args.warps, int(args.generic), args.bulk, args.host_chunks)
    env.set_pipeline(args.variant, args.ctas)
    cfg = env.cfg
    env.reset()
    # actions: a pre-filled U(low, high) tensor re-read from HBM every step
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = env.sample_actions(g).contiguous()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(min(args.warmup, 24)):
        env.step(actions)
    barrier()

    # one episode's worth of launches captured in a CUDA graph (removes the Python / driver launch cost
    # from the timed region; the kernels and their work are unchanged)
    gsteps = 0
    if args.graph_steps > 0:     # the largest graph length <= --graph-steps that divides K exactly
        gsteps = next((g for g in range(min(args.graph_steps, args.steps), 1, -1) if args.steps % g == 0), 0)
    graph = None
    if gsteps:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                env.step(actions)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(gsteps):
                env.step(actions)

    def run_steps(n):
        if graph is None:
            for _ in range(n):
                env.step(actions)
        else:
            for _ in range(n // gsteps):
                graph.replay()

    run_steps(max(args.warmup - 24, gsteps or 1) // (gsteps or 1) * (gsteps or 1))
    barrier()

    # ---- device-resident throughput: K launches bracketed by CUDA events on the launching stream ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    run_steps(args.steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = args.steps if graph is not None else env.launch_count - launches0
    clock_window = "timed region"
    if ms >= 150.0:
        clocks = sampler.stop()
    else:
        # the timed region is shorter than a few nvidia-smi samples: keep sampling while the same kernel
        # runs on (untimed), so that the clock record is taken under this load
        clock_window = "timed region + %d untimed steps of the same kernel" % 0
        extra = 0
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            run_steps(gsteps or 8)
            torch.cuda.synchronize(dev)
            extra += gsteps or 8
        clock_window = "timed region + %d untimed steps of the same kernel" % extra
        clocks = sampler.stop()
    clocks["window"] = clock_window
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_envs = E * n_gpus
    value = total_envs * args.steps / (ms_max * 1e-3)
    assert env.error_flags() == 0

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D actions, step, D2H obs/reward/done ----
    a_h = actions.cpu().pin_memory()
    o_h = torch.empty(E, cfg.obs_dim, dtype=torch.float32).pin_memory()
    r_h = torch.empty(E, dtype=torch.float32).pin_memory()
    d_h = torch.empty(E, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        env.step_host(a_h, o_h, r_h, d_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        env.step_host(a_h, o_h, r_h, d_h)     # synchronises the stream before returning
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_envs * args.e2e_steps / float(te.item())
    h2d = a_h.numel() * a_h.element_size()
    d2h = o_h.numel() * 4 + r_h.numel() * 4 + d_h.numel()

    # ---- optional episode-return statistics: the only collective, off the step path ----
    from smart_nanogrid_gym_b200.sharding import ReturnStats
    mean_ret = ReturnStats.from_returns(env.last_return).all_reduce(device=dev).mean

    if rank == 0:
        bytes_step = algorithmic_bytes_per_env_step(cfg.n_spots, int(cfg.batt), int(cfg.pv))
        kernel_ms = ms / max(launches, 1)
        achieved = bytes_step * E / (kernel_ms * 1e-3) / 1e9
        peak, how = measured_peak_gbs()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s, %d envs per GPU" % (wl["name"], E),
                       "envs_per_gpu": E, "total_envs": total_envs, "parallelism": "env-sharded x%d, no collective" % n_gpus,
                       "l2": "inputs larger than L2 (%.0f MB touched per step), no flush" % (bytes_step * E / 1e6),
                       "launch": ("CUDA graph of %d step launches, replayed" % gsteps) if graph is not None else "one launch per step",
                       "mean_episode_return": mean_ret},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps, "path": "sng_step_host: pinned host buffers, H2D + step + D2H + sync",
                    "numa_local_cpus": numa_cpus},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(E), "peak_source": how,
                         "algorithmic_bytes_per_env_step": bytes_step, "kernel_ms": kernel_ms},
        }
        if args.rollout > 0:
            line["rollout_collection"] = rollout_leg(env, args.rollout, dev)
        if not args.no_cpu and n_gpus == 1:
            from oracle import oracle as orc
            orc.build()
            threads = os.cpu_count() or 1
            v, el, steps = cpu_port_throughput(args.ref_envs, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d envs x %d steps in %.1f s, float64 C port of the reference step "
                                              "(oracle/), all host threads" % (args.ref_envs, steps, el)}
        emit(line)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
