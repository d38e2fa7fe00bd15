#!/usr/bin/env python
"""bench.py -- batched env-steps/sec of the nanogrid step on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 2400 --warmup 240
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU restatement of the reference on the host cores

Headline workload (BASELINE config 4 per GPU): N = 10 charging spots, PV + battery, bounded / sparse, 1 h
steps (24 per episode), fused step + auto-reset + in-kernel Philox schedule sampling,
`--envs` environments per GPU (default 1,048,576: the working set of one step, ~0.3 GB, is larger
than the 126 MB L2, so no flush is needed between iterations).  One "step" = one sng_step launch
over all envs of the rank (the launches of one episode are captured in a CUDA graph and replayed,
`--graph-steps 0` launches them one by one).  Envs shard trivially: rank r owns global env ids
[r*E, (r+1)*E) and there is no collective on the step path (scaling = weak).

Beside the headline the same JSON line carries `legs`: every other BASELINE configuration measured in the same
run (untimed for the headline) -- `c4_strong` (1,048,576 envs SPLIT over the ranks: 131,072 per GPU at 8),
`c5` (64 spots, 96 steps, 262,144 envs per GPU), `c3` / `c3_sb3` (PPO rollout collection at 65,536 envs, policy in the
loop: fresh-init N = 10 network / the reference's shipped checkpoint), `c2` (4,096 envs, per-step launches and the
multi-step sng_rollout launch), `rollout_kernel` (sng_rollout with pre-supplied actions at 65,536 / 131,072 envs) and
`generic` (stations off the reference's observation shape: 5-step horizon, two-day PV, no PV) -- each with value,
ms_per_step, roofline (measured DRAM traffic from profiles/roofline_traffic.json) and clocks.  `e2e` carries the box's
concurrent pinned-memcpy ceiling for the same bytes; `cpu_baseline` (the C port) and `cpu_baseline_live` (the live
Python reference from baseline/_ref) are timed on the host cores at N = 1.
"""
import argparse
import json
import math
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
    # the shipped SB3 checkpoint's station (solvers/RL/models/PPO-b-pv-bounded-sparse-4ch-1h)
    "n4": dict(kw=dict(number_of_chargers=4, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse",
                       time_interval="1h"), envs=65536,
               name="N=4 spots, PV+battery, 24-step episodes (the station of the reference's shipped PPO checkpoint)"),
    # SURVEY 8f row 4 (generalised station): configurations that run the generic runtime-N kernel
    "c4_h5": dict(kw=dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse",
                          time_interval="1h", hours_ahead=5), envs=1048576,
                  name="C4 station with a 5-step forecast horizon (generic kernel)"),
    "c4_pv2d": dict(kw=dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse",
                            time_interval="1h", number_of_days_to_predict=2, cycle_pv_days=True), envs=1048576,
                    name="C4 station with two-day PV (episode k reads day k % 2; generic kernel)"),
    "c4_nopv": dict(kw=dict(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse",
                            time_interval="1h", pv_system_available_in_model=False), envs=1048576,
                    name="C4 station without PV (generic kernel)"),
}
ENV_KW = WORKLOADS["c4"]["kw"]
METRIC = "batched env-steps/sec"
UNIT = "env-steps/s"
L2_BYTES = 126e6


def algorithmic_bytes_per_env_step(n_spots, batt, pv, horizon=3):
    """4A + 4D + 4 + 1 + 8N + 8b + 4 + 12N (float32 build)."""
    a = n_spots + batt
    d = (1 + horizon) * (1 + pv) + 2 * n_spots + batt
    return 4 * a + 4 * d + 4 + 1 + 8 * n_spots + 8 * batt + 4 + 12 * n_spots


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes(workload, n_envs):
    """dram read+write bytes per launch of the step kernel for this workload and size, from the ncu captures that
    scripts/ncu_traffic.sh takes on a B200 and scripts/ncu_traffic_update.py folds into profiles/roofline_traffic.json
    (steady state: launch >= 30 of the run is profiled, so dirty lines of the previous launch are counted)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fp:
            t = json.load(fp)
        e = t.get("entries", {}).get("%s:%d" % (workload, int(n_envs)))
        if e:
            return float(e["dram_bytes_per_launch"])
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


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port and the live Python reference): the only code here that may touch oracle/
# ------------------------------------------------------------------------------------------------------------
class CpuPort:
    """The float64 oracle (a C port of the reference step, oracle/) stepping `n_envs` envs of the C4 station on
    `threads` host threads: full episodes with re-sampling at every episode end."""

    def __init__(self, n_envs, threads):
        import numpy as np
        from oracle.oracle import OracleBatch
        from smart_nanogrid_gym_b200.config import NanogridConfig
        self.np = np
        self.cfg = NanogridConfig(**ENV_KW)
        self.n = n_envs
        self.ob = OracleBatch(self.cfg, n_envs, n_threads=threads)
        lo, hi = self.cfg.action_bounds()
        self.a = np.random.default_rng(0).uniform(lo, hi, size=(n_envs, self.cfg.act_dim))
        self.obs = np.empty((n_envs, self.cfg.obs_dim), np.float32)
        self.rew = np.empty(n_envs)
        self.done = np.empty(n_envs, np.uint8)
        self.episode = 0
        self.ob.sample(0, 0, 0)
        self.ob.observe()

    def step(self):
        self.ob.step_noalloc(self.a, self.obs, self.rew, self.done)
        if self.done[0]:
            self.episode += 1
            self.ob.sample(0, 0, self.episode)
            self.ob.observe()


def cpu_port_throughput(n_envs, seconds, threads):
    port = CpuPort(n_envs, threads)
    T = port.cfg.n_steps
    for _ in range(T):  # warm-up episode
        port.step()
    steps, t0 = 0, time.perf_counter()
    while True:
        for _ in range(T):
            port.step()
        steps += T
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return n_envs * steps / el, el, steps


def pick_ref_envs(requested, want):
    """Env count of the CPU arm: the GPU arm's own size when the oracle's dense float64 arrays (~8.1 KB per env at
    N = 10) fit comfortably in host memory, else a bounded sample."""
    if requested > 0:
        return requested
    try:
        import psutil
        if psutil.virtual_memory().available > 3 * want * 8200:
            return want
    except Exception:  # noqa: BLE001
        pass
    return 65536


def _live_worker(n_spots, seconds, stub_io, q):
    """One process of the live-reference leg: the UNMODIFIED Python env (baseline/_ref) under the import shims."""
    import numpy as np
    from oracle import ref_loader as rl
    rl._loaded.clear()
    rl.load_reference(stub_io=stub_io)
    env = rl.make_ref_env(number_of_chargers=n_spots)
    rl.seed_reference(os.getpid())
    lo, hi = env.action_space.low, env.action_space.high
    rng = np.random.default_rng(os.getpid())
    env.reset()
    for _ in range(24):
        env.step(rng.uniform(lo, hi).astype(np.float32))
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        env.reset()
        for _ in range(24):
            env.step(rng.uniform(lo, hi).astype(np.float32))
        steps += 24
    q.put((steps, time.perf_counter() - t0))


def live_reference_throughput(seconds):
    """north_star / BASELINE.md section 4: the reference Python env itself stepped on this box's host cores, one
    process per core, uniform random float32 actions, reset() included -- as shipped (per-episode JSON dumps) and
    I/O-stubbed.  Needs the installed reference under baseline/_ref (see __graft_entry__.build)."""
    import multiprocessing as mp
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "smart_nanogrid_gym")):
        return {"unavailable": "baseline/_ref not installed (run __graft_entry__.build() where /root/reference exists)"}
    os.environ["SNG_REFERENCE_ROOT"] = ref_root
    os.environ.setdefault("PYTHONBREAKPOINT", "0")
    procs = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    out = {"kind": "reference (live Python env from baseline/_ref under the gym/path shims of oracle/ref_loader.py)",
           "processes": procs, "unit": UNIT, "station": "C4 station: N=10 spots, PV+battery, 24-step episodes, reset included"}
    for tag, stub, secs in (("io_stubbed", True, seconds), ("as_shipped", False, max(seconds / 2, 5.0))):
        q = ctx.Queue()
        ps = [ctx.Process(target=_live_worker, args=(10, secs, stub, q)) for _ in range(procs)]
        for p in ps:
            p.start()
        res, deadline = [], time.perf_counter() + secs + 120
        while len(res) < procs and time.perf_counter() < deadline:
            try:
                res.append(q.get(timeout=1.0))
            except Exception:  # noqa: BLE001 -- queue.Empty: stop early if the workers died
                if not any(p.is_alive() for p in ps) and q.empty():
                    break
        for p in ps:
            p.join(timeout=30)
        if len(res) != procs:
            out[tag] = {"unavailable": "%d of %d workers reported" % (len(res), procs)}
            continue
        out[tag] = {"value": sum(s / t for s, t in res), "seconds": secs, "env_steps": sum(s for s, _ in res)}
    if "value" in out.get("io_stubbed", {}):
        out["value"] = out["io_stubbed"]["value"]
    return out


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


def workload_config(wl_key, envs_per_gpu, n_gpus, bytes_step):
    """`config` of the JSON line: names the workload only, identical for the b200 and the reference arm."""
    wl = WORKLOADS[wl_key]
    return {"workload": "%s, %d envs per GPU" % (wl["name"], envs_per_gpu),
            "envs_per_gpu": envs_per_gpu, "total_envs": envs_per_gpu * n_gpus,
            "parallelism": "env-sharded x%d, no collective" % n_gpus,
            "l2": "inputs larger than L2 (%.0f MB touched per step), no flush" % (bytes_step * envs_per_gpu / 1e6)}


def run_reference_arm(args):
    """--impl reference: the reference's own algorithm on the host CPU.  The reference is pure Python (nothing to
    compile into oracle/_ref), so this times the oracle port (oracle/nanogrid_oracle.c, pinned bit-exactly to the
    live reference) on all host threads, on the b200 arm's own configuration when it fits in host memory.  The K
    timed steps are repeated until at least --ref-seconds have been measured (K steps of a 65,536-env sample last
    under 0.1 s)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = os.cpu_count() or 1
    wl = WORKLOADS["c4"]
    E = args.envs or wl["envs"]
    n_envs = pick_ref_envs(args.ref_envs, E)
    port = CpuPort(n_envs, threads)
    for _ in range(args.warmup):
        port.step()
    reps, el = 0, 0.0
    while True:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            port.step()
        el += time.perf_counter() - t0
        reps += 1
        if el >= args.ref_seconds or reps >= 10000:
            break
    value = n_envs * args.steps * reps / el
    sample = ("%d envs (%s) x %d steps x %d repeats = %.1f s, fused re-sampling at episode ends, float64 C port of the "
              "reference step on %d host threads" % (n_envs, "the b200 arm's size" if n_envs == E else "a bounded sample of the workload",
                                                     args.steps, reps, el, threads))
    bytes_step = algorithmic_bytes_per_env_step(10, 1, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / (args.steps * reps) * (E / n_envs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config("c4", E, max(args.gpus, 1), bytes_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "repeats": reps, "timed_seconds": el, "sample_envs": n_envs,
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


# ------------------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------------------
class Ctx:
    """Per-process GPU context of the bench: device, ranks, barrier, max-over-ranks reduction."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.n_gpus = self.world if self.world > 1 else 1
        self.numa_cpus = bind_to_gpu_numa_node(self.local_rank) if self.world > 1 else 0
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.peak, self.peak_source = measured_peak_gbs()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


class StepLoop:
    """`env.step(actions)` launches, optionally captured `gsteps` at a time in a CUDA graph and replayed (removes the
    Python / driver launch cost from the timed region; the kernels and their work are unchanged)."""

    def __init__(self, ctx, env, actions, graph_steps, total_steps=None):
        torch = ctx.torch
        self.ctx, self.env, self.actions = ctx, env, actions
        self.gsteps = 0
        if graph_steps > 0:
            if total_steps is None:
                self.gsteps = graph_steps
            else:   # the largest graph length <= graph_steps that divides the timed step count exactly
                self.gsteps = next((g for g in range(min(graph_steps, total_steps), 1, -1) if total_steps % g == 0), 0)
        self.graph = None
        if self.gsteps:
            side = torch.cuda.Stream(device=ctx.dev)
            side.wait_stream(torch.cuda.current_stream(ctx.dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    env.step(actions)
            torch.cuda.current_stream(ctx.dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                for _ in range(self.gsteps):
                    env.step(actions)

    def run(self, n):
        if self.graph is None:
            for _ in range(n):
                self.env.step(self.actions)
        else:
            for _ in range(n // self.gsteps):
                self.graph.replay()

    def round_up(self, n):
        g = self.gsteps or 1
        return max((n + g - 1) // g, 1) * g


def timed_ms(ctx, fn):
    """fn() bracketed by CUDA events on the launching stream, a barrier + device synchronise on both sides."""
    torch = ctx.torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    fn()
    ev1.record()
    ctx.barrier()
    return ev0.elapsed_time(ev1)


def launch_floor_us(ctx, n=2400):
    """Device time per dependent kernel launch of an EMPTY kernel replayed from a CUDA graph: the floor under any
    per-step launch (kernel-to-kernel dependency latency), the honest denominator for batches that live in L2."""
    import ctypes as C
    from smart_nanogrid_gym_b200 import _native as nat
    torch = ctx.torch
    lib = nat.lib()

    def launch():
        nat.check(lib.sng_null_launch(C.c_void_p(torch.cuda.current_stream(ctx.dev).cuda_stream)))

    side = torch.cuda.Stream(device=ctx.dev)
    side.wait_stream(torch.cuda.current_stream(ctx.dev))
    with torch.cuda.stream(side):
        launch()
    torch.cuda.current_stream(ctx.dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(24):
            launch()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(ctx.dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n // 24):
        graph.replay()
    ev1.record()
    torch.cuda.synchronize(ctx.dev)
    return 1e3 * ev0.elapsed_time(ev1) / (n // 24 * 24)


def roofline_of(ctx, wl_key, cfg, E, kernel_ms, floor_us=None):
    bytes_step = algorithmic_bytes_per_env_step(cfg.n_spots, int(cfg.batt), int(cfg.pv), cfg.hours_ahead)
    achieved = bytes_step * E / (kernel_ms * 1e-3) / 1e9
    traffic = ncu_traffic_bytes(wl_key, E)
    r = {"bound": "hbm", "achieved": achieved, "peak": ctx.peak, "unit": "GB/s", "frac": achieved / ctx.peak,
         "traffic": traffic, "peak_source": ctx.peak_source, "algorithmic_bytes_per_env_step": bytes_step,
         "kernel_ms": kernel_ms}
    if traffic:
        r["traffic_frac"] = traffic / (kernel_ms * 1e-3) / 1e9 / ctx.peak       # real DRAM throughput over the copy peak
    if bytes_step * E < L2_BYTES:
        r["l2_resident"] = True
        r["note"] = ("the step's working set (%.0f MB) fits in the 126 MB L2: the HBM roofline is not the bound here, "
                     "kernel-to-kernel launch latency is" % (bytes_step * E / 1e6))
        if floor_us:
            r["launch_floor_us"] = floor_us
            r["launch_floor_frac"] = floor_us / (kernel_ms * 1e3)
    return r


def make_env(ctx, wl_key, E, gid0, args=None):
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    env = BatchedSmartNanogridEnv(E, device=ctx.dev, seed=0, env_gid0=gid0, precision="float32", auto_reset=True,
                                  **WORKLOADS[wl_key]["kw"])
    if args is not None:
        env.set_tuning(args.warps, int(args.generic), args.bulk, args.host_chunks)
        env.set_pipeline(args.variant, args.ctas)
    env.reset()
    return env


PDL_ARG = [-1]       # --pdl, set by main()
ROLLOUT_PDL = [None]  # --rollout-pdl, set by main(): programmatic dependent launch inside the rollout loop (collect_rollout's pdl)


def launch_mode_for(cfg, E):
    """Programmatic dependent launch for the per-step launches of small batches (measured on a B200, C4 station:
    4,096 envs 4.73 -> 4.44 us per step, 65,536 envs 6.58 -> 6.34; 131,072 envs 9.99 -> 10.24; 1,048,576: no change)."""
    if PDL_ARG[0] >= 0:
        return PDL_ARG[0]
    return 1 if algorithmic_bytes_per_env_step(cfg.n_spots, int(cfg.batt), int(cfg.pv), cfg.hours_ahead) * E < 32e6 else 0


def step_leg(ctx, wl_key, E_local, gid0, floor_us, min_ms=200.0, graph_steps=24):
    """One leg: the step kernel over `E_local` envs of this rank (global ids from gid0), per-step launches replayed
    from a CUDA graph, timed for >= min_ms with clocks sampled."""
    torch = ctx.torch
    env = make_env(ctx, wl_key, E_local, gid0)
    g = torch.Generator(device=ctx.dev).manual_seed(1234 + ctx.rank)
    actions = env.sample_actions(g).contiguous()
    pdl = launch_mode_for(env.cfg, E_local)
    env.set_launch_mode(pdl)
    loop = StepLoop(ctx, env, actions, graph_steps)
    loop.run(loop.round_up(48))
    ms_probe = timed_ms(ctx, lambda: loop.run(loop.round_up(48)))
    ms_probe = ctx.max_over_ranks(ms_probe)
    n = loop.round_up(int(min(max(min_ms / max(ms_probe / loop.round_up(48), 1e-4), 48), 200000)))
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ms = timed_ms(ctx, lambda: loop.run(n))
    clocks = sampler.stop()
    ms_max = ctx.max_over_ranks(ms)
    total_envs = int(ctx.sum_over_ranks(E_local))
    assert env.error_flags() == 0
    out = {"value": total_envs * n / (ms_max * 1e-3), "unit": UNIT, "ms_per_step": ms_max / n, "steps": n,
           "envs_per_gpu": E_local, "total_envs": total_envs, "gpu_launches": n,
           "launch": "CUDA graph of %d sng_step launches%s, replayed" % (loop.gsteps, " (programmatic dependent launch)" if pdl else ""),
           "roofline": roofline_of(ctx, wl_key, env.cfg, E_local, ms / n, floor_us), "clocks": clocks}
    env.close()
    return out


def rollout_kernel_leg(ctx, wl_key, E, floor_us, n_steps=24, min_ms=150.0):
    """sng_rollout: `n_steps` consecutive steps in ONE launch with pre-supplied actions [n_steps][E][A]: no per-step
    launch latency at all (the batch sizes that live in L2)."""
    torch = ctx.torch
    env = make_env(ctx, wl_key, E, ctx.rank * E)
    cfg = env.cfg
    g = torch.Generator(device=ctx.dev).manual_seed(77 + ctx.rank)
    acts = env.random_actions(77, 0, n_steps)      # sng_sample_actions: the random policy, keyed by global env id and step
    obs = torch.empty(n_steps, E, cfg.obs_dim, device=ctx.dev)
    rew = torch.empty(n_steps, E, device=ctx.dev)
    done = torch.empty(n_steps, E, device=ctx.dev, dtype=torch.uint8)
    for _ in range(3):
        env.rollout(acts, obs, rew, done)
    ms_probe = ctx.max_over_ranks(timed_ms(ctx, lambda: env.rollout(acts, obs, rew, done)))
    reps = int(min(max(min_ms / max(ms_probe, 1e-3), 3), 20000))

    def run():
        for _ in range(reps):
            env.rollout(acts, obs, rew, done)

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ms = timed_ms(ctx, run)
    clocks = sampler.stop()
    ms_max = ctx.max_over_ranks(ms)
    assert env.error_flags() == 0
    steps = reps * n_steps
    out = {"value": E * ctx.n_gpus * steps / (ms_max * 1e-3), "unit": UNIT, "ms_per_step": ms_max / steps, "steps": steps,
           "envs_per_gpu": E, "total_envs": E * ctx.n_gpus, "gpu_launches": reps,
           "launch": "sng_rollout: %d steps per kernel launch, actions pre-supplied, obs / reward / done slabs written per step" % n_steps,
           "roofline": roofline_of(ctx, wl_key, cfg, E, ms / steps, floor_us), "clocks": clocks}
    env.close()
    return out


def load_shipped_policy(ctx, obs_dim, act_dim):
    """The reference's shipped PPO checkpoint (tests/golden/sb3_ppo_4ch_policy.npz, extracted from
    solvers/RL/models/PPO-b-pv-bounded-sparse-4ch-1h/999600.zip by tests/golden/generate_golden.py)."""
    from smart_nanogrid_gym_b200.rollout import MlpPolicy
    path = os.path.join(ROOT, "tests", "golden", "sb3_ppo_4ch_policy.npz")
    if not os.path.exists(path) or not hasattr(MlpPolicy, "from_sb3_state_dict"):
        return None
    import numpy as np
    z = np.load(path)
    sd = {k: ctx.torch.tensor(z[k]) for k in z.files}
    pol = MlpPolicy.from_sb3_state_dict(sd)
    if (pol.pi[0].in_features, pol.action_net.out_features) != (obs_dim, act_dim):
        return None
    return pol.to(ctx.dev)


def pattern_ceiling_leg(ctx, E):
    """What the step kernel's ACCESS PATTERN alone costs on this machine: sng_debug_traffic_skeleton performs the same loads
    and stores from the same launch geometry and occupancy, without the arithmetic.  The copy-bandwidth roofline assumes
    two perfectly sequential streams; the step moves five read streams and five write streams in 128-byte to 3.7 KB
    granules per warp.  Reported beside the roofline (rank 0's GPU), not instead of it."""
    import ctypes as C
    torch = ctx.torch
    env = make_env(ctx, "c4", E, ctx.rank * E)
    env.reset()
    g = torch.Generator(device=ctx.dev).manual_seed(7)
    env.step(env.sample_actions(g))
    lib, h = env._lib, env._h
    stream = C.c_void_p(torch.cuda.current_stream(ctx.dev).cuda_stream)

    def run(n, variant=0):
        for _ in range(n):
            rc = lib.sng_debug_traffic_skeleton(h, variant, stream)
            if rc != 0:
                raise RuntimeError("sng_debug_traffic_skeleton: %d" % rc)
    run(30)
    n = 300
    ms = timed_ms(ctx, lambda: run(n)) / n
    what_if = {}
    for name, variant in (("planes_interleaved_per_spot_r2_layout", 1), ("loads_only", 2), ("obs_rows_by_copy_engine", 4), ("stores_only", 3)):
        run(10, variant)
        what_if[name] = timed_ms(ctx, lambda: run(n, variant)) / n
    traffic = ncu_traffic_bytes("c4", E)
    out = {"ms_per_launch": ms, "launches": n, "envs_per_gpu": E,
           "what": "the step kernel's loads and stores without its arithmetic (same geometry, occupancy and byte counts)",
           "what_if_ms": what_if}
    if traffic:
        out["gbs_real_traffic"] = traffic / (ms * 1e-3) / 1e9
        out["frac_of_copy_peak"] = out["gbs_real_traffic"] / ctx.peak
    env.close()
    return out


def rollout_leg(ctx, wl_key, E, n_steps, shipped=False, fuse_step=False):
    """BASELINE config 3: PPO rollout collection, policy in the loop.  Reported beside the headline, not as it."""
    torch = ctx.torch
    from smart_nanogrid_gym_b200.rollout import GraphedRollout, MlpPolicy, RolloutBuffer
    env = make_env(ctx, wl_key, E, ctx.rank * E)
    torch.manual_seed(0)
    policy = load_shipped_policy(ctx, env.cfg.obs_dim, env.cfg.act_dim) if shipped else None
    if shipped and policy is None:
        env.close()
        return {"unavailable": "shipped checkpoint fixture not found"}
    if policy is None:
        policy = MlpPolicy(env.cfg.obs_dim, env.cfg.act_dim).to(ctx.dev)
    buf = RolloutBuffer(n_steps, env.num_envs, env.cfg.obs_dim, env.cfg.act_dim, ctx.dev)
    obs = env.reset()
    starts = torch.ones(env.num_envs, dtype=torch.uint8, device=ctx.dev)
    rpdl = (ROLLOUT_PDL[0] or "policy") if policy.fused_supported() else None   # measured: policy 29.5, off 30.5, both 31.6 us per step
    rpdl = None if rpdl == "off" else rpdl
    fuse_step = bool(fuse_step and policy.fused_supported() and env.supports_policy_step())
    collect = GraphedRollout(env, policy, buf, pdl=rpdl or False, fuse_step=fuse_step)
    state = [obs, starts]

    def run(reps):
        for _ in range(reps):
            state[0], state[1] = collect(state[0], state[1])

    run(2)
    ms_probe = ctx.max_over_ranks(timed_ms(ctx, lambda: run(2))) / 2
    reps = int(min(max(250.0 / max(ms_probe, 1e-3), 5), 2000))
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ms = timed_ms(ctx, lambda: run(reps))
    clocks = sampler.stop()
    ms_max = ctx.max_over_ranks(ms)
    steps = reps * n_steps
    # policy-forward share: the fused kernel alone on one observation slab
    t_pol = None
    if policy.fused_supported():
        o = buf.observations[0]
        noise = torch.randn(E, env.cfg.act_dim, device=ctx.dev)
        low, high = env.action_low.float(), env.action_high.float()

        policy.pack_weights()

        def pol():
            for _ in range(50):
                policy.fused_forward(o, noise, low, high, buf.raw_actions[0], buf.actions[0], buf.values[0], buf.log_probs[0], repack=False)
        pol()
        t_pol = timed_ms(ctx, pol) / 50
    flops = 2.0 * 2 * (env.cfg.obs_dim * 64 + 64 * 64) + 2.0 * 64 * (env.cfg.act_dim + 1)    # per env: actor + critic + heads
    out = {"value": E * ctx.n_gpus * steps / (ms_max * 1e-3), "unit": UNIT, "ms_per_step": ms_max / steps, "steps": steps,
           "n_steps_per_rollout": n_steps, "envs_per_gpu": E, "total_envs": E * ctx.n_gpus,
           "launch": "one CUDA graph per rollout (n_steps x %s + bootstrap value + GAE)%s" % (
               "[ONE kernel: policy forward + env step, sng_policy_step]" if fuse_step else "[policy kernel, step kernel]",
               ", programmatic dependent launch: %s" % rpdl if rpdl else ""),
           "policy": "%s tanh MLP %d-64-64-%d actor + critic%s, actions sampled and clipped to the Box, GAE by sng_gae" % (
               "the reference's shipped SB3 PPO checkpoint:" if shipped else "fresh-init (torch seed 0)",
               env.cfg.obs_dim, env.cfg.act_dim, " (fused sng_policy_forward kernel: %s)" % policy.fused_kind()
               if policy.fused_supported() and hasattr(policy, "fused_kind") else ""),
           "mean_step_reward": float(buf.rewards.mean()), "clocks": clocks}
    if t_pol is not None:
        out["policy_forward_ms"] = t_pol
        out["policy_forward_tflops"] = flops * E / (t_pol * 1e-3) / 1e12
    env.close()
    return out


def rollout_sharded_leg(ctx, wl_key, E, n_steps, shards=2):
    """BASELINE config 3 with the batch cut into `shards` env shards collected side by side (ShardedGraphedRollout):
    the policy kernel of one shard overlaps the step kernel of the other."""
    torch = ctx.torch
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    from smart_nanogrid_gym_b200.rollout import MlpPolicy, RolloutBuffer, ShardedGraphedRollout
    per = E // shards
    envs = [BatchedSmartNanogridEnv(per, device=ctx.dev, seed=0, env_gid0=ctx.rank * E + k * per, precision="float32",
                                    auto_reset=True, **WORKLOADS[wl_key]["kw"]) for k in range(shards)]
    cfg = envs[0].cfg
    torch.manual_seed(0)
    policy = MlpPolicy(cfg.obs_dim, cfg.act_dim).to(ctx.dev)
    bufs = [RolloutBuffer(n_steps, per, cfg.obs_dim, cfg.act_dim, ctx.dev) for _ in range(shards)]
    obs = [e.reset() for e in envs]
    starts = [torch.ones(per, dtype=torch.uint8, device=ctx.dev) for _ in range(shards)]
    collect = ShardedGraphedRollout(envs, policy, bufs)
    state = [obs, starts]

    def run(reps):
        for _ in range(reps):
            state[0], state[1] = collect(state[0], state[1])

    run(2)
    ms_probe = ctx.max_over_ranks(timed_ms(ctx, lambda: run(2))) / 2
    reps = int(min(max(250.0 / max(ms_probe, 1e-3), 5), 2000))
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ms = timed_ms(ctx, lambda: run(reps))
    clocks = sampler.stop()
    ms_max = ctx.max_over_ranks(ms)
    steps = reps * n_steps
    out = {"value": per * shards * ctx.n_gpus * steps / (ms_max * 1e-3), "unit": UNIT, "ms_per_step": ms_max / steps, "steps": steps,
           "n_steps_per_rollout": n_steps, "envs_per_gpu": per * shards, "total_envs": per * shards * ctx.n_gpus, "shards_per_gpu": shards,
           "launch": "one CUDA graph per rollout with %d parallel branches (one per env shard): n_steps x [policy kernel, step kernel] "
                     "+ bootstrap value + GAE each" % shards,
           "policy": "fresh-init (torch seed 0) tanh MLP %d-64-64-%d actor + critic (%s), exploration noise drawn in the kernel, "
                     "GAE by sng_gae" % (cfg.obs_dim, cfg.act_dim, policy.fused_kind()),
           "mean_step_reward": float(torch.stack([b.rewards.mean() for b in bufs]).mean()), "clocks": clocks}
    for e in envs:
        e.close()
    return out


def e2e_leg(ctx, env, actions, steps):
    """End to end through the C ABI with HOST buffers (pinned): H2D actions, step, D2H obs/reward/done, sync; and the
    PCIe ceiling beside it: plain pinned cudaMemcpyAsync of the same byte counts in both directions at once, all
    ranks concurrently."""
    torch = ctx.torch
    cfg, E = env.cfg, env.num_envs
    a_h = actions.cpu().pin_memory()
    o_h = torch.empty(E, cfg.obs_dim, dtype=torch.float32).pin_memory()
    r_h = torch.empty(E, dtype=torch.float32).pin_memory()
    d_h = torch.empty(E, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        env.step_host(a_h, o_h, r_h, d_h)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        env.step_host(a_h, o_h, r_h, d_h)     # synchronises the stream before returning
    torch.cuda.synchronize(ctx.dev)
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    h2d = a_h.numel() * a_h.element_size()
    d2h = o_h.numel() * 4 + r_h.numel() * 4 + d_h.numel()
    # ---- the memcpy ceiling of this box for the same bytes ----
    s_in, s_out = torch.cuda.Stream(device=ctx.dev), torch.cuda.Stream(device=ctx.dev)
    a_d, o_d, r_d, d_d = env.actions, env.obs, env.reward, env.done

    def copies():
        with torch.cuda.stream(s_in):
            a_d.copy_(a_h, non_blocking=True)
        with torch.cuda.stream(s_out):
            o_h.copy_(o_d, non_blocking=True)
            r_h.copy_(r_d, non_blocking=True)
            d_h.copy_(d_d, non_blocking=True)

    for _ in range(3):
        copies()
    torch.cuda.synchronize(ctx.dev)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        copies()
    torch.cuda.synchronize(ctx.dev)
    copy_s = ctx.max_over_ranks(time.perf_counter() - t0)
    total = E * ctx.n_gpus
    ceiling = total * steps / copy_s
    value = total * steps / e2e_s
    return {"value": value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
            "path": "sng_step_host: pinned host buffers, H2D + step + D2H + sync",
            "numa_local_cpus": ctx.numa_cpus,
            "pcie_ceiling": {"value": ceiling, "unit": UNIT,
                             "d2h_gbs_per_gpu": d2h * steps / copy_s / 1e9, "h2d_gbs_per_gpu": h2d * steps / copy_s / 1e9,
                             "how": "pinned cudaMemcpyAsync of the same bytes, H2D and D2H on two streams at once, "
                                    "all %d ranks concurrently, max over ranks" % ctx.n_gpus},
            "pcie_peak_gbs": (h2d + d2h) * steps / copy_s / 1e9, "frac": value / ceiling}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2400)
    ap.add_argument("--warmup", type=int, default=240)
    ap.add_argument("--graph-steps", type=int, default=24, help="steps per captured CUDA graph (0 = plain launches)")
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (0 = the workload's size)")
    ap.add_argument("--total-envs", type=int, default=0, help="STRONG scaling: this many envs in total, split over the "
                    "ranks by sharding.shard_range (overrides --envs; the line then says scaling = strong)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-envs", type=int, default=0, help="env count of the CPU arms (0 = the b200 arm's size if it fits in host memory, else 65,536)")
    ap.add_argument("--ref-seconds", type=float, default=2.5, help="minimum measured time of the --impl reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="wall time of the cpu_baseline leg")
    ap.add_argument("--live-seconds", type=float, default=20.0, help="wall time of the live-reference leg (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--warps", type=int, default=0, help="tuning: warps (blocks of 32 envs) per CTA (0 = auto)")
    ap.add_argument("--generic", action="store_true", help="tuning: use the generic runtime-N kernel")
    ap.add_argument("--bulk", type=int, default=1, help="tuning: row staging: -1 scalar, 0 vector loads/stores, 1 copy-engine loads + vector stores, 3 copy engine both ways")
    ap.add_argument("--host-chunks", type=int, default=0, help="tuning: env chunks of the pipelined host path")
    ap.add_argument("--variant", type=int, default=0, help="tuning: 0 default, 1 persistent pipelined kernel, 2 one lane per env even for large stations, 3 two lanes per env, 4 one lane per SPOT at every batch size, 5 never one lane per spot")
    ap.add_argument("--ctas", type=int, default=0, help="tuning: cap on resident CTAs per SM (pipelined kernel)")
    ap.add_argument("--pdl", type=int, default=-1, help="step-kernel launch mode: 0 ordinary, 1 programmatic dependent launch; -1 = auto "
                    "(1 for small batches, where kernel-to-kernel latency is a visible share of a step)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--rollout-pdl", default="", help="programmatic dependent launch inside the rollout loop of the c3 legs: "
                    "off, policy (default), step, both (optionally +x: policy CTAs claim their SM's whole shared memory)")
    ap.add_argument("--legs", default="all", help="'all', 'none' or a comma list of c4_strong,pattern_ceiling,c5,c3,c3_fused,c3_sharded,c3_full_waves,c3_sb3,c2,rollout_kernel,generic")
    ap.add_argument("--rollout", type=int, default=0, help="legacy: same as --legs c3 with this many steps per rollout")
    args = ap.parse_args()
    PDL_ARG[0] = args.pdl
    ROLLOUT_PDL[0] = args.rollout_pdl or None
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    ctx = Ctx()
    torch = ctx.torch
    from smart_nanogrid_gym_b200.sharding import ReturnStats, shard_range

    wl = WORKLOADS[args.workload]
    strong = args.total_envs > 0
    if strong:
        lo, hi = shard_range(args.total_envs, ctx.world, ctx.rank)
        E, gid0 = hi - lo, lo
    else:
        E = args.envs or wl["envs"]
        gid0 = ctx.rank * E
    env = make_env(ctx, args.workload, E, gid0, args)
    cfg = env.cfg
    env.set_launch_mode(launch_mode_for(cfg, E))
    # actions: a pre-filled U(low, high) tensor re-read from HBM every step
    g = torch.Generator(device=ctx.dev).manual_seed(1234 + ctx.rank)
    actions = env.sample_actions(g).contiguous()

    for _ in range(min(args.warmup, 24)):
        env.step(actions)
    ctx.barrier()
    loop = StepLoop(ctx, env, actions, args.graph_steps, args.steps)
    gs = loop.gsteps or 1
    loop.run(max(args.warmup - 24, gs) // gs * gs)
    ctx.barrier()

    # ---- device-resident throughput: K launches bracketed by CUDA events on the launching stream ----
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    launches0 = env.launch_count
    ms = timed_ms(ctx, lambda: loop.run(args.steps))
    launches = args.steps if loop.graph is not None else env.launch_count - launches0
    clock_window = "timed region"
    if ms >= 150.0:
        clocks = sampler.stop()
    else:
        # the timed region is shorter than a few nvidia-smi samples: keep sampling while the same kernel
        # runs on (untimed), so that the clock record is taken under this load
        extra = 0
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            loop.run(gs if loop.graph is not None else 8)
            torch.cuda.synchronize(ctx.dev)
            extra += gs if loop.graph is not None else 8
        clock_window = "timed region + %d untimed steps of the same kernel" % extra
        clocks = sampler.stop()
    clocks["window"] = clock_window
    ms_max = ctx.max_over_ranks(ms)
    total_envs = int(ctx.sum_over_ranks(E))
    value = total_envs * args.steps / (ms_max * 1e-3)
    assert env.error_flags() == 0

    e2e = e2e_leg(ctx, env, actions, args.e2e_steps)

    # ---- optional episode-return statistics: the only collective, off the step path ----
    mean_ret = ReturnStats.from_returns(env.last_return).all_reduce(device=ctx.dev).mean
    env.close()

    # ---- the other BASELINE configurations, same run, untimed for the headline ----
    want = args.legs
    if args.rollout > 0 and want in ("none", ""):
        want = "c3"
    names = ["c4_strong", "pattern_ceiling", "c5", "c3", "c3_sb3", "c2", "rollout_kernel", "generic"] if want == "all" else [x for x in want.split(",") if x and x != "none"]
    legs = {}
    if names:
        floor_us = launch_floor_us(ctx)
        legs["launch_floor_us"] = floor_us
        for name in names:
            try:
                if name == "c4_strong":
                    lo, hi = shard_range(1048576, ctx.world, ctx.rank)
                    legs[name] = step_leg(ctx, "c4", hi - lo, lo, floor_us)
                    legs[name]["scaling"] = "strong"
                    legs[name]["what"] = "BASELINE config 4 as stated: 1,048,576 envs SPLIT over %d GPU(s)" % ctx.n_gpus
                elif name == "c5":
                    legs[name] = step_leg(ctx, "c5", WORKLOADS["c5"]["envs"], ctx.rank * WORKLOADS["c5"]["envs"], floor_us)
                    legs[name]["what"] = "BASELINE config 5: 64 spots, 96 steps, 262,144 envs per GPU"
                elif name == "c3":
                    legs[name] = rollout_leg(ctx, "c4", 65536, args.rollout or 24)
                    legs[name]["what"] = "BASELINE config 3: PPO rollout collection over 65,536 envs per GPU, N=10 station"
                elif name == "c3_full_waves":
                    # by name only: 4 x 148 tiles of 128 envs, i.e. every policy CTA gets exactly four tiles (65,536 envs are
                    # 512 tiles: 3.46 per SM, four rounds for 68 SMs and three for the other 80)
                    legs[name] = rollout_leg(ctx, "c4", 4 * 148 * 128, args.rollout or 24)
                    legs[name]["what"] = "BASELINE config 3 at 75,776 envs per GPU (a whole number of policy tiles per SM): what the tile quantisation at 65,536 envs costs"
                elif name == "c3_fused":
                    legs[name] = rollout_leg(ctx, "c4", 65536, args.rollout or 24, fuse_step=True)
                    legs[name]["what"] = "BASELINE config 3 with ONE launch per rollout step (policy forward + env step fused, sng_policy_step)"
                elif name == "c3_sharded":
                    legs[name] = rollout_sharded_leg(ctx, "c4", 65536, args.rollout or 24, shards=2)
                    legs[name]["what"] = "BASELINE config 3, 65,536 envs per GPU as two 32,768-env shards collected side by side"
                elif name == "c3_sb3":
                    legs[name] = rollout_leg(ctx, "n4", 65536, 24, shipped=True)
                    legs[name]["what"] = "BASELINE config 3 with the reference's shipped policy (N=4 station)"
                elif name == "c2":
                    legs[name] = step_leg(ctx, "c4", 4096, ctx.rank * 4096, floor_us, min_ms=100.0)
                    legs[name]["what"] = "BASELINE config 2: 4,096 envs per GPU, one sng_step launch per step"
                    legs["c2_rollout_kernel"] = rollout_kernel_leg(ctx, "c4", 4096, floor_us, min_ms=100.0)
                    legs["c2_rollout_kernel"]["what"] = "BASELINE config 2 through sng_rollout (24 steps per launch)"
                    for k in (name, "c2_rollout_kernel"):
                        legs[k]["kernel"] = "step_lanes_kernel: one lane per charging spot, two envs per warp (the default for small batches; bit-identical to the one-block-per-warp kernel)"
                elif name == "generic":
                    for wl in ("c4_h5", "c4_pv2d", "c4_nopv"):
                        legs["generic_" + wl] = step_leg(ctx, wl, WORKLOADS[wl]["envs"], ctx.rank * WORKLOADS[wl]["envs"], floor_us, min_ms=120.0)
                        legs["generic_" + wl]["what"] = "SURVEY 8f row 4: " + WORKLOADS[wl]["name"] + ", 1,048,576 envs per GPU"
                elif name == "pattern_ceiling":
                    legs[name] = pattern_ceiling_leg(ctx, WORKLOADS["c4"]["envs"])
                elif name == "rollout_kernel":
                    for n in (65536, 131072):
                        legs["rollout_kernel_%d" % n] = rollout_kernel_leg(ctx, "c4", n, floor_us)
            except Exception as exc:  # noqa: BLE001 -- a failing leg must not take the headline down; it is reported
                legs[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
                ctx.torch.cuda.synchronize(ctx.dev)

    if ctx.rank == 0:
        bytes_step = algorithmic_bytes_per_env_step(cfg.n_spots, int(cfg.batt), int(cfg.pv))
        kernel_ms = ms / max(launches, 1)
        config = workload_config(args.workload, E, ctx.n_gpus, bytes_step)
        if strong:
            config["workload"] = "%s, %d envs in total split over %d GPU(s)" % (wl["name"], total_envs, ctx.n_gpus)
            config["total_envs"] = total_envs
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "launch": ("CUDA graph of %d step launches, replayed" % loop.gsteps) if loop.graph is not None else "one launch per step",
            "mean_episode_return": mean_ret,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": roofline_of(ctx, args.workload, cfg, E, kernel_ms),
        }
        if legs:
            line["legs"] = legs
        if not args.no_cpu and ctx.n_gpus == 1:
            from oracle import oracle as orc
            orc.build()
            threads = os.cpu_count() or 1
            n_ref = pick_ref_envs(args.ref_envs, 65536)
            v, el, steps = cpu_port_throughput(n_ref, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d envs x %d steps in %.1f s, float64 C port of the reference step "
                                              "(oracle/), all host threads" % (n_ref, steps, el)}
            if args.live_seconds > 0:
                line["cpu_baseline_live"] = live_reference_throughput(args.live_seconds)
        emit(line)
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
