"""On-device PPO rollout collection around the batched step (BASELINE config 3).

The reference trains with Stable-Baselines3 PPO on ONE env (solvers/RL/ppo_train.py:89-102): SB3's
`collect_rollouts` loops `policy(obs) -> clip -> env.step -> buffer.add` and then computes GAE.  Here the
same loop runs for E envs without leaving the GPU: the step kernel writes observations / rewards / dones
straight into the rollout-buffer slices (zero-copy `out=`), the policy is a torch module of SB3's default
MlpPolicy shape (tanh 64-64 actor and critic, state-independent log-std) whose inference runs as ONE fused
kernel (`sng_policy_forward`; plain torch ops as the fallback and for training), and advantages / returns come
from the `sng_gae` kernel.  Not a trainer: it is the caller-side row next to the hot path.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import _native as nat


class MlpPolicy(nn.Module):
    """SB3 ActorCriticPolicy defaults for Box actions: separate tanh 64-64 networks, DiagGaussian head."""

    def __init__(self, obs_dim: int, act_dim: int, hidden: int = 64, log_std_init: float = 0.0):
        super().__init__()
        self.pi = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())
        self.vf = nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())
        self.action_net = nn.Linear(hidden, act_dim)
        self.value_net = nn.Linear(hidden, 1)
        self.log_std = nn.Parameter(torch.full((act_dim,), float(log_std_init)))

    def forward(self, obs: torch.Tensor, noise: torch.Tensor | None = None):
        """-> (actions, values, log_probs); `noise` ~ N(0, 1) of the action shape (None = deterministic)."""
        mean = self.action_net(self.pi(obs))
        values = self.value_net(self.vf(obs)).squeeze(-1)
        if noise is None:
            noise = torch.zeros_like(mean)
        actions = mean + noise * self.log_std.exp()
        log_probs = (-0.5 * noise.pow(2) - self.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
        return actions, values, log_probs

    def predict_values(self, obs: torch.Tensor) -> torch.Tensor:
        return self.value_net(self.vf(obs)).squeeze(-1)

    # ---- Stable-Baselines3 checkpoints (what solvers/RL/ppo_train.py:100-102 saves, predictor.py:72 loads) ----
    # SB3 ActorCriticPolicy.state_dict() key -> attribute path here; the mapping is 1:1 and lossless.
    SB3_KEYS = {
        "mlp_extractor.policy_net.0.weight": "pi.0.weight", "mlp_extractor.policy_net.0.bias": "pi.0.bias",
        "mlp_extractor.policy_net.2.weight": "pi.2.weight", "mlp_extractor.policy_net.2.bias": "pi.2.bias",
        "mlp_extractor.value_net.0.weight": "vf.0.weight", "mlp_extractor.value_net.0.bias": "vf.0.bias",
        "mlp_extractor.value_net.2.weight": "vf.2.weight", "mlp_extractor.value_net.2.bias": "vf.2.bias",
        "action_net.weight": "action_net.weight", "action_net.bias": "action_net.bias",
        "value_net.weight": "value_net.weight", "value_net.bias": "value_net.bias", "log_std": "log_std",
    }

    @classmethod
    def from_sb3_state_dict(cls, sd) -> "MlpPolicy":
        """Build the policy from an SB3 `policy.pth` state dict (MlpPolicy, default net_arch, tanh)."""
        missing = [k for k in cls.SB3_KEYS if k not in sd]
        extra = [k for k in sd if k not in cls.SB3_KEYS]
        if missing or extra:
            raise ValueError("not a default SB3 MlpPolicy state dict: missing %s, unexpected %s" % (missing, extra))
        w0 = sd["mlp_extractor.policy_net.0.weight"]
        pol = cls(int(w0.shape[1]), int(sd["action_net.weight"].shape[0]), hidden=int(w0.shape[0]))
        own = pol.state_dict()
        mapped = {dst: torch.as_tensor(sd[src]).to(own[dst].dtype) for src, dst in cls.SB3_KEYS.items()}
        for k, v in mapped.items():
            if tuple(v.shape) != tuple(own[k].shape):
                raise ValueError("shape of %s does not fit: %s vs %s" % (k, tuple(v.shape), tuple(own[k].shape)))
        pol.load_state_dict(mapped, strict=True)
        return pol

    @classmethod
    def from_sb3_zip(cls, path: str) -> "MlpPolicy":
        """Load the policy of a Stable-Baselines3 PPO checkpoint (`model.save(...)` zip, e.g. the reference's
        solvers/RL/models/PPO-b-pv-bounded-sparse-4ch-1h/999600.zip): reads `policy.pth` only."""
        import io
        import zipfile
        with zipfile.ZipFile(path) as z:
            sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        return cls.from_sb3_state_dict(sd)

    def to_sb3_state_dict(self):
        own = self.state_dict()
        return {src: own[dst].detach().clone() for src, dst in self.SB3_KEYS.items()}

    # ---- fused inference (libsng.so: sng_policy_forward) ----------------------------------
    def _mlp_struct(self):
        """ctypes view of the parameters (rebuilt on every call: optimisers may replace .data)."""
        m = nat.SngMlp()
        m.struct_size = C.sizeof(nat.SngMlp)
        m.obs_dim, m.hidden, m.act_dim = self.pi[0].in_features, self.pi[0].out_features, self.action_net.out_features
        ptr = lambda t: t.detach().data_ptr()  # noqa: E731
        m.w_pi0, m.b_pi0, m.w_pi1, m.b_pi1 = ptr(self.pi[0].weight), ptr(self.pi[0].bias), ptr(self.pi[2].weight), ptr(self.pi[2].bias)
        m.w_act, m.b_act, m.log_std = ptr(self.action_net.weight), ptr(self.action_net.bias), ptr(self.log_std)
        m.w_vf0, m.b_vf0, m.w_vf1, m.b_vf1 = ptr(self.vf[0].weight), ptr(self.vf[0].bias), ptr(self.vf[2].weight), ptr(self.vf[2].bias)
        m.w_val, m.b_val = ptr(self.value_net.weight), ptr(self.value_net.bias)
        return m

    def fused_supported(self) -> bool:
        """The tensor-core kernel (sng_policy_forward_packed) takes any observation width up to 30 and action width up
        to 16 with SB3's hidden width of 64."""
        p = self.pi[0].weight
        return (p.is_cuda and p.dtype == torch.float32 and self.pi[0].out_features == 64 and
                self.pi[0].in_features <= 30 and self.action_net.out_features <= 16)

    def cuda_core_supported(self) -> bool:
        """Shapes of the FP32 CUDA-core kernel (sng_policy_forward), kept as an independent second implementation."""
        return self.fused_supported() and (self.pi[0].in_features, self.action_net.out_features) in ((17, 5), (25, 9), (29, 11))

    def fused_kind(self) -> str:
        return "tcgen05 tf32x3, activations in TMEM"

    @torch.no_grad()
    def pack_weights(self):
        """Split the FP32 weights into tf32 hi / lo parts in the tensor cores' operand layout (sng_policy_pack: one
        tiny launch).  Call after every weight update; collect_rollout does it once per rollout."""
        lib = nat.lib()
        dev = self.pi[0].weight.device
        if getattr(self, "_packed", None) is None or self._packed.device != dev:
            self._packed = torch.empty(lib.sng_policy_packed_bytes() // 4, dtype=torch.float32, device=dev)
        m = self._mlp_struct()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        nat.check(lib.sng_policy_pack(C.byref(m), C.c_void_p(self._packed.data_ptr()), stream))
        return self._packed

    @torch.no_grad()
    def fused_forward(self, obs, noise, low, high, raw_actions, actions, values, log_probs, repack: bool = True,
                      cuda_cores: bool = False, rng=None, noise_out=None):
        """One launch: values, sampled + clipped actions and log-probs written into the given [E, ...] buffers
        (noise None = deterministic; actions None = values only).  Inference only (no autograd graph).
        repack=False reuses the weight image of the last pack_weights() (weights unchanged since);
        cuda_cores=True runs the FP32 CUDA-core kernel instead of the tensor-core one.
        rng=(seed, step_counter, step_offset, env_gid0) draws the noise inside the kernel (sng_policy_forward_sampled:
        Philox keyed by seed and step = step_counter[0] + step_offset, step_counter a one-element int64 CUDA tensor);
        noise_out [E, A] then optionally receives the standard normals that were used."""
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        if cuda_cores:
            m = self._mlp_struct()
            nat.check(nat.lib().sng_policy_forward(C.byref(m), p(obs), p(noise), p(low), p(high), p(raw_actions), p(actions),
                                                   p(values), p(log_probs), obs.shape[0], stream))
            return
        if repack or getattr(self, "_packed", None) is None:
            self.pack_weights()
        if rng is not None:
            seed, counter, offset, gid0 = rng
            assert noise is None and counter.dtype == torch.int64 and counter.is_cuda and counter.numel() == 1
            nat.check(nat.lib().sng_policy_forward_sampled(p(self._packed), self.pi[0].in_features, self.action_net.out_features,
                                                           p(obs), int(seed) & (2 ** 64 - 1), p(counter), int(offset), int(gid0),
                                                           p(low), p(high), p(raw_actions), p(actions), p(values), p(log_probs),
                                                           p(noise_out), obs.shape[0], stream))
            return
        nat.check(nat.lib().sng_policy_forward_packed(p(self._packed), self.pi[0].in_features, self.action_net.out_features,
                                                      p(obs), p(noise), p(low), p(high), p(raw_actions), p(actions),
                                                      p(values), p(log_probs), obs.shape[0], stream))


class RolloutBuffer:
    """[n_steps, E, ...] device tensors, SB3 RolloutBuffer field for field.  `observations` has n_steps + 1
    slabs: slab s is what the policy saw at step s, slab s + 1 is written by the step kernel itself."""

    def __init__(self, n_steps: int, n_envs: int, obs_dim: int, act_dim: int, device, gamma: float = 0.99,
                 gae_lambda: float = 0.95):
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=device)  # noqa: E731
        self.n_steps, self.n_envs, self.gamma, self.gae_lambda = n_steps, n_envs, gamma, gae_lambda
        self.observations = z(n_steps + 1, n_envs, obs_dim)
        self.actions = z(n_steps, n_envs, act_dim)            # what the env executed (clipped to the Box)
        self.raw_actions = z(n_steps, n_envs, act_dim)        # what the policy sampled (SB3 stores these)
        self.rewards = z(n_steps, n_envs)
        # episode_starts[s + 1] IS dones[s] (SB3 carries `_last_episode_starts = dones`): two views of one
        # [n_steps + 1, E] array, so the step kernel's done flags land in both without a copy per step
        self._flags = z(n_steps + 1, n_envs, dtype=torch.uint8)
        self.episode_starts = self._flags[:n_steps]
        self.dones = self._flags[1:]
        self.values = z(n_steps, n_envs)
        self.log_probs = z(n_steps, n_envs)
        self.advantages = z(n_steps, n_envs)
        self.returns = z(n_steps, n_envs)
        self.last_values = z(n_envs)

    def compute_returns_and_advantage(self, last_values: torch.Tensor, last_dones: torch.Tensor):
        """SB3 RolloutBuffer.compute_returns_and_advantage on the device (sng_gae kernel)."""
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        lv = last_values.detach().float().contiguous()
        ld = last_dones.to(torch.uint8).contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(self.rewards.device).cuda_stream)
        nat.check(nat.lib().sng_gae(p(self.rewards), p(self.values), p(self.episode_starts), p(lv), p(ld),
                                    p(self.advantages), p(self.returns), self.n_steps, self.n_envs,
                                    C.c_float(self.gamma), C.c_float(self.gae_lambda), stream))
        return self.advantages, self.returns


@torch.no_grad()
def collect_rollout(env, policy: MlpPolicy, buf: RolloutBuffer, obs: torch.Tensor, episode_starts: torch.Tensor,
                    generator: torch.Generator | None = None, deterministic: bool = False, fused: bool = True,
                    rng_seed: int | None = None, pack: bool = True, advance_counter: bool = True, pdl: bool = False,
                    fuse_step: bool = False):
    """SB3 OnPolicyAlgorithm.collect_rollouts for a BatchedSmartNanogridEnv: n_steps policy + env steps,
    then GAE.  `obs` [E, D] is the current observation (from reset() or the previous rollout),
    `episode_starts` [E] u8.  Returns (last_obs, last_dones) to carry into the next call.
    Exploration noise: torch.randn (with `generator`) per step, or -- `rng_seed` given, fused kernel -- drawn inside
    the policy kernel (Philox keyed by rng_seed and the policy's step counter: no noise tensor, no RNG launch).
    pack=False / advance_counter=False: the caller packs the weights / advances the step counter itself (several
    shards collected side by side share both, see ShardedGraphedRollout).
    fuse_step (fused kernel only, env.supports_policy_step()): ONE launch per rollout step -- the policy kernel's io warps
    run the env step of their tile themselves (env.policy_step, sng_policy_step); bit-identical to the two launches.
    pdl (fused kernel only): programmatic dependent launch inside the loop -- a kernel starts while its predecessor is
    draining (block scheduling, barrier / TMEM set-up, the policy's weight image, the step's state loads, none of which
    the predecessor writes) and waits for it before reading what it wrote.  "policy": only the policy kernel is
    launched that way (behind the step kernel); "step": only the step kernel; True / "both": both; a trailing "+x"
    ("policy+x", ...) makes the policy CTAs claim their SM's whole shared memory, so that early step CTAs cannot become
    resident next to a running policy CTA.  DESIGN.md section 6 has the measurements."""
    low, high = env.action_low.float(), env.action_high.float()
    buf.observations[0].copy_(obs)
    buf.episode_starts[0].copy_(episode_starts.to(torch.uint8))    # episode_starts[s + 1] aliases dones[s]
    fused = fused and policy.fused_supported()
    in_kernel_noise = fused and not deterministic and rng_seed is not None
    if fused and pack:
        policy.pack_weights()          # once per rollout: the weights do not change while it is collected
    if in_kernel_noise and (getattr(policy, "rng_counter", None) is None or policy.rng_counter.device != obs.device):
        policy.rng_counter = torch.zeros(1, dtype=torch.int64, device=obs.device)   # rollout steps drawn so far
    fuse_step = bool(fuse_step and fused and env.supports_policy_step())
    pdl = pdl if fused else False
    mode = "both" if pdl is True else (pdl or "")
    exclusive = mode.endswith("+x")
    mode = mode[:-2] if exclusive else mode
    if mode not in ("", "both", "policy", "step"):
        raise ValueError("pdl must be False, True, 'both', 'policy' or 'step' (optionally + '+x')")
    if mode in ("both", "step"):
        env.set_launch_mode(2)         # the step's predecessor is the policy kernel: state loads ahead of the wait
    policy_mode = (1 if mode in ("both", "policy") else 0) | (2 if exclusive else 0)
    if policy_mode:
        nat.check(nat.lib().sng_policy_set_launch_mode(policy_mode & 2))   # the first call follows the weight packing: ordinary launch
    try:
        _rollout_steps(env, policy, buf, low, high, fused, in_kernel_noise, rng_seed, deterministic, generator, policy_mode, fuse_step)
        last_obs = buf.observations[buf.n_steps]
        if fused:
            policy.fused_forward(last_obs, None, None, None, None, None, buf.last_values, None, repack=False)
        else:
            buf.last_values.copy_(policy.predict_values(last_obs))
    finally:
        if mode in ("both", "step"):
            env.set_launch_mode(0)
        if policy_mode:
            nat.check(nat.lib().sng_policy_set_launch_mode(0))
    if in_kernel_noise and advance_counter:
        policy.rng_counter += buf.n_steps
    buf.compute_returns_and_advantage(buf.last_values, buf.dones[buf.n_steps - 1])
    return last_obs, buf.dones[buf.n_steps - 1]


def _rollout_steps(env, policy, buf, low, high, fused, in_kernel_noise, rng_seed, deterministic, generator, policy_mode=0,
                   fuse_step=False):
    for s in range(buf.n_steps):
        o = buf.observations[s]
        if s == 1 and policy_mode & 1:     # from the second call on the policy kernel follows a step kernel
            nat.check(nat.lib().sng_policy_set_launch_mode(policy_mode))
        if fuse_step:                      # one kernel: policy forward, sampling, clipping, env step, next observation
            noise = None
            if not in_kernel_noise:
                noise = torch.zeros(buf.n_envs, buf.actions.shape[2], device=o.device) if deterministic else \
                    torch.randn(buf.n_envs, buf.actions.shape[2], device=o.device, generator=generator)
            env.policy_step(policy._packed, o, low, high, buf.raw_actions[s], buf.actions[s], buf.values[s], buf.log_probs[s],
                            out=(buf.observations[s + 1], buf.rewards[s], buf.dones[s]), noise=noise,
                            rng=(rng_seed, policy.rng_counter, s) if in_kernel_noise else None)
            continue
        if in_kernel_noise:
            policy.fused_forward(o, None, low, high, buf.raw_actions[s], buf.actions[s], buf.values[s], buf.log_probs[s],
                                 repack=False, rng=(rng_seed, policy.rng_counter, s, env.env_gid0))
        else:
            noise = None if deterministic else torch.randn(buf.n_envs, buf.actions.shape[2], device=o.device, generator=generator)
            if fused:      # one kernel: both networks, heads, sampling, clipping, log-probs, straight into the buffer slabs
                policy.fused_forward(o, noise, low, high, buf.raw_actions[s], buf.actions[s], buf.values[s], buf.log_probs[s], repack=False)
            else:
                a, v, lp = policy(o, noise)
                buf.raw_actions[s].copy_(a)
                torch.clamp(a, low, high, out=buf.actions[s])      # SB3 clips Box actions before env.step
                buf.values[s].copy_(v)
                buf.log_probs[s].copy_(lp)
        # the kernel writes the next observation, the reward and the done flag (= the next step's episode start)
        # into the buffer slabs
        env.step(buf.actions[s], out=(buf.observations[s + 1], buf.rewards[s], buf.dones[s]))


class GraphedRollout:
    """collect_rollout captured ONCE in a CUDA graph (n_steps x [policy MLP, clip, step kernel] + GAE) and
    replayed: at 65,536 envs the eager loop is bound by ~20 small launches per step, the graph is not.
    Sampling noise is drawn inside the policy kernel (rng_seed; the step counter is a device word the graph advances),
    or -- rng_seed=None -- by torch's default CUDA generator (graph-safe); `deterministic=True` uses the mean.
    pdl / fuse_step: see collect_rollout; the default launches the policy kernel programmatically behind the step kernel
    (measured fastest, bit-identical)."""

    def __init__(self, env, policy: MlpPolicy, buf: RolloutBuffer, deterministic: bool = False, fused: bool = True,
                 rng_seed: int | None = 0, pdl="policy", fuse_step: bool = False):
        dev = buf.rewards.device
        self.env, self.policy, self.buf = env, policy, buf
        self.obs_in = torch.zeros(buf.n_envs, buf.observations.shape[2], device=dev)
        self.starts_in = torch.zeros(buf.n_envs, dtype=torch.uint8, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up outside capture (lazy inits, cuBLAS workspaces)
            self.obs_in.copy_(env.obs)
            collect_rollout(env, policy, buf, self.obs_in, self.starts_in, deterministic=deterministic, fused=fused, rng_seed=rng_seed, pdl=pdl, fuse_step=fuse_step)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.last_obs, self.last_dones = collect_rollout(env, policy, buf, self.obs_in, self.starts_in,
                                                             deterministic=deterministic, fused=fused, rng_seed=rng_seed, pdl=pdl, fuse_step=fuse_step)

    def __call__(self, obs: torch.Tensor, episode_starts: torch.Tensor):
        self.obs_in.copy_(obs)
        self.starts_in.copy_(episode_starts)
        self.graph.replay()
        return self.last_obs, self.last_dones


class ShardedGraphedRollout:
    """One rollout over K env shards of ONE GPU, collected side by side: shard k's policy kernel (tensor cores, one CTA
    per SM) runs while shard k'ish step kernel (a few small CTAs per SM, HBM / L2 traffic) runs -- the two kernels bound by
    different resources overlap instead of alternating with a launch gap in between.  Captured once as a CUDA graph with
    K parallel branches.  The shards are ordinary BatchedSmartNanogridEnv objects with consecutive env_gid0 (RNG streams
    and exploration noise are keyed by global env id, so K shards produce exactly what one env of the summed size does:
    tests/test_gpu_rollout.py)."""

    def __init__(self, envs, policy: MlpPolicy, bufs, deterministic: bool = False, rng_seed: int | None = 0):
        assert len(envs) == len(bufs) and len(envs) >= 1
        dev = bufs[0].rewards.device
        self.envs, self.policy, self.bufs = list(envs), policy, list(bufs)
        self.n_steps = bufs[0].n_steps
        self.obs_in = [torch.zeros(b.n_envs, b.observations.shape[2], device=dev) for b in bufs]
        self.starts_in = [torch.zeros(b.n_envs, dtype=torch.uint8, device=dev) for b in bufs]
        self.streams = [torch.cuda.Stream(device=dev) for _ in envs]
        in_kernel = (not deterministic) and rng_seed is not None and policy.fused_supported()
        if in_kernel and (getattr(policy, "rng_counter", None) is None or policy.rng_counter.device != dev):
            policy.rng_counter = torch.zeros(1, dtype=torch.int64, device=dev)

        def body():
            main = torch.cuda.current_stream(dev)
            policy.pack_weights()
            fork = torch.cuda.Event()
            fork.record(main)
            outs = []
            for k, st in enumerate(self.streams):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    outs.append(collect_rollout(self.envs[k], policy, self.bufs[k], self.obs_in[k], self.starts_in[k],
                                                deterministic=deterministic, rng_seed=rng_seed, pack=False, advance_counter=False))
                    join = torch.cuda.Event()
                    join.record(st)
                main.wait_event(join)
            if in_kernel:
                policy.rng_counter += self.n_steps
            return outs

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up outside capture
            for k, e in enumerate(self.envs):
                self.obs_in[k].copy_(e.obs)
            body()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outs = body()

    def __call__(self, obs_list, starts_list):
        for k in range(len(self.envs)):
            self.obs_in[k].copy_(obs_list[k])
            self.starts_in[k].copy_(starts_list[k])
        self.graph.replay()
        return [o[0] for o in self.outs], [o[1] for o in self.outs]


def gae_reference(rewards, values, episode_starts, last_values, last_dones, gamma, gae_lambda):
    """Plain torch float64 restatement of SB3's loop (the checker of the sng_gae kernel in tests)."""
    n = rewards.shape[0]
    adv = torch.zeros_like(rewards, dtype=torch.float64)
    last_gae = torch.zeros_like(last_values, dtype=torch.float64)
    for t in reversed(range(n)):
        if t == n - 1:
            non_terminal = 1.0 - last_dones.double()
            next_values = last_values.double()
        else:
            non_terminal = 1.0 - episode_starts[t + 1].double()
            next_values = values[t + 1].double()
        delta = rewards[t].double() + gamma * next_values * non_terminal - values[t].double()
        last_gae = delta + gamma * gae_lambda * non_terminal * last_gae
        adv[t] = last_gae
    return adv, adv + values.double()
