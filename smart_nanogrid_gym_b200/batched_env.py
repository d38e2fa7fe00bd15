"""`BatchedSmartNanogridEnv`: E nanogrid environments stepped by one CUDA kernel launch.

Host-side mirror of the reference's gym interface (envs/smart_nanogrid_environment.py):
same constructor keyword arguments, `reset` / `step` semantics, spaces, reward and
termination rules -- for `num_envs` environments at once, on torch CUDA tensors that the
C-ABI extension borrows zero-copy.  There is no CPU path: without libsng.so or a GPU this raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from .config import NanogridConfig
from .schedule import ScheduleRecords, MAX_VEHICLES
from .spaces import Box


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedSmartNanogridEnv:
    """num_envs environments on one GPU.

    Parameters beyond the reference constructor's:
      num_envs, device     batch size and CUDA device (this process's GPU)
      seed                 base seed of the counter-based (Philox) schedule sampler
      env_gid0             global id of env 0 (multi-GPU sharding: streams are keyed by global id,
                           so results do not depend on how envs are split over GPUs)
      precision            "float32" (production) or "float64" (validation build)
      auto_reset           finished envs restart inside the same step (VecEnv semantics);
                           `terminal_obs` then holds the last observation of the finished episode
    Tensors returned by reset()/step() are views of internal buffers and are overwritten by the
    next call (clone them to keep them), unless `out=` buffers are supplied.
    """

    def __init__(self, num_envs: int, device="cuda:0", seed: int = 0, env_gid0: int = 0, precision: str = "float32",
                 auto_reset: bool = True, want_terminal_obs: bool = False, want_diagnostics: bool = False,
                 config: Optional[NanogridConfig] = None, **kwargs):
        self.cfg = config if config is not None else NanogridConfig(**kwargs)
        if not torch.cuda.is_available():
            raise nat.NativeError("BatchedSmartNanogridEnv needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.num_envs = int(num_envs)
        self.seed_value = int(seed)
        self.env_gid0 = int(env_gid0)
        self.precision = {"float32": nat.SNG_F32, "float64": nat.SNG_F64}[precision]
        self.real = torch.float32 if self.precision == nat.SNG_F32 else torch.float64
        self.auto_reset = bool(auto_reset)
        self._lib = nat.lib()
        cfg = self.cfg
        E, N, A, D = self.num_envs, cfg.n_spots, cfg.act_dim, cfg.obs_dim
        self._c, self._tabs = nat.make_config(cfg, E, env_gid0, self.precision, auto_reset)
        self.layout = nat.query_layout(self._c)
        assert self.layout.act_dim == A and self.layout.obs_dim == D
        dev = self.device
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.actions = z(E, A, dtype=self.real)
        self.obs = z(E, D, dtype=torch.float32)
        self.reward = z(E, dtype=self.real)
        self.done = z(E, dtype=torch.uint8)
        self.terminal_obs = z(E, D, dtype=torch.float32) if want_terminal_obs else None
        # per-spot state: one array of real-sized words, a structure of arrays, plane-major and blocked by 32 envs:
        # [3 planes: header | SoC | requested SoC][ceil(E/32)][N][32] (include/sng.h sng_buffers.spot)
        B = self.layout.env_block
        self._blocks = (E + B - 1) // B
        self._word = torch.int32 if self.precision == nat.SNG_F32 else torch.int64
        self._spot = z(self.layout.spot_planes, self._blocks, N, B, dtype=self._word)
        self._envst = z(E * self.layout.envst_bytes, dtype=torch.uint8)
        self._plan = None
        self.err = z(E, dtype=torch.int32)
        self.diag = z(E, self.layout.diag_count, dtype=self.real) if want_diagnostics else None
        self.spot_power = z(E, N, dtype=self.real) if want_diagnostics else None   # per-spot power of the last step
        self.last_return = z(E, dtype=self.real)
        self.truncated = z(E, dtype=torch.bool)
        low, high = cfg.action_bounds()
        self.action_space = Box(low=low, high=high, shape=(A,), dtype=np.float32)
        self.observation_space = Box(low=np.zeros(D, np.float32), high=np.ones(D, np.float32), dtype=np.float32)
        self.action_low = torch.tensor(low, device=dev, dtype=self.real)
        self.action_high = torch.tensor(high, device=dev, dtype=self.real)
        h = C.c_void_p()
        nat.check(self._lib.sng_create(C.byref(self._c), dev.index or 0, C.byref(h)))
        self._h = h
        self._bound_actions = self.actions
        self._bound_out = (self.obs, self.reward, self.done)
        self._bind()
        self._pending = None

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _bind(self):
        b = nat.SngBuffers()
        b.struct_size = C.sizeof(nat.SngBuffers)
        obs, rew, done = self._bound_out
        b.actions, b.obs, b.reward, b.done = (_ptr(self._bound_actions), _ptr(obs), _ptr(rew), _ptr(done))
        b.terminal_obs, b.spot, b.envst = _ptr(self.terminal_obs), _ptr(self._spot), _ptr(self._envst)
        b.plan, b.err, b.diag, b.last_return = _ptr(self._plan), _ptr(self.err), _ptr(self.diag), _ptr(self.last_return)
        b.spot_power = _ptr(self.spot_power)
        nat.check(self._lib.sng_bind(self._h, C.byref(b)))

    def _ensure_plan(self):
        if self._plan is None:
            n = self.num_envs * self.cfg.n_spots * MAX_VEHICLES * self.layout.plan_rec_bytes
            self._plan = torch.zeros(n, dtype=torch.uint8, device=self.device)
            self._bind()

    def _check_tensor(self, t, shape, dtype, name):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device and t.dtype == dtype and
                tuple(t.shape) == tuple(shape) and t.is_contiguous()):
            raise ValueError("%s must be a contiguous %s CUDA tensor of shape %s on %s" % (name, dtype, tuple(shape), self.device))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sng_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def set_tuning(self, warps_per_cta=0, use_generic_kernel=0, use_bulk_copy=1, host_chunks=0):
        """Experiment / test knobs of the step launch (include/sng.h sng_set_tuning)."""
        nat.check(self._lib.sng_set_tuning(self._h, warps_per_cta, use_generic_kernel, use_bulk_copy, host_chunks))

    def set_pipeline(self, kernel_variant=0, ctas_per_sm=0):
        """0 default (four lanes per env for stations of more than 32 spots; for the default station shapes one lane per SPOT
        for small batches and two lanes per env for 10-spot batches of up to 16,384 / 65,536 envs per step / rollout launch),
        1 persistent pipelined kernel, 2 always one lane per env, 3 two lanes per env (64 and 10 spots), 4 one lane per spot at
        every batch size, 5 like 0 without the two small-batch forms (include/sng.h)."""
        nat.check(self._lib.sng_set_pipeline(self._h, kernel_variant, ctas_per_sm))

    def set_launch_mode(self, mode=0):
        """0 ordinary launches; 1 programmatic dependent launch of the step kernel (hides the kernel-to-kernel latency of
        small batches); 2 the same with the state loads ahead of the wait -- only when the previous kernel in the stream
        does not write this env's state (a policy kernel, not another step of this env).  include/sng.h."""
        nat.check(self._lib.sng_set_launch_mode(self._h, int(mode)))

    def _plane(self, f):
        """Plane f of the blocked per-spot state -> a de-blocked [E, N] copy (words)."""
        return self._spot[f].permute(0, 2, 1).reshape(-1, self.cfg.n_spots)[:self.num_envs].contiguous()

    @property
    def soc(self):
        """[E, N] SoC column the next step starts from (a de-blocked copy of the kernel state)."""
        return self._plane(1).view(self.real)

    def spot_state(self):
        """Decoded per-spot state as numpy arrays [E, N]: arrival, departure, capacity, next arrival
        (255 = none), requested SoC, SoC."""
        h = self._plane(0).cpu().numpy().astype(np.int64)
        return dict(arr=(h & 0xFF).astype(np.int32), dep=((h >> 8) & 0xFF).astype(np.int32),
                    cap=((h >> 16) & 0xFF).astype(np.int32), next=((h >> 24) & 0xFF).astype(np.int32),
                    req=self._plane(2).view(self.real).cpu().numpy().astype(np.float64),
                    soc=self.soc.cpu().numpy().astype(np.float64))

    @property
    def launch_count(self) -> int:
        return int(self._lib.sng_launch_count(self._h))

    # ------------------------------------------------------------------ gym surface
    def seed(self, seed=None):
        """The reference's seed() is a no-op (…environment.py:362-365); here it sets the Philox base seed
        used by the next reset()."""
        if seed is not None:
            self.seed_value = int(seed)
        return [self.seed_value]

    def reset(self, seed: Optional[int] = None, mask: Optional[torch.Tensor] = None, reset_battery: bool = False):
        """Start a new sampled episode in every env (or in the envs selected by `mask`).
        Returns the observation tensor [E, D] (reference: reset() -> (obs, {}), …environment.py:311-351).
        The battery SoC is kept across resets like the reference does.
        A masked reset leaves the handle-wide settings alone: it cannot change the seed and is refused while a
        loaded schedule is being replayed (the other envs would silently switch to sampling)."""
        self.cfg.validate_modes()
        if seed is not None:
            self.seed_value = int(seed)
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        if self._bound_out[0] is not self.obs:
            self._bound_out = (self.obs, self.reward, self.done)
            self._bind()
        nat.check(self._lib.sng_reset(self._h, C.c_uint64(self.seed_value & (2 ** 64 - 1)), _ptr(m), int(reset_battery),
                                      self._stream()))
        return self.obs

    def step(self, actions: Optional[torch.Tensor] = None, out=None):
        """One step of every env.  `actions` [E, A] (CUDA tensor of the env's real dtype) is borrowed
        zero-copy; pass None to use `self.actions` filled in place.  `out=(obs, reward, done)` lets the
        kernel write straight into caller buffers (e.g. rollout-buffer slices).
        Returns (obs, reward, terminated, truncated, info) -- reference: …environment.py:140-188."""
        rebind = False
        a = self.actions if actions is None else actions
        if a is not self._bound_actions:
            self._check_tensor(a, (self.num_envs, self.cfg.act_dim), self.real, "actions")
            self._bound_actions = a
            rebind = True
        o = (self.obs, self.reward, self.done) if out is None else tuple(out)
        if any(x is not y for x, y in zip(o, self._bound_out)):
            self._check_tensor(o[0], (self.num_envs, self.cfg.obs_dim), torch.float32, "out[0] (obs)")
            self._check_tensor(o[1], (self.num_envs,), self.real, "out[1] (reward)")
            self._check_tensor(o[2], (self.num_envs,), torch.uint8, "out[2] (done)")
            self._bound_out = o
            rebind = True
        if rebind:
            self._bind()
        nat.check(self._lib.sng_step(self._h, self._stream()))
        return o[0], o[1], o[2], self.truncated, {}

    def supports_policy_step(self) -> bool:
        """sng_policy_step (policy forward + env step in one launch) exists for the reference's default station -- PV with
        three steps ahead, battery, every vehicle requesting SoC 1.0 -- at 4 and 10 spots, float32, sampled schedules,
        batches that are a multiple of 128 envs."""
        c = self.cfg
        return (self.precision == nat.SNG_F32 and c.pv and c.batt and c.hours_ahead == 3 and c.n_spots in (4, 10) and
                not c.enable_requested_state_of_charge and int(self._c.pv_days) <= 1 and self._plan is None and
                self.num_envs % 128 == 0)

    def policy_step(self, packed, obs, low, high, raw_actions, actions, values, log_probs, out, noise=None, rng=None,
                    noise_out=None):
        """ONE launch for a rollout step: the tensor-core policy forward on `obs` (weight image `packed` from
        MlpPolicy.pack_weights) fused with this env's step on the clipped actions it samples (include/sng.h
        sng_policy_step).  `out=(obs_next, reward, done)` receive the step's results; raw_actions / actions / values /
        log_probs what SB3's rollout buffer stores.  noise [E, A] standard normals, or rng=(seed, step_counter, step_offset)
        to draw them in the kernel.  Bit-identical to MlpPolicy.fused_forward followed by step()."""
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        o = tuple(out)
        self._check_tensor(obs, (self.num_envs, self.cfg.obs_dim), torch.float32, "obs")
        self._check_tensor(o[0], (self.num_envs, self.cfg.obs_dim), torch.float32, "out[0] (obs)")
        self._check_tensor(o[1], (self.num_envs,), self.real, "out[1] (reward)")
        self._check_tensor(o[2], (self.num_envs,), torch.uint8, "out[2] (done)")
        self._check_tensor(actions, (self.num_envs, self.cfg.act_dim), torch.float32, "actions")
        self._check_tensor(raw_actions, (self.num_envs, self.cfg.act_dim), torch.float32, "raw_actions")
        seed, counter, offset = (0, None, 0) if rng is None else rng
        nat.check(self._lib.sng_policy_step(self._h, p(packed), p(obs), p(noise), int(seed) & (2 ** 64 - 1), p(counter), int(offset),
                                            p(low), p(high), p(raw_actions), p(actions), p(values), p(log_probs), p(noise_out),
                                            p(o[0]), p(o[1]), p(o[2]), self._stream()))
        return o[0], o[1], o[2], self.truncated, {}

    # SB3 VecEnv-style aliases
    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        a, self._pending = self._pending, None
        return self.step(a)

    def step_host(self, actions_host: torch.Tensor, obs_host: torch.Tensor, reward_host: torch.Tensor,
                  done_host: torch.Tensor):
        """The gym-facing call with HOST (ideally pinned) tensors: H2D copy of the actions, the step,
        D2H copies of obs / reward / done, and a stream sync -- all inside the extension."""
        if self._bound_actions is not self.actions or self._bound_out[0] is not self.obs:
            self._bound_actions, self._bound_out = self.actions, (self.obs, self.reward, self.done)
            self._bind()
        for t in (actions_host, obs_host, reward_host, done_host):
            if t.is_cuda or not t.is_contiguous():
                raise ValueError("step_host expects contiguous host tensors")
        nat.check(self._lib.sng_step_host(self._h, _ptr(actions_host), _ptr(obs_host), _ptr(reward_host),
                                          _ptr(done_host), self._stream()))
        return obs_host, reward_host, done_host

    def rollout(self, actions: torch.Tensor, obs=None, reward=None, done=None):
        """n consecutive steps in ONE kernel launch.  actions [n, E, A] -> obs [n, E, D], reward [n, E], done [n, E]."""
        n = actions.shape[0]
        E, A, D = self.num_envs, self.cfg.act_dim, self.cfg.obs_dim
        self._check_tensor(actions, (n, E, A), self.real, "actions")
        obs = torch.empty(n, E, D, dtype=torch.float32, device=self.device) if obs is None else obs
        reward = torch.empty(n, E, dtype=self.real, device=self.device) if reward is None else reward
        done = torch.empty(n, E, dtype=torch.uint8, device=self.device) if done is None else done
        self._check_tensor(obs, (n, E, D), torch.float32, "obs")
        self._check_tensor(reward, (n, E), self.real, "reward")
        self._check_tensor(done, (n, E), torch.uint8, "done")
        nat.check(self._lib.sng_rollout(self._h, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done), int(n),
                                        self._stream()))
        return obs, reward, done

    # ------------------------------------------------------------------ schedules
    def load_schedule(self, rec: ScheduleRecords, pv_shift=None, soc_b=None):
        """Replay mode: run the given schedules instead of sampling (the parity-test ingestion path;
        reference: ChargingStation.load_initial_values, charging_station.py:119-136).
        Returns the reset observation."""
        self.cfg.validate_modes()
        E, N = self.num_envs, self.cfg.n_spots
        if rec.arr.shape[:2] != (E, N):
            raise ValueError("schedule shape %s does not match (num_envs, spots) = %s" % (rec.arr.shape[:2], (E, N)))
        rec.validate(self.cfg.n_steps)
        self._ensure_plan()
        if self._bound_out[0] is not self.obs:
            self._bound_out = (self.obs, self.reward, self.done)
            self._bind()
        keep = [np.ascontiguousarray(rec.arr, np.int32), np.ascontiguousarray(rec.dep, np.int32),
                np.ascontiguousarray(rec.cap, np.int32), np.ascontiguousarray(rec.soc0, np.float64),
                np.ascontiguousarray(rec.req, np.float64), np.ascontiguousarray(rec.n_veh, np.int32)]
        v = nat.SngScheduleView()
        v.struct_size = C.sizeof(nat.SngScheduleView)
        v.n_slots = rec.arr.shape[2]
        v.arr, v.dep, v.cap, v.soc0, v.req, v.n_veh = [k.ctypes.data for k in keep]
        if pv_shift is not None:
            ps = np.ascontiguousarray(np.broadcast_to(np.asarray(pv_shift, np.float64), (E,)))
            keep.append(ps)
            v.pv_shift = ps.ctypes.data
        if soc_b is not None:
            sb = np.ascontiguousarray(np.broadcast_to(np.asarray(soc_b, np.float64), (E,)))
            keep.append(sb)
            v.soc_b = sb.ctypes.data
        nat.check(self._lib.sng_load_schedule(self._h, C.byref(v), self._stream()))
        return self.obs

    def _decode_plan(self) -> ScheduleRecords:
        E, N, V = self.num_envs, self.cfg.n_spots, MAX_VEHICLES
        raw = self._plan.cpu().numpy()
        if self.precision == nat.SNG_F32:
            dt = np.dtype([("hdr", "<u4"), ("soc0", "<f4"), ("req", "<f4")])
        else:
            dt = np.dtype([("hdr", "<u4"), ("pad", "<u4"), ("soc0", "<f8"), ("req", "<f8")])
        r = raw.view(dt).reshape(E, N, V)
        hdr = r["hdr"]
        arr = (hdr & 0xFF).astype(np.int32)
        valid = arr != 0xFF
        rec = ScheduleRecords(arr=np.where(valid, arr, 0), dep=np.where(valid, (hdr >> 8) & 0xFF, 0).astype(np.int32),
                              cap=np.where(valid, (hdr >> 16) & 0xFF, 0).astype(np.int32),
                              soc0=np.where(valid, r["soc0"], 0).astype(np.float64),
                              req=np.where(valid, r["req"], 0).astype(np.float64),
                              n_veh=valid.sum(axis=2).astype(np.int32))
        return rec

    def sample_plan(self) -> ScheduleRecords:
        """Whole-day schedule of the current (sampled) episode of every env, generated on the GPU from the
        same Philox streams the in-step sampler uses (the `initial_values.json` content of the reference)."""
        self._ensure_plan()
        nat.check(self._lib.sng_sample_plan(self._h, self._stream()))
        torch.cuda.synchronize(self.device)
        return self._decode_plan()

    # ------------------------------------------------------------------ state access
    def env_state(self):
        """Decoded per-env scalars: battery SoC, pv_shift, running episode return, episode, t."""
        raw = self._envst.cpu().numpy()
        if self.precision == nat.SNG_F32:
            dt = np.dtype([("soc_b", "<f4"), ("pv_shift", "<f4"), ("ep_ret", "<f4"), ("t_ep", "<u4")])
        else:
            dt = np.dtype([("soc_b", "<f8"), ("pv_shift", "<f8"), ("ep_ret", "<f8"), ("t_ep", "<u4"), ("pad", "<u4")])
        r = raw.view(dt)
        return dict(soc_b=r["soc_b"].astype(np.float64), pv_shift=r["pv_shift"].astype(np.float64),
                    ep_ret=r["ep_ret"].astype(np.float64), t=(r["t_ep"] & 0xFF).astype(np.int32),
                    episode=(r["t_ep"] >> 8).astype(np.int64))

    def state_dict(self):
        sd = dict(spot=self._spot.clone(), envst=self._envst.clone(), err=self.err.clone(),
                  last_return=self.last_return.clone(), seed=self.seed_value, obs=self.obs.clone())
        if self._plan is not None:
            sd["plan"] = self._plan.clone()
        return sd

    def load_state_dict(self, sd):
        """Restores a state captured by state_dict() on an env that has been reset() / load_schedule()d
        in the same mode (sampling vs replay)."""
        self._spot.copy_(sd["spot"])
        self._envst.copy_(sd["envst"])
        self.err.copy_(sd["err"])
        self.last_return.copy_(sd["last_return"])
        self.obs.copy_(sd["obs"])
        if "plan" in sd:
            self._ensure_plan()
            self._plan.copy_(sd["plan"])
        self.seed_value = int(sd["seed"])

    def error_flags(self) -> int:
        out = C.c_uint32(0)
        nat.check(self._lib.sng_error_flags(self._h, C.byref(out), self._stream()))
        return int(out.value)

    def check_errors(self):
        """Raise what the reference would have raised inside step()."""
        f = self.error_flags()
        if f & nat.FLAG_NEG_DEMAND:
            # central_management_system.py:158-159
            raise ValueError("Error: If V2X mode is not enabled, then power_demand cannot be less than 0!")
        if f & nat.FLAG_BATT_SOC_GT1:
            raise ValueError("Error: Battery SOC greater than 1!")  # penaliser.py:111
        if f & nat.FLAG_NAN_ACTION:
            raise ValueError("NaN action")

    def sample_actions(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """Uniform actions from the action box, on the device."""
        u = torch.rand(self.num_envs, self.cfg.act_dim, device=self.device, dtype=self.real, generator=generator)
        return self.action_low + (self.action_high - self.action_low) * u

    def random_actions(self, seed: int, step0: int = 0, n_steps: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The random policy as a counter-based draw (include/sng.h sng_sample_actions): actions [n_steps, E, A] uniform in
        the action box, a function of (seed, GLOBAL env id, step0 + s) only -- the same for any sharding, batch size or
        n_steps.  Feed the slab to `rollout`, or slab[s] to `step`."""
        E, A = self.num_envs, self.cfg.act_dim
        out = torch.empty(n_steps, E, A, dtype=self.real, device=self.device) if out is None else out
        self._check_tensor(out, (n_steps, E, A), self.real, "out")
        nat.check(self._lib.sng_sample_actions(self._h, int(seed), int(step0), int(n_steps), _ptr(out), self._stream()))
        return out

    def rbc_actions(self, obs: torch.Tensor) -> torch.Tensor:
        """The reference's rule-based controller (solvers/RBC/rbc.py:6-29) with generic offsets
        (SURVEY 8c): per spot 0 if no vehicle, 1 if it departs within 3 h, else the mean of the current
        and next-step normalised radiation; battery action 0."""
        cfg = self.cfg
        off = (1 + int(cfg.pv)) * (1 + cfg.hours_ahead) + cfg.n_spots     # departure entries follow the SoC entries
        dep = obs[:, off:off + cfg.n_spots].to(torch.float64)
        rad = ((obs[:, 0].to(torch.float64) + obs[:, 2].to(torch.float64)) / 2)[:, None].expand_as(dep)
        a = torch.where(dep == 0, torch.zeros_like(dep), torch.where(dep < 0.16667, torch.ones_like(dep), rad))
        if cfg.batt:
            a = torch.cat([a, torch.zeros_like(a[:, :1])], dim=1)
        return a.to(self.real)
