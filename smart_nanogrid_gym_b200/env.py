"""`SmartNanogridEnv`: the reference's single-environment gym API on top of the CUDA step.

Drop-in for `smart_nanogrid_gym.envs.smart_nanogrid_environment.SmartNanogridEnv`
(envs/smart_nanogrid_environment.py:31-369): identical constructor keywords, `reset()` returns
`(obs float32[D], {})`, `step(a)` returns `(obs, reward float, terminated bool, False, {})`,
`seed/render/close` exist.  It runs one env (E = 1) of the batched engine in the float64
validation build, so the arithmetic follows the reference's float64 numpy scalars.
It is the compatibility surface, not the fast path -- trainers should drive
`BatchedSmartNanogridEnv` directly.
"""
from __future__ import annotations

import numpy as np
import torch

from .batched_env import BatchedSmartNanogridEnv
from .config import NanogridConfig

ENV_ID = "SmartNanogridEnv-v0"  # smart_nanogrid_gym/__init__.py:4-8
MAX_EPISODE_STEPS = 200


class SmartNanogridEnv:
    metadata = {"render_modes": []}

    def __init__(self, price_model=0, number_of_chargers=8, pv_system_available_in_model=True,
                 battery_system_available_in_model=True, vehicle_to_everything=False,
                 enable_different_vehicle_battery_capacities=True, enable_requested_state_of_charge=False,
                 algorithm_used='', environment_mode='', time_interval='', charging_mode='',
                 vehicle_uncharged_penalty_mode='', *, device="cuda:0", seed=0, precision="float64",
                 restore_requested_soc_on_reload=False):
        self.cfg = NanogridConfig(
            price_model=price_model, number_of_chargers=number_of_chargers,
            pv_system_available_in_model=pv_system_available_in_model,
            battery_system_available_in_model=battery_system_available_in_model,
            vehicle_to_everything=vehicle_to_everything,
            enable_different_vehicle_battery_capacities=enable_different_vehicle_battery_capacities,
            enable_requested_state_of_charge=enable_requested_state_of_charge, algorithm_used=algorithm_used,
            environment_mode=environment_mode, time_interval=time_interval, charging_mode=charging_mode,
            vehicle_uncharged_penalty_mode=vehicle_uncharged_penalty_mode)
        self.NUMBER_OF_CHARGERS = self.cfg.n_spots
        self.TIME_INTERVAL = self.cfg.dt
        self.ALGORITHM_USED, self.ENVIRONMENT_MODE = algorithm_used, environment_mode
        self._restore_req = bool(restore_requested_soc_on_reload)
        # auto_reset off: like the reference, the caller resets after `terminated`
        self._b = BatchedSmartNanogridEnv(1, device=device, seed=seed, precision=precision, auto_reset=False,
                                          config=self.cfg)
        self.action_space = self._b.action_space
        self.observation_space = self._b.observation_space
        self.total_amount_of_states = self.cfg.obs_dim
        self.timestep = None
        self.simulated_single_day = False
        self.info = None
        self._last_plan = None
        dt = torch.float64 if precision == "float64" else torch.float32
        self._a = torch.zeros(1, self.cfg.act_dim, dtype=dt).pin_memory()
        self._o = torch.zeros(1, self.cfg.obs_dim, dtype=torch.float32).pin_memory()
        self._r = torch.zeros(1, dtype=dt).pin_memory()
        self._d = torch.zeros(1, dtype=torch.uint8).pin_memory()

    def reset(self, generate_new_initial_values=True, algorithm_used='', environment_mode='', seed=None, **kwargs):
        """…environment.py:311-351.  `generate_new_initial_values=False` replays the last generated
        schedule like the reference's reload of `initial_values.json`; as in the reference
        (charging_station.py:119-136, quirk Q7) the requested SoC is then lost unless
        `restore_requested_soc_on_reload=True` was given."""
        self.ALGORITHM_USED = algorithm_used if algorithm_used else self.ALGORITHM_USED
        self.ENVIRONMENT_MODE = environment_mode if environment_mode else self.ENVIRONMENT_MODE
        self.timestep = 0
        self.simulated_single_day = False
        if generate_new_initial_values or self._last_plan is None:
            obs = self._b.reset(seed=seed)
            self._last_plan = self._b.sample_plan()
        else:
            import copy
            rec = copy.deepcopy(self._last_plan)
            if not self._restore_req:
                rec.req[:] = 0.0
            st = self._b.env_state()
            # a new pv_shift is drawn on every reset (…environment.py:349); take the sampler's next draw
            obs = self._b.reset(seed=seed)
            shift = self._b.env_state()["pv_shift"]
            obs = self._b.load_schedule(rec, pv_shift=shift, soc_b=st["soc_b"])
        return obs[0].cpu().numpy().copy(), {}

    def step(self, actions):
        """…environment.py:140-188."""
        a = np.asarray(actions, dtype=np.float64).reshape(1, -1)
        self._a.copy_(torch.from_numpy(a).to(self._a.dtype))
        self._b.step_host(self._a, self._o, self._r, self._d)
        self._b.check_errors()
        terminated = bool(self._d[0].item())
        self.timestep = 0 if terminated else self.timestep + 1
        self.simulated_single_day = terminated
        self.info = {}
        return self._o[0].numpy().copy(), float(self._r[0].item()), terminated, False, self.info

    def load_schedule(self, rec, pv_shift=None, soc_b=None):
        """Replay a given schedule (records for one env); returns the reset observation."""
        self.timestep = 0
        obs = self._b.load_schedule(rec, pv_shift=pv_shift, soc_b=soc_b)
        self._last_plan = rec
        return obs[0].cpu().numpy().copy(), {}

    def render(self, mode="human"):
        pass

    def seed(self, seed=None):
        return self._b.seed(seed)

    def close(self):
        self._b.close()


def make(env_id=ENV_ID, **kwargs):
    """Stand-in for gym.make('SmartNanogridEnv-v0', **kwargs) (solvers/RL/ppo_train.py:89)."""
    if env_id != ENV_ID:
        raise ValueError("unknown environment id %r" % env_id)
    return SmartNanogridEnv(**kwargs)


def register_with_gym():
    """Register the id with gymnasium / gym when one of them is installed (neither is in the build image)."""
    for mod in ("gymnasium", "gym"):
        try:
            m = __import__(mod + ".envs.registration", fromlist=["register"])
            m.register(id=ENV_ID, entry_point="smart_nanogrid_gym_b200.env:SmartNanogridEnv",
                       max_episode_steps=MAX_EPISODE_STEPS)
            return mod
        except Exception:  # noqa: BLE001
            continue
    return None
