"""CPU-only checks of the drop-in boundary: libsng.so loads and exports every symbol include/sng.h
declares, the ctypes mirrors match the compiled structs, and the host-side configuration mirrors the
reference constructor.  No compute calls (there is no GPU here and the library has no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from smart_nanogrid_gym_b200 import NanogridConfig, _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sng.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sng_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    nat.build()
    L = C.CDLL(nat.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(nat.EXPORTS) == declared


def test_ctypes_mirrors_match_compiled_structs():
    L = nat.lib()
    assert L.sng_abi_version() == 3
    for which, st in enumerate((nat.SngConfig, nat.SngLayout, nat.SngBuffers, nat.SngScheduleView)):
        assert L.sng_sizeof(which) == C.sizeof(st)


@pytest.mark.parametrize("kw,act,obs", [
    (dict(number_of_chargers=10), 11, 29), (dict(number_of_chargers=4), 5, 17),
    (dict(number_of_chargers=64, time_interval="15min"), 65, 137),
    (dict(number_of_chargers=8, pv_system_available_in_model=False, battery_system_available_in_model=False), 8, 20),
])
def test_layout_query_matches_reference_space_shapes(kw, act, obs):
    """sng_query_layout needs no device: dims follow envs/smart_nanogrid_environment.py:90-118."""
    cfg = NanogridConfig(charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse", **kw)
    for prec, real_bytes, plan_bytes, es_bytes in ((nat.SNG_F32, 4, 12, 16), (nat.SNG_F64, 8, 24, 32)):
        c, keep = nat.make_config(cfg, 1000, precision=prec)
        lay = nat.query_layout(c)
        assert (lay.act_dim, lay.obs_dim) == (act, obs) == (cfg.act_dim, cfg.obs_dim)
        assert (lay.real_bytes, lay.plan_rec_bytes, lay.envst_bytes) == (real_bytes, plan_bytes, es_bytes)
        assert lay.env_block == 32 and lay.spot_planes == 3 and lay.plan_slots == 8


def test_create_without_gpu_fails_loudly():
    """No CUDA device -> sng_create returns an error (there is no CPU fallback to fall into)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = NanogridConfig(number_of_chargers=10, charging_mode="bounded", vehicle_uncharged_penalty_mode="sparse")
    c, keep = nat.make_config(cfg, 64)
    h = C.c_void_p()
    rc = nat.lib().sng_create(C.byref(c), 0, C.byref(h))
    assert rc != 0 and b"CUDA" in nat.lib().sng_last_error()
    from smart_nanogrid_gym_b200 import BatchedSmartNanogridEnv
    with pytest.raises(nat.NativeError):
        BatchedSmartNanogridEnv(64, config=cfg)


def test_config_mirrors_reference_constructor_and_quirks():
    cfg = NanogridConfig()
    assert cfg.n_spots == 8 and cfg.charging_mode == "" and cfg.vehicle_uncharged_penalty_mode == ""
    with pytest.raises(ValueError):
        cfg.validate_modes()              # the reference only fails at reset / first charge
    with pytest.raises(ValueError):
        NanogridConfig(price_model=5)     # broken in the reference as well (Q11)
    lo, hi = NanogridConfig(number_of_chargers=4, vehicle_to_everything=True).action_bounds()
    assert lo.tolist() == [-1] * 5 and hi.tolist() == [1] * 5
    c15 = NanogridConfig(number_of_chargers=64, time_interval="15min")
    assert c15.n_steps == 96 and c15.dt == 0.25 and c15.table_len >= 96 + 3
    assert np.isclose(cfg.price.max(), 0.190933333) and np.isclose(cfg.price.min(), 0.114946666)
