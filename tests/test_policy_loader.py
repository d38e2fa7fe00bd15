"""The reference's shipped PPO checkpoint (solvers/RL/models/PPO-b-pv-bounded-sparse-4ch-1h/999600.zip, loaded by
solvers/predictor.py:72) maps losslessly onto rollout.MlpPolicy.  CPU only."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ZIP = "/root/reference/solvers/RL/models/PPO-b-pv-bounded-sparse-4ch-1h/999600.zip"


def _fixture():
    z = np.load(os.path.join(GOLD, "sb3_ppo_4ch_policy.npz"))
    return {k: torch.tensor(z[k]) for k in z.files}


def _sb3_forward_f64(sd, obs):
    """What SB3's ActorCriticPolicy computes for MlpPolicy (mlp_extractor -> action_net / value_net), in float64."""
    w = {k: v.double().numpy() for k, v in sd.items()}
    x = obs.astype(np.float64)
    h = np.tanh(x @ w["mlp_extractor.policy_net.0.weight"].T + w["mlp_extractor.policy_net.0.bias"])
    h = np.tanh(h @ w["mlp_extractor.policy_net.2.weight"].T + w["mlp_extractor.policy_net.2.bias"])
    mean = h @ w["action_net.weight"].T + w["action_net.bias"]
    g = np.tanh(x @ w["mlp_extractor.value_net.0.weight"].T + w["mlp_extractor.value_net.0.bias"])
    g = np.tanh(g @ w["mlp_extractor.value_net.2.weight"].T + w["mlp_extractor.value_net.2.bias"])
    return mean, (g @ w["value_net.weight"].T + w["value_net.bias"])[:, 0]


def test_shipped_checkpoint_maps_losslessly():
    from smart_nanogrid_gym_b200.rollout import MlpPolicy
    sd = _fixture()
    pol = MlpPolicy.from_sb3_state_dict(sd)
    assert (pol.pi[0].in_features, pol.pi[0].out_features, pol.action_net.out_features) == (17, 64, 5)   # N = 4: D = 17, A = 5
    back = pol.to_sb3_state_dict()
    assert sorted(back) == sorted(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k                        # bit for bit
    assert sum(p.numel() for p in pol.parameters()) == sum(v.numel() for v in sd.values())   # nothing left unmapped
    obs = np.random.default_rng(0).random((64, 17)).astype(np.float32)
    mean_ref, val_ref = _sb3_forward_f64(sd, obs)
    with torch.no_grad():
        mean, val, _ = pol(torch.tensor(obs), None)
    assert np.allclose(mean.numpy(), mean_ref, rtol=1e-5, atol=1e-5) and np.allclose(val.numpy(), val_ref, rtol=1e-5, atol=1e-4)


def test_malformed_state_dicts_are_rejected():
    from smart_nanogrid_gym_b200.rollout import MlpPolicy
    sd = _fixture()
    bad = dict(sd)
    del bad["log_std"]
    with pytest.raises(ValueError):
        MlpPolicy.from_sb3_state_dict(bad)
    bad = dict(sd)
    bad["features_extractor.cnn.0.weight"] = torch.zeros(1)
    with pytest.raises(ValueError):
        MlpPolicy.from_sb3_state_dict(bad)
    bad = dict(sd)
    bad["action_net.bias"] = torch.zeros(6)
    with pytest.raises(ValueError):
        MlpPolicy.from_sb3_state_dict(bad)


@pytest.mark.skipif(not os.path.exists(ZIP), reason="reference tree not present (GPU box)")
def test_zip_loader_equals_the_committed_fixture():
    from smart_nanogrid_gym_b200.rollout import MlpPolicy
    pol = MlpPolicy.from_sb3_zip(ZIP)
    sd = _fixture()
    back = pol.to_sb3_state_dict()
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
