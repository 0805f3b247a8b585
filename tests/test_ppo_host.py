"""CPU suite: the PPO network mirrors the reference's (ppo.py:91-131) — shipped 10-PM weights load (with the
`_orig_mod.` prefix of ppo.py:142,166) and reproduce the reference's own forward values (SURVEY §8c anchors, measured
with the unmodified reference Network in eager fp32 on CPU)."""
import os

import numpy as np
import pytest
import torch

WEIGHTS = "/root/reference/weights-10"
ANCHORS = {   # file: (value at reset obs, logits[0:4] at reset obs, sum of logits at reset obs)
    "ppo-wr.pt": (-52.131874, (-12.823983, -20.025423, -21.213934, -17.742081), -2890.28714),
    "ppo-ut.pt": (650.797668, (10.212225, 11.898891, 10.473968, 9.680696), -3209.63155),
    "ppo-kl.pt": (-81.077576, (-11.224595, -9.200848, -14.120452, -14.593458), -2782.41826),
}


@pytest.mark.parametrize("name", sorted(ANCHORS))
def test_shipped_weights_reproduce_reference_forward(name):
    path = os.path.join(WEIGHTS, name)
    if not os.path.exists(path):
        pytest.skip("reference weights not available on this box")
    from vmgym.ppo import Network
    net = Network(110, 30, 12, 512)
    sd = torch.load(path, map_location="cpu")
    assert all(k.startswith("_orig_mod.") for k in sd)
    net.load_state_dict({k[len("_orig_mod."):]: v for k, v in sd.items()})
    obs = torch.zeros(1, 110)
    obs[0, :30] = 11.0                       # reset observation of config/10.yml: every slot empty (P+1 = 11)
    with torch.no_grad():
        value = net.get_value(obs).item()
        logits = net.actor(obs)[0]
    want_v, want_l, want_sum = ANCHORS[name]
    assert value == pytest.approx(want_v, abs=2e-4)
    assert np.allclose(logits[:4].numpy(), want_l, atol=2e-4)
    assert logits.sum().item() == pytest.approx(want_sum, abs=5e-2)
    assert net.get_det_action(obs).shape == (1, 30)


def test_network_shape_and_init():
    from vmgym.ppo import Network
    torch.manual_seed(0)
    net = Network(1100, 300, 102, 512)
    n_actor = sum(p.numel() for p in net.actor.parameters())
    n_critic = sum(p.numel() for p in net.critic.parameters())
    assert (n_actor, n_critic) == (16524168, 826881)
    w = net.actor[0].weight
    assert torch.allclose(w @ w.T, 2.0 * torch.eye(512), atol=1e-3)      # orthogonal rows, gain sqrt(2)
    assert float(net.actor[4].weight.norm()) < float(net.critic[4].weight.norm()) * 200
    assert all(float(m.bias.abs().max()) == 0.0 for m in net.modules() if isinstance(m, torch.nn.Linear))
