"""CPU tier: the reference's checkpoint format (pql/utils/model_util.py:24-41,
evaluator.py:112-119).  The state_dict key names recorded from the reference's own modules
(tests/golden/learner_*.npz) must be exactly the keys of the pql_b200 modules, and a checkpoint
round-trips through torch.save / torch.load into fresh modules."""
import os

import numpy as np
import torch

from pql_b200.models import DistributionalDoubleQ, DoubleQ, TanhMLPPolicy
from pql_b200.utils import load_model, save_model


def _ref_names(golden_dir, tag, prefix):
    g = np.load(os.path.join(golden_dir, f"learner_{tag}.npz"))
    return sorted(k[len(prefix):] for k in g.files if k.startswith(prefix))


def test_state_dict_keys_are_the_references(golden_dir):
    assert sorted(DoubleQ(88, 16).state_dict()) == _ref_names(golden_dir, "doubleq", "critic.")
    assert sorted(TanhMLPPolicy(88, 16).state_dict()) == _ref_names(golden_dir, "doubleq", "actor.")
    assert sorted(DistributionalDoubleQ(88, 16, device="cpu").state_dict()) == _ref_names(golden_dir, "c51", "critic.")
    sd = DoubleQ(211, 20).state_dict()
    assert sd["net_q1.net.0.weight"].shape == (512, 231) and sd["net_q2.net.6.weight"].shape == (1, 128)


def test_checkpoint_round_trip(tmp_path):
    torch.manual_seed(0)
    actor, critic = TanhMLPPolicy(9, 3), DoubleQ(9, 3)
    rms = (torch.randn(9), torch.rand(9) + 0.5, 1e-4)
    path = str(tmp_path / "model.pth")
    ck = save_model(path, actor.state_dict(), critic.state_dict(), rms, wandb_run=None, description="best")
    assert sorted(ck) == ["actor", "critic", "obs_rms"]
    raw = torch.load(path, map_location="cpu")
    assert sorted(raw["critic"]) == sorted(critic.state_dict()) and torch.equal(raw["obs_rms"][0], rms[0])
    a2, c2 = TanhMLPPolicy(9, 3), DoubleQ(9, 3)
    assert load_model(a2, "actor", path) and load_model(c2, "critic", path)
    for (k, v), (k2, v2) in zip(actor.state_dict().items(), a2.state_dict().items()):
        assert k == k2 and torch.equal(v, v2)
    # the flat kernel arena follows load_state_dict (the kernels read the arena, not the parameters)
    assert torch.equal(c2.arena.flat, critic.arena.flat)
    assert not load_model(a2, "actor_3", path)
    save_model(path, actor, critic, None)
    assert not load_model(object(), "obs_rms", path)
