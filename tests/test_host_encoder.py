"""Host-side checks of the ResNet stem mirror (no GPU): same module tree / state_dict keys / shapes as the reference's
models/encoder.py ResNet (taken from the golden file its own code produced), the `pretrained` key filter of
models/layers.py:119-133, and no CPU fallback."""
import os

import pytest
import torch


def test_state_dict_matches_reference_module_tree(golden_dir):
    from dns_slam_b200 import encoder
    g = torch.load(os.path.join(golden_dir, "stem_tiny.pt"), weights_only=False)
    enc = encoder.ResNet()
    sd = enc.state_dict()
    assert sorted(sd.keys()) == sorted(g["state0"].keys())
    for k, v in g["state0"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    enc.load_state_dict(g["state0"])          # strict


def test_pretrained_filter_ignores_foreign_keys():
    from dns_slam_b200 import encoder
    w = torch.randn(64, 3, 7, 7)
    full = {"conv1.weight": w, "bn1.weight": torch.full((64,), 2.0), "layer1.0.conv1.weight": torch.randn(64, 64, 3, 3),
            "fc.weight": torch.randn(1000, 512)}
    enc = encoder.ResNet(state_dict=full)
    assert torch.equal(enc.conv_blocks.conv1.weight.detach(), w)
    assert torch.equal(enc.conv_blocks.bn1.weight.detach(), torch.full((64,), 2.0))
    with pytest.raises(RuntimeError):
        encoder.ResNet18(pretrained=True)      # no download here: the weights must be handed in


def test_no_cpu_fallback():
    from dns_slam_b200 import encoder
    enc = encoder.ResNet()
    with pytest.raises(RuntimeError):
        enc(torch.rand(1, 1, 8, 8, 3))
