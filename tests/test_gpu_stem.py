"""ResNet stem (models/encoder.py:4-17 over models/layers.py:52-114) through ``dns_stem_fwd``:
against the golden vectors of the reference's own encoder (tests/golden/stem_tiny.pt) and, at the Replica frame
size, against the CPU oracle.  Tolerance 1e-4 (fp32 convolution, 147 terms; statistics in fp64)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import close  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    return torch.device("cuda:0")


def _encoder(state, dev):
    from dns_slam_b200 import encoder
    enc = encoder.ResNet().to(dev)
    enc.load_state_dict(state)
    return enc


def test_stem_golden(golden_dir):
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "stem_tiny.pt"), weights_only=False)
    enc = _encoder(g["state0"], dev)
    assert sorted(enc.state_dict().keys()) == sorted(g["state0"].keys())     # same module tree as the reference
    bn = enc.conv_blocks.bn1
    out1 = enc(g["frames1"].to(dev))
    close(out1, g["out1"], rtol=1e-4, atol=1e-5, name="out1")
    for k in ("running_mean", "running_var"):
        close(getattr(bn, k), g["state1"]["conv_blocks.bn1." + k], rtol=1e-5, atol=1e-6, name=k)
    assert int(bn.num_batches_tracked) == int(g["state1"]["conv_blocks.bn1.num_batches_tracked"])
    out2 = enc(g["frames2"].to(dev))
    close(out2, g["out2"], rtol=1e-4, atol=1e-5, name="out2")
    for k in ("running_mean", "running_var"):
        close(getattr(bn, k), g["state2"]["conv_blocks.bn1." + k], rtol=1e-5, atol=1e-6, name=k + "2")
    enc.eval()
    close(enc(g["frames1"].to(dev)), g["out_eval"], rtol=1e-4, atol=1e-5, name="eval")
    # the channels-last form is the same data
    enc.train()
    cl = enc.forward_cl(g["frames2"].to(dev))
    assert cl.shape == (1, 8, 68, 64) and cl.is_contiguous()


@pytest.mark.parametrize("shape", [(2, 680, 1200), (1, 460, 620), (3, 1, 1), (1, 7, 129)])
def test_stem_against_oracle(shape):
    """Replica / ScanNet frame sizes and degenerate ones against the CPU restatement."""
    from oracle import reference_path as rp
    dev = _dev()
    n, H, W = shape
    g = torch.Generator().manual_seed(n * 1000 + H)
    frames = torch.rand(1, n, H, W, 3, generator=g)
    torch.manual_seed(7)
    from dns_slam_b200 import encoder
    enc = encoder.ResNet().to(dev)
    with torch.no_grad():
        enc.conv_blocks.bn1.weight.copy_(torch.rand(64, generator=g) + 0.5)
        enc.conv_blocks.bn1.bias.copy_(torch.randn(64, generator=g) * 0.3)
    s = {k: v.detach().cpu().clone() for k, v in enc.state_dict().items()}
    rm, rv = s["conv_blocks.bn1.running_mean"], s["conv_blocks.bn1.running_var"]
    want = rp.stem_forward(frames, s["conv_blocks.conv1.weight"], s["conv_blocks.bn1.weight"],
                           s["conv_blocks.bn1.bias"], rm, rv, True)
    got = enc(frames.to(dev))
    assert got.shape == want.shape
    if n * ((H - 1) // 2 + 1) * ((W - 1) // 2 + 1) > 1:
        close(got, want, rtol=1e-4, atol=2e-5, name="features")
        close(enc.conv_blocks.bn1.running_mean, rm, rtol=1e-5, atol=1e-6, name="running_mean")
        close(enc.conv_blocks.bn1.running_var, rv, rtol=1e-4, atol=1e-6, name="running_var")


def test_stem_feeds_feature_matching():
    """The channels-last stem output is what the gather consumes: same result as the reference's NCHW route."""
    from dns_slam_b200 import encoder, fused
    dev = _dev()
    torch.manual_seed(3)
    enc = encoder.ResNet().to(dev)
    frames = torch.rand(1, 2, 40, 56, 3, device=dev)
    nchw = enc(frames)[0]
    enc2 = encoder.ResNet().to(dev)
    enc2.load_state_dict({k: v for k, v in enc.state_dict().items()})
    enc2.conv_blocks.bn1.running_mean.zero_(); enc2.conv_blocks.bn1.running_var.fill_(1)
    cl = enc2.forward_cl(frames)
    assert torch.equal(fused.channels_last(nchw), cl)
