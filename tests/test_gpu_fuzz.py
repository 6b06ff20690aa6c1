"""Shape fuzz through the C ABI: on seeded random (mode, rays, samples, classes) the tcgen05 path and the fp32 SIMT
path of dns_render_fwd_bwd must agree (losses, predictions, ray gradients, flat parameter gradient).  Covers the row
tilings of the two-thread ray kernel (T = 128 / 256), ragged last tiles, S = 1 and S = 256, 1..101 classes."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [("map", 64, 5, 1), ("map", 700, 200, 101), ("map", 257, 128, 40), ("map", 1500, 1, 1), ("track", 100, 47, 2),
         ("map", 1500, 129, 101), ("track", 1171, 256, 9), ("track", 700, 2, 2), ("map", 700, 127, 5), ("map", 1, 256, 5),
         ("track", 257, 1, 101), ("map", 1500, 33, 9), ("map", 2, 2, 1), ("track", 3, 13, 40),
         ("map", 300, 5, 128), ("track", 300, 6, 128), ("map", 40, 1, 128)]        # largest per-ray state of the ray kernel


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("mode,N,S,C", CASES)
def test_tensor_core_path_matches_simt_path(mode, N, S, C):
    import contextlib
    from dns_slam_b200 import bench_util, fused, step as stepmod
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    dec, samples = bench_util.synthetic_batch("tiny", mode, N, S, C, dev, seed=N + 7 * S + C, n_frames=1)
    out = {}
    for tc in (0, 1):
        with (contextlib.nullcontext() if tc else fused.simt_path()):     # per-call switch (dns_render_args.use_simt)
            if mode == "map":
                smp = {k: v for k, v in samples.items() if k != "mask"}
                ms = stepmod.MappingStep(dec, 5e-3)
                out[tc] = (ms.forward_backward(smp), ms.grad.clone())
            else:
                out[tc] = (stepmod.TrackingStep(dec).forward_backward(samples), None)
    (o0, g0), (o1, g1) = out[0], out[1]
    if not torch.isfinite(o0[0][:7]).all():          # e.g. a fully masked tracking batch: NaN like the reference, both paths
        assert not torch.isfinite(o1[0][:7]).all()
        return
    assert _rel(o1[0][:7], o0[0][:7]) < 1e-4
    for k in ("color", "depth", "var", "logits"):
        assert _rel(o1[1][k], o0[1][k]) < 1e-4, k
    assert _rel(o1[2], o0[2]) < 3e-3 and _rel(o1[3], o0[3]) < 3e-3       # ReLU-mask flips at ~0 pre-activations
    if g0 is not None:
        for k in ("table", "coarse", "color", "logit", "experts"):
            a, n = dec.layout[k]
            assert _rel(g1[a:a + n], g0[a:a + n]) < 3e-3, k
