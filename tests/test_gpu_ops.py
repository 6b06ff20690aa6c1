"""Operator-surface parity on the GPU (through the C ABI): OneBlob, HashGrid, Network, Adam,
sampling and the feature gather, each against the CPU oracle on identical seeded inputs.
Tolerances: bit exact for indices / z / rays; 1e-3 relative (BASELINE.json) for values."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import close, rel_err  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    return torch.device("cuda:0")


def test_library_loaded_from_tree():
    from dns_slam_b200 import _lib
    _lib.lib()
    assert os.path.exists(_lib.LIB_PATH)


@pytest.mark.parametrize("P", [1, 257, 5000])
def test_oneblob(P):
    from oracle import tcnn_standin as otc
    from dns_slam_b200 import tcnn
    dev = _dev()
    g = torch.Generator().manual_seed(P)
    x = torch.rand(P, 3, generator=g) * 1.2 - 0.1
    enc_o = otc.Encoding(3, {"otype": "OneBlob", "n_bins": 16})
    enc = tcnn.Encoding(3, {"otype": "OneBlob", "n_bins": 16})
    xo = x.clone().requires_grad_(True)
    xg = x.to(dev).requires_grad_(True)
    yo, yg = enc_o(xo), enc(xg)
    close(yg, yo, rtol=1e-4, atol=1e-6, name="oneblob fwd")
    w = torch.randn(P, 48, generator=g)
    (yo * w).sum().backward()
    (yg * w.to(dev)).sum().backward()
    # sums of 48 signed terms of magnitude ~30: compare norm-wise (1e-3 bar) and element-wise against the scale
    assert rel_err(xg.grad, xo.grad) < 1e-4
    close(xg.grad, xo.grad, rtol=1e-3, atol=1e-4 * float(xo.grad.abs().max()), name="oneblob bwd")


@pytest.mark.parametrize("hash_size,res", [(13, 124), (16, 592)])
def test_hashgrid(hash_size, res):
    from oracle import tcnn_standin as otc
    from dns_slam_b200 import tcnn, _lib, grid
    import ctypes as C
    dev = _dev()
    g = torch.Generator().manual_seed(res)
    P = 3000
    x = torch.rand(P, 3, generator=g)
    x[:8] = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.5, 0.5, 0.5], [1.0, 0.0, 0.3],
                          [-0.01, 0.2, 1.02], [0.999999, 0.000001, 0.5], [0.25, 0.75, 1.0], [0.0, 1.0, 0.0]])
    cfg = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": hash_size,
           "base_resolution": 16, "per_level_scale": float(np.exp2(np.log2(res / 16) / 15))}
    enc_o = otc.Encoding(3, cfg, seed=3)
    enc = tcnn.Encoding(3, cfg, seed=3)
    # host tables of product and oracle agree
    for k in ("res", "size", "hashed"):
        assert list(enc_o.impl.tables[k]) == list(enc.tables[k])
    assert np.array_equal(np.asarray(enc.tables["scale"], np.float32), enc_o.impl.tables["scale"])
    with torch.no_grad():
        enc_o.params.mul_(3000.0)
        enc.params.copy_(enc_o.params)
    # bit-exact indices
    idx_o, _ = enc_o.impl.corner_indices(x)
    idx_g = torch.empty(P, 16, 8, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().dns_hashgrid_indices(C.byref(enc.gstruct), _lib.ptr(x.to(dev).contiguous()), P,
                                               _lib.ptr(idx_g), _lib.stream()))
    assert torch.equal(idx_g.cpu().to(torch.int64) & 0xFFFFFFFF, idx_o)
    xo = x.clone().requires_grad_(True)
    xg = x.to(dev).requires_grad_(True)
    yo, yg = enc_o(xo), enc(xg)
    close(yg, yo, rtol=1e-4, atol=1e-6, name="grid fwd")
    w = torch.randn(P, 32, generator=g)
    (yo * w).sum().backward()
    (yg * w.to(dev)).sum().backward()
    assert rel_err(enc.params.grad, enc_o.params.grad) < 1e-4
    close(xg.grad, xo.grad, rtol=1e-3, atol=1e-2 * float(xo.grad.abs().mean()), name="grid dx")


@pytest.mark.parametrize("n_in,n_out,P", [(80, 33, 1000), (112, 3, 777), (112, 40, 129), (112, 32, 64), (80, 101, 300)])
def test_network(n_in, n_out, P):
    from oracle import tcnn_standin as otc
    from dns_slam_b200 import tcnn
    dev = _dev()
    g = torch.Generator().manual_seed(n_out)
    cfg = {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 32,
           "n_hidden_layers": 1}
    net_o = otc.Network(n_in, n_out, cfg, seed=7)
    net = tcnn.Network(n_in, n_out, cfg, seed=7)
    close(net.params, net_o.params, rtol=0, atol=0, name="init")
    x = torch.randn(P, n_in, generator=g)
    xo = x.clone().requires_grad_(True)
    xg = x.to(dev).requires_grad_(True)
    yo, yg = net_o(xo), net(xg)
    close(yg, yo, rtol=1e-4, atol=1e-5, name="mlp fwd")
    w = torch.randn(P, n_out, generator=g)
    (yo * w).sum().backward()
    (yg * w.to(dev)).sum().backward()
    assert rel_err(xg.grad, xo.grad) < 1e-4
    close(xg.grad, xo.grad, rtol=1e-3, atol=1e-4 * float(xo.grad.abs().max()), name="mlp dx")
    assert rel_err(net.params.grad, net_o.params.grad) < 1e-4


def test_adam_matches_torch():
    from dns_slam_b200 import fused
    dev = _dev()
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(10007, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-3)
    p = p0.to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        gr = torch.randn(10007, generator=g) * (0.1 if step % 2 else 10.0)
        ref.grad = gr.clone()
        opt.step()
        fused.adam_step(p, gr.to(dev), m, v, 5e-3, step)
    close(p, ref, rtol=1e-5, atol=1e-6, name="adam")


def test_sample_rays_bit_exact(golden_dir):
    """z values, rays and gathered pixels against the reference-generated golden (tracking case)."""
    from oracle import cases
    from dns_slam_b200 import fused, slam, synthetic as syn
    from gpu_util import frame_to
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "tracking_tiny.pt"), weights_only=False)
    meta = g["meta"]
    inp = cases.tracking_inputs(meta)
    cam, s = inp["cam"], syn.SHAPES[meta["shape"]]
    tape = g["tape"]
    fr = frame_to(inp["frame"], dev)
    R = slam.get_rotation_from_quad(g["quad"].to(dev))
    out = fused.sample_rays(cam, inp["bound"], fr, tape[0][1].to(dev), (20, cam["H"] - 20, 20, cam["W"] - 20), R,
                            g["T"].to(dev), meta["n_samples"], meta["n_surface"],
                            fused.fix_surface_draw(tape[1][1], meta["n_surface"]), tape[2][1], want_pts=True)
    gs = g["samples"]
    assert torch.equal(out["gt_label"].cpu(), gs["gt_label"])
    assert torch.equal(out["gt_depth"].cpu(), gs["gt_depth"])
    assert torch.equal(out["gt_color"].cpu(), gs["gt_color"])
    assert torch.equal(out["rays_o"].cpu(), gs["rays_o"])
    # R is produced by torch ops on the GPU; rays_d is bit exact GIVEN R (checked below with the CPU R)
    Rc = slam.get_rotation_from_quad(g["quad"])
    out2 = fused.sample_rays(cam, inp["bound"], fr, tape[0][1].to(dev), (20, cam["H"] - 20, 20, cam["W"] - 20),
                             Rc.to(dev), g["T"].to(dev), meta["n_samples"], meta["n_surface"],
                             fused.fix_surface_draw(tape[1][1], meta["n_surface"]), tape[2][1], want_pts=True)
    assert torch.equal(out2["rays_d"].cpu(), gs["rays_d"])
    assert torch.equal(out2["z_vals"].cpu(), gs["z_vals"])
    assert torch.equal(out2["pts"].cpu(), gs["pts"])
    assert torch.equal(((out2["gt_depth"] > 0.01) * out2["inside"]).cpu(), gs["mask"].bool())
    assert bool((out2["z_vals"][:, 1:] >= out2["z_vals"][:, :-1]).all())


def test_sample_rays_batch_equals_per_frame_calls(golden_dir):
    """dns_sample_rays_batch (all target frames of an iteration in one pair of launches) gives bit for bit what one
    dns_sample_rays call per frame gives: different ray counts per frame (one frame empty), own poses, own scratch."""
    from oracle import cases
    from dns_slam_b200 import fused, slam, synthetic as syn
    from gpu_util import frame_to
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "tracking_tiny.pt"), weights_only=False)
    meta = g["meta"]
    inp = cases.tracking_inputs(meta)
    cam = inp["cam"]
    fr = frame_to(inp["frame"], dev)
    window = (0, cam["H"], 0, cam["W"])
    gen = torch.Generator().manual_seed(3)
    ns, nf = meta["n_samples"], meta["n_surface"]
    S = ns + nf
    counts = (37, 0, 130, 5)
    frames = []
    for f, n in enumerate(counts):
        q = torch.randn(4, generator=gen)
        frames.append(dict(idx=torch.randint(0, cam["H"] * cam["W"], (n,), generator=gen).to(dev),
                           R=slam.get_rotation_from_quad(q).to(dev), T=torch.randn(3, generator=gen).to(dev),
                           ts=fused.fix_surface_draw(torch.rand(nf, generator=gen), nf).to(dev),
                           tz=torch.rand(nf, generator=gen).to(dev)))

    def buffers():
        return [dict(gt_color=torch.full((n, 3), -1.0, device=dev), gt_depth=torch.full((n,), -1.0, device=dev),
                     gt_label=torch.full((n,), -1, dtype=torch.int64, device=dev), rays_o=torch.full((n, 3), -1.0, device=dev),
                     rays_d=torch.full((n, 3), -1.0, device=dev), z_vals=torch.full((n, S), -1.0, device=dev),
                     inside=torch.full((n,), 7, dtype=torch.uint8, device=dev), pixel=torch.full((n,), -1, dtype=torch.int64, device=dev),
                     scratch=torch.full((2,), -1.0, device=dev)) for n in counts]
    one, many = buffers(), buffers()
    pend = []
    for f, n in enumerate(counts):
        for outs, defer in ((one, None), (many, pend)):
            if n == 0 and defer is None:
                continue
            fused.sample_rays(cam, inp["bound"], fr, frames[f]["idx"], window, frames[f]["R"], frames[f]["T"], ns, nf,
                              frames[f]["ts"], frames[f]["tz"], out=outs[f], defer=defer)
    assert len(pend) == len(counts)
    fused.sample_rays_flush(pend)
    assert not pend
    for f, n in enumerate(counts):
        for k in one[f]:
            if n == 0 and k == "scratch":
                continue     # an empty frame is not touched by either path
            assert torch.equal(one[f][k], many[f][k]), (f, k)
    assert float(one[2]["scratch"][0]) > 0.0   # the frame's max depth went through


def test_map_step_result_kernel():
    """dns_map_step_result against the scalar arithmetic it replaces (loss vector, per-frame scratch, running flags)."""
    from dns_slam_b200 import _lib
    dev = _dev()
    L = _lib.lib()
    f32 = torch.float32
    gen = torch.Generator().manual_seed(1)
    F = 4
    losses = torch.rand(8, generator=gen).to(dev)
    sm = torch.rand(1, generator=gen).to(dev)
    scratch = torch.rand(F, 2, generator=gen).to(dev)
    scratch[:, 1] = torch.tensor([0.0, 2.0, 0.0, 1.0])
    loss_vec = torch.full((9,), -5.0, device=dev)
    result = torch.zeros(11 + 2 * F, device=dev)
    result[10 + 2 * F] = 1e9
    w, lam = 0.5, 1e-3
    for rep in range(2):
        _lib.check(L.dns_map_step_result(1, _lib.ptr(losses, f32), _lib.ptr(sm, f32), w, lam * w, 1.0, _lib.ptr(scratch, f32), F,
                                         _lib.ptr(loss_vec, f32), _lib.ptr(result, f32), _lib.stream()))
        want = torch.cat((losses, sm * w)).clone()
        want[6] = losses[6] + (lam * w) * sm[0]
        assert torch.allclose(loss_vec, want, rtol=0, atol=1e-7)
        _lib.check(L.dns_map_step_result(2, None, None, 0.0, 0.0, 0.25, _lib.ptr(scratch, f32), F, _lib.ptr(loss_vec, f32),
                                         _lib.ptr(result, f32), _lib.stream()))
        want[7] = want[7] * 0.25
        assert torch.allclose(result[:9], want, rtol=0, atol=1e-7)
        assert torch.equal(result[9:9 + 2 * F], scratch.reshape(-1))
        assert float(result[9 + 2 * F]) == 3.0 * (rep + 1)           # rays outside the bound accumulate over the steps
        assert abs(float(result[10 + 2 * F]) - float(want[7])) < 1e-7
    # without a TV term, both phases in one call
    _lib.check(L.dns_map_step_result(3, _lib.ptr(losses, f32), None, 0.0, 0.0, 1.0, _lib.ptr(scratch, f32), F,
                                     _lib.ptr(loss_vec, f32), _lib.ptr(result, f32), _lib.stream()))
    assert torch.allclose(result[:8], losses, rtol=0, atol=0)


def test_feature_gather(golden_dir):
    from oracle import cases, reference_path as rp
    from dns_slam_b200 import fused
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "tracking_tiny.pt"), weights_only=False)
    inp = cases.tracking_inputs(g["meta"])
    cam = inp["cam"]
    pts = g["samples"]["pts"].reshape(-1, 3)
    w2c = torch.stack((torch.inverse(inp["poses"][2]), torch.inverse(inp["poses"][3])), 0)
    captured = {}

    def fake_merge(p, o, code):
        captured["code"] = code
        return code.mean(0)[:, :32]

    _, uv_o, mask_o = rp.feature_matching(cam["H"], cam["W"], cam["K"], pts, w2c, inp["feats"], fake_merge)
    code, uv, mask = fused.feature_gather(cam["H"], cam["W"], cam["K"], pts.to(dev), w2c.to(dev),
                                          fused.channels_last(inp["feats"].to(dev)))
    same = (uv.cpu() == uv_o).all(-1) & (mask.cpu() == mask_o)
    # Index work is bit exact EXCEPT where the projected pixel is a rounding tie: the reference's projection is two
    # matmuls (common.py:650-655) whose fp32 summation order belongs to the BLAS in use (cuBLAS there, MKL in the oracle,
    # an fma chain here), so a coordinate within an fp32 ulp-scale of k + 0.5 (or of an image border of the mask test) can
    # round either way.  Every differing (view, point) is shown to be such a tie with a float64 projection.
    P64 = torch.cat((pts.double(), torch.ones(pts.shape[0], 1, dtype=torch.float64)), -1)
    camd = torch.matmul(w2c.double(), P64.t())
    camd = torch.cat((camd[:, 0:1], -camd[:, 1:2], -camd[:, 2:3]), 1)
    img = torch.matmul(cam["K"].double()[None], camd)
    uvd = (img[:, :2, :] / (img[:, 2:3, :] + 1e-5)).permute(0, 2, 1)                       # [R,P,2] before rounding
    frac = (uvd - torch.floor(uvd) - 0.5).abs().min(-1)[0]                                # distance to a rounding tie
    Wd, Hd = cam["W"], cam["H"]
    edge = torch.stack((uvd[..., 0].abs(), (uvd[..., 0] - (Wd - 1)).abs(), uvd[..., 1].abs(), (uvd[..., 1] - (Hd - 1)).abs(),
                        camd[:, 2, :].abs() * 1e3), -1).min(-1)[0]                         # distance to a mask threshold
    tol = 1e-4 * uvd.abs().amax(-1).clamp(min=1.0)                                         # ~fp32 noise of the projection
    tie = (frac < tol) | ((edge - 0.5).abs() < tol) | (edge < tol)
    assert bool(tie[~same].all()), f"{int((~same & ~tie).sum())} differing pixels are not rounding ties"
    assert float(same.float().mean()) > 0.999
    sel = same.unsqueeze(-1).expand_as(captured["code"])
    close(code.cpu()[sel], captured["code"][sel], rtol=1e-4, atol=1e-5, name="feature code")


@pytest.mark.parametrize("shape,n_ids", [((37, 53), 7), ((680, 1200), 40), ((1, 1), 3), ((64, 129), 300)])
def test_class_tables_kernel_is_the_stable_sort(shape, n_ids):
    """a2 (utils/common.py:312-322: torch.unique(label) + torch.nonzero per class) on the device: ``dns_class_tables`` is a
    stable counting sort, bit identical to the torch formulation (ascending classes, ascending pixels inside a class),
    including absent ids, a single-pixel class and more ids than a warp."""
    from dns_slam_b200 import slam
    dev = _dev()
    g = torch.Generator().manual_seed(shape[0] * 1000 + n_ids)
    present = torch.randperm(n_ids, generator=g)[:max(1, (2 * n_ids) // 3)]
    label = present[torch.randint(len(present), shape, generator=g)]
    if label.numel() > 10:
        lone = [c for c in range(n_ids) if c not in present.tolist()]
        if lone:
            label[shape[0] // 2, shape[1] // 3] = lone[0]          # a class with exactly one pixel
    want = slam.class_tables(label)                                # torch formulation on the host
    got = slam.class_tables(label.to(dev), n_ids=n_ids)
    for w, g_, name in zip(want, got, ("classes", "order", "starts", "counts")):
        assert torch.equal(w, g_.cpu()), name
    assert got.classes_h == want.classes_h and got.starts_h == want.starts_h and got.counts_h == want.counts_h


def test_class_tables_kernel_rejects_out_of_range_labels():
    from dns_slam_b200 import slam
    dev = _dev()
    label = torch.randint(5, (16, 16)).to(dev)
    label[3, 3] = 5
    with pytest.raises(ValueError):
        slam.class_tables(label, n_ids=5)
    label[3, 3] = -1
    with pytest.raises(ValueError):
        slam.class_tables(label, n_ids=5)
