"""The fused pixel-feature branch (dns_featmerge_fwd / dns_featmerge_bwd) against the oracle restatement of
utils/common.py:645-679 (feature_matching) + models/decoder.py:67-77 (Merge) + the truncation mask of
slams/tracking.py:167-171: features, the gradient of the Merge weights and the gradient that reaches the rays through the
OneBlob of the points.  Tolerance 1e-3 relative (BASELINE.json north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import frame_to, product_decoder_from_oracle, rel_err  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    return torch.device("cuda:0")


def _case(shape, n_frames, n_views, n_rays, seed, n_class=6):
    """Rays of n_frames synthetic frames sampled by the CUDA sampler + reference views + feature maps."""
    from oracle.make_golden import build_models
    from dns_slam_b200 import fused, synthetic as syn
    dev = _dev()
    gen = torch.Generator().manual_seed(seed)
    bound, odec, _ = build_models(shape, n_class, seed)
    dec = product_decoder_from_oracle(shape, odec, n_class=n_class)
    cam = syn.camera(shape)
    poses = syn.trajectory(shape, 2 * n_frames + 2)
    H, W = cam["H"], cam["W"]
    parts, w2c, feats, ray_start = [], [], [], [0]
    for f in range(n_frames):
        c2w = poses[2 * f + 1]
        fr = frame_to(syn.frame(shape, c2w, gen, n_class=n_class), dev)
        idx = torch.randint(H * W, (n_rays,), generator=gen).to(dev)
        s = fused.sample_rays(cam, dec.bound, fr, idx, (0, H, 0, W), c2w[:3, :3].to(dev), c2w[:3, 3].to(dev), 8, 7,
                              fused.fix_surface_draw(torch.rand(7, generator=gen), 7), torch.rand(7, generator=gen))
        parts.append(s)
        refs = [poses[2 * f], poses[2 * f + 2], c2w][:n_views]
        w2c.append(torch.stack([torch.inverse(p) for p in refs], 0))
        feats.append(syn.pixel_features(shape, n_views, gen))
        ray_start.append(ray_start[-1] + n_rays)
    cat = {k: torch.cat([p[k] for p in parts], 0).contiguous() for k in ("rays_o", "rays_d", "z_vals", "gt_depth")}
    return dict(dev=dev, bound=bound, odec=odec, dec=dec, cam=cam, cat=cat, w2c=w2c, feats=feats, ray_start=ray_start)


def _oracle(c, apply_trunc, weight):
    """features, d(merge params), d(rays_o), d(rays_d) of  sum(features * weight)  through the oracle."""
    from oracle import reference_path as rp
    cam, odec = c["cam"], c["odec"]
    ro = c["cat"]["rays_o"].cpu().requires_grad_(True)
    rd = c["cat"]["rays_d"].cpu().requires_grad_(True)
    z, gd = c["cat"]["z_vals"].cpu(), c["cat"]["gt_depth"].cpu()
    for p in odec.parameters():
        p.grad = None
    outs = []
    for f in range(len(c["w2c"])):
        r0, r1 = c["ray_start"][f], c["ray_start"][f + 1]
        pts = ro[r0:r1, None, :] + rd[r0:r1, None, :] * z[r0:r1, :, None]
        code, _, _ = rp.feature_matching(cam["H"], cam["W"], cam["K"], pts.flatten(0, 1), c["w2c"][f], c["feats"][f], odec.merge)
        code = code.reshape(r1 - r0, z.shape[1], -1)
        if apply_trunc:
            code = code * rp.trunc_mask(z[r0:r1], gd[r0:r1])[..., None]
        outs.append(code)
    feat = torch.cat(outs, 0)
    (feat * weight).sum().backward()
    return feat.detach(), odec.merge.decoder.params.grad.clone(), ro.grad, rd.grad


@pytest.mark.parametrize("n_frames,n_views,apply_trunc", [(1, 2, True), (2, 3, True), (1, 1, False), (3, 3, True)])
def test_featmerge_vs_oracle(n_frames, n_views, apply_trunc):
    from dns_slam_b200 import fused
    c = _case("tiny", n_frames, n_views, 37, seed=5 + n_frames)
    dev, dec, cat = c["dev"], c["dec"], c["cat"]
    w2c = torch.cat(c["w2c"], 0).to(dev)
    cam_o = torch.inverse(torch.cat(c["w2c"], 0))[:, :3, 3].to(dev)
    views = fused.Views(w2c, cam_o, [fused.channels_last(f.to(dev)) for f in c["feats"]], c["ray_start"])
    ro = cat["rays_o"].clone().requires_grad_(True)
    rd = cat["rays_d"].clone().requires_grad_(True)
    mp = dec.merge.decoder.params
    mp.grad = None
    feat = fused.feature_merge(c["cam"], dec.merge.bound, views, ro, rd, cat["z_vals"], cat["gt_depth"], mp, apply_trunc)
    weight = torch.randn(feat.shape, generator=torch.Generator().manual_seed(9))
    (feat * weight.to(dev)).sum().backward()
    of, og, oro, ord_ = _oracle(c, apply_trunc, weight)
    # exact zeros outside the band; 1e-3 inside (a handful of samples sit on a rounding tie of the projected pixel:
    # those rows differ in WHICH pixel is fetched, not in arithmetic -- bounded separately below)
    got = feat.detach().cpu()
    assert torch.equal(got == 0, of == 0) or ((got == 0) != (of == 0)).float().mean() < 1e-3
    row_err = (got - of).flatten(0, 1).norm(dim=-1) / (of.flatten(0, 1).norm(dim=-1) + 1e-6)
    bad = row_err > 1e-3
    assert bad.float().mean() < 2e-3, f"{int(bad.sum())} of {bad.numel()} rows differ"
    if not bool(bad.any()):
        assert rel_err(mp.grad, og) < 1e-3
        assert rel_err(ro.grad, oro) < 1e-3
        assert rel_err(rd.grad, ord_) < 1e-3
    else:   # tie rows change the gathered feature, which feeds dW1 (not the ray gradients' OneBlob columns much)
        assert rel_err(mp.grad, og) < 2e-2
        assert rel_err(ro.grad, oro) < 2e-2


def test_featmerge_matches_operator_chain_at_replica_size():
    """Mid-size Replica-shaped batch: the fused branch against the operator kernels it replaces (dns_feature_gather +
    dns_merge_fwd/bwd + truncation mask in torch) -- same arithmetic, band compaction and weight-gradient accumulation in
    TMEM being the differences."""
    from dns_slam_b200 import fused, slam
    c = _case("replica", 2, 3, 1500, seed=21, n_class=40)
    dev, dec, cat = c["dev"], c["dec"], c["cat"]
    w2c = torch.cat(c["w2c"], 0).to(dev)
    c2w = torch.inverse(torch.cat(c["w2c"], 0)).to(dev)
    feats = [fused.channels_last(f.to(dev)) for f in c["feats"]]
    views = fused.Views(w2c, c2w[:, :3, 3].contiguous(), feats, c["ray_start"])
    mp = dec.merge.decoder.params
    weight = torch.randn(cat["z_vals"].shape + (32,), generator=torch.Generator().manual_seed(2)).to(dev)
    res = []
    for fused_path in (True, False):
        ro = cat["rays_o"].clone().requires_grad_(True)
        rd = cat["rays_d"].clone().requires_grad_(True)
        mp.grad = None
        if fused_path:
            feat = fused.feature_merge(c["cam"], dec.merge.bound, views, ro, rd, cat["z_vals"], cat["gt_depth"], mp)
        else:
            parts = []
            for f in range(2):
                r0, r1 = c["ray_start"][f], c["ray_start"][f + 1]
                pts = ro[r0:r1, None, :] + rd[r0:r1, None, :] * cat["z_vals"][r0:r1, :, None]
                code = fused.feature_matching(c["cam"]["H"], c["cam"]["W"], c["cam"]["K"].to(dev), pts.flatten(0, 1),
                                              w2c[3 * f:3 * f + 3], feats[f], dec.merge, refer_c2w=c2w[3 * f:3 * f + 3])
                parts.append(code.reshape(r1 - r0, -1, 32) * slam.trunc_mask(cat["z_vals"][r0:r1], cat["gt_depth"][r0:r1])[..., None])
            feat = torch.cat(parts, 0)
        (feat * weight).sum().backward()
        res.append((feat.detach(), mp.grad.clone(), ro.grad.clone(), rd.grad.clone()))
    (f1, g1, o1, d1), (f0, g0, o0, d0) = res
    assert float((f1 != 0).float().mean()) > 0.2, "band is empty: the case tests nothing"
    assert rel_err(f1, f0) < 1e-5
    assert rel_err(g1, g0) < 1e-4
    assert rel_err(o1, o0) < 1e-4 and rel_err(d1, d0) < 1e-4


def test_featmerge_backward_same_with_stash_and_recompute():
    """The backward either bulk-copies the operand tiles the forward stashed or gathers + encodes again: same gradients;
    a stash that is too small for the band falls back to recomputing (decided on the device)."""
    from dns_slam_b200 import fused
    c = _case("tiny", 2, 3, 400, seed=31)
    dev, dec, cat = c["dev"], c["dec"], c["cat"]
    w2c = torch.cat(c["w2c"], 0).to(dev)
    cam_o = torch.inverse(torch.cat(c["w2c"], 0))[:, :3, 3].to(dev)
    views = fused.Views(w2c, cam_o, [fused.channels_last(f.to(dev)) for f in c["feats"]], c["ray_start"])
    mp = dec.merge.decoder.params.detach()
    N, S = cat["z_vals"].shape
    d_f = torch.randn(N, S, 32, generator=torch.Generator().manual_seed(3)).to(dev)
    res = []
    for stash in (None, fused.featmerge_stash(N, S, 3, dev), torch.empty(fused.FEATMERGE_TILE_BYTES, dtype=torch.uint8, device=dev)):
        feat, ws = fused.featmerge_raw(c["cam"], dec.merge.bound, views, cat["rays_o"], cat["rays_d"], cat["z_vals"],
                                       cat["gt_depth"], mp, stash=stash)
        d_p, d_o, d_d = torch.zeros_like(mp), torch.zeros(N, 3, device=dev), torch.zeros(N, 3, device=dev)
        fused.featmerge_bwd_raw(c["cam"], dec.merge.bound, views, cat["rays_o"], cat["rays_d"], cat["z_vals"], cat["gt_depth"],
                                mp, d_f, ws, d_p, d_o, d_d, stash=stash)
        res.append((feat, d_p, d_o, d_d))
    assert int(ws[:4].view(torch.int32)[0]) > 42, "band too small to exercise more than one tile"
    for other in res[1:]:
        assert torch.equal(other[0], res[0][0])
        for a, b in zip(other[1:], res[0][1:]):
            assert rel_err(a, b) < 1e-5
