"""Inference path against the CPU oracle: the full-frame render of Mapper.frame_vis
(slams/mapping.py:636-690) and the free-point query of Mesher.eval_points (slams/meshing.py:461-498), both
through ``dns_render_fwd_bwd`` with ``forward_only`` set.  Colours / depth / occupancy within 1e-3; labels must
agree wherever the oracle's top-2 logit margin is not at rounding level."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import close, frame_to, product_decoder_from_oracle  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    return torch.device("cuda:0")


def _inputs():
    from oracle import cases
    meta = dict(shape="tiny", n_class=5, seed=11, tgt_ids=[1, 3], refer_idx=[[0, 2, -1], [2, 4, -1]])
    return meta, cases.mapping_inputs(meta)


def _labels_agree(got, logits_want, name):
    want = torch.argmax(logits_want, -1)
    top2 = torch.topk(logits_want, 2, -1)[0]
    sure = (top2[..., 0] - top2[..., 1]) > 1e-3 * top2.abs().amax(-1).clamp_min(1e-6)
    assert bool((got.cpu()[sure] == want[sure]).all()), name
    assert float(sure.float().mean()) > 0.9


def test_render_frame_vs_oracle():
    from oracle import reference_path as rp
    from dns_slam_b200 import fused, inference
    dev = _dev()
    meta, inp = _inputs()
    cam, bound, odec, oexp = inp["cam"], inp["bound"], inp["decoder"], inp["experts"]
    dec = product_decoder_from_oracle(meta["shape"], odec, oexp, n_class=meta["n_class"])
    frame, c2w = inp["frames"][0], inp["poses"][1]
    refer_w2c = torch.inverse(inp["poses"][0])
    feats = inp["feats"][0][:1]
    tape = rp.DrawTape(seed=5)
    n_batch = 173                                          # ragged chunks, so the per-chunk class rule is exercised
    with torch.no_grad():
        want = rp.frame_vis_render(cam, bound, odec, oexp, frame, c2w, refer_w2c[None], feats, 8, 5, tape, n_batch)
    t_surface, t_zero = tape.items[0][1], tape.items[1][1]
    got = inference.render_frame(cam, dec, frame_to(frame, dev), c2w, refer_w2c, fused.channels_last(feats.to(dev)),
                                 8, 5, t_surface, t_zero, n_pts_batch=n_batch)
    close(got[0], want[0], rtol=1e-3, atol=1e-4, name="color")
    close(got[1], want[1], rtol=1e-3, atol=1e-4, name="depth")
    assert got[2].shape == want[2].shape and got[2].dtype == torch.int64
    assert float((got[2].cpu() == want[2]).float().mean()) > 0.98      # argmax ties at rounding level aside


def test_eval_points_vs_oracle():
    from oracle import reference_path as rp
    from dns_slam_b200 import inference
    dev = _dev()
    meta, inp = _inputs()
    bound, odec, oexp = inp["bound"], inp["decoder"], inp["experts"]
    dec = product_decoder_from_oracle(meta["shape"], odec, oexp, n_class=meta["n_class"])
    g = torch.Generator().manual_seed(3)
    P = 1000
    lo, hi = bound[:, 0].float(), bound[:, 1].float()
    pts = lo + (hi - lo) * (torch.rand(P, 3, generator=g) * 1.2 - 0.1)          # ~40 % outside the bound
    pix = torch.randn(P, 32, generator=g) * 0.3
    lab = torch.randint(0, meta["n_class"], (P,), generator=g)
    for stage in ("fine", "coarse"):
        with torch.no_grad():
            wv, wl = rp.eval_points(odec, oexp, bound, pts.to(bound.dtype), pix, lab, stage)
        gv, gl = inference.eval_points(dec, pts.to(dev), pix.to(dev), lab.to(dev), stage)
        close(gv, wv, rtol=1e-3, atol=1e-4, name=f"values {stage}")
        if stage == "fine":
            inside = wl >= 0
            assert torch.equal(gl.cpu() >= 0, inside)
            with torch.no_grad():
                _, logits = odec.out_fn(*_pe_and_feat(odec, oexp, bound, pts, pix, lab))
            _labels_agree(gl.cpu()[inside], logits[inside], "labels")
        else:
            assert gl is None


def _pe_and_feat(odec, oexp, bound, pts, pix, lab):
    from oracle import reference_path as rp
    p = (pts.to(bound.dtype) - bound[:, 0]) / (bound[:, 1] - bound[:, 0])
    pe, grid = odec.pe_fn(p)
    lat = rp.fine_fn(oexp, odec.hidden_dim, pe, lab, grid)
    return pe, torch.cat((lat[:, 1:], pix), -1)


def test_eval_points_missing_expert_is_an_error():
    from dns_slam_b200 import inference
    dev = _dev()
    meta, inp = _inputs()
    experts = {c: e for c, e in inp["experts"].items() if c != 2}
    dec = product_decoder_from_oracle(meta["shape"], inp["decoder"], experts, n_class=meta["n_class"])
    pts = torch.zeros(64, 3, device=dev)
    with pytest.raises(ValueError):
        inference.eval_points(dec, pts, torch.zeros(64, 32, device=dev), torch.full((64,), 2, device=dev), "fine")


def test_get_2d_feature_vs_reference_golden(golden_dir):
    """``inference.get_2d_feature`` (the producer of eval_points' pixel features in a mesh extraction) against the output
    of the reference's own Mesher.get_2d_feature (slams/meshing.py:294-377, tests/golden/get_2d_feature_tiny.pt): labels
    bit exact, merged features 1e-3."""
    import os
    from oracle.make_golden import build_models
    from dns_slam_b200 import fused, inference, synthetic as syn
    from gpu_util import product_decoder_from_oracle
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(golden_dir, "get_2d_feature_tiny.pt"), weights_only=False)
    meta = g["meta"]
    gen = torch.Generator().manual_seed(meta["seed"])
    bound, odec, _ = build_models(meta["shape"], meta["n_class"], meta["seed"])
    dec = product_decoder_from_oracle(meta["shape"], odec, n_class=meta["n_class"])
    cam = syn.camera(meta["shape"])
    poses = syn.trajectory(meta["shape"], 6)
    kfs = []
    for i, ft in zip(meta["kf_pose"], g["features"]):
        fr = syn.frame(meta["shape"], poses[i], gen, n_class=meta["n_class"])
        syn.pixel_features(meta["shape"], 1, gen)
        kfs.append({"est_c2w": poses[i].clone(), "gt_label": fr["label"].to(dev), "gt_depth": fr["depth"].to(dev),
                    "features_cl": fused.channels_last(ft.to(dev))})
    pix, lab = inference.get_2d_feature(cam, dec, g["points"].to(dev), kfs)
    same = lab.cpu() == g["label_pts"]
    # a projected pixel on a rounding tie (or a point on the rim of a key frame's image) may fall either way: the
    # projection is a batched fp32 matmul whose summation order differs between the CPU and the GPU BLAS
    assert float(same.float().mean()) > 0.995
    want = g["pixel_pts"]
    # per point: a key frame that sees the point on one side of a tie and not on the other changes the AVERAGE over the key
    # frames, so the bound is per row: 99 % of the points within 1e-3, median at fp32 noise
    got = pix.cpu()
    assert float((((got != 0).any(-1)) == ((want != 0).any(-1))).float().mean()) > 0.995
    err = ((got - want).norm(dim=-1) / (want.norm(dim=-1) + 1e-6))[(want != 0).any(-1)].sort()[0]
    assert float(err[len(err) // 2]) < 1e-5 and float(err[int(len(err) * 0.99)]) < 1e-3, (float(err[len(err) // 2]), float(err[int(len(err) * 0.99)]))
