"""The oracle restatement (oracle/reference_path.py) against golden vectors produced by the
reference's own Python (oracle/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import cases, reference_path as rp
from oracle.make_golden import grad_summary


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.detach().float(), b.detach().float(), rtol=rtol, atol=atol)


def test_small_functions(golden_dir):
    g = load(golden_dir, "kernels.pt")
    close(rp.quad2rotation(g["quad"]), g["R"])
    d, v, rgb, w = rp.raw2nerf_color(g["raw"], g["z"])
    close(d, g["depth_map"]); close(v, g["depth_var"]); close(rgb, g["rgb_map"]); close(w, g["weights"])
    fs, op = rp.get_opacity_loss(g["z"], g["gd"], g["occ"], 0.05)
    close(fs, g["fs"]); close(op, g["op"])
    z = rp.sample_along_rays(g["sar_depth"], 32, 15, g["sar_far"], rp.DrawTape(g["sar_tape"]))
    assert torch.equal(z, g["sar_z"])          # bit exact: drives the hash indices


def test_tracking_iteration(golden_dir):
    g = load(golden_dir, "tracking_tiny.pt")
    o = cases.run_tracking(g["meta"], g["quad"], g["T"], g["tape"])
    # integer / index work and everything feeding the hash indices: bit exact
    for k in ("gt_label", "z_vals", "rays_o", "rays_d", "pts", "gt_depth", "gt_color"):
        assert torch.equal(o["samples"][k].detach(), g["samples"][k]), k
    assert torch.equal(o["samples"]["mask"], g["samples"]["mask"].bool())
    close(o["samples"]["features"], g["samples"]["features"])
    for k in ("color", "depth", "var", "logits"):
        close(o["pred"][k], g["pred"][k])
    for k in ("p", "d", "l", "total"):
        close(o["loss"][k], g["loss"][k])
    for k in ("quad", "T", "coarse", "color", "logit", "merge"):
        close(o["grad"][k], g["grad"][k], rtol=1e-4, atol=1e-7)
    s, gs = grad_summary(o["grad"]["table"]), g["grad"]["table"]
    assert int(s["nnz"]) == int(gs["nnz"])
    close(s["strided"], gs["strided"], rtol=1e-4, atol=1e-9)
    close(s["abssum"], gs["abssum"], rtol=1e-5)


def test_mapping_iteration(golden_dir):
    g = load(golden_dir, "mapping_tiny.pt")
    o = cases.run_mapping(g["meta"], g["quad"], g["T"], g["tape"])
    for k in ("gt_label", "z_vals", "rays_o", "rays_d", "pts", "gt_depth", "gt_color"):
        assert torch.equal(o["samples"][k].detach(), g["samples"][k]), k
    close(o["samples"]["features"], g["samples"]["features"])
    for k in ("color", "depth", "var", "logits", "fine", "coarse"):
        close(o["pred"][k], g["pred"][k])
    for k in ("p", "d", "l", "lt", "sm", "fs", "op", "total"):
        close(o["loss"][k], g["loss"][k])
    for k in ("coarse", "color", "logit", "merge"):
        close(o["grad"][k], g["grad"][k], rtol=1e-4, atol=1e-7)
    for c, ge in g["grad"]["experts"].items():
        if ge is None:
            assert o["grad"]["experts"][c] is None or float(o["grad"]["experts"][c].abs().sum()) == 0
        else:
            close(o["grad"]["experts"][c], ge, rtol=1e-4, atol=1e-7)
    for i in range(len(g["quad"])):
        if g["grad"]["quad"][i] is None:
            assert o["grad"]["quad"][i] is None
        else:
            close(o["grad"]["quad"][i], g["grad"]["quad"][i], rtol=1e-4, atol=1e-7)
            close(o["grad"]["T"][i], g["grad"]["T"][i], rtol=1e-4, atol=1e-7)
    s, gs = grad_summary(o["grad"]["table"]), g["grad"]["table"]
    assert int(s["nnz"]) == int(gs["nnz"])
    close(s["strided"], gs["strided"], rtol=1e-4, atol=1e-9)


def test_hash_tables_match_survey_appendix_b():
    """Level resolutions of SURVEY Appendix B (computed with the reference's own formulas)."""
    import numpy as np
    from oracle.tcnn_standin import grid_level_tables
    t = grid_level_tables(16, 16, np.exp2(np.log2(592 / 16) / 15), 16)
    assert list(t["res"]) == [16, 21, 26, 33, 42, 54, 68, 87, 110, 140, 178, 227, 288, 366, 466, 593]
    assert t["n_entries"] == 853312 and list(t["hashed"]) == [0] * 4 + [1] * 12
    t = grid_level_tables(16, 16, np.exp2(np.log2(231 / 16) / 15), 20)
    assert t["n_entries"] == 7333944 and list(t["hashed"]) == [0] * 11 + [1] * 5


def test_stem(golden_dir):
    """oracle.stem_forward against the reference's own models/encoder.py ResNet (training-mode bn1, two calls,
    then eval mode).
    atol 1e-5 on the features: the convolution's summation order depends on the host thread count."""
    g = load(golden_dir, "stem_tiny.pt")
    s = {k: v.clone() for k, v in g["state0"].items()}
    w, gam, bet = s["conv_blocks.conv1.weight"], s["conv_blocks.bn1.weight"], s["conv_blocks.bn1.bias"]
    rm, rv = s["conv_blocks.bn1.running_mean"], s["conv_blocks.bn1.running_var"]
    close(rp.stem_forward(g["frames1"], w, gam, bet, rm, rv, True), g["out1"], atol=1e-5)
    close(rm, g["state1"]["conv_blocks.bn1.running_mean"]); close(rv, g["state1"]["conv_blocks.bn1.running_var"])
    close(rp.stem_forward(g["frames2"], w, gam, bet, rm, rv, True), g["out2"], atol=1e-5)
    close(rm, g["state2"]["conv_blocks.bn1.running_mean"]); close(rv, g["state2"]["conv_blocks.bn1.running_var"])
    close(rp.stem_forward(g["frames1"], w, gam, bet, rm, rv, False), g["out_eval"], atol=1e-5)


def test_get_2d_feature_restatement_matches_reference(golden_dir):
    """oracle/reference_path.get_2d_feature against the output of the reference's own Mesher.get_2d_feature
    (slams/meshing.py:294-377; oracle/make_golden.py get_2d_feature_case)."""
    import os
    import torch
    from oracle import reference_path as rp
    from oracle.make_golden import build_models
    from dns_slam_b200 import synthetic as syn
    g = torch.load(os.path.join(golden_dir, "get_2d_feature_tiny.pt"), weights_only=False)
    meta = g["meta"]
    gen = torch.Generator().manual_seed(meta["seed"])
    bound, odec, _ = build_models(meta["shape"], meta["n_class"], meta["seed"])
    cam = syn.camera(meta["shape"])
    poses = syn.trajectory(meta["shape"], 6)
    kfs = []
    for i, ft in zip(meta["kf_pose"], g["features"]):
        fr = syn.frame(meta["shape"], poses[i], gen, n_class=meta["n_class"])
        syn.pixel_features(meta["shape"], 1, gen)        # keeps the generator in step with the golden script
        kfs.append({"est_c2w": poses[i].clone(), "gt_label": fr["label"], "gt_depth": fr["depth"], "features": ft})
    with torch.no_grad():
        pix, lab = rp.get_2d_feature(cam, odec, g["points"], kfs)
    assert torch.equal(lab, g["label_pts"])
    torch.testing.assert_close(pix, g["pixel_pts"], rtol=1e-5, atol=1e-6)
