"""Fused-path parity on the GPU (through the C ABI): one tracking iteration and one mapping
iteration against (a) the golden vectors produced by the reference's own Python and (b) the CPU
oracle on the same seeded inputs.  Bit exact for indices / z / rays; <= 1e-3 relative
(BASELINE.json north_star) for rendered rgb / depth, losses and gradients."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import close, frame_to, product_decoder_from_oracle, rel_err, split_mapping_tape  # noqa: E402

TOL = 1e-3


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    return torch.device("cuda:0")


def _check_grad(name, got, want, tol=TOL):
    e = rel_err(got, want)
    assert e < tol, f"{name}: relative error {e:.3e}"


def test_tracking_iteration_vs_golden(golden_dir):
    from oracle import cases
    from dns_slam_b200 import fused, slam, synthetic as syn
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "tracking_tiny.pt"), weights_only=False)
    meta = g["meta"]
    s = syn.SHAPES[meta["shape"]]
    inp = cases.tracking_inputs(meta)
    dec = product_decoder_from_oracle(meta["shape"], inp["decoder"], n_class=meta["n_class"])
    trk = slam.TrackerCore(inp["cam"], dec, s["tracking_pixels"], meta["n_samples"], meta["n_surface"],
                           s["lambda_color"], s["lambda_depth"], s["lambda_label"])
    quad = g["quad"].to(dev).requires_grad_(True)
    T = g["T"].to(dev).requires_grad_(True)
    cur_c2w = slam.c2w_from_quad_T(quad, T)
    est_w2c = torch.stack((torch.inverse(inp["poses"][meta["refer_index"]]).to(dev), torch.inverse(cur_c2w)), 0)
    cur = dict(frame_to(inp["frame"], dev), est_quad=quad, est_T=T)
    tape = g["tape"]
    draws = dict(idx=tape[0][1], t_surface=tape[1][1], t_zero=tape[2][1])
    ld, preds, samples = trk.iteration(cur, {"est_w2c": est_w2c}, fused.channels_last(inp["feats"].to(dev)), draws)
    ld["total"].backward()
    gs = g["samples"]
    for k in ("gt_label", "gt_depth", "gt_color", "z_vals"):
        assert torch.equal(samples[k].detach().cpu(), gs[k]), k
    assert torch.equal(samples["mask"].cpu(), gs["mask"].bool())
    close(samples["rays_d"], gs["rays_d"], rtol=1e-6, atol=1e-6, name="rays_d")
    close(samples["features"], gs["features"], rtol=TOL, atol=1e-4, name="features")
    for k in ("color", "depth", "var", "logits"):
        close(preds[k], g["pred"][k], rtol=TOL, atol=1e-5, name=k)
    for k, gk in (("p_loss", "p"), ("d_loss", "d"), ("l_loss", "l"), ("total", "total")):
        close(ld[k], g["loss"][gk], rtol=TOL, atol=1e-6, name=k)
    _check_grad("quad", quad.grad, g["grad"]["quad"])
    _check_grad("T", T.grad, g["grad"]["T"])
    _check_grad("coarse", dec.coarse_fn.decoder.params.grad, g["grad"]["coarse"])
    _check_grad("color", dec.out_fn.color_decoder.params.grad, g["grad"]["color"])
    _check_grad("logit", dec.out_fn.logit_decoder.params.grad, g["grad"]["logit"])
    _check_grad("merge", dec.merge.decoder.params.grad, g["grad"]["merge"])
    tg = dec.pe_fn.grid_fn.params.grad.cpu()
    gt = g["grad"]["table"]
    assert int((tg != 0).sum()) == int(gt["nnz"])
    _check_grad("table (strided)", tg[::int(gt["stride"])], gt["strided"])
    # and the full table gradient against the oracle run
    o = cases.run_tracking(meta, g["quad"], g["T"], tape, inp)
    _check_grad("table", tg, o["grad"]["table"])


def test_mapping_iteration_vs_golden(golden_dir):
    from oracle import cases
    from dns_slam_b200 import fused, slam, synthetic as syn
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "mapping_tiny.pt"), weights_only=False)
    meta = g["meta"]
    s = syn.SHAPES[meta["shape"]]
    inp = cases.mapping_inputs(meta)
    dec = product_decoder_from_oracle(meta["shape"], inp["decoder"], inp["experts"], n_class=meta["n_class"])
    mp = slam.MapperCore(inp["cam"], dec, s["mapping_pixels"], meta["n_samples"], meta["n_surface"],
                         lambdas=dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"],
                                      lt=meta["lambda_lt"], fs=s["lambda_fs"], op=s["lambda_opacity"]),
                         opacity_sigma=s["opacity_sigma"], smooth_pts=s["smooth_pts"], lambda_sm=meta["lambda_sm"])
    quad_list = [q.to(dev).requires_grad_(i != 0) for i, q in enumerate(g["quad"])]
    T_list = [t.to(dev).requires_grad_(i != 0) for i, t in enumerate(g["T"])]
    frames = [frame_to(f, dev) for f in inp["frames"]]
    draws, tv_draws = split_mapping_tape(g["tape"], inp["frames"], s["mapping_pixels"] // len(frames))
    target = dict(kf_idx=meta["tgt_ids"], frames=frames,
                  class_tables=[slam.class_tables(f["label"]) for f in frames])
    refer = dict(kf_idx=meta["refer_idx"], est_c2w=[[c.to(dev) for c in row] for row in inp["refer_c2w"]])
    feats = [fused.channels_last(f.to(dev)) for f in inp["feats"]]
    ld, preds, samples = mp.iteration(target, quad_list, T_list, refer, feats, draws, tv_draws, want_latents=True)
    ld["total"].backward()
    gs = g["samples"]
    for k in ("gt_label", "gt_depth", "gt_color", "z_vals"):
        assert torch.equal(samples[k].detach().cpu(), gs[k]), k
    close(samples["rays_d"], gs["rays_d"], rtol=1e-6, atol=1e-6, name="rays_d")
    close(samples["features"], gs["features"], rtol=TOL, atol=1e-4, name="features")
    for k in ("color", "depth", "var", "logits", "fine", "coarse"):
        close(preds[k], g["pred"][k], rtol=TOL, atol=2e-5, name=k)
    for k, gk in (("p_loss", "p"), ("d_loss", "d"), ("l_loss", "l"), ("lt_loss", "lt"), ("fs_loss", "fs"),
                  ("opacity_loss", "op"), ("smooth_loss", "sm"), ("total", "total")):
        close(ld[k], g["loss"][gk], rtol=TOL, atol=1e-7, name=k)
    _check_grad("coarse", dec.coarse_fn.decoder.params.grad, g["grad"]["coarse"])
    _check_grad("color", dec.out_fn.color_decoder.params.grad, g["grad"]["color"])
    _check_grad("logit", dec.out_fn.logit_decoder.params.grad, g["grad"]["logit"])
    _check_grad("merge", dec.merge.decoder.params.grad, g["grad"]["merge"])
    eg = dec.expert_params.grad
    for c, ge in g["grad"]["experts"].items():
        if ge is None:
            assert float(eg[c].abs().sum()) == 0.0
        else:
            _check_grad(f"expert {c}", eg[c], ge)
    for i in range(1, len(quad_list)):
        _check_grad(f"quad[{i}]", quad_list[i].grad, g["grad"]["quad"][i])
        _check_grad(f"T[{i}]", T_list[i].grad, g["grad"]["T"][i])
    o = cases.run_mapping(meta, g["quad"], g["T"], g["tape"], inp)
    _check_grad("table", dec.pe_fn.grid_fn.params.grad, o["grad"]["table"])


def test_tv_alone_vs_oracle():
    from oracle import reference_path as rp
    from oracle.make_golden import build_models
    from dns_slam_b200 import fused
    dev = _dev()
    bound, odec, _ = build_models("tiny", 6, 77)
    dec = product_decoder_from_oracle("tiny", odec, n_class=6)
    g = torch.Generator().manual_seed(5)
    r3, r113 = torch.rand(3, generator=g), torch.rand(1, 1, 1, 3, generator=g)
    for sp in (8, 12):
        tape = rp.DrawTape([("rand", r3), ("rand", r113)])
        for p in odec.parameters():
            p.grad = None
        lo = rp.smoothness(odec, bound, sp, tape)
        lo.backward()
        dec.zero_grad()
        lg = fused.tv_loss(dec, sp, r3, r113)
        lg.backward()
        close(lg, lo, rtol=TOL, atol=1e-9, name="tv loss")
        _check_grad("tv coarse", dec.coarse_fn.decoder.params.grad, odec.coarse_fn.decoder.params.grad)
        _check_grad("tv table", dec.pe_fn.grid_fn.params.grad, odec.pe_fn.grid_fn.params.grad)


def test_operator_surface_equals_fused(golden_dir):
    """The reference-style composition pe_fn -> coarse_fn -> out_fn -> raw2nerf_color written with
    torch ops on the operator kernels gives the same render as the fused call."""
    from oracle import cases, reference_path as rp
    from dns_slam_b200 import _lib, fused
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "tracking_tiny.pt"), weights_only=False)
    inp = cases.tracking_inputs(g["meta"])
    dec = product_decoder_from_oracle(g["meta"]["shape"], inp["decoder"], n_class=g["meta"]["n_class"])
    samples = {k: v.to(dev) for k, v in g["samples"].items()}
    samples["mask"] = samples["mask"].bool()
    with torch.no_grad():
        rgb, depth, var, logits = rp.tracker_renderer(dec, dec.bound, samples)   # oracle glue, CUDA operators
        ld, preds = fused.render_and_loss(dec, samples, _lib.MODE_TRACK)
    close(preds["color"], rgb, rtol=TOL, atol=1e-5, name="rgb")
    close(preds["depth"], depth, rtol=TOL, atol=1e-5, name="depth")
    close(preds["var"], var, rtol=TOL, atol=1e-6, name="var")
    close(preds["logits"], logits, rtol=TOL, atol=1e-5, name="logits")


@pytest.mark.parametrize("mode,N,S,C", [("track", 1024, 96, 40), ("map", 4096, 47, 40)])
def test_full_size_properties(mode, N, S, C):
    """BASELINE.json shapes (configs 2 and 3): size-independent properties of the fused call --
    gradients are linear in the loss weights, rendered colours are convex combinations, and the
    loss dictionary is finite; the table gradient only touches entries some sample maps to."""
    from dns_slam_b200 import _lib, bench_util, fused
    dev = _dev()
    dec, samples = bench_util.synthetic_batch("replica", mode, N, S, C, dev, seed=3)
    m = _lib.MODE_TRACK if mode == "track" else _lib.MODE_MAP

    def run(scale):
        dec.zero_grad()
        lam = dict(p=5.0 * scale, d=5.0 * scale, l=0.1 * scale, lt=10.0 * scale, fs=10.0 * scale, op=10.0 * scale)
        ld, preds = fused.render_and_loss(dec, samples, m, lambdas=lam, opacity_sigma=0.05)
        ld["total"].backward()
        return ld, preds, dec.pe_fn.grid_fn.params.grad.clone(), dec.out_fn.color_decoder.params.grad.clone()

    ld1, p1, gt1, gc1 = run(1.0)
    ld2, p2, gt2, gc2 = run(2.0)
    assert all(bool(torch.isfinite(v).all()) for v in ld1.values())
    close(ld2["total"], 2 * ld1["total"], rtol=1e-4, atol=0, name="loss linear in lambda")
    assert rel_err(gt2, 2 * gt1) < 1e-3 and rel_err(gc2, 2 * gc1) < 1e-3
    assert float(p1["color"].min()) >= 0.0 and float(p1["color"].max()) <= 1.0
    z = samples["z_vals"]
    assert bool((p1["depth"] >= z.min(1)[0] - 1e-4).all()) and bool((p1["depth"] <= z.max(1)[0] + 1e-4).all())
    assert 0 < int((gt1 != 0).sum()) <= N * S * 16 * 8 * 2


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ranks_emulated_on_one_gpu(world):
    """the ray-sharded call with GLOBAL denominators and the global
    class rule, run shard by shard on ONE GPU (no collectives), sums to the single-call result."""
    from dns_slam_b200 import bench_util, fused, step as stepmod
    dev = _dev()
    N, S, C = 601, 47, 12
    dec, samples = bench_util.synthetic_batch("tiny", "map", N, S, C, dev, seed=5, n_frames=2)
    samples = {k: v for k, v in samples.items() if k != "mask"}
    ms = stepmod.MappingStep(dec, 5e-3, dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0))
    full = ms.forward_backward(samples)
    g_full, l_full = ms.grad.clone(), full[0].clone()
    shards = [stepmod.shard_bounds(N, world, r) for r in range(world)]
    local = [{k: v[lo:hi].contiguous() for k, v in samples.items()} for lo, hi in shards]
    counts = sum(fused.render_counts(ms._config(s)) for s in local)
    assert torch.equal(counts.cpu(), fused.render_counts(ms._config(samples)).cpu())
    g_sum, l_sum = torch.zeros_like(g_full), torch.zeros(8, device=dev)
    d_o = []
    for (lo, hi), s in zip(shards, local):
        cfg = ms._config(s).shard(N, lo, samples["gt_label"], counts)
        out = ms.forward_backward(s, cfg=cfg)
        g_sum += ms.grad
        l_sum += out[0]
        d_o.append(out[2])
    close(l_sum[:7], l_full[:7], rtol=1e-4, atol=1e-7, name="sharded losses")
    assert rel_err(g_sum, g_full) < 1e-4
    close(torch.cat(d_o, 0), full[2], rtol=1e-3, atol=1e-6 * float(full[2].abs().max()) + 1e-9, name="d_rays_o")
