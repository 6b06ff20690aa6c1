"""Whole-loop drop-ins against the CPU oracle: the pose-optimisation loop of
Tracker.run (slams/tracking.py:304-346) and the optimisation loop of Mapper.optimize
(slams/mapping.py:868-910), a few iterations each with identical draws.  Adam amplifies rounding
differences, so the bar is 2e-3 on the loss trajectory / poses after 3 steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import close, frame_to, product_decoder_from_oracle, rel_err  # noqa: E402

META_T = dict(shape="tiny", n_class=6, seed=31, n_samples=32, n_surface=15, pose_index=3, refer_index=2)
META_M = dict(shape="tiny", n_class=6, seed=37, n_samples=32, n_surface=15, tgt_ids=[1, 4, 6],
              refer_idx=[[0, 4, -1], [1, 2, -1], [4, 5, -1]], lambda_lt=10.0, lambda_sm=0.05)


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_tracking_loop_vs_oracle():
    from oracle import cases, reference_path as rp
    from dns_slam_b200 import fused, slam, synthetic as syn
    dev = _dev()
    s = syn.SHAPES["tiny"]
    inp = cases.tracking_inputs(META_T)
    cam = inp["cam"]
    n_it, lr = 3, 1e-3
    g = torch.Generator().manual_seed(5)
    n_win = (cam["H"] - 40) * (cam["W"] - 40)
    draws = [dict(idx=torch.randint(n_win, (s["tracking_pixels"],), generator=g), t_surface=torch.rand(15, generator=g),
                  t_zero=torch.rand(15, generator=g)) for _ in range(n_it)]
    est = inp["poses"][3].clone()
    est[:3, 3] += torch.tensor([0.02, -0.01, 0.015])
    refer_w2c = torch.inverse(inp["poses"][2])
    # ---- oracle loop (CPU)
    odec, bound = inp["decoder"], inp["bound"]
    quad = rp.quad_from_matrix(est[:3, :3].numpy()).clone().requires_grad_(True)
    T = est[:3, 3].clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [T], "lr": lr}, {"params": [quad], "lr": lr}])
    hist_o = []
    for it in range(n_it):
        opt.zero_grad()
        w2c = torch.stack((refer_w2c, torch.inverse(rp.c2w_from_quad_T(quad, T))), 0)
        tape = rp.DrawTape([("randint", draws[it]["idx"]), ("rand", draws[it]["t_surface"]), ("rand", draws[it]["t_zero"])])
        smp = rp.tracker_get_target_samples(cam, bound, odec, inp["frame"], quad, T, w2c, inp["feats"],
                                            s["tracking_pixels"], 32, 15, tape)
        pc, pd, pv, pl = rp.tracker_renderer(odec, bound, smp)
        p, d, l = rp.tracking_losses(smp, pc, pd, pv, pl)
        loss = s["lambda_color"] * p + s["lambda_depth"] * d + s["lambda_label"] * l
        hist_o.append(loss.detach())
        loss.backward()
        opt.step()
    # ---- native loop
    dec = product_decoder_from_oracle("tiny", odec, n_class=6)
    trk = slam.TrackerCore(cam, dec, s["tracking_pixels"], 32, 15, s["lambda_color"], s["lambda_depth"], s["lambda_label"])
    best, best_loss, hist = slam.track_frame(trk, frame_to(inp["frame"], dev), refer_w2c, fused.channels_last(inp["feats"].to(dev)),
                                             est, n_it, lr, lambda it: draws[it])
    close(hist, torch.stack(hist_o), rtol=2e-3, atol=1e-5, name="tracking loss trajectory")
    k = int(torch.argmin(torch.stack(hist_o)))
    assert abs(float(best_loss) - float(hist_o[k])) < 2e-3 * abs(float(hist_o[k]))
    # CUDA-graph replay of the same loop (frozen decoder: tracking only moves the pose) == eager loop
    n2 = 6
    g2 = torch.Generator().manual_seed(6)
    draws2 = [dict(idx=torch.randint(n_win, (s["tracking_pixels"],), generator=g2), t_surface=torch.rand(15, generator=g2),
                   t_zero=torch.rand(15, generator=g2)) for _ in range(n2)]
    trk2 = slam.TrackerCore(cam, dec, s["tracking_pixels"], 32, 15, s["lambda_color"], s["lambda_depth"], s["lambda_label"],
                            freeze_decoder=True)
    fr, fcl = frame_to(inp["frame"], dev), fused.channels_last(inp["feats"].to(dev))
    b_e, l_e, h_e = slam.track_frame(trk2, fr, refer_w2c, fcl, est, n2, lr, lambda it: draws2[it])
    b_g, l_g, h_g = slam.track_frame(trk2, fr, refer_w2c, fcl, est, n2, lr, lambda it: draws2[it], use_graph=True, native=False)
    close(h_g, h_e, rtol=1e-4, atol=1e-6, name="graph vs eager loss trajectory")
    close(b_g, b_e, rtol=1e-5, atol=1e-6, name="graph vs eager best pose")
    # the captured loop is REUSED for the next frame (new images, new start pose, fresh Adam state): no new capture
    loops = trk2._track_loops
    graph_before = next(iter(loops.values())).graph
    fr2 = dict(fr)
    fr2["color"] = (fr["color"] * 0.7 + 0.1).contiguous()
    fr2["depth"] = (fr["depth"] * 1.03).contiguous()
    est2 = est.clone()
    est2[:3, 3] += torch.tensor([-0.01, 0.02, 0.0])
    b_e2, l_e2, h_e2 = slam.track_frame(trk2, fr2, refer_w2c, fcl, est2, n2, lr, lambda it: draws2[it])
    b_g2, l_g2, h_g2 = slam.track_frame(trk2, fr2, refer_w2c, fcl, est2, n2, lr, lambda it: draws2[it], use_graph=True,
                                        native=False)
    assert len(loops) == 1 and next(iter(loops.values())).graph is graph_before
    close(h_g2, h_e2, rtol=1e-4, atol=1e-6, name="reused graph vs eager loss trajectory")
    close(b_g2, b_e2, rtol=1e-5, atol=1e-6, name="reused graph vs eager best pose")
    assert not torch.equal(h_g2, h_g)
    # the NATIVE loop (step.TrackingFrameStep: no autograd, no PyTorch kernels on the data path) -- the default fast path --
    # against the oracle trajectory and against the eager loops, on both frames with ONE cached step object
    trk3 = slam.TrackerCore(cam, dec, s["tracking_pixels"], 32, 15, s["lambda_color"], s["lambda_depth"], s["lambda_label"],
                            freeze_decoder=True)
    b_n, l_n, h_n = slam.track_frame(trk3, fr, refer_w2c, fcl, est, n_it, lr, lambda it: draws[it], use_graph=True)
    close(h_n, torch.stack(hist_o), rtol=2e-3, atol=1e-5, name="native tracking loss trajectory vs oracle")
    assert abs(float(l_n) - float(hist_o[k])) < 2e-3 * abs(float(hist_o[k]))
    b_n1, l_n1, h_n1 = slam.track_frame(trk3, fr, refer_w2c, fcl, est, n2, lr, lambda it: draws2[it], native=True)
    close(h_n1, h_e, rtol=1e-4, atol=1e-6, name="native vs eager loss trajectory")
    close(b_n1, b_e, rtol=1e-5, atol=1e-6, name="native vs eager best pose")
    close(l_n1, l_e, rtol=1e-5, atol=1e-7, name="native vs eager best loss")
    b_n2, l_n2, h_n2 = slam.track_frame(trk3, fr2, refer_w2c, fcl, est2, n2, lr, lambda it: draws2[it], native=True)
    assert len(trk3._track_steps) == 2          # (3 iterations) and (6 iterations): one step object per loop length
    close(h_n2, h_e2, rtol=1e-4, atol=1e-6, name="native vs eager loss trajectory, next frame")
    close(b_n2, b_e2, rtol=1e-5, atol=1e-6, name="native vs eager best pose, next frame")


def test_mapping_loop_vs_oracle():
    from oracle import cases, reference_path as rp
    from dns_slam_b200 import fused, slam, synthetic as syn
    dev = _dev()
    s = syn.SHAPES["tiny"]
    meta = META_M
    inp = cases.mapping_inputs(meta)
    cam, bound = inp["cam"], inp["bound"]
    n_it, lr, cam_lr = 2, 5e-3, 5e-4
    n_t = len(meta["tgt_ids"])
    npf = s["mapping_pixels"] // n_t
    g = torch.Generator().manual_seed(9)
    draws, tv = [], []
    for _ in range(n_it):
        per = []
        for fr in inp["frames"]:
            lab = fr["label"].reshape(-1)
            classes, counts = torch.unique(lab, return_counts=True)
            n_c, n_k = classes.numel(), (npf // 3) // classes.numel()
            cd = []
            for c in range(n_c):
                m = (npf // 3) - n_k * (n_c - 1) if c == 0 else n_k
                if int(counts[c]) != 1:
                    cd.append(torch.randint(int(counts[c]), (m,), generator=g))
            per.append(dict(idx_uniform=torch.randint(cam["H"] * cam["W"], (npf // 3 * 2,), generator=g), class_draws=cd,
                            t_surface=torch.rand(15, generator=g), t_zero=torch.rand(15, generator=g)))
        draws.append(per)
        tv.append((torch.rand(3, generator=g), torch.rand(1, 1, 1, 3, generator=g)))
    est = [inp["poses"][i].clone() for i in meta["tgt_ids"]]
    for n, e in enumerate(est):
        e[:3, 3] += 0.01 * (n + 1)
    lam = dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=10.0, fs=s["lambda_fs"], op=s["lambda_opacity"])
    # ---- oracle loop (CPU)
    odec, oexp = inp["decoder"], inp["experts"]
    ql = [rp.quad_from_matrix(e[:3, :3].numpy()).clone().requires_grad_(i != 0) for i, e in enumerate(est)]
    Tl = [e[:3, 3].clone().requires_grad_(i != 0) for i, e in enumerate(est)]
    net = list(odec.parameters()) + [e.params for e in oexp.values()]
    opt = torch.optim.Adam([{"params": net, "lr": lr}, {"params": ql[1:], "lr": cam_lr}, {"params": Tl[1:], "lr": cam_lr}])
    losses_o = []
    for it in range(n_it):
        opt.zero_grad()
        items = []
        for d in draws[it]:
            items.append(("randint", d["idx_uniform"]))
            items += [("randint", c) for c in d["class_draws"]]
            items += [("rand", d["t_surface"]), ("rand", d["t_zero"])]
        items += [("rand", tv[it][0]), ("rand", tv[it][1])]
        tape = rp.DrawTape(items)
        smp = rp.mapper_get_target_samples(cam, bound, odec, inp["frames"], ql, Tl, meta["refer_idx"], meta["tgt_ids"],
                                           inp["refer_c2w"], inp["feats"], s["mapping_pixels"], 32, 15, tape)
        pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(odec, oexp, bound, smp)
        p, d, l, lt, fs, op = rp.mapping_losses(smp, pc, pd, pl, fine, coarse, s["opacity_sigma"])
        sm = rp.smoothness(odec, bound, s["smooth_pts"], tape)
        loss = lam["p"] * p + lam["d"] * d + lam["l"] * l + lam["lt"] * lt + meta["lambda_sm"] * sm + lam["fs"] * fs + lam["op"] * op
        losses_o.append(loss.detach())
        loss.backward()
        opt.step()
    # ---- native loop
    # the oracle decoder has already been stepped: rebuild identical initial weights for the native side
    inp2 = cases.mapping_inputs(meta)
    dec = product_decoder_from_oracle("tiny", inp2["decoder"], inp2["experts"], n_class=6)
    mp = slam.MapperCore(cam, dec, s["mapping_pixels"], 32, 15, lambdas=lam, opacity_sigma=s["opacity_sigma"],
                         smooth_pts=s["smooth_pts"], lambda_sm=meta["lambda_sm"])
    frames = [frame_to(f, dev) for f in inp["frames"]]
    target = dict(kf_idx=meta["tgt_ids"], frames=frames, class_tables=[slam.class_tables(f["label"]) for f in frames])
    refer = dict(kf_idx=meta["refer_idx"], est_c2w=[[c.to(dev) for c in row] for row in inp["refer_c2w"]])
    feats = [fused.channels_last(f.to(dev)) for f in inp["feats"]]
    hist = []
    orig_iter = mp.iteration

    def rec(*a, **k):
        out = orig_iter(*a, **k)
        hist.append(out[0]["total"].detach())
        return out
    mp.iteration = rec
    quad_list, T_list, ld = slam.map_optimize(mp, target, refer, feats, est, n_it, lr, cam_lr, True, [],
                                              lambda it: draws[it], lambda it: tv[it])
    close(torch.stack(hist), torch.stack(losses_o), rtol=2e-3, atol=1e-5, name="mapping loss trajectory")
    for i in range(1, n_t):
        close(quad_list[i], ql[i], rtol=2e-3, atol=2e-5, name=f"quad[{i}] after BA")
        close(T_list[i], Tl[i], rtol=2e-3, atol=2e-5, name=f"T[{i}] after BA")
    # parameters after two Adam steps: the update direction is sign-like, so compare the net displacement
    assert rel_err(dec.coarse_fn.decoder.params, odec.coarse_fn.decoder.params) < 1e-3
    assert rel_err(dec.pe_fn.grid_fn.params, odec.pe_fn.grid_fn.params) < 1e-3
    assert mp.last_path == "eager"
    # ---- the NATIVE loop (step.MappingFrameStep behind slam.map_optimize: the default fast path) against the same oracle
    inp3 = cases.mapping_inputs(meta)
    dec3 = product_decoder_from_oracle("tiny", inp3["decoder"], inp3["experts"], n_class=6)
    mp3 = slam.MapperCore(cam, dec3, s["mapping_pixels"], 32, 15, lambdas=lam, opacity_sigma=s["opacity_sigma"],
                          smooth_pts=s["smooth_pts"], lambda_sm=meta["lambda_sm"])
    hist3 = []
    q3, t3, ld3 = slam.map_optimize(mp3, target, refer, feats, est, n_it, lr, cam_lr, True, [], lambda it: draws[it],
                                    lambda it: tv[it], use_graph=True, history=hist3)
    assert mp3.last_path == "native", "a sampled ray left the bound: the native loop fell back"
    close(torch.stack(hist3), torch.stack(losses_o), rtol=2e-3, atol=1e-5, name="native mapping loss trajectory")
    close(ld3["total"], losses_o[-1], rtol=2e-3, atol=1e-5, name="native last loss")
    for i in range(1, n_t):
        close(q3[i], ql[i], rtol=2e-3, atol=2e-5, name=f"native quad[{i}] after BA")
        close(t3[i], Tl[i], rtol=2e-3, atol=2e-5, name=f"native T[{i}] after BA")
    close(q3[0], ql[0], rtol=1e-6, atol=1e-7, name="the oldest frame stays fixed")
    assert rel_err(dec3.coarse_fn.decoder.params, odec.coarse_fn.decoder.params) < 1e-3
    assert rel_err(dec3.pe_fn.grid_fn.params, odec.pe_fn.grid_fn.params) < 1e-3
    assert rel_err(dec3.merge.decoder.params, odec.merge.decoder.params) < 1e-3


def test_mapping_loop_cuda_graph_equals_eager():
    """CUDA-graph replay of the mapping loop (static shapes, device-side TV offsets) == the eager loop."""
    from dns_slam_b200 import bench_util, slam, synthetic as syn
    dev = _dev()
    s = syn.SHAPES["tiny"]
    n_it = 7
    sc = bench_util.slam_scene("tiny", 6, dev, seed=3, n_target=3)
    md, tv = bench_util.mapping_draws(sc, s["mapping_pixels"], n_it, seed=4)
    target = dict(kf_idx=sc["kf_idx"], frames=sc["frames"], class_tables=sc["class_tables"])
    refer = dict(kf_idx=sc["refer_idx"], est_c2w=sc["refer_c2w"])
    est = [sc["poses"][2 * f + 1].clone() for f in range(3)]
    lam = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
    res = []
    for use_graph, native, path in ((False, None, "eager"), (True, False, "graph"), (True, None, "native")):
        dec = bench_util.make_decoder("tiny", 6, dev, seed=1)
        mp = slam.MapperCore(sc["cam"], dec, s["mapping_pixels"], 32, 15, lambdas=lam, opacity_sigma=0.05,
                             smooth_pts=s["smooth_pts"], lambda_sm=0.05)
        ql, tl, ld = slam.map_optimize(mp, target, refer, sc["feats"], est, n_it, 5e-3, 5e-4, True, [],
                                       lambda it: md[it], lambda it: tv[it], use_graph=use_graph, native=native)
        if use_graph:
            assert getattr(mp, "last_graph_ok", False), "fast path fell back to eager (rays outside the bound)"
        assert mp.last_path == path
        res.append((ql, tl, ld, dec.flat.clone()))
    (q0, t0, l0, f0) = res[0]
    for (q1, t1, l1, f1), path in zip(res[1:], ("graph", "native")):
        close(l1["total"], l0["total"], rtol=1e-3, atol=1e-6, name=f"last loss ({path})")
        close(l1["smooth_loss"], l0["smooth_loss"], rtol=1e-3, atol=1e-6, name=f"smoothness ({path})")
        for i in range(1, 3):
            close(q1[i], q0[i], rtol=1e-3, atol=1e-5, name=f"quad ({path})")
            close(t1[i], t0[i], rtol=1e-3, atol=1e-5, name=f"T ({path})")
        assert rel_err(f1, f0) < 1e-3, path


def test_graph_capture_after_an_eager_loop_on_the_same_decoder():
    """Regression (found by the ScanNet-shaped sequence of examples/synthetic_slam.py): the eager loop used to return
    its last loss dictionary with the autograd graph attached; held by the caller, it kept the AccumulateGrad nodes
    of the default stream alive and the next capture failed with cudaErrorStreamCaptureImplicit."""
    from dns_slam_b200 import bench_util, slam, synthetic as syn
    dev = _dev()
    s = syn.SHAPES["tiny"]
    n_it = 6
    sc = bench_util.slam_scene("tiny", 6, dev, seed=5, n_target=2)
    md, tv = bench_util.mapping_draws(sc, s["mapping_pixels"], n_it, seed=6)
    target = dict(kf_idx=sc["kf_idx"], frames=sc["frames"], class_tables=sc["class_tables"])
    refer = dict(kf_idx=sc["refer_idx"], est_c2w=sc["refer_c2w"])
    est = [sc["poses"][2 * f + 1].clone() for f in range(2)]
    dec = bench_util.make_decoder("tiny", 6, dev, seed=2)
    mp = slam.MapperCore(sc["cam"], dec, s["mapping_pixels"], 32, 15,
                         lambdas=dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0), opacity_sigma=0.05,
                         smooth_pts=s["smooth_pts"], lambda_sm=0.05)
    args = (mp, target, refer, sc["feats"], est, n_it, 5e-3, 5e-4, True, [], lambda it: md[it], lambda it: tv[it])
    _, _, kept = slam.map_optimize(*args, use_graph=False)          # the caller keeps this dictionary
    assert all(v.grad_fn is None and not v.requires_grad for v in kept.values())
    _, _, ld = slam.map_optimize(*args, use_graph=True, native=False)
    assert getattr(mp, "last_graph_ok", False)
    assert torch.isfinite(ld["total"]).all() and torch.isfinite(kept["total"]).all()


def test_feature_gather_rejects_a_view_count_mismatch():
    """The C entry point only sees raw pointers: a feature tensor with fewer views than poses would be read out of
    bounds, so the wrapper refuses it."""
    from dns_slam_b200 import fused
    dev = _dev()
    pts = torch.rand(10, 3, device=dev)
    w2c = torch.eye(4, device=dev).repeat(2, 1, 1)
    K = torch.tensor([[60.0, 0, 39.5], [0, 60.0, 29.5], [0, 0, 1]])
    with pytest.raises(ValueError):
        fused.feature_gather(60, 80, K, pts, w2c, torch.rand(1, 30, 40, 64, device=dev))
    code, uv, mask = fused.feature_gather(60, 80, K, pts, w2c, torch.rand(2, 30, 40, 64, device=dev))
    assert code.shape == (2, 10, 64)


def test_decoder_init_vs_reference_golden(golden_dir):
    """``slam.decoder_init`` against ``Mapper.decoder_init`` of the reference itself (slams/mapping.py:764-836 run by
    oracle/make_golden.py decoder_init_case for 3 iterations on the recorded draws): every parameter after the last
    Adam step -- hash table, coarse / colour / logit / Merge nets and the freshly created class experts."""
    import os
    from oracle import make_golden as mg
    from dns_slam_b200 import fused, slam, synthetic as syn
    from gpu_util import frame_to, product_decoder_from_oracle, rel_err
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(golden_dir, "decoder_init_tiny.pt"), weights_only=False)
    meta = g["meta"]
    s = syn.SHAPES[meta["shape"]]
    gen = torch.Generator().manual_seed(meta["seed"])
    bound, odec, oexp = mg.build_models(meta["shape"], meta["n_class"], meta["seed"], expert_classes=meta["decoder_idx"])
    dec = product_decoder_from_oracle(meta["shape"], odec, oexp, n_class=meta["n_class"])
    cam = syn.camera(meta["shape"])
    pose = syn.trajectory(meta["shape"], 6)[meta["pose_index"]]
    fr = frame_to(syn.frame(meta["shape"], pose, gen, n_class=meta["n_class"]), dev)
    feats = fused.channels_last(g["features"].to(dev))
    mp = slam.MapperCore(cam, dec, s["mapping_pixels"], 32, 15,
                         lambdas=dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=0.0, fs=s["lambda_fs"],
                                      op=s["lambda_opacity"]),
                         opacity_sigma=s["opacity_sigma"], smooth_pts=s["smooth_pts"], lambda_sm=meta["lambda_sm"])
    # recorded draw order per iteration: one randint per class of decoder_idx with more than one pixel, rand x2
    # (sample_along_rays), rand(3) + rand(1,1,1,3) (smoothness)
    tape, per = g["tape"], len(g["tape"]) // meta["n_iters"]
    its = []
    for it in range(meta["n_iters"]):
        chunk = tape[it * per:(it + 1) * per]
        ints = [t for k, t in chunk if k == "randint"]
        rands = [t for k, t in chunk if k == "rand"]
        its.append((dict(class_draws=ints, t_surface=rands[0], t_zero=rands[1]), (rands[2], rands[3])))
    table0 = dec.pe_fn.grid_fn.params.detach().clone()
    slam.decoder_init(mp, meta["decoder_idx"], fr, slam.class_tables(fr["label"]), pose, feats, s["lr"],
                      lambda it: its[it][0], lambda it: its[it][1], n_iters=meta["n_iters"], n_rays=meta["n_rays"])
    sd = dec.state_dict()
    for k, want in g["params"].items():
        if k in sd and want.numel() == sd[k].numel():   # three Adam steps: every element moves by ~lr per step
            assert rel_err(sd[k], want) < 5e-3, k
    tab = dec.pe_fn.grid_fn.params.detach().cpu()
    gt = g["table"]
    moved = rel_err(tab[::int(gt["stride"])] - table0.cpu()[::int(gt["stride"])], gt["strided"] - table0.cpu()[::int(gt["stride"])])
    assert moved < 5e-3, f"table update {moved:.2e}"
    for c, want in g["experts"].items():
        assert rel_err(dec.expert_params[c][:want.numel()], want) < 5e-3, f"expert {c}"


def test_fast_mapping_loops_fall_back_when_rays_leave_the_bound():
    """mapping.py:525 drops the rays whose depth lies beyond the scene bound; the static-shape fast loops (native step,
    captured graph) cannot, so they must notice -- in whatever iteration it happens -- restore the decoder and hand the call
    to the compacting eager loop (a flag that only saw the warm-up iterations and the last replay).  One frame's
    depth is stretched so that a part of its pixels fails the inside test."""
    from dns_slam_b200 import bench_util, slam, synthetic as syn
    dev = _dev()
    s = syn.SHAPES["tiny"]
    n_it = 12
    sc = bench_util.slam_scene("tiny", 6, dev, seed=11, n_target=2)
    frames = [dict(f) for f in sc["frames"]]
    d = frames[1]["depth"].clone()
    d[: d.shape[0] // 6] *= 3.0                      # a band of pixels whose surface point is outside the bound
    frames[1]["depth"] = d.contiguous()
    md, tv = bench_util.mapping_draws(sc, s["mapping_pixels"], n_it, seed=12)
    target = dict(kf_idx=sc["kf_idx"], frames=frames, class_tables=sc["class_tables"])
    refer = dict(kf_idx=sc["refer_idx"], est_c2w=sc["refer_c2w"])
    est = [sc["poses"][2 * f + 1].clone() for f in range(2)]
    lam = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
    res = {}
    for name, kw in (("eager", dict(use_graph=False)), ("native", dict(use_graph=True)), ("graph", dict(use_graph=True, native=False))):
        dec = bench_util.make_decoder("tiny", 6, dev, seed=3)
        mp = slam.MapperCore(sc["cam"], dec, s["mapping_pixels"], 32, 15, lambdas=lam, opacity_sigma=0.05,
                             smooth_pts=s["smooth_pts"], lambda_sm=0.05)
        ql, tl, ld = slam.map_optimize(mp, target, refer, sc["feats"], est, n_it, 5e-3, 5e-4, True, [],
                                       lambda it: md[it], lambda it: tv[it], **kw)
        assert mp.last_path == "eager", f"{name}: rays outside the bound went unnoticed"
        if name != "eager":
            assert mp.last_graph_ok is False
        res[name] = (ld["total"].clone(), dec.flat.clone(), [q.clone() for q in ql])
    for name in ("native", "graph"):       # the fallback restarts from the saved state: identical to the eager call
        close(res[name][0], res["eager"][0], rtol=1e-5, atol=1e-7, name=f"{name} fallback loss")
        assert rel_err(res[name][1], res["eager"][1]) < 1e-5
        close(res[name][2][1], res["eager"][2][1], rtol=1e-5, atol=1e-7, name=f"{name} fallback pose")


def test_native_tracking_flags_a_label_outside_the_semantic_head():
    """torch's cross_entropy raises on a label >= C (tracking.py:203); the native loop reads the kernel's flag once."""
    from dns_slam_b200 import bench_util, slam, synthetic as syn
    dev = _dev()
    s = syn.SHAPES["tiny"]
    sc = bench_util.slam_scene("tiny", 6, dev, seed=13, n_target=2)
    dec = bench_util.make_decoder("tiny", 6, dev, seed=4)
    trk = slam.TrackerCore(sc["cam"], dec, s["tracking_pixels"], 32, 15, freeze_decoder=True)
    td = bench_util.tracking_draws(sc["cam"], s["tracking_pixels"], 4, seed=1)
    fr = dict(sc["frames"][1])
    est, refer_w2c = sc["poses"][3].clone(), torch.inverse(sc["poses"][1])
    feats2 = sc["feats"][1][:2].contiguous()
    best, loss, hist = slam.track_frame(trk, fr, refer_w2c, feats2, est, 4, 1e-3, lambda it: td[it], native=True)
    assert torch.isfinite(hist).all() and float(loss) == float(hist.min())
    bad = dict(fr)
    bad["label"] = torch.full_like(fr["label"], 6)
    with pytest.raises(ValueError):
        slam.track_frame(trk, bad, refer_w2c, feats2, est, 4, 1e-3, lambda it: td[it], native=True)


def test_native_loops_cover_the_optimizer_variants():
    """Two switches of the reference's optimiser set-up on the native loops against the autograd loops: the tracker's
    ``seperate_LR`` (translation at 0.2 x cam_lr, slams/tracking.py:119-124) and a mapping call without bundle adjustment
    (``BA = False``: no pose parameter group, slams/mapping.py:457-466)."""
    from dns_slam_b200 import bench_util, slam, synthetic as syn
    dev = _dev()
    s = syn.SHAPES["tiny"]
    sc = bench_util.slam_scene("tiny", 6, dev, seed=21, n_target=2)
    dec = bench_util.make_decoder("tiny", 6, dev, seed=5)
    # ---- tracking, separate learning rates
    trk = slam.TrackerCore(sc["cam"], dec, s["tracking_pixels"], 32, 15, freeze_decoder=True)
    n_it = 5
    td = bench_util.tracking_draws(sc["cam"], s["tracking_pixels"], n_it, seed=2)
    est = sc["poses"][3].clone()
    est[:3, 3] += torch.tensor([0.015, -0.01, 0.02])
    args = (trk, sc["frames"][1], torch.inverse(sc["poses"][1]), sc["feats"][1][:2].contiguous(), est, n_it, 2e-3, lambda it: td[it])
    b_e, l_e, h_e = slam.track_frame(*args, seperate_LR=True)
    b_n, l_n, h_n = slam.track_frame(*args, seperate_LR=True, native=True)
    close(h_n, h_e, rtol=1e-4, atol=1e-6, name="native vs eager loss trajectory (seperate_LR)")
    close(b_n, b_e, rtol=1e-5, atol=1e-6, name="native vs eager best pose (seperate_LR)")
    b_1, _, h_1 = slam.track_frame(*args, seperate_LR=False, native=True)
    assert not torch.equal(h_1, h_n), "the translation learning rate must matter"
    # ---- mapping without bundle adjustment: poses stay where they were, the decoder moves as in the eager loop
    md, tv = bench_util.mapping_draws(sc, s["mapping_pixels"], 4, seed=3)
    target = dict(kf_idx=sc["kf_idx"], frames=sc["frames"], class_tables=sc["class_tables"])
    refer = dict(kf_idx=sc["refer_idx"], est_c2w=sc["refer_c2w"])
    est_l = [sc["poses"][2 * f + 1].clone() for f in range(2)]
    res = []
    for native in (False, True):
        d2 = bench_util.make_decoder("tiny", 6, dev, seed=6)
        mp = slam.MapperCore(sc["cam"], d2, s["mapping_pixels"], 32, 15, lambdas=dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0),
                             opacity_sigma=0.05, smooth_pts=s["smooth_pts"], lambda_sm=0.05)
        ql, tl, ld = slam.map_optimize(mp, target, refer, sc["feats"], est_l, 4, 5e-3, 5e-4, False, [], lambda it: md[it],
                                       lambda it: tv[it], native=native)
        assert mp.last_path == ("native" if native else "eager")
        res.append((ql, tl, ld["total"].clone(), d2.flat.clone()))
    for i in range(2):
        close(res[1][0][i], res[0][0][i], rtol=1e-6, atol=1e-7, name="pose untouched without BA")
        close(res[1][0][i].cpu(), slam.quad_from_matrix(est_l[i][:3, :3]), rtol=1e-6, atol=1e-7, name="pose = start pose")
    close(res[1][2], res[0][2], rtol=1e-3, atol=1e-6, name="last loss without BA")
    assert rel_err(res[1][3], res[0][3]) < 1e-3
