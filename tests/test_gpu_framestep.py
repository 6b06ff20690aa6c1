"""MappingFrameStep (frames + draws boundary): one whole mapping iteration -- sampling incl. the class-balanced draw,
pixel-feature branch, render, 7 losses, every gradient incl. Merge weights and camera poses, Adam -- against the oracle
iteration body (oracle/cases.run_mapping <-> slams/mapping.py:884-909) on the draws the step generated, and the 2-rank
sharded step (two threads on one GPU, reductions done by a lock-step communicator) against the unsharded call chain on
the concatenated batch.  1e-3 relative (north_star); indices / z bit exact."""
import os
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import frame_to, product_decoder_from_oracle, rel_err  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    return torch.device("cuda:0")


def _build(golden_dir, n_rays, rank=0, world=1, comm=None, with_tv=True, is_BA=True, decoder=None):
    from oracle import cases
    from dns_slam_b200 import fused, slam, step, synthetic as syn
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "mapping_tiny.pt"), weights_only=False)
    meta = g["meta"]
    s = syn.SHAPES[meta["shape"]]
    inp = cases.mapping_inputs(meta)
    dec = decoder or product_decoder_from_oracle(meta["shape"], inp["decoder"], inp["experts"], n_class=meta["n_class"])
    frames = [frame_to(f, dev) for f in inp["frames"]]
    tables = [slam.class_tables(f["label"]) for f in frames]
    feats = [fused.channels_last(f.to(dev)) for f in inp["feats"]]
    est = [torch.eye(4) for _ in frames]
    for c, q, t in zip(est, g["quad"], g["T"]):
        c[:3, :3] = slam.quad2rotation(q[None])[0]
        c[:3, 3] = t
    st = step.MappingFrameStep(dec, inp["cam"], frames, tables, feats, est, meta["refer_idx"], inp["refer_c2w"], meta["tgt_ids"],
                               n_rays, meta["n_samples"], meta["n_surface"], lr=s["lr"], BA_cam_lr=s["BA_cam_lr"], is_BA=is_BA,
                               lambdas=dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=meta["lambda_lt"],
                                            fs=s["lambda_fs"], op=s["lambda_opacity"]),
                               opacity_sigma=s["opacity_sigma"], smooth_pts=s["smooth_pts"], lambda_sm=meta["lambda_sm"],
                               with_tv=with_tv, comm=comm, rank=rank, world=world)
    # the step keeps quaternions as given (the golden ones are not normalised); overwrite the matrix round trip
    with torch.no_grad():
        st.quats.copy_(torch.stack(g["quad"], 0))
        st.trans.copy_(torch.stack(g["T"], 0))
    return st, g, meta, inp, s


@pytest.mark.parametrize("n_rays,simt", [(48, True), (300, False), (1200, False)])
def test_frame_step_vs_oracle_iteration(golden_dir, n_rays, simt):
    """``simt``: the fused core runs its fp32 SIMT kernels.  On the 45-ray batch ONE hidden unit whose pre-activation is
    zero to within the bf16 hi/lo rounding is 6e-3 of the whole table gradient, so the tiny case
    pins the arithmetic on the SIMT core and the larger ones the tcgen05 core."""
    import contextlib
    from oracle import cases
    from dns_slam_b200 import fused, synthetic as syn
    st, g, meta, inp, s = _build(golden_dir, n_rays)
    dec = st.dec
    buf, tape = st.make_host_draws(torch.Generator().manual_seed(4), return_tape=True)
    st.upload(buf)
    flat0 = dec.flat.detach().clone()
    q0, t0 = st.quats.clone(), st.trans.clone()
    with (fused.simt_path() if simt else contextlib.nullcontext()):
        res = st.step().cpu()
    st.check(res)
    old = syn.SHAPES["tiny"]["mapping_pixels"]
    syn.SHAPES["tiny"]["mapping_pixels"] = n_rays
    try:
        o = cases.run_mapping(meta, g["quad"], g["T"], tape, inp)
    finally:
        syn.SHAPES["tiny"]["mapping_pixels"] = old
    os_ = o["samples"]
    b = st.batch
    for k in ("gt_label", "gt_depth", "gt_color", "z_vals"):
        assert torch.equal(b[k].cpu(), os_[k]), k
    assert torch.equal(b["pixel"].cpu(), os_["_idx"]), "class-balanced / uniform pixel indices"
    assert rel_err(b["rays_d"], os_["rays_d"]) < 1e-6
    # the step hands its pixel features over BAND ONLY (dns_featmerge_args.no_zero_fill / dns_render_args.features_band_only):
    # rows of samples outside the truncation band are neither written nor read -- the oracle holds zeros there
    from dns_slam_b200 import slam
    band = slam.trunc_mask(b["z_vals"], b["gt_depth"])[..., None]
    assert float(band.mean()) > 0.05 and float((os_["features"].detach() * (1 - band.cpu())).abs().max()) == 0.0
    assert rel_err(st.features * band, os_["features"]) < 1e-3
    for i, k in enumerate(("p", "d", "l", "lt", "fs", "op")):
        torch.testing.assert_close(res[i], o["loss"][k].detach().float(), rtol=1e-3, atol=1e-7, msg=k)
    torch.testing.assert_close(res[8], o["loss"]["sm"].detach().float(), rtol=1e-3, atol=1e-9)
    torch.testing.assert_close(res[6], o["loss"]["total"].detach().float(), rtol=1e-3, atol=1e-7)
    gv = st._flat_views(st.grad)
    for k in ("table", "coarse", "color", "logit", "merge"):
        assert rel_err(gv[k], o["grad"][k]) < 1e-3, k
    for c, ge in o["grad"]["experts"].items():
        if ge is None:
            assert float(gv["experts"][c].abs().sum()) == 0.0
        else:
            assert rel_err(gv["experts"][c], ge) < 1e-3, f"expert {c}"
    for f in range(1, st.F):
        assert rel_err(st.d_quats[f], o["grad"]["quad"][f]) < 1e-3, f"quad {f}"
        assert rel_err(st.d_trans[f], o["grad"]["T"][f]) < 1e-3, f"T {f}"
    # first Adam step (torch defaults): p - lr * g / (|g| + eps); frame 0 stays fixed (mapping.py:457)
    gflat = st.grad
    want = flat0 - s["lr"] * gflat / (gflat.abs() + 1e-8)
    assert rel_err(dec.flat - flat0, want - flat0) < 1e-4
    assert torch.equal(st.quats[0], q0[0]) and torch.equal(st.trans[0], t0[0])
    dq = st.d_quats[1:]
    assert rel_err(st.quats[1:] - q0[1:], -s["BA_cam_lr"] * dq / (dq.abs() + 1e-8)) < 1e-3


def test_frame_step_flags_missing_expert_and_bad_label(golden_dir):
    st, g, meta, inp, s = _build(golden_dir, 48)
    st.upload(st.make_host_draws(torch.Generator().manual_seed(1)))
    st.dec.class_to_expert[int(st.frames[0]["label"][5, 5])] = -1      # 'Fine decoders does NOT have class'
    res = st.step().cpu()
    with pytest.raises(ValueError, match="Fine decoders"):
        st.check(res)


class _LockstepComm:
    """Two threads = two ranks on one GPU.  Compute sections are serialised by ``lock`` (one rank at a time touches the
    library and the shared scratch workspace); every collective releases it, meets the other rank at a barrier and is
    evaluated by each rank from the published tensors."""

    def __init__(self, world):
        self.world, self.bar, self.slots = world, threading.Barrier(world), {}
        self.local, self.lock = threading.local(), threading.Lock()

    def _meet(self):
        self.lock.release()
        try:
            self.bar.wait()
        finally:
            self.lock.acquire()

    def _exchange(self, t, fn):
        self.slots[self.local.rank] = t
        self._meet()
        torch.cuda.synchronize()
        out = fn([self.slots[k].clone() for k in range(self.world)])
        torch.cuda.synchronize()
        self._meet()          # nobody overwrites a published tensor before every rank has read it
        return out

    def all_reduce_sum(self, t):
        t.copy_(self._exchange(t, lambda xs: torch.stack(xs).sum(0)))
        return t

    def all_reduce_max(self, t):
        t.copy_(self._exchange(t, lambda xs: torch.stack(xs).max(0)[0]))
        return t

    def all_gather(self, t, sizes=None):
        return self._exchange(t, lambda xs: torch.cat(xs))


def test_sharded_frame_step_sums_to_unsharded_chain(golden_dir):
    from dns_slam_b200 import _lib, fused
    world, n_rays = 2, 300
    comm = _LockstepComm(world)
    st0, g, meta, inp, s = _build(golden_dir, n_rays, rank=0, world=world, comm=comm)
    # both ranks share ONE decoder object here; Adam is switched off so the second rank sees the same weights
    st1 = _build(golden_dir, n_rays, rank=1, world=world, comm=comm, decoder=st0.dec)[0]
    steps = [st0, st1]
    for r, st in enumerate(steps):   # own pixel draws per rank, the batch-level draws (surface offsets, TV lattice) shared
        st.upload(st.make_host_draws(torch.Generator().manual_seed(10 + r), shared_gen=torch.Generator().manual_seed(77)))
        st.adam.step = lambda: None
    torch.cuda.synchronize()
    errs = []

    def run(r):
        comm.local.rank = r
        with comm.lock:
            try:
                steps[r].step()
                torch.cuda.synchronize()
            except Exception as e:      # noqa: BLE001
                errs.append(e)
                comm.bar.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    assert torch.equal(st0.packed, st1.packed), "ranks disagree after the all-reduce"
    # the same frame must see the same max depth on both ranks (z sampling, common.py:581,591)
    assert torch.equal(st0.scratch[:, 0], st1.scratch[:, 0])
    # ---- unsharded chain on the rank-major concatenation: (rank, frame) pairs become the frames of one call
    from dns_slam_b200 import slam as slam_mod
    dec, cam = st0.dec, st0.cam
    cat = {k: torch.cat([st.batch[k] for st in steps], 0).contiguous() for k in st0.batch}
    ray_start, w2c, cam_o, feats = [0], [], [], []
    for st in steps:
        for f in range(st.F):
            ray_start.append(ray_start[-1] + st.ray_start[f + 1] - st.ray_start[f])
        w2c.append(st.w2c), cam_o.append(st.cam_o), feats.extend(st.feats)
    views = fused.Views(torch.cat(w2c, 0), torch.cat(cam_o, 0), feats, ray_start)
    mp = dec.view("merge")
    feat, ws = fused.featmerge_raw(cam, dec.merge.bound, views, cat["rays_o"], cat["rays_d"], cat["z_vals"], cat["gt_depth"], mp)
    band = slam_mod.trunc_mask(cat["z_vals"], cat["gt_depth"])[..., None]      # band-only hand-over inside the steps
    assert torch.equal(feat, torch.cat([st.features for st in steps], 0) * band)
    packed = torch.zeros_like(st0.packed)
    gv = st0._flat_views(packed[:dec.flat.numel()])
    p = st0._flat_views(dec.flat)
    cfg = fused.RenderConfig(_lib.MODE_MAP, dec.bound, dec.pe_fn.grid_fn.gstruct, cat["z_vals"], cat["gt_color"], cat["gt_depth"],
                             cat["gt_label"], None, dec.class_to_expert, dec.n_class, st0.lambdas, opacity_trunc=st0.opacity_sigma)
    losses, _, d_o, d_d, d_f = fused.render_raw(cfg, p["table"], p["coarse"], p["color"], p["logit"], p["experts"], cat["rays_o"],
                                                cat["rays_d"], feat, gv, True, True)
    fused.featmerge_bwd_raw(cam, dec.merge.bound, views, cat["rays_o"], cat["rays_d"], cat["z_vals"], cat["gt_depth"], mp, d_f,
                            ws, gv["merge"], d_o, d_d)
    sm = fused.tv_raw(dec.pe_fn.grid_fn.gstruct, dec.bound, p["table"], p["coarse"], st0.smooth_pts, None, None, st0.lambda_sm,
                      gv["table"], gv["coarse"], oj_dev=st0.draw_view(st0.draws_dev, "tv", torch.float64))
    nfl = dec.flat.numel()
    for k in ("table", "coarse", "color", "logit", "merge", "experts"):
        assert rel_err(st0._flat_views(st0.grad)[k], gv[k]) < 2e-4, k
    got = st0.loss_vec.cpu()
    for i in range(6):
        torch.testing.assert_close(got[i], losses[i].cpu(), rtol=2e-4, atol=1e-7)
    # pose gradients: per frame, both ranks' rays
    F = st0.F
    dq = torch.zeros(F, 4, device=dec.flat.device)
    dt = torch.zeros(F, 3, device=dec.flat.device)
    off = 0
    for st in steps:
        n = st.n_local
        q1, t1 = torch.empty(F, 4, device=dq.device), torch.empty(F, 3, device=dq.device)
        fused.pose_grad_raw(cam, st.window, d_o[off:off + n].contiguous(), d_d[off:off + n].contiguous(), st.batch["pixel"],
                            st.ray_start, st.quats, q1, t1, torch.empty(12 * F, device=dq.device))
        dq += q1
        dt += t1
        off += n
    assert rel_err(st0.d_quats, dq) < 2e-4 and rel_err(st0.d_trans, dt) < 2e-4
    assert nfl + 7 * F + 12 == st0.packed.numel()
