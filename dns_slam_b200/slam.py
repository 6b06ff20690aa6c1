"""Host-side mirror of the reference's iteration bodies, on the native kernels.

    TrackerCore.get_target_samples / iteration   <->  slams/tracking.py:128-186, 313-340
    MapperCore.get_target_samples / iteration     <->  slams/mapping.py:471-588, 881-910

Same names, argument meaning and loss dictionary as the reference; every random draw is an
explicit input (``draws``), so that the CPU oracle and the GPU consume identical pixel indices
and surface offsets (oracle patch P5).  Orchestration (keyframes, process plumbing, logging,
meshing) is out of scope.
"""
import torch

from . import _lib, fused

BOTTOM = [0.0, 0.0, 0.0, 1.0]


def quad2rotation(quad):
    """utils/common.py:406-429 (device-safe)."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
            two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr),
            two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)]
    return torch.stack(rows, -1).reshape(quad.shape[0], 3, 3)


def get_rotation_from_quad(quad):
    return quad2rotation(quad.unsqueeze(0))[0] if quad.dim() == 1 else quad2rotation(quad)


def c2w_from_quad_T(quad, T):
    R = get_rotation_from_quad(quad)
    return torch.cat([torch.cat((R, T[:, None]), -1), fused.bottom_row(quad.device)], 0)


def trunc_mask(z, gt_depth):
    """tracking.py:167-170."""
    d = gt_depth[:, None]
    front = (z < d * 0.95).to(z.dtype)
    back = (z > d * 1.05).to(z.dtype)
    return (1.0 - front) * (1.0 - back) * (d > 0.0).to(z.dtype)


class ClassTables(tuple):
    """(classes, order, starts, counts) with host copies of the small per-class lists made ONCE per frame,
    so that the per-iteration index draw needs no device->host read."""

    def __new__(cls, classes, order, starts, counts):
        self = super().__new__(cls, (classes, order, starts, counts))
        self.classes_h, self.starts_h, self.counts_h = classes.tolist(), starts.tolist(), counts.tolist()
        return self


def class_tables(label_win):
    """Per-frame tables for the class-balanced draw (common.py:312-322): labels do not change
    between iterations, so ``unique`` / ``nonzero`` are done once per frame, not per iteration.
    Returns (classes ascending, sorted pixel indices, start offsets, counts)."""
    flat = label_win.reshape(-1)
    order = torch.sort(flat, stable=True)[1]
    classes, counts = torch.unique_consecutive(flat[order], return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    return ClassTables(classes, order, starts, counts)


def class_balanced_indices(tables, n, draws):
    """common.py:315-330 with the per-class randint draws supplied in order (a class with one
    pixel consumes no draw)."""
    classes, order, starts, counts = tables
    counts_h, starts_h = tables.counts_h, tables.starts_h
    n_class = len(counts_h)
    n_k = n // n_class
    out, di = [], 0
    for c in range(n_class):
        m = n - n_k * (n_class - 1) if c == 0 else n_k
        if counts_h[c] == 1:
            out.append(order[starts_h[c]].reshape(1).repeat(m))
        else:
            out.append(order[starts_h[c] + draws[di].to(order.device)])
            di += 1
    return torch.cat(out, -1), di


class TrackerCore:
    """The part of ``Tracker`` that sits on the hot path."""

    def __init__(self, cam, decoder, n_pixels, n_samples_ray=32, n_surface_ray=15, lambda_p=5.0,
                 lambda_d=5.0, lambda_l=0.1, freeze_decoder=False):
        self.cam, self.decoder, self.freeze_decoder = cam, decoder, freeze_decoder
        self.H, self.W = cam["H"], cam["W"]
        self.K = cam["K"].to(decoder.bound.device)
        self.n_pixels, self.n_samples_ray, self.n_surface_ray = n_pixels, n_samples_ray, n_surface_ray
        self.lambda_p, self.lambda_d, self.lambda_l = lambda_p, lambda_d, lambda_l

    def get_target_samples(self, cur_frames, refer_frames, features_cl, draws):
        """draws = dict(idx [n] int64, t_surface [15], t_zero [15])."""
        quad, T = cur_frames["est_quad"], cur_frames["est_T"]
        R = get_rotation_from_quad(quad)
        window = (20, self.H - 20, 20, self.W - 20)
        idx = draws["idx"].to(quad.device)
        s = fused.sample_rays(self.cam, self.decoder.bound, cur_frames, idx, window, R, T, self.n_samples_ray,
                              self.n_surface_ray, fused.fix_surface_draw(draws["t_surface"], self.n_surface_ray),
                              draws["t_zero"])
        dirs = fused.pixel_dirs(self.cam, idx, window)
        rays_o, rays_d = fused.attach_pose_grad(s["rays_o"], s["rays_d"], dirs, R, T)
        z = s["z_vals"]
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
        code = fused.feature_matching(self.H, self.W, self.K, pts.flatten(0, 1), refer_frames["est_w2c"].detach(),
                                      features_cl, self.decoder.merge)
        code = code.reshape(pts.shape[0], pts.shape[1], -1) * trunc_mask(z, s["gt_depth"])[..., None]
        mask = (s["gt_depth"] > 0.01) * s["inside"]
        return {"gt_color": s["gt_color"], "gt_depth": s["gt_depth"], "gt_label": s["gt_label"],
                "rays_o": rays_o, "rays_d": rays_d, "pts": pts, "z_vals": z, "mask": mask, "features": code}

    def iteration(self, cur_frames, refer_frames, features_cl, draws):
        """Body of tracking.py:322-329; returns (loss dict, preds, samples)."""
        samples = self.get_target_samples(cur_frames, refer_frames, features_cl, draws)
        ld, preds = fused.render_and_loss(self.decoder, samples, _lib.MODE_TRACK,
                                          lambdas=dict(p=self.lambda_p, d=self.lambda_d, l=self.lambda_l),
                                          freeze_decoder=self.freeze_decoder)
        return ld, preds, samples


class MapperCore:
    """The part of ``Mapper`` that sits on the hot path."""

    def __init__(self, cam, decoder, n_pixels, n_samples_ray=32, n_surface_ray=15, lambdas=None,
                 opacity_sigma=0.05, smooth_pts=64, lambda_sm=1e-5):
        self.cam, self.decoder = cam, decoder
        self.H, self.W = cam["H"], cam["W"]
        self.K = cam["K"].to(decoder.bound.device)
        self.n_pixels, self.n_samples_ray, self.n_surface_ray = n_pixels, n_samples_ray, n_surface_ray
        self.lambdas = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
        self.lambdas.update(lambdas or {})
        self.opacity_sigma, self.smooth_pts, self.lambda_sm = opacity_sigma, smooth_pts, lambda_sm

    def get_target_samples(self, target_frames, quad_list, T_list, refer_frames, features_cl, draws,
                           assume_inside=False):
        """target_frames: dict(kf_idx, frames=[dict(color,depth,label)], class_tables=[...]);
        refer_frames: dict(kf_idx=[[...]], est_c2w=[[...]]); features_cl[i]: [R,h,w,64];
        draws[i] = dict(idx_uniform, class_draws=[...], t_surface, t_zero)."""
        n_t = len(target_frames["frames"])
        n_pixels = self.n_pixels // n_t
        window = (0, self.H, 0, self.W)
        acc = {k: [] for k in ("gt_color", "gt_depth", "gt_label", "rays_o", "rays_d", "z_vals", "mask", "features")}
        target_idx = target_frames["kf_idx"]
        for i in range(n_t):
            fr = target_frames["frames"][i]
            R = get_rotation_from_quad(quad_list[i])
            T = T_list[i]
            cur_c2w = c2w_from_quad_T(quad_list[i], T)
            dev = R.device
            idx1 = draws[i]["idx_uniform"].to(dev)
            idx2, _ = class_balanced_indices(target_frames["class_tables"][i], n_pixels // 3, draws[i]["class_draws"])
            idx = torch.cat((idx1, idx2), 0)
            s = fused.sample_rays(self.cam, self.decoder.bound, fr, idx, window, R, T, self.n_samples_ray,
                                  self.n_surface_ray, fused.fix_surface_draw(draws[i]["t_surface"], self.n_surface_ray),
                                  draws[i]["t_zero"])
            dirs = fused.pixel_dirs(self.cam, idx, window)
            rays_o, rays_d = fused.attach_pose_grad(s["rays_o"], s["rays_d"], dirs, R, T)
            z = s["z_vals"]
            pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
            w2c = []
            for k, rid in enumerate(refer_frames["kf_idx"][i]):
                if rid == -1:
                    c2w = cur_c2w.detach()
                elif rid in target_idx:
                    t = target_idx.index(rid)
                    c2w = c2w_from_quad_T(quad_list[t], T_list[t]).detach()
                else:
                    c2w = refer_frames["est_c2w"][i][k].detach()
                w2c.append(fused.rigid_inverse(c2w))
            code = fused.feature_matching(self.H, self.W, self.K, pts.flatten(0, 1), torch.stack(w2c, 0),
                                          features_cl[i], self.decoder.merge)
            code = code.reshape(pts.shape[0], pts.shape[1], -1) * trunc_mask(z, s["gt_depth"])[..., None]
            for k, v in (("gt_color", s["gt_color"]), ("gt_depth", s["gt_depth"]), ("gt_label", s["gt_label"]),
                         ("rays_o", rays_o), ("rays_d", rays_d), ("z_vals", z), ("mask", s["inside"]),
                         ("features", code)):
                acc[k].append(v)
        cat = {k: torch.cat(v, 0) for k, v in acc.items()}
        m = cat.pop("mask")
        if assume_inside:          # graph replay: no compaction; the flag is checked once after the loop
            ok = m.all()
            self.inside_ok = ok if getattr(self, "inside_ok", None) is None else self.inside_ok & ok
            return cat
        if bool(m.all()):          # one small D2H read; the reference syncs here too (mapping.py:576)
            return cat
        return {k: v[m] for k, v in cat.items()}

    def iteration(self, target_frames, quad_list, T_list, refer_frames, features_cl, draws, tv_draws,
                  lambda_lt=None, want_latents=False, assume_inside=False):
        """Body of mapping.py:884-907; returns (loss dict incl. smooth_loss and total, preds, samples)."""
        samples = self.get_target_samples(target_frames, quad_list, T_list, refer_frames, features_cl, draws,
                                          assume_inside=assume_inside)
        lam = dict(self.lambdas)
        if lambda_lt is not None:
            lam["lt"] = lambda_lt
        ld, preds = fused.render_and_loss(self.decoder, samples, _lib.MODE_MAP, lambdas=lam,
                                          opacity_sigma=self.opacity_sigma, want_latents=want_latents)
        sm = fused.tv_loss(self.decoder, self.smooth_pts, tv_draws[0], tv_draws[1])
        ld["smooth_loss"] = sm.detach()
        ld["total"] = ld["total"] + self.lambda_sm * sm
        return ld, preds, samples


# ----------------------------------------------------------------------------------------
# whole-loop drop-ins: the optimisation loops around the iteration bodies
# ----------------------------------------------------------------------------------------
def quad_from_matrix(R):
    """Rotation matrix -> quaternion (w,x,y,z) (utils/common.py:485-504 without ``mathutils``; the sign
    is immaterial because quad2rotation uses 2/|q|^2)."""
    import numpy as np
    R = np.asarray(R.detach().cpu() if isinstance(R, torch.Tensor) else R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return torch.tensor(q, dtype=torch.float32)


def uniq_class_indices(tables, n, class_list, draws):
    """common.py:375-393 (get_samples_by_uniq_class): n rays spread over the GIVEN classes; class 0 of
    the list takes the remainder, a class with one pixel is repeated, an absent class is skipped."""
    classes, order, starts, counts = tables
    lookup = {int(c): k for k, c in enumerate(tables.classes_h)}
    counts_h, starts_h = tables.counts_h, tables.starts_h
    n_class = len(class_list)
    n_k = n // n_class
    out, di = [], 0
    for pos, cid in enumerate(class_list):
        m = n - n_k * (n_class - 1) if pos == 0 else n_k
        k = lookup.get(int(cid))
        if k is None:
            continue
        if counts_h[k] == 1:
            out.append(order[starts_h[k]].reshape(1).repeat(m))
        else:
            out.append(order[starts_h[k] + draws[di].to(order.device)])
            di += 1
    return torch.cat(out, -1), di


def track_frame(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR=False,
                use_graph=False):
    """The pose-optimisation loop of ``Tracker.run`` (slams/tracking.py:304-346): Adam over
    (translation, quaternion), the best-loss pose is kept.  The ``loss < current_min_loss`` test that
    costs the reference a host sync per iteration (tracking.py:331) stays on the device.
    ``draws_fn(it)`` -> dict(idx, t_surface, t_zero).  Returns (best [quad|T] 7-vector, best loss, losses).
    ``use_graph``: capture ONE iteration (sampling, feature matching, fused render + backward, Adam, best
    pose) in a CUDA graph after three eager warm-up iterations and replay it -- the loop is launch bound
    (hundreds of tiny launches per iteration), not GPU bound."""
    if use_graph:
        return _track_frame_graph(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR)
    dev = tracker.decoder.bound.device
    quad = quad_from_matrix(est_c2w[:3, :3]).to(dev).requires_grad_(True)
    T = est_c2w[:3, 3].detach().clone().to(dev).requires_grad_(True)
    opt = torch.optim.Adam([{"params": [T], "lr": cam_lr * (0.2 if seperate_LR else 1.0)},
                            {"params": [quad], "lr": cam_lr}])
    best_loss = torch.full((), 1e10, device=dev)
    best = torch.cat((quad, T), 0).detach().clone()
    cur = dict(frame, est_quad=quad, est_T=T)
    history = []
    for it in range(n_iters):
        opt.zero_grad()
        cur_w2c = fused.rigid_inverse(c2w_from_quad_T(quad, T))
        est_w2c = torch.stack((refer_w2c.to(dev), cur_w2c), 0)
        ld, _, _ = tracker.iteration(cur, {"est_w2c": est_w2c}, features_cl, draws_fn(it))
        loss = ld["total"]
        with torch.no_grad():
            better = loss < best_loss
            best_loss = torch.where(better, loss, best_loss)
            best = torch.where(better, torch.cat((quad, T), 0), best)
        history.append(loss.detach())
        loss.backward()
        opt.step()
    return best, best_loss, torch.stack(history)


def _track_frame_graph(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR):
    dev = tracker.decoder.bound.device
    quad = quad_from_matrix(est_c2w[:3, :3]).to(dev).requires_grad_(True)
    T = est_c2w[:3, 3].detach().clone().to(dev).requires_grad_(True)
    opt = torch.optim.Adam([{"params": [T], "lr": cam_lr * (0.2 if seperate_LR else 1.0)},
                            {"params": [quad], "lr": cam_lr}], capturable=True)
    d0 = draws_fn(0)
    static = {k: v.to(dev).clone() for k, v in d0.items()}
    best_loss = torch.full((), 1e10, device=dev)
    best = torch.cat((quad, T), 0).detach().clone()
    hist = torch.zeros(n_iters, device=dev)
    slot = torch.zeros((), dtype=torch.int64, device=dev)
    refer = refer_w2c.to(dev)
    cur = dict(frame, est_quad=quad, est_T=T)

    def one():
        opt.zero_grad(set_to_none=True)
        est_w2c = torch.stack((refer, fused.rigid_inverse(c2w_from_quad_T(quad, T))), 0)
        ld, _, _ = tracker.iteration(cur, {"est_w2c": est_w2c}, features_cl, static)
        loss = ld["total"]
        with torch.no_grad():
            better = loss < best_loss
            best.copy_(torch.where(better, torch.cat((quad, T), 0), best))
            best_loss.copy_(torch.where(better, loss, best_loss))
            hist.index_copy_(0, slot.reshape(1), loss.detach().reshape(1))
            slot.add_(1)
        loss.backward()
        opt.step()

    n_warm = min(3, n_iters)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for it in range(n_warm):
            for k, v in draws_fn(it).items():
                static[k].copy_(v)
            one()
    torch.cuda.current_stream().wait_stream(side)
    if n_iters > n_warm:
        for k, v in draws_fn(n_warm).items():
            static[k].copy_(v)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one()
        for it in range(n_warm, n_iters):          # capture only records: the captured iteration is replayed too
            for k, v in draws_fn(it).items():
                static[k].copy_(v, non_blocking=True)
            graph.replay()
    return best, best_loss, hist


def map_optimize(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr, is_BA,
                 new_decoders, draws_fn, tv_draws_fn, use_graph=False):
    """The optimisation loop of ``Mapper.optimize`` (slams/mapping.py:868-910): one Adam over the decoder
    (hash grid + MLPs), the class experts present and -- when ``is_BA`` -- quaternion / translation of every
    target frame but the oldest (mapping.py:457); lambda_lt follows the schedule of mapping.py:898-904.
    Returns (quad_list, T_list, loss dict of the last iteration).
    ``use_graph``: after three eager iterations ONE iteration (4 frames of sampling + feature matching, fused
    render/backward, TV, Adam) is captured in a CUDA graph and replayed with fresh draws.  Replay needs static
    shapes, so it assumes every sampled ray passes the inside test of mapping.py:525 (checked once at the end;
    the call falls back to the eager loop from the saved state otherwise) and a constant lambda_lt."""
    if use_graph and len(new_decoders) == 0 and n_iters > 4:
        out = _map_optimize_graph(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr,
                                  is_BA, draws_fn, tv_draws_fn)
        if out is not None:
            return out
    dec = mapper.decoder
    dev = dec.bound.device
    n_t = len(target_frames["frames"])
    quad_list, T_list = [], []
    for f in range(n_t):
        c2w = est_c2w_list[f]
        q = quad_from_matrix(c2w[:3, :3]).to(dev)
        t = c2w[:3, 3].detach().clone().to(dev)
        if (n_t == 1 or f != 0) and is_BA:
            q.requires_grad_(True)
            t.requires_grad_(True)
        quad_list.append(q)
        T_list.append(t)
    net = [p for p in dec.parameters() if p.requires_grad and p.numel() > 0]
    cam_lr = BA_cam_lr * float(is_BA)
    opt = torch.optim.Adam([{"params": net, "lr": lr},
                            {"params": [q for q in quad_list if q.requires_grad], "lr": cam_lr},
                            {"params": [t for t in T_list if t.requires_grad], "lr": cam_lr}])
    ld = None
    for it in range(n_iters):
        opt.zero_grad()
        lam_lt = (10.0 if it > n_iters // 2 else 0.0) if len(new_decoders) > 0 else 10.0
        ld, _, _ = mapper.iteration(target_frames, quad_list, T_list, refer_frames, features_cl, draws_fn(it),
                                    tv_draws_fn(it), lambda_lt=lam_lt)
        ld["total"].backward()
        opt.step()
    return quad_list, T_list, ld


def decoder_init(mapper, decoder_idx, frame, class_table, cur_c2w, features_cl, lr, draws_fn, tv_draws_fn,
                 n_iters=100, n_rays=300):
    """``Mapper.decoder_init`` (slams/mapping.py:764-836): warm-up of freshly created class experts on the
    current frame: rays spread over the new classes (get_samples_by_uniq_class), losses p + d + l + fs +
    opacity + TV (no latent term), Adam over the decoder and the new experts."""
    dec = mapper.decoder
    dev = dec.bound.device
    for c in decoder_idx:
        dec.activate_expert(c)
    net = [p for p in dec.parameters() if p.requires_grad and p.numel() > 0]
    opt = torch.optim.Adam([{"params": net, "lr": lr}])
    R, T = cur_c2w[:3, :3].to(dev), cur_c2w[:3, 3].to(dev)
    window = (0, mapper.H, 0, mapper.W)
    w2c = torch.inverse(cur_c2w.to(dev)).unsqueeze(0)
    lam = dict(mapper.lambdas, lt=0.0)
    ld = None
    for it in range(n_iters):
        opt.zero_grad()
        d = draws_fn(it)
        idx, _ = uniq_class_indices(class_table, n_rays, decoder_idx, d["class_draws"])
        s = fused.sample_rays(mapper.cam, dec.bound, frame, idx, window, R, T, mapper.n_samples_ray,
                              mapper.n_surface_ray, fused.fix_surface_draw(d["t_surface"], mapper.n_surface_ray),
                              d["t_zero"])
        z = s["z_vals"]
        pts = s["rays_o"][:, None, :] + s["rays_d"][:, None, :] * z[:, :, None]
        code = fused.feature_matching(mapper.H, mapper.W, mapper.K, pts.flatten(0, 1), w2c, features_cl, dec.merge)
        samples = {"gt_color": s["gt_color"], "gt_depth": s["gt_depth"], "gt_label": s["gt_label"],
                   "rays_o": s["rays_o"], "rays_d": s["rays_d"], "z_vals": z,
                   "features": code.reshape(pts.shape[0], pts.shape[1], -1)}   # no trunc mask here (mapping.py:807-809)
        ld, _ = fused.render_and_loss(dec, samples, _lib.MODE_MAP, lambdas=lam, opacity_sigma=mapper.opacity_sigma)
        tv = tv_draws_fn(it)
        sm = fused.tv_loss(dec, mapper.smooth_pts, tv[0], tv[1])
        (ld["total"] + mapper.lambda_sm * sm).backward()
        opt.step()
    return ld


def _to_static(d, dev):
    if isinstance(d, torch.Tensor):
        return d.to(dev).clone()
    if isinstance(d, dict):
        return {k: _to_static(v, dev) for k, v in d.items()}
    return [_to_static(v, dev) for v in d]


def _copy_static(dst, src):
    if isinstance(dst, torch.Tensor):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_static(dst[k], src[k])
    else:
        for a, b in zip(dst, src):
            _copy_static(a, b)


def _map_optimize_graph(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr, is_BA,
                        draws_fn, tv_draws_fn):
    dec = mapper.decoder
    dev = dec.bound.device
    n_t = len(target_frames["frames"])
    saved = dec.flat.detach().clone()
    quad_list, T_list = [], []
    for f in range(n_t):
        c2w = est_c2w_list[f]
        q = quad_from_matrix(c2w[:3, :3]).to(dev)
        t = c2w[:3, 3].detach().clone().to(dev)
        if (n_t == 1 or f != 0) and is_BA:
            q.requires_grad_(True)
            t.requires_grad_(True)
        quad_list.append(q)
        T_list.append(t)
    net = [p for p in dec.parameters() if p.requires_grad and p.numel() > 0]
    cam_lr = BA_cam_lr * float(is_BA)
    groups = [{"params": net, "lr": lr}]
    if any(q.requires_grad for q in quad_list):
        groups += [{"params": [q for q in quad_list if q.requires_grad], "lr": cam_lr},
                   {"params": [t for t in T_list if t.requires_grad], "lr": cam_lr}]
    opt = torch.optim.Adam(groups, capturable=True)
    static_d = _to_static(draws_fn(0), dev)
    static_tv = _to_static(list(tv_draws_fn(0)), dev)
    mapper.inside_ok = None
    last = {}

    def one():
        opt.zero_grad(set_to_none=True)
        ld, _, _ = mapper.iteration(target_frames, quad_list, T_list, refer_frames, features_cl, static_d, static_tv,
                                    lambda_lt=10.0, assume_inside=True)
        ld["total"].backward()
        opt.step()
        for k, v in ld.items():
            if k not in last:
                last[k] = torch.zeros_like(v.detach())
            last[k].copy_(v.detach())

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for it in range(3):
            _copy_static(static_d, draws_fn(it))
            _copy_static(static_tv, list(tv_draws_fn(it)))
            one()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    _copy_static(static_d, draws_fn(3))
    _copy_static(static_tv, list(tv_draws_fn(3)))
    with torch.cuda.graph(graph):
        one()
    for it in range(3, n_iters):
        _copy_static(static_d, draws_fn(it))
        _copy_static(static_tv, list(tv_draws_fn(it)))
        graph.replay()
    ok = bool(mapper.inside_ok)          # the single host read of the loop
    mapper.inside_ok = None
    mapper.last_graph_ok = ok
    if not ok:                           # some ray left the bound: static shapes were wrong, redo eagerly
        with torch.no_grad():
            dec.flat.copy_(saved)
        return None
    return quad_list, T_list, dict(last)
