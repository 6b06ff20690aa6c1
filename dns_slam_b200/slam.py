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
    bottom = torch.tensor([BOTTOM], dtype=torch.float32, device=quad.device)
    return torch.cat([torch.cat((R, T[:, None]), -1), bottom], 0)


def trunc_mask(z, gt_depth):
    """tracking.py:167-170."""
    d = gt_depth[:, None]
    front = (z < d * 0.95).to(z.dtype)
    back = (z > d * 1.05).to(z.dtype)
    return (1.0 - front) * (1.0 - back) * (d > 0.0).to(z.dtype)


def class_tables(label_win):
    """Per-frame tables for the class-balanced draw (common.py:312-322): labels do not change
    between iterations, so ``unique`` / ``nonzero`` are done once per frame, not per iteration.
    Returns (classes ascending, sorted pixel indices, start offsets, counts)."""
    flat = label_win.reshape(-1)
    order = torch.sort(flat, stable=True)[1]
    classes, counts = torch.unique_consecutive(flat[order], return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    return classes, order, starts, counts


def class_balanced_indices(tables, n, draws):
    """common.py:315-330 with the per-class randint draws supplied in order (a class with one
    pixel consumes no draw)."""
    classes, order, starts, counts = tables
    n_class = classes.numel()
    n_k = n // n_class
    counts_h, starts_h = counts.tolist(), starts.tolist()
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
                 lambda_d=5.0, lambda_l=0.1):
        self.cam, self.decoder = cam, decoder
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
                                          lambdas=dict(p=self.lambda_p, d=self.lambda_d, l=self.lambda_l))
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

    def get_target_samples(self, target_frames, quad_list, T_list, refer_frames, features_cl, draws):
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
                w2c.append(torch.inverse(c2w))
            code = fused.feature_matching(self.H, self.W, self.K, pts.flatten(0, 1), torch.stack(w2c, 0),
                                          features_cl[i], self.decoder.merge)
            code = code.reshape(pts.shape[0], pts.shape[1], -1) * trunc_mask(z, s["gt_depth"])[..., None]
            for k, v in (("gt_color", s["gt_color"]), ("gt_depth", s["gt_depth"]), ("gt_label", s["gt_label"]),
                         ("rays_o", rays_o), ("rays_d", rays_d), ("z_vals", z), ("mask", s["inside"]),
                         ("features", code)):
                acc[k].append(v)
        cat = {k: torch.cat(v, 0) for k, v in acc.items()}
        m = cat.pop("mask")
        if bool(m.all()):          # one small D2H read; the reference syncs here too (mapping.py:576)
            return cat
        return {k: v[m] for k, v in cat.items()}

    def iteration(self, target_frames, quad_list, T_list, refer_frames, features_cl, draws, tv_draws,
                  lambda_lt=None, want_latents=False):
        """Body of mapping.py:884-907; returns (loss dict incl. smooth_loss and total, preds, samples)."""
        samples = self.get_target_samples(target_frames, quad_list, T_list, refer_frames, features_cl, draws)
        lam = dict(self.lambdas)
        if lambda_lt is not None:
            lam["lt"] = lambda_lt
        ld, preds = fused.render_and_loss(self.decoder, samples, _lib.MODE_MAP, lambdas=lam,
                                          opacity_sigma=self.opacity_sigma, want_latents=want_latents)
        sm = fused.tv_loss(self.decoder, self.smooth_pts, tv_draws[0], tv_draws[1])
        ld["smooth_loss"] = sm.detach()
        ld["total"] = ld["total"] + self.lambda_sm * sm
        return ld, preds, samples
