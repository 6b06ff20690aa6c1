"""Inference path on the fused kernels: full-frame render and free-point queries.

    render_frame   <->  Mapper.frame_vis           slams/mapping.py:636-690 (without the plotting)
    eval_points    <->  Mesher.eval_points         slams/meshing.py:461-498
    query_points   <->  the batched loops of Mesher.get_mesh, slams/meshing.py:640-655

All three run ``dns_render_fwd_bwd`` with ``forward_only`` set: the point / ray kernels of training, stopped
after the predictions (no backward kernels, no stashes).  Marching cubes, mesh cleaning and the key-frame
projection that produces the per-point pixel features (``get_2d_feature``) are outside the hot path.
"""
import torch

from . import _lib, fused


def render_forward(decoder, samples, mode, want_latents=False, strict=True, forward_only=1):
    """Predictions of ``Tracker.renderer`` / ``Mapper.renderer`` (tracking.py:188-214, mapping.py:603-635)
    without autograd.  samples: rays_o, rays_d, z_vals, gt_label (MAP: classes of the class rule), features."""
    z = samples["z_vals"].contiguous()
    n = z.shape[0]
    dev = z.device
    zeros3 = samples.get("gt_color")
    zeros3 = torch.zeros(n, 3, device=dev) if zeros3 is None else zeros3.contiguous()
    gd = samples.get("gt_depth")
    gd = torch.zeros(n, device=dev) if gd is None else gd.contiguous()
    lab = samples.get("gt_label")
    lab = torch.zeros(n, dtype=torch.int64, device=dev) if lab is None else lab.contiguous()
    is_map = mode == _lib.MODE_MAP
    cfg = fused.RenderConfig(mode, decoder.bound, decoder.pe_fn.grid_fn.gstruct, z, zeros3, gd, lab, None,
                             decoder.class_to_expert if is_map else None, decoder.n_class,
                             dict(p=0.0, d=0.0, l=0.0, lt=0.0, fs=0.0, op=0.0), want_latents=want_latents)
    feats = samples.get("features")
    losses, preds, _, _, _ = fused.render_raw(
        cfg, decoder.view("table"), decoder.view("coarse"), decoder.view("color"), decoder.view("logit"),
        decoder.expert_params.detach() if is_map else None, samples["rays_o"].contiguous(),
        samples["rays_d"].contiguous(), None if feats is None else feats.contiguous(), None, False, False,
        forward_only=forward_only)
    if strict and is_map:      # mapping.py:594-595 / meshing.py:449-450: a class without expert is an error
        flag = float(losses[7])
        if flag < 0:
            raise ValueError("Fine decoders does NOT have class" if flag <= -2 else "label outside [0, n_class_ids)")
    return preds


def render_frame(cam, decoder, frame, c2w, refer_w2c, feats_cl, n_samples_ray, n_surface_ray, t_surface, t_zero,
                 n_pts_batch=100000):
    """Every pixel of ``frame`` (dict color [H,W,3], depth [H,W], label [H,W]) seen from ``c2w``; ONE reference
    view (``refer_w2c`` [4,4], ``feats_cl`` [1,h,w,64]) feeds the pixel features, as frame_vis does.  The renderer
    runs per chunk of ``n_pts_batch`` rays like the reference, so the class rule class(p) = label[p mod n]
    (mapping.py:612-613) sees the same n.  Returns (color [H,W,3], depth [H,W], label [H,W] int64)."""
    H, W = cam["H"], cam["W"]
    dev = decoder.bound.device
    c2w = c2w.to(dev)
    idx = torch.arange(H * W, device=dev)
    s = fused.sample_rays(cam, decoder.bound, frame, idx, (0, H, 0, W), c2w[:3, :3].contiguous(), c2w[:3, 3].contiguous(),
                          n_samples_ray, n_surface_ray, fused.fix_surface_draw(t_surface, n_surface_ray), t_zero)
    w2c = refer_w2c.to(dev).reshape(1, 4, 4)
    K = cam["K"].to(dev)
    cols, deps, labs = [], [], []
    with torch.no_grad():
        for a in range(0, H * W, n_pts_batch):
            b = min(a + n_pts_batch, H * W)
            ro, rd, z = s["rays_o"][a:b], s["rays_d"][a:b], s["z_vals"][a:b]
            pts = ro[:, None, :] + rd[:, None, :] * z[:, :, None]
            code = fused.feature_matching(H, W, K, pts.flatten(0, 1), w2c, feats_cl, decoder.merge)
            smp = dict(rays_o=ro, rays_d=rd, z_vals=z, gt_label=s["gt_label"][a:b],
                       features=code.reshape(b - a, z.shape[1], -1))
            p = render_forward(decoder, smp, _lib.MODE_MAP)
            cols.append(p["color"])
            deps.append(p["depth"])
            labs.append(torch.argmax(p["logits"], -1))
    return torch.cat(cols, 0).reshape(H, W, 3), torch.cat(deps, 0).reshape(H, W), torch.cat(labs, 0).reshape(H, W)


def eval_points(decoder, pts, pixel_pts, gt_label_pts=None, stage="fine"):
    """meshing.py:461-498.  pts [P,3] world coordinates, pixel_pts [P,32] merged pixel features, gt_label_pts [P]
    class of every point (stage 'fine').  Returns (values [P,4] = rgb | occupancy, -100 outside the bound;
    labels [P] int64, -1 outside, or None for stage 'coarse')."""
    bound = decoder.bound
    pts = pts.to(torch.float32).contiguous()
    P = pts.shape[0]
    dev = pts.device
    b = bound.to(pts.dtype)
    inside = ((pts < b[:, 1]) & (pts > b[:, 0])).all(-1)
    fine = stage != "coarse"
    smp = dict(rays_o=pts, rays_d=torch.zeros_like(pts), z_vals=torch.zeros(P, 1, device=dev),
               features=pixel_pts.to(torch.float32).reshape(P, 1, -1))
    if fine:
        smp["gt_label"] = gt_label_pts.to(torch.int64)
    p = render_forward(decoder, smp, _lib.MODE_MAP if fine else _lib.MODE_TRACK, want_latents=True, forward_only=2)
    values = torch.cat((p["color"], p["fine"][:, 0:1]), -1)
    values[~inside, 3] = -100
    if not fine:
        return values, None
    labels = torch.argmax(p["logits"], -1)
    labels[~inside] = -1
    return values, labels


def query_points(decoder, points, feature_fn, stage="fine", points_batch_size=500000):
    """The batched evaluation loop of Mesher.get_mesh (meshing.py:640-655): ``feature_fn(pts) -> (pixel_pts,
    label_pts)`` plays get_2d_feature.  Returns (occupancy [P], labels [P] or None) on the device."""
    occ, labs = [], []
    for a in range(0, points.shape[0], points_batch_size):
        pts = points[a:a + points_batch_size]
        pixel_pts, label_pts = feature_fn(pts)
        v, l = eval_points(decoder, pts, pixel_pts, label_pts, stage)
        occ.append(v[:, 3])
        labs.append(l)
    return torch.cat(occ, 0), (torch.cat(labs, 0) if labs and labs[0] is not None else None)


def get_2d_feature(cam, decoder, points, keyframes):
    """``Mesher.get_2d_feature`` (slams/meshing.py:294-377): merged pixel features [P,32] and labels [P] of free points
    from the key frames that see them -- the producer of ``eval_points``' ``pixel_pts`` / ``gt_label_pts`` in a mesh
    extraction.  keyframes: list of dict(est_c2w [4,4], gt_label [H,W], gt_depth [H,W], features_cl [1,h,w,64] = the
    channels-last stem output of the key frame's colour image, ``encoder.ResNet.forward_cl``).  Projection, masks, rounding
    and the truncation test follow the reference line by line (fp32 torch ops on the device); the 209 MB/key-frame
    up-sample of meshing.py:356 is replaced by the 4-tap bilinear fetch of the half-resolution map at the rounded pixel
    (what ``dns_feature_gather`` does) and Merge runs as the fused kernel (``dns_merge_fwd``)."""
    from . import fused
    H, W = cam["H"], cam["W"]
    dev = points.device
    K = cam["K"].to(dev, torch.float32)
    P = points.shape[0]
    points = points.to(torch.float32)
    pixel_pts = torch.zeros(P, 32, device=dev)
    label_pts = torch.zeros(P, device=dev)
    count_pts = torch.zeros(P, device=dev)
    homo = torch.cat([points, torch.ones_like(points[:, :1])], dim=1).reshape(-1, 4, 1)
    with torch.no_grad():
        for kf in keyframes:
            c2w = kf["est_c2w"].to(dev, torch.float32)
            w2c = torch.inverse(c2w)
            cam_cord = (w2c @ homo)[:, :3]
            cam_cord[:, 0] *= -1
            uv = K @ cam_cord
            z = uv[:, -1:] + 1e-8
            uv = uv[:, :2] / z
            seen = (uv[:, 0] < W) & (uv[:, 0] > 0) & (uv[:, 1] < H) & (uv[:, 1] > 0)
            seen = (seen & (z[:, :, 0] < 0)).reshape(-1)
            uv_ = uv[seen, :, 0]
            if uv_.numel() == 0:
                continue
            p = points[seen, :]
            uv_ = torch.round(uv_).to(torch.int64)
            ui, vi = uv_[:, 0].clamp(0, W - 1), uv_[:, 1].clamp(0, H - 1)
            label_seen = kf["gt_label"].to(dev)[vi, ui]
            depth_seen = kf["gt_depth"].to(dev)[vi, ui]
            depth_proj = -z[seen].reshape(-1)
            trunc = ((~(depth_proj < depth_seen * 0.95)) & (~(depth_proj > depth_seen * 1.05))).to(torch.float32)
            # F.interpolate(..., bilinear, align_corners=True) at the integer pixel (vi, ui) of the half-resolution map
            fm = kf["features_cl"][0]
            h, w = fm.shape[0], fm.shape[1]
            fy = vi.to(torch.float32) * (float(h - 1) / float(H - 1) if H > 1 else 0.0)
            fx = ui.to(torch.float32) * (float(w - 1) / float(W - 1) if W > 1 else 0.0)
            y0, x0 = fy.to(torch.int64), fx.to(torch.int64)
            y1, x1 = y0 + (y0 < h - 1).to(torch.int64), x0 + (x0 < w - 1).to(torch.int64)
            ly1, lx1 = (fy - y0.to(torch.float32))[:, None], (fx - x0.to(torch.float32))[:, None]
            ly0, lx0 = 1.0 - ly1, 1.0 - lx1
            ft = ly0 * (lx0 * fm[y0, x0] + lx1 * fm[y0, x1]) + ly1 * (lx0 * fm[y1, x0] + lx1 * fm[y1, x1])
            refer_p = (p - c2w[:3, 3][None, :])[None].contiguous()
            code = fused.merge_fused(refer_p, ft[None].contiguous(), decoder.merge.decoder.params.detach(), decoder.merge.bound)
            count_pts[seen] += trunc
            pixel_pts[seen, :] += code * trunc[:, None]
            label_pts[seen] = label_seen.to(torch.float32)
        ok = count_pts > 0
        pixel_pts[ok, :] = pixel_pts[ok, :] / count_pts[ok, None]
    return pixel_pts, label_pts
