"""Fused path: sampling, render + loss + backward, TV smoothness, feature matching, Adam.

Python mirrors of the reference call sites, each a thin wrapper over ONE C-ABI call:

    sample_rays        get_samples / far plane / sample_along_rays  (common.py:296-304,561-599;
                                                                      tracking.py:137-160)
    render_and_loss    Tracker.renderer + losses + backward          (tracking.py:188-214,85-96)
                       Mapper.renderer + losses + backward           (mapping.py:590-635,110-126,891-909)
    tv_loss            Mapper.smoothness                             (mapping.py:129-159)
    feature_matching   utils.common.feature_matching                 (common.py:645-679)
    adam_step          torch.optim.Adam.step over a flat buffer      (mapping.py:464-466,910)

``render_and_loss`` is an autograd Function: forward runs the fused forward AND backward
kernels (the loss is a scalar whose normalisers depend on inputs only), ``backward`` scales the
stored gradients by the incoming gradient, so ``loss.backward()`` fills ``.grad`` of the
decoder parameters and of the pose leaves exactly like the reference loop.
"""
import ctypes as C

import torch

from . import _lib

LOSS_KEYS = ("p_loss", "d_loss", "l_loss", "lt_loss", "fs_loss", "opacity_loss", "total", "n_valid")

_ws_cache = {}
_tlin_cache = {}
_use_simt = False


class simt_path:
    """``with fused.simt_path():`` -- the fused calls issued inside run the fp32 SIMT kernels (``use_simt`` of
    dns_render_args / dns_tv_args) instead of the tcgen05 ones.  A/B reference of the parity tests; the library itself
    holds no mode state."""

    def __enter__(self):
        global _use_simt
        self.prev, _use_simt = _use_simt, True
        return self

    def __exit__(self, *exc):
        global _use_simt
        _use_simt = self.prev


def raise_on_flag(losses):
    """ONE device->host read of the fused call's error flag (losses[7] < 0): the reference raises on the spot
    (slams/mapping.py:594-595 'Fine decoders does NOT have class'; torch's cross_entropy on a label outside the head)."""
    flag = float(losses[7])
    if flag < 0:
        raise ValueError({-1: "label outside [0, n_class_ids)", -2: "Fine decoders does NOT have class",
                          -3: "label outside [0, n_class) of the semantic head"}.get(int(flag), f"render error {flag}"))


def workspace(nbytes, device):
    """Cached, growing scratch buffer (caller-owned memory of the C ABI)."""
    key = (device.type, device.index)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _ws_cache[key] = buf = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=device)
    return buf


# ----------------------------------------------------------------------------------------
# sampling
# ----------------------------------------------------------------------------------------
def fix_surface_draw(t, n_surface):
    """common.py:572-573: force one offset to 0.5 unless the draw already holds one."""
    forced = t.clone()
    forced[n_surface // 2 + 1:n_surface // 2 + 2].fill_(0.5)      # fill kernel: no H2D scalar copy (graph capturable)
    return torch.where((t == 0.5).any(), t, forced)      # no host sync


def sample_rays(cam, bound, frame, idx, window, R, T, n_samples, n_surface, t_surface, t_zero,
                want_pts=False, t_lin=None, class_order=None, slot_base=None, n_direct=None, want_pixel=False, out=None,
                phase=0, defer=None):
    """frame: dict(color [H,W,3] f32, depth [H,W] f32, label [H,W] i64) on the GPU; idx: flat
    window indices [n] int64 (the draws of common.py:274 / :327); window = (H0, H1, W0, W1).
    ``class_order`` / ``slot_base`` / ``n_direct``: rays >= n_direct are class-balanced draws resolved on the device
    (``order[slot_base[j] + idx]``, common.py:315-330).  ``out``: pre-allocated output slices (dict) to write into.
    ``defer``: a list -- the filled argument block is appended to it instead of being launched; ``sample_rays_flush``
    then runs all deferred frames in ONE pair of launches (``dns_sample_rays_batch``).
    Returns the ``samples`` fields of tracking.py:177-185 plus ``inside`` (and ``pixel`` with want_pixel)."""
    dev = frame["color"].device
    n = idx.numel()
    S = n_samples + n_surface
    a = _lib.SampleArgs()
    H0, H1, W0, W1 = window
    a.n, a.H, a.W, a.H0, a.W0, a.Ww = n, cam["H"], cam["W"], H0, W0, W1 - W0
    a.n_uniform, a.n_surface = n_samples, n_surface
    a.fx, a.fy, a.cx, a.cy = cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    _lib.fill_bound(a.bound, bound)
    if t_lin is None:
        key = (n_samples, dev.type, dev.index)
        t_lin = _tlin_cache.get(key)
        if t_lin is None:   # torch.linspace evaluated on the CPU as the oracle does, uploaded once
            t_lin = (torch.linspace(0.0, 1.0, steps=n_samples) if n_samples > 0 else torch.zeros(0)).to(dev)
            _tlin_cache[key] = t_lin
    keep = [frame["color"], frame["depth"], frame["label"], idx.contiguous(),
            R.detach().to(dev, torch.float32).contiguous(), T.detach().to(dev, torch.float32).contiguous(),
            t_lin.to(dev, torch.float32).contiguous(), t_surface.to(dev, torch.float32).contiguous(),
            t_zero.to(dev, torch.float32).contiguous()]
    a.color, a.depth = _lib.ptr(keep[0], torch.float32), _lib.ptr(keep[1], torch.float32)
    a.label, a.index = _lib.ptr(keep[2], torch.int64), _lib.ptr(keep[3], torch.int64)
    a.R, a.T, a.t_lin, a.t_surface, a.t_zero = (_lib.ptr(k) for k in keep[4:])
    raw = out is not None
    if out is None:
        out = dict(gt_color=torch.empty(n, 3, device=dev), gt_depth=torch.empty(n, device=dev),
                   gt_label=torch.empty(n, dtype=torch.int64, device=dev), rays_o=torch.empty(n, 3, device=dev),
                   rays_d=torch.empty(n, 3, device=dev), z_vals=torch.empty(n, S, device=dev),
                   inside=torch.empty(n, dtype=torch.uint8, device=dev))
        if want_pts:
            out["pts"] = torch.empty(n, S, 3, device=dev)
        if want_pixel:
            out["pixel"] = torch.empty(n, dtype=torch.int64, device=dev)
        out["scratch"] = torch.empty(2, device=dev)
    for k in ("gt_color", "gt_depth", "gt_label", "rays_o", "rays_d", "z_vals", "inside"):
        setattr(a, k, _lib.ptr(out[k]))
    a.pts = _lib.ptr(out.get("pts"), allow_none=True)
    a.pixel = _lib.ptr(out.get("pixel"), torch.int64, allow_none=True)
    a.scratch = _lib.ptr(out["scratch"])
    if class_order is not None:
        keep += [class_order, slot_base]
        a.order, a.slot_base = _lib.ptr(class_order, torch.int64), _lib.ptr(slot_base, torch.int32)
        a.n_direct = int(n_direct)
    else:
        a.n_direct = n
    a.phase = int(phase)
    if defer is not None:
        defer.append((a, keep))
        return out
    _lib.check(_lib.lib().dns_sample_rays(C.byref(a), _lib.stream()))
    if not raw:
        out["inside"] = out["inside"].bool()
        out.pop("scratch")
    return out


def sample_rays_flush(deferred):
    """Launch the frames collected with ``sample_rays(..., defer=list)``: one gather and one z-value kernel for all of them."""
    if not deferred:
        return
    arr = (_lib.SampleArgs * len(deferred))(*[a for a, _ in deferred])
    _lib.check(_lib.lib().dns_sample_rays_batch(arr, len(deferred), _lib.stream()))
    deferred.clear()


def pixel_dirs(cam, idx, window):
    """Camera-frame directions of the sampled pixels (common.py:257-258)."""
    H0, H1, W0, W1 = window
    Ww = W1 - W0
    i = (W0 + idx % Ww).to(torch.float32)
    j = (H0 + torch.div(idx, Ww, rounding_mode="floor")).to(torch.float32)
    return torch.stack([(i - cam["cx"]) / cam["fx"], -(j - cam["cy"]) / cam["fy"], -torch.ones_like(i)], -1)


def attach_pose_grad(rays_o, rays_d, dirs, R, T):
    """Values stay the kernel's (bit exact); gradients flow to R / T through the torch expression
    of get_rays_from_uv (common.py:262-263):  v + (e - e.detach()) == v exactly."""
    e_d = torch.sum(dirs[:, None, :] * R, -1)
    e_o = T.expand(e_d.shape)
    return rays_o + (e_o - e_o.detach()), rays_d + (e_d - e_d.detach())


# ----------------------------------------------------------------------------------------
# fused render + loss + backward
# ----------------------------------------------------------------------------------------
class RenderConfig:
    """Non-differentiable inputs and knobs of one fused call."""

    def __init__(self, mode, bound, gstruct, z_vals, gt_color, gt_depth, gt_label, mask=None,
                 class_to_expert=None, n_class=40, lambdas=None, opacity_trunc=0.05, opacity_sigma=0.05,
                 want_latents=False):
        self.mode, self.bound, self.gstruct = mode, bound, gstruct
        self.z_vals, self.gt_color, self.gt_depth, self.gt_label = z_vals, gt_color, gt_depth, gt_label
        self.mask, self.class_to_expert, self.n_class = mask, class_to_expert, n_class
        lam = dict(p=5.0, d=5.0, l=0.1, lt=0.0, fs=0.0, op=0.0)
        lam.update(lambdas or {})
        self.lam = lam
        self.opacity_trunc, self.opacity_sigma = opacity_trunc, opacity_sigma
        self.want_latents = want_latents
        # ray sharding: set by shard(); None => single-GPU call
        self.n_rays_total = self.ray_offset = self.gt_label_all = self.global_counts = None

    def shard(self, n_rays_total, ray_offset, gt_label_all, global_counts):
        """This call holds rays [ray_offset, ray_offset + N) of a batch of n_rays_total rays."""
        self.n_rays_total, self.ray_offset = int(n_rays_total), int(ray_offset)
        self.gt_label_all, self.global_counts = gt_label_all, global_counts
        return self


def _fill_inputs(a, cfg):
    f32, i64 = torch.float32, torch.int64
    N, S = cfg.z_vals.shape
    a.mode, a.n_rays, a.n_samples, a.n_class = cfg.mode, N, S, cfg.n_class
    a.opacity_trunc, a.opacity_sigma = cfg.opacity_trunc, cfg.opacity_sigma
    a.z_vals, a.gt_color = _lib.ptr(cfg.z_vals, f32), _lib.ptr(cfg.gt_color, f32)
    a.gt_depth, a.gt_label = _lib.ptr(cfg.gt_depth, f32), _lib.ptr(cfg.gt_label, i64)
    mask8 = None
    if cfg.mode == _lib.MODE_TRACK and cfg.mask is not None:
        mask8 = cfg.mask.to(torch.uint8).contiguous()
    a.mask = _lib.ptr(mask8, allow_none=True)
    if cfg.n_rays_total is not None:
        a.n_rays_total, a.ray_offset = cfg.n_rays_total, cfg.ray_offset
        a.gt_label_all = _lib.ptr(cfg.gt_label_all, i64)
        a.global_counts = _lib.ptr(cfg.global_counts, torch.int32, allow_none=True)
    return mask8


def render_counts(cfg):
    """Local {n_mask, n_depth>0, n_front, n_band} (int32[4] on the device): all-reduce them across
    ranks and hand the sum to ``RenderConfig.shard`` (``dns_render_counts``)."""
    a = _lib.RenderArgs()
    keep = _fill_inputs(a, cfg)
    out = torch.zeros(4, dtype=torch.int32, device=cfg.z_vals.device)
    _lib.check(_lib.lib().dns_render_counts(C.byref(a), _lib.ptr(out), _lib.stream()))
    del keep
    return out


def render_raw(cfg, table, coarse, color, logit, experts, rays_o, rays_d, features, grads, need_drays,
               need_dfeat, forward_only=0, ws=None, features_band_only=False):
    """One ``dns_render_fwd_bwd`` call.  ``grads``: dict with optional device buffers
    ``table/coarse/color/logit/experts`` that are ACCUMULATED into (None => no parameter grads).
    Returns (losses[8], preds dict, d_rays_o, d_rays_d, d_features)."""
    dev = rays_o.device
    N, S = cfg.z_vals.shape
    Cn = cfg.n_class
    a = _lib.RenderArgs()
    mask8 = _fill_inputs(a, cfg)
    need_dparams = grads is not None
    a.need_dparams, a.need_drays, a.need_dfeat = int(need_dparams), int(need_drays), int(need_dfeat)
    a.forward_only = int(forward_only)
    a.use_simt = int(_use_simt)
    a.features_band_only = int(bool(features_band_only) and not _use_simt)
    _lib.fill_bound(a.bound, cfg.bound)
    a.lambda_p, a.lambda_d, a.lambda_l = cfg.lam["p"], cfg.lam["d"], cfg.lam["l"]
    a.lambda_lt, a.lambda_fs, a.lambda_op = cfg.lam["lt"], cfg.lam["fs"], cfg.lam["op"]
    a.grid = cfg.gstruct
    f32, i64 = torch.float32, torch.int64
    a.rays_o, a.rays_d = _lib.ptr(rays_o, f32), _lib.ptr(rays_d, f32)
    a.features = _lib.ptr(features, f32, allow_none=True)
    a.table, a.coarse = _lib.ptr(table, f32), _lib.ptr(coarse, f32)
    a.color, a.logit = _lib.ptr(color, f32), _lib.ptr(logit, f32)
    n_ids = 1
    if cfg.mode == _lib.MODE_MAP:
        if experts is None or cfg.class_to_expert is None:
            raise ValueError("mapping mode needs the expert bank and class_to_expert")
        a.experts = _lib.ptr(experts, f32)
        a.n_experts = experts.shape[0]
        a.class_to_expert = _lib.ptr(cfg.class_to_expert, torch.int32)
        n_ids = cfg.class_to_expert.numel()
    a.n_class_ids = n_ids
    preds = dict(color=torch.empty(N, 3, device=dev), depth=torch.empty(N, device=dev),
                 var=torch.empty(N, device=dev), logits=torch.empty(N, Cn, device=dev))
    a.pred_color, a.pred_depth = _lib.ptr(preds["color"]), _lib.ptr(preds["depth"])
    a.pred_var, a.pred_logits = _lib.ptr(preds["var"]), _lib.ptr(preds["logits"])
    if cfg.want_latents:
        preds["fine"] = torch.empty(N * S, 33, device=dev)
        a.fine = _lib.ptr(preds["fine"])
        if cfg.mode == _lib.MODE_MAP:
            preds["coarse"] = torch.empty(N * S, 33, device=dev)
            a.coarse_out = _lib.ptr(preds["coarse"])
    losses = torch.empty(8, device=dev)
    a.losses = _lib.ptr(losses)
    if need_dparams:
        a.d_table, a.d_coarse = _lib.ptr(grads["table"], f32), _lib.ptr(grads["coarse"], f32)
        a.d_color, a.d_logit = _lib.ptr(grads["color"], f32), _lib.ptr(grads["logit"], f32)
        if cfg.mode == _lib.MODE_MAP:
            a.d_experts = _lib.ptr(grads["experts"], f32)
    d_o = torch.empty(N, 3, device=dev) if need_drays else None
    d_d = torch.empty(N, 3, device=dev) if need_drays else None
    d_f = torch.empty(N, S, 32, device=dev) if need_dfeat else None
    a.d_rays_o, a.d_rays_d = _lib.ptr(d_o, allow_none=True), _lib.ptr(d_d, allow_none=True)
    a.d_features = _lib.ptr(d_f, allow_none=True)
    L = _lib.lib()
    nbytes = L.dns_render_workspace_bytes(cfg.mode, N, S, Cn, n_ids)
    if ws is None:          # shared growing scratch; a caller that captures the call in a CUDA graph passes its own
        ws = workspace(nbytes, dev)
    elif ws.numel() < nbytes:
        raise ValueError("render_raw: the caller's workspace is too small")
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.check(L.dns_render_fwd_bwd(C.byref(a), _lib.stream()))
    return losses, preds, d_o, d_d, d_f


def render_workspace_bytes(mode, n_rays, n_samples, n_class, n_class_ids=1):
    return int(_lib.lib().dns_render_workspace_bytes(mode, n_rays, n_samples, n_class, n_class_ids))


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, table, coarse, color, logit, experts, rays_o, rays_d, features):
        need = ctx.needs_input_grad
        need_dparams = any(need[1:6])
        need_drays = need[6] or need[7]
        need_dfeat = features is not None and need[8]
        grads = None
        if need_dparams:
            grads = dict(table=torch.zeros_like(table), coarse=torch.zeros_like(coarse),
                         color=torch.zeros_like(color), logit=torch.zeros_like(logit),
                         experts=torch.zeros_like(experts) if experts is not None else None)
        feats = features.detach().to(torch.float32).contiguous() if features is not None else None
        losses, preds, d_o, d_d, d_f = render_raw(
            cfg, table.detach(), coarse.detach(), color.detach(), logit.detach(),
            experts.detach() if experts is not None else None, rays_o.detach().contiguous(),
            rays_d.detach().contiguous(), feats, grads, need_drays, need_dfeat)
        ctx.grads, ctx.d_rays, ctx.d_f = grads, (d_o, d_d), d_f
        total = losses[6].clone()
        outs = (total, losses, preds["color"], preds["depth"], preds["var"], preds["logits"],
                preds.get("fine"), preds.get("coarse"))
        ctx.mark_non_differentiable(*[o for o in outs[1:] if o is not None])
        return outs

    @staticmethod
    def backward(ctx, g_total, *unused):
        g, need = ctx.grads, ctx.needs_input_grad

        def sc(t, flag):
            return g_total * t if (flag and t is not None) else None

        gp = g or {}
        return (None, sc(gp.get("table"), need[1]), sc(gp.get("coarse"), need[2]), sc(gp.get("color"), need[3]),
                sc(gp.get("logit"), need[4]), sc(gp.get("experts"), need[5]), sc(ctx.d_rays[0], need[6]),
                sc(ctx.d_rays[1], need[7]), sc(ctx.d_f, need[8]))


def render_and_loss(decoder, samples, mode, n_class=None, lambdas=None, opacity_sigma=0.05, want_latents=False,
                    freeze_decoder=False, strict=False):
    """Fused drop-in for ``renderer(samples)`` + the loss block of the iteration bodies.

    ``samples``: the dict of tracking.py:177-185 / mapping.py:579-586 (``pts`` is not needed: points
    are rebuilt from rays and z).  Returns ``(loss_dict, preds)`` where ``loss_dict['total']`` is
    differentiable w.r.t. the decoder parameters, ``samples['rays_o'/'rays_d']`` and
    ``samples['features']``; the other entries are detached scalars (mapping.py:942-947 keys)."""
    cfg = RenderConfig(mode, decoder.bound, decoder.pe_fn.grid_fn.gstruct, samples["z_vals"].contiguous(),
                       samples["gt_color"].contiguous(), samples["gt_depth"].contiguous(),
                       samples["gt_label"].contiguous(), samples.get("mask"),
                       decoder.class_to_expert if mode == _lib.MODE_MAP else None,
                       n_class or decoder.n_class, lambdas, opacity_trunc=opacity_sigma, want_latents=want_latents)
    experts = decoder.expert_params if mode == _lib.MODE_MAP else None
    feats = samples.get("features")
    prm = [decoder.pe_fn.grid_fn.params, decoder.coarse_fn.decoder.params, decoder.out_fn.color_decoder.params,
           decoder.out_fn.logit_decoder.params, experts]
    if freeze_decoder:   # tracking only moves the pose (tracking.py:108-126): skip every parameter gradient
        prm = [p.detach() if p is not None else None for p in prm]
    out = _RenderFn.apply(cfg, prm[0], prm[1], prm[2], prm[3], prm[4], samples["rays_o"], samples["rays_d"], feats)
    total, losses = out[0], out[1]
    if strict:   # one device->host read: the reference raises on the spot (mapping.py:594-595)
        raise_on_flag(losses)
    ld = {k: losses[i] for i, k in enumerate(LOSS_KEYS)}
    ld["total"] = total
    preds = dict(color=out[2], depth=out[3], var=out[4], logits=out[5], fine=out[6], coarse=out[7])
    return ld, preds


# ----------------------------------------------------------------------------------------
# TV smoothness
# ----------------------------------------------------------------------------------------
def tv_offsets(bound, sample_points, rand3, rand113, voxel_size=0.1, margin=0.05):
    """float64 offset / jitter exactly as mapping.py:133-140 builds them (host side, tiny)."""
    b = bound.detach().double().cpu()          # (a device bound costs a synchronising read: loops pass a host copy)
    volume = b[:, 1] - b[:, 0]
    offset_max = volume - (sample_points - 1) * voxel_size - 2 * margin
    offset = rand3.detach().cpu().to(offset_max) * offset_max + margin
    jitter = rand113.detach().cpu().reshape(3).to(volume)
    return offset, jitter


def tv_offsets_device(bound, sample_points, rand3, rand113, voxel_size=0.1, margin=0.05):
    """Same float64 arithmetic as ``tv_offsets`` evaluated on the device: [offset(3) | jitter(3)] float64 tensor
    (no host round trip, CUDA-graph capturable)."""
    b = bound.detach().double()
    volume = b[:, 1] - b[:, 0]
    offset_max = volume - (sample_points - 1) * voxel_size - 2 * margin
    offset = rand3.to(offset_max) * offset_max + margin
    return torch.cat((offset, rand113.reshape(3).to(volume))).contiguous()


def tv_raw(gstruct, bound, table, coarse, sample_points, offset, jitter, lambda_sm, d_table, d_coarse, oj_dev=None):
    dev = table.device
    a = _lib.TvArgs()
    a.n, a.smooth_pts, a.voxel = sample_points - 1, sample_points, 0.1
    _lib.fill_bound(a.bound, bound)
    if oj_dev is not None:
        a.offset_jitter_dev = _lib.ptr(oj_dev, torch.float64)
    else:
        for k in range(3):
            a.offset[k], a.jitter[k] = float(offset[k]), float(jitter[k])
    a.lambda_sm = lambda_sm
    a.need_dparams = int(d_table is not None)
    a.use_simt = int(_use_simt)
    a.grid = gstruct
    a.table, a.coarse = _lib.ptr(table, torch.float32), _lib.ptr(coarse, torch.float32)
    loss = torch.empty(1, device=dev)
    a.loss = _lib.ptr(loss)
    a.d_table, a.d_coarse = _lib.ptr(d_table, allow_none=True), _lib.ptr(d_coarse, allow_none=True)
    L = _lib.lib()
    ws = workspace(L.dns_tv_workspace_bytes(a.n), dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.check(L.dns_tv_fwd_bwd(C.byref(a), _lib.stream()))
    return loss[0]


class _TvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, coarse, gstruct, bound, sample_points, offset, jitter, oj_dev=None):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d_t = torch.zeros_like(table) if need else None
        d_c = torch.zeros_like(coarse) if need else None
        loss = tv_raw(gstruct, bound, table.detach(), coarse.detach(), sample_points, offset, jitter, 1.0, d_t, d_c, oj_dev)
        ctx.g = (d_t, d_c)
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        d_t, d_c = ctx.g
        return (g * d_t if d_t is not None else None, g * d_c if d_c is not None else None, None, None, None, None, None,
                None)


def tv_loss(decoder, sample_points, rand3, rand113):
    """Mapper.smoothness(sample_points) with the two CPU draws of mapping.py:138,140 as inputs."""
    if rand3.is_cuda:    # draws already on the device: offsets computed there (no host sync, graph capturable)
        oj = tv_offsets_device(decoder.bound, sample_points, rand3, rand113)
        return _TvFn.apply(decoder.pe_fn.grid_fn.params, decoder.coarse_fn.decoder.params,
                           decoder.pe_fn.grid_fn.gstruct, decoder.bound, sample_points, None, None, oj)
    offset, jitter = tv_offsets(decoder.bound, sample_points, rand3, rand113)
    return _TvFn.apply(decoder.pe_fn.grid_fn.params, decoder.coarse_fn.decoder.params,
                       decoder.pe_fn.grid_fn.gstruct, decoder.bound, sample_points, offset, jitter)


# ----------------------------------------------------------------------------------------
# pixel-feature branch
# ----------------------------------------------------------------------------------------
def channels_last(features):
    """[R,C,h,w] -> contiguous [R,h,w,C]; done once per frame, not per iteration."""
    return features.permute(0, 2, 3, 1).contiguous()


def feature_gather(H, W, K, pts, refer_w2c, feats_cl):
    """Projection + rounding + masks + bilinear fetch (common.py:646-670,676).  Non-differentiable
    (the reference's ``torch.round`` cuts the gradient).  Returns code [R,P,C], uv [R,P,2], mask [R,P]."""
    dev = pts.device
    P, R = pts.shape[0], refer_w2c.shape[0]
    Rf, h, w, Cc = feats_cl.shape
    if Rf != R:      # the C entry point only sees raw pointers: a short feature tensor would be read out of bounds
        raise ValueError(f"feature_gather: {R} reference views but feature maps of {Rf} views")
    code = torch.empty(R, P, Cc, device=dev)
    uv = torch.empty(R, P, 2, dtype=torch.int64, device=dev)
    mask = torch.empty(R, P, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().dns_feature_gather(
        _lib.ptr(pts.detach().to(torch.float32).contiguous()), P,
        _lib.ptr(refer_w2c.detach().to(torch.float32).contiguous()), R,
        _lib.ptr(K.detach().to(dev, torch.float32).contiguous()), H, W, _lib.ptr(feats_cl, torch.float32), Cc, h, w,
        _lib.ptr(code), _lib.ptr(uv), _lib.ptr(mask), _lib.stream()))
    return code, uv, mask.bool()


_bottom_cache = {}


def bottom_row(device):
    """[[0,0,0,1]] on ``device``, uploaded once (an H2D copy per call would break CUDA-graph capture)."""
    key = (device.type, device.index)
    b = _bottom_cache.get(key)
    if b is None:
        b = _bottom_cache[key] = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=torch.float32).to(device)
    return b


def rigid_inverse(M):
    """Inverse of rigid transforms [..,4,4] ([R t; 0 1] -> [R^T, -R^T t]); replaces the cuSOLVER-backed
    ``torch.inverse`` of tracking.py:318 / common.py:672 on the native path (same result to fp32 rounding,
    no library launch, CUDA-graph capturable)."""
    R = M[..., :3, :3].transpose(-1, -2)
    t = -(R @ M[..., :3, 3:4])
    top = torch.cat((R, t), -1)
    bottom = M[..., 3:4, :].detach() * 0 + bottom_row(M.device)
    return torch.cat((top, bottom), -2)


class _MergeFn(torch.autograd.Function):
    """``Merge.forward`` (models/decoder.py:67-77) as two fused tcgen05 kernels (dns_merge_fwd / dns_merge_bwd):
    gradients reach ``refer_p`` (and through it the points / poses) and the Merge weights; the gathered
    features are constants (the reference rounds the pixel coordinates, utils/common.py:657)."""

    @staticmethod
    def forward(ctx, refer_p, code, params, bound):
        R, P, _ = refer_p.shape
        dev = refer_p.device
        rp = refer_p.detach().to(torch.float32).contiguous()
        cd = code.detach().to(torch.float32).contiguous()
        L = _lib.lib()
        keep = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[2])
        # the activation images must survive until the backward: a buffer of their own, not the shared scratch
        ws = torch.empty(int(L.dns_merge_workspace_bytes(R * P)), dtype=torch.uint8, device=dev)
        out = torch.empty(P, 32, device=dev)
        b = ((C.c_double * 2) * 3)()
        _lib.fill_bound(b, bound)
        _lib.check(L.dns_merge_fwd(_lib.ptr(rp), _lib.ptr(cd), _lib.ptr(params.detach(), torch.float32), P, R, b,
                                   _lib.ptr(out), int(keep), ws.data_ptr(), ws.numel(), _lib.stream()))
        ctx.save_for_backward(rp, ws, params)
        ctx.bound, ctx.shape = bound, (R, P)
        return out

    @staticmethod
    def backward(ctx, d_out):
        rp, ws, params = ctx.saved_tensors
        R, P = ctx.shape
        d_rp = torch.empty_like(rp)
        d_par = torch.zeros_like(params) if ctx.needs_input_grad[2] else None
        b = ((C.c_double * 2) * 3)()
        _lib.fill_bound(b, ctx.bound)
        _lib.check(_lib.lib().dns_merge_bwd(_lib.ptr(rp), _lib.ptr(d_out.to(torch.float32).contiguous()), P, R, b,
                                            _lib.ptr(d_rp), _lib.ptr(d_par, allow_none=True), ws.data_ptr(), ws.numel(),
                                            _lib.stream()))
        return (d_rp if ctx.needs_input_grad[0] else None), None, d_par, None


def merge_fused(refer_p, code, params, bound):
    """refer_p [R,P,3] (points minus the reference camera centres), code [R,P,64] -> [P,32]."""
    return _MergeFn.apply(refer_p, code, params, bound)


def feature_matching(H, W, K, pts_, refer_w2c, feats_cl, merge_fn, refer_c2w=None):
    """utils.common.feature_matching with channels-last features (no 209 MB/view up-sample).  ``refer_c2w``: the
    poses ``refer_w2c`` was inverted from, when the caller has them (the reference inverts back, common.py:672)."""
    code, _, _ = feature_gather(H, W, K, pts_, refer_w2c.contiguous(), feats_cl)
    refer_o = (rigid_inverse(refer_w2c) if refer_c2w is None else refer_c2w)[:, :3, 3]
    refer_p = pts_[None, :, :] - refer_o[:, None, :]
    return merge_fn(refer_p, refer_o, code)


class Views:
    """Reference views of a ray batch for the fused pixel-feature branch: rays [ray_start[f], ray_start[f+1]) belong to
    target frame f, whose R views are rows f*R.. of ``w2c`` [F*R,4,4] / ``cam_o`` [F*R,3] and ``feats[f]``
    ([R,h,w,64] channels-last, contiguous fp32 CUDA)."""

    def __init__(self, w2c, cam_o, feats, ray_start):
        self.w2c = w2c.detach().to(torch.float32).contiguous()
        self.cam_o = cam_o.detach().to(torch.float32).contiguous()
        self.feats = list(feats)
        self.ray_start = [int(x) for x in ray_start]
        self.F = len(self.feats)
        self.R = self.w2c.shape[0] // self.F
        for f in self.feats:
            if f.shape[0] != self.R or f.shape[-1] != 64:
                raise ValueError(f"feature maps must be [R={self.R},h,w,64] channels-last, got {tuple(f.shape)}")
        if self.F > _lib.MAX_FRAMES or self.R > 8:
            raise ValueError("at most 8 target frames and 8 views per frame in one fused call")


FEATMERGE_TILE_BYTES = 57344      # one stashed operand tile (128 / R band samples x R views)


def featmerge_stash(n_rays, n_samples, n_views, device, band_fraction=0.5):
    """Stash for the operand tiles of ``dns_featmerge_fwd`` sized for ``band_fraction`` of the samples (the truncation
    band holds about a third; a band that does not fit simply makes the backward recompute)."""
    ppt = 128 // n_views
    tiles = int(n_rays * n_samples * band_fraction) // ppt + 2
    return torch.empty(tiles * FEATMERGE_TILE_BYTES, dtype=torch.uint8, device=device)


def _featmerge_args(cam, bound, K, views, rays_o, rays_d, z_vals, gt_depth, params, apply_trunc, ws, stash=None):
    a = _lib.FeatMergeArgs()
    N, S = z_vals.shape
    a.n_rays, a.n_samples, a.n_frames, a.n_views = N, S, views.F, views.R
    for f, r in enumerate(views.ray_start):
        a.ray_start[f] = r
    a.H, a.W = cam["H"], cam["W"]
    a.h, a.w = views.feats[0].shape[1], views.feats[0].shape[2]
    a.apply_trunc = int(apply_trunc)
    _lib.fill_bound(a.bound, bound)
    f32 = torch.float32
    a.K, a.w2c, a.cam_o = _lib.ptr(K, f32), _lib.ptr(views.w2c, f32), _lib.ptr(views.cam_o, f32)
    for f, t in enumerate(views.feats):
        a.feats[f] = _lib.ptr(t, f32)
    a.rays_o, a.rays_d = _lib.ptr(rays_o, f32), _lib.ptr(rays_d, f32)
    a.z_vals, a.gt_depth = _lib.ptr(z_vals, f32), _lib.ptr(gt_depth, f32)
    a.params = _lib.ptr(params, f32)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    if stash is not None:
        a.stash, a.stash_bytes = stash.data_ptr(), stash.numel()
    return a


def _K_dev(cam, dev):
    """Device copy of the intrinsics, cached ON the camera dict (a cache keyed by id() could alias a freed tensor)."""
    cache = cam.setdefault("_K_dev", {})
    key = (dev.type, dev.index)
    k = cache.get(key)
    if k is None:
        k = cache[key] = cam["K"].detach().to(dev, torch.float32).contiguous()
    return k


def featmerge_raw(cam, bound, views, rays_o, rays_d, z_vals, gt_depth, params, apply_trunc=True, ws=None, out=None,
                  stash=None, zero_fill=True):
    """One ``dns_featmerge_fwd`` call.  Returns (features [N,S,32], workspace) -- the workspace holds the band row list
    and must be handed to ``featmerge_bwd_raw`` (like ``stash``, the optional operand-tile stash)."""
    dev = z_vals.device
    N, S = z_vals.shape
    L = _lib.lib()
    if ws is None:
        ws = torch.empty(int(L.dns_featmerge_workspace_bytes(N, S)), dtype=torch.uint8, device=dev)
    a = _featmerge_args(cam, bound, _K_dev(cam, dev), views, rays_o, rays_d, z_vals, gt_depth, params, apply_trunc, ws, stash)
    if out is None:
        out = torch.empty(N, S, 32, device=dev)
    a.features = _lib.ptr(out, torch.float32)
    a.no_zero_fill = int(not zero_fill and apply_trunc)   # rows outside the band stay untouched: render_raw(features_band_only=True)
    _lib.check(L.dns_featmerge_fwd(C.byref(a), _lib.stream()))
    return out, ws


def featmerge_bwd_raw(cam, bound, views, rays_o, rays_d, z_vals, gt_depth, params, d_features, ws, d_params, d_rays_o,
                      d_rays_d, apply_trunc=True, stash=None):
    """One ``dns_featmerge_bwd`` call: ACCUMULATES into d_params / d_rays_o / d_rays_d (None = not wanted)."""
    a = _featmerge_args(cam, bound, _K_dev(cam, z_vals.device), views, rays_o, rays_d, z_vals, gt_depth, params,
                        apply_trunc, ws, stash)
    a.d_features = _lib.ptr(d_features, torch.float32)
    a.need_dparams, a.need_drays = int(d_params is not None), int(d_rays_o is not None)
    a.d_params = _lib.ptr(d_params, torch.float32, allow_none=True)
    a.d_rays_o, a.d_rays_d = _lib.ptr(d_rays_o, allow_none=True), _lib.ptr(d_rays_d, allow_none=True)
    _lib.check(_lib.lib().dns_featmerge_bwd(C.byref(a), _lib.stream()))


class _FeatMergeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, params, cam, bound, views, z_vals, gt_depth, apply_trunc):
        ro, rd = rays_o.detach().contiguous(), rays_d.detach().contiguous()
        pr = params.detach()
        N, S = z_vals.shape
        stash = featmerge_stash(N, S, views.R, z_vals.device, 0.5 if apply_trunc else 1.0) if any(ctx.needs_input_grad[:3]) \
            else torch.empty(0, dtype=torch.uint8, device=z_vals.device)
        out, ws = featmerge_raw(cam, bound, views, ro, rd, z_vals, gt_depth, pr, apply_trunc,
                                stash=stash if stash.numel() else None)
        ctx.save_for_backward(ro, rd, pr, z_vals, gt_depth, ws, stash)
        ctx.meta = (cam, bound, views, apply_trunc)
        return out

    @staticmethod
    def backward(ctx, d_out):
        ro, rd, pr, z_vals, gt_depth, ws, stash = ctx.saved_tensors
        cam, bound, views, apply_trunc = ctx.meta
        need = ctx.needs_input_grad
        want_r = need[0] or need[1]
        d_p = torch.zeros_like(pr) if need[2] else None
        d_o = torch.zeros_like(ro) if want_r else None
        d_d = torch.zeros_like(rd) if want_r else None
        if want_r or need[2]:
            featmerge_bwd_raw(cam, bound, views, ro, rd, z_vals, gt_depth, pr, d_out.to(torch.float32).contiguous(), ws,
                              d_p, d_o, d_d, apply_trunc, stash=stash if stash.numel() else None)
        return (d_o if need[0] else None, d_d if need[1] else None, d_p, None, None, None, None, None, None)


def feature_merge(cam, bound, views, rays_o, rays_d, z_vals, gt_depth, merge_params, apply_trunc=True):
    """``feature_matching`` + ``decoder.merge`` + truncation mask of the iteration bodies (tracking.py:163-171,
    mapping.py:549-557) as ONE fused call each way: -> features [N,S,32], differentiable w.r.t. the rays (through the
    OneBlob of the points) and the Merge weights."""
    return _FeatMergeFn.apply(rays_o, rays_d, merge_params, cam, bound, views, z_vals.contiguous(),
                              gt_depth.contiguous(), bool(apply_trunc))


def pose_prepare(quats, trans, view_src=None, fixed_w2c=None, fixed_cam_o=None):
    """``dns_pose_prepare``: R [F,3,3] of un-normalised quaternions (common.py:406-429) and, with ``view_src``, the
    world-to-camera matrices / camera centres of the reference views (mapping.py:534-551)."""
    dev = quats.device
    F = quats.shape[0]
    R = torch.empty(F, 3, 3, device=dev)
    V = 0 if view_src is None else view_src.numel()
    w2c = torch.empty(V, 4, 4, device=dev) if V else None
    cam_o = torch.empty(V, 3, device=dev) if V else None
    f32 = torch.float32
    _lib.check(_lib.lib().dns_pose_prepare(
        _lib.ptr(quats.detach().contiguous(), f32), _lib.ptr(trans.detach().contiguous(), f32), F,
        _lib.ptr(view_src, torch.int32, allow_none=True), _lib.ptr(fixed_w2c, f32, allow_none=True),
        _lib.ptr(fixed_cam_o, f32, allow_none=True), V, _lib.ptr(R), _lib.ptr(w2c, allow_none=True),
        _lib.ptr(cam_o, allow_none=True), _lib.stream()))
    return R, w2c, cam_o


def pose_grad_raw(cam, window, d_rays_o, d_rays_d, pixel, ray_start, quats, d_quats, d_trans, scratch):
    H0, H1, W0, W1 = window
    rs = (C.c_int32 * len(ray_start))(*[int(x) for x in ray_start])
    f32 = torch.float32
    _lib.check(_lib.lib().dns_pose_grad(
        _lib.ptr(d_rays_o, f32), _lib.ptr(d_rays_d, f32), _lib.ptr(pixel, torch.int64), len(ray_start) - 1, rs, H0, W0,
        W1 - W0, cam["fx"], cam["fy"], cam["cx"], cam["cy"], _lib.ptr(quats, f32, allow_none=True),
        _lib.ptr(d_quats, f32, allow_none=True), _lib.ptr(d_trans, f32, allow_none=True), _lib.ptr(scratch, f32),
        _lib.stream()))


class _PoseRaysFn(torch.autograd.Function):
    """Identity on the sampler's rays in the forward pass (values stay the kernel's, bit exact); the backward is the
    closed form of ``rays_d = R(q) dirs, rays_o = T`` (common.py:257-263, 406-429) in two small kernels
    (``dns_pose_grad``) instead of the autograd chain of the quaternion formula."""

    @staticmethod
    def forward(ctx, quats, trans, rays_o, rays_d, pixel, ray_start, cam, window):
        ctx.save_for_backward(quats.detach().contiguous(), pixel)
        ctx.meta = (ray_start, cam, window)
        return rays_o.view_as(rays_o), rays_d.view_as(rays_d)

    @staticmethod
    def backward(ctx, d_o, d_d):
        quats, pixel = ctx.saved_tensors
        ray_start, cam, window = ctx.meta
        F = quats.shape[0]
        dq, dt = torch.empty_like(quats), torch.empty(F, 3, device=quats.device)
        pose_grad_raw(cam, window, d_o.contiguous(), d_d.contiguous(), pixel, ray_start, quats, dq, dt,
                      torch.empty(12 * F, device=quats.device))
        return dq, dt, None, None, None, None, None, None


def pose_rays(quats, trans, rays_o, rays_d, pixel, ray_start, cam, window):
    """Attach the pose gradient to sampled rays: quats [F,4] / trans [F,3] (stacked leaves), rays of F frames."""
    return _PoseRaysFn.apply(quats, trans, rays_o, rays_d, pixel, list(ray_start), cam, window)


# ----------------------------------------------------------------------------------------
# Adam
# ----------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8):
    """In-place torch.optim.Adam step (defaults) over flat fp32 buffers."""
    _lib.check(_lib.lib().dns_adam_step(_lib.ptr(params, torch.float32), _lib.ptr(grads, torch.float32),
                                        _lib.ptr(exp_avg, torch.float32), _lib.ptr(exp_avg_sq, torch.float32),
                                        params.numel(), lr, betas[0], betas[1], eps, step, _lib.stream()))


class FusedAdam:
    """``torch.optim.Adam`` (defaults) over parameter groups ``[{"params": [...], "lr": lr, "flat": buf?}, ...]`` in
    ONE ``dns_adam_multi`` launch per step -- the groups of slams/tracking.py:119-124 and slams/mapping.py:464-466.

    * Gradients are gathered into buffers owned by the optimiser (``step`` copies every ``p.grad`` into its slice
      with one ``_foreach_copy_``; autograd itself sees ``p.grad = None`` after ``zero_grad``, exactly as with a torch
      optimiser), so the segment table is uploaded once and a CUDA graph that captured ``zero_grad`` / ``backward`` /
      ``step`` replays correctly; the step counter is a device integer.  ``inplace=True`` pre-sets ``p.grad``
      to views of the buffers instead (autograd accumulates in place, no copies).
    * ``flat``: a contiguous buffer that the group's parameters tile exactly (``Decoder.flat``): the group becomes
      one segment, its gradient one flat buffer (one memset per iteration).
    * ``rows`` = ``Decoder.expert_rows()``: the class-expert bank at the tail of ``flat`` is a stack of independent
      parameter tensors in the reference; a row that receives no gradient in an iteration (no sample of that class) is
      skipped -- no moment decay, no step count -- exactly as ``torch.optim.Adam`` skips a parameter whose ``.grad`` is
      None, and every row has its own step count.  Other parameters always get a gradient in the loops here.
    """

    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-8, inplace=False):
        import numpy as np
        self.betas, self.eps = betas, eps
        self.inplace = bool(inplace)
        self._views = []          # (parameter, its slice of the gradient buffer)
        self.groups = [g for g in groups if len(g["params"]) > 0]
        dev = self.groups[0]["params"][0].device
        segs, self._zero, self._keep = [], [], []
        for g in self.groups:
            ps, lr, flat = list(g["params"]), float(g["lr"]), g.get("flat")
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
            covered = flat is not None and sum(p.numel() for p in ps) == flat.numel() and all(
                0 <= p.data_ptr() - flat.data_ptr() <= 4 * (flat.numel() - p.numel()) for p in ps)
            n = flat.numel() if covered else sum(p.numel() for p in ps)
            gbuf, m, v = (torch.zeros(n, device=dev) for _ in range(3))
            self._zero.append(gbuf)
            self._keep += [m, v]
            if covered:
                for p in ps:
                    off = (p.data_ptr() - flat.data_ptr()) // 4
                    self._views.append((p, gbuf[off:off + p.numel()].view_as(p)))
                rows = g.get("rows")      # (offset, n_rows, row_len): the class-expert bank, independent tensors per row
                if rows is not None and rows[0] + rows[1] * rows[2] == n:
                    a = rows[0]
                    steps = torch.zeros(rows[1], dtype=torch.int32, device=dev)
                    self._keep.append(steps)
                    segs.append((flat.data_ptr(), gbuf.data_ptr(), m.data_ptr(), v.data_ptr(), a, lr, 0, 0))
                    segs.append((flat.data_ptr() + 4 * a, gbuf.data_ptr() + 4 * a, m.data_ptr() + 4 * a, v.data_ptr() + 4 * a,
                                 n - a, lr, rows[2], steps.data_ptr()))
                else:
                    segs.append((flat.data_ptr(), gbuf.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr, 0, 0))
            else:
                off = 0
                for p in ps:
                    k = p.numel()
                    self._views.append((p, gbuf[off:off + k].view_as(p)))
                    segs.append((p.data_ptr(), gbuf.data_ptr() + 4 * off, m.data_ptr() + 4 * off,
                                 v.data_ptr() + 4 * off, k, lr, 0, 0))
                    off += k
        tab = np.zeros(len(segs), dtype=ADAM_SEG_DTYPE)
        for i, sg in enumerate(segs):
            tab[i] = sg
        self.table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        self.n_segs, self.max_n = len(segs), max(sg[4] for sg in segs)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.inplace:
            for p, view in self._views:
                p.grad = view

    def reset_state(self):
        """Fresh moments and step counter (a new optimiser object per frame / per optimize() call in the reference),
        in place: a CUDA graph that captured ``step`` keeps replaying on the same buffers."""
        for t in self._keep:
            t.zero_()
        self.step_dev.zero_()

    def zero_grad(self, set_to_none=True):
        if self.inplace:
            for g in self._zero:
                g.zero_()
        else:
            for p, _ in self._views:
                p.grad = None

    def step(self):
        if not self.inplace:
            dst, src = [], []
            for p, view in self._views:
                if p.grad is None:
                    view.zero_()
                else:
                    dst.append(view)
                    src.append(p.grad)
            if dst:
                torch._foreach_copy_(dst, src)
        _lib.check(_lib.lib().dns_adam_multi(_lib.ptr(self.table), self.n_segs, self.max_n,
                                             _lib.ptr(self.step_dev, torch.int32), self.betas[0], self.betas[1],
                                             self.eps, _lib.stream()))


def _adam_seg_dtype():
    import numpy as np
    return np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("lr", "<f4"), ("row_len", "<i4"),
                     ("row_steps", "<u8")])


ADAM_SEG_DTYPE = _adam_seg_dtype()      # mirror of dns_adam_seg (56 bytes)


def split_expert_rows(flat, grad, lr, rows):
    """One flat (params, grads, lr) segment -> [plain part, expert rows]: ``rows`` = (offset, n_rows, row_len) of the class
    expert bank inside the flat buffer (``Decoder.expert_rows()``).  The experts are independent parameter tensors in the
    reference (slams/mapping.py:445-446), so Adam treats each row on its own (``dns_adam_seg.row_len``)."""
    off, n_rows, row_len = rows
    if off + n_rows * row_len != flat.numel():
        raise ValueError("the expert bank must be the tail of the flat buffer")
    return [(flat[:off], grad[:off], lr, 0), (flat[off:], grad[off:], lr, row_len)]


class AdamSegments:
    """Raw-buffer Adam (torch defaults) over segments ``[(params, grads, lr), ...]`` of flat fp32 CUDA tensors in ONE
    ``dns_adam_multi`` launch per step; owns the moments and a device-side step counter (fresh state per object, as
    the reference builds a new optimiser per ``optimize()`` call, slams/mapping.py:438-468)."""

    def __init__(self, segments, betas=(0.9, 0.999), eps=1e-8):
        import numpy as np
        self.betas, self.eps = betas, eps
        self.segments = [(sg[0], sg[1], float(sg[2]), int(sg[3]) if len(sg) > 3 else 0) for sg in segments if sg[0].numel() > 0]
        dev = self.segments[0][0].device
        tab = np.zeros(len(self.segments), dtype=ADAM_SEG_DTYPE)
        self._keep = []
        for i, (p, g, lr, row_len) in enumerate(self.segments):
            if not (p.is_cuda and g.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()):
                raise RuntimeError("AdamSegments: contiguous fp32 CUDA tensors only (no CPU fallback)")
            m, v = torch.zeros_like(p), torch.zeros_like(p)
            steps = torch.zeros(p.numel() // row_len, dtype=torch.int32, device=dev) if row_len else None
            self._keep += [m, v, steps]
            tab[i] = (p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, row_len,
                      steps.data_ptr() if row_len else 0)
        self.table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        self.n_segs, self.max_n = len(self.segments), max(sg[0].numel() for sg in self.segments)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)

    def step(self):
        _lib.check(_lib.lib().dns_adam_multi(_lib.ptr(self.table), self.n_segs, self.max_n,
                                             _lib.ptr(self.step_dev, torch.int32), self.betas[0], self.betas[1],
                                             self.eps, _lib.stream()))

    def reset_state(self):
        """Fresh moments and step counts (a new optimiser per frame / per optimize() call in the reference)."""
        for t in self._keep:
            if t is not None:
                t.zero_()
        self.step_dev.zero_()


use_torch_adam = False      # A/B hook of tests / measurements: torch.optim.Adam instead of FusedAdam


def make_adam(groups, capturable=False):
    """The optimiser of the tracking / mapping loops: ``FusedAdam`` (``fused.use_torch_adam = True`` selects
    ``torch.optim.Adam``, the A/B reference)."""
    if use_torch_adam:
        gs = [{"params": g["params"], "lr": g["lr"]} for g in groups if len(g["params"]) > 0]
        return torch.optim.Adam(gs, capturable=capturable)
    return FusedAdam(groups)
