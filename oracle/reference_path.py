"""TEST INFRASTRUCTURE ONLY -- CPU (PyTorch, fp32/fp64) restatement of the DNS-SLAM hot path.

Each function cites the reference lines it follows (paths relative to /root/reference).
The only deliberate deviations are the documented oracle patches of SURVEY.md section 8c:

  P1  slams/mapping.py:152   ``occ.reshape(pts_shape[:3], 1)`` -> ``reshape(*pts_shape[:3], 1)``
  P2  utils/common.py:419    ``.to(quad.get_device())``        -> ``device=quad.device``
  P3  utils/common.py:461-504 matrix -> quaternion without ``mathutils``
  P4  models/layers.py:125   no S3 download: pixel feature maps are synthetic inputs
  P5  every ``torch.randint`` / ``torch.rand`` draw is taken from a ``DrawTape`` so the CPU
      oracle and the CUDA path consume identical pixel indices and surface offsets.

Pinned against the reference's own code by tests/test_oracle_golden.py (see oracle/__init__.py).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import tcnn_standin as tcnn


# --------------------------------------------------------------------------------------
# P5: explicit random draws
# --------------------------------------------------------------------------------------
class DrawTape:
    """Replays recorded draws in call order, or draws fresh ones from a seeded generator
    (and records them).  ``kind`` is 'randint' or 'rand'."""

    def __init__(self, items=None, seed=None):
        self.replay = items is not None
        self.items = list(items) if items is not None else []
        self.pos = 0
        self.gen = torch.Generator().manual_seed(0 if seed is None else seed)

    def _next(self, kind, shape, high=None):
        if self.replay:
            k, v = self.items[self.pos]
            self.pos += 1
            assert k == kind and tuple(v.shape) == tuple(shape), (k, kind, v.shape, shape)
            return v.clone()
        if kind == "randint":
            v = torch.randint(high, shape, generator=self.gen)
        else:
            v = torch.rand(shape, generator=self.gen)
        self.items.append((kind, v.clone()))
        return v

    def randint(self, high, shape):
        return self._next("randint", tuple(shape), high)

    def rand(self, shape):
        return self._next("rand", tuple(shape))


# --------------------------------------------------------------------------------------
# scene bound  (slams/dns_slam.py:100-107)
# --------------------------------------------------------------------------------------
def load_bound(bound, scale=1.0, bound_divisible=0.32):
    b = torch.from_numpy(np.array(bound, dtype=np.float64) * scale)
    b[:, 1] = (((b[:, 1] - b[:, 0]) / bound_divisible).int() + 1) * bound_divisible + b[:, 0]
    return b


# --------------------------------------------------------------------------------------
# poses  (utils/common.py:406-458, patches P2 / P3)
# --------------------------------------------------------------------------------------
def quad2rotation(quad):
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [
        1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
        two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr),
        two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2),
    ]
    return torch.stack(rows, -1).reshape(quad.shape[0], 3, 3)


def get_rotation_from_quad(quad):
    if quad.dim() == 1:
        return quad2rotation(quad.unsqueeze(0))[0]
    return quad2rotation(quad)


def get_camera_from_tensor(inputs):
    one = inputs.dim() == 1
    if one:
        inputs = inputs.unsqueeze(0)
    rt = torch.cat([quad2rotation(inputs[:, :4]), inputs[:, 4:, None]], 2)
    return rt[0] if one else rt


def quad_from_matrix(R):
    """P3: rotation matrix -> quaternion (w,x,y,z); sign is immaterial (R uses 2/|q|^2)."""
    R = np.asarray(R, dtype=np.float64)
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


def c2w_from_quad_T(quad, T):
    R = get_rotation_from_quad(quad)
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=torch.float32, device=quad.device)
    return torch.cat([torch.cat((R, T[:, None]), -1), bottom], 0)


# --------------------------------------------------------------------------------------
# pixel / ray sampling  (utils/common.py:248-361)
# --------------------------------------------------------------------------------------
def uv_from_flat(idx, H0, W0, Ww):
    """i (x) and j (y) pixel coordinates, fp32, of flat window indices (common.py:288-291)."""
    i = (W0 + (idx % Ww)).to(torch.float32)
    j = (H0 + torch.div(idx, Ww, rounding_mode="floor")).to(torch.float32)
    return i, j


def rays_from_uv(i, j, R, T, fx, fy, cx, cy):
    """common.py:248-264; the 3-term sum is evaluated left to right."""
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    p = dirs.reshape(-1, 1, 3) * R
    rays_d = (p[..., 0] + p[..., 1]) + p[..., 2]
    rays_o = T.expand(rays_d.shape)
    return rays_o, rays_d


def uniform_indices(H0, H1, W0, W1, n, tape):
    """common.py:274: one device randint over the window."""
    return tape.randint((H1 - H0) * (W1 - W0), (n,))


def class_balanced_indices(label_win, n, tape):
    """common.py:307-330: per class (ascending label) ``nonzero`` + randint; class 0 of the
    sorted list takes the remainder; a class with exactly one pixel is repeated, no draw."""
    flat = label_win.reshape(-1)
    classes = torch.unique(flat, sorted=True)
    n_class = classes.numel()
    n_k = n // n_class
    out = []
    for c in range(n_class):
        m = n - n_k * (n_class - 1) if c == 0 else n_k
        members = torch.nonzero(flat == classes[c]).reshape(-1)
        k = members.numel()
        if k == 1:
            out.append(members.repeat(m))
        else:
            out.append(members[tape.randint(k, (m,))])
    return torch.cat(out, -1)


def gather_window(image5, H0, H1, W0, W1, idx):
    return image5[H0:H1, W0:W1].reshape(-1, image5.shape[-1])[idx]


# --------------------------------------------------------------------------------------
# far plane + depth-guided z sampling  (tracking.py:148-159, mapping.py:519-530,
# utils/common.py:561-599)
# --------------------------------------------------------------------------------------
def far_plane(rays_o, rays_d, bound, gt_depth):
    with torch.no_grad():
        o = rays_o.detach().unsqueeze(-1)
        d = rays_d.detach().unsqueeze(-1)
        t = (bound.unsqueeze(0) - o) / d                    # float64 [N,3,2]
        far_bb, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
        inside = far_bb >= gt_depth
        far_bb = far_bb.unsqueeze(-1) + 0.01
    return far_bb, inside


def sample_along_rays(gt_depth, n_samples, n_surface, far_bb, tape):
    gt_depth = gt_depth.reshape(-1, 1)
    nz = (gt_depth > 0).squeeze(-1)
    t_surf = tape.rand((n_surface,)).to(gt_depth.device)
    if not torch.any(t_surf == 0.5):
        t_surf[n_surface // 2 + 1] = 0.5
    d_nz = gt_depth[nz].reshape(-1, 1).repeat(1, n_surface)
    z_near = torch.zeros(gt_depth.shape[0], n_surface, device=gt_depth.device)
    z_near[nz, :] = (0.95 * d_nz * (1.0 - t_surf) + 1.05 * d_nz * t_surf).to(z_near.dtype)
    far = torch.max(gt_depth)
    t_zero = tape.rand((n_surface,)).to(gt_depth.device)
    z_near[~nz, :] = (0.001 * (1.0 - t_zero) + far * t_zero).to(z_near.dtype)
    if n_samples > 0:
        near = gt_depth.repeat(1, n_samples) * 0.001
        far = torch.clamp(far_bb, 0, torch.max(gt_depth * 1.2))
        t_vals = torch.linspace(0.0, 1.0, steps=n_samples, device=gt_depth.device)
        z = near * (1.0 - t_vals) + far * t_vals
        z, _ = torch.sort(torch.cat([z, z_near], -1), -1)
    else:
        z, _ = torch.sort(z_near, -1)
    return z.float()


def trunc_mask(z, gt_depth):
    """tracking.py:167-170 / mapping.py:553-556."""
    d = gt_depth[:, None]
    front = (z < d * 0.95).to(z.dtype)
    back = (z > d * 1.05).to(z.dtype)
    valid = (d > 0.0).to(z.dtype)
    return (1.0 - front) * (1.0 - back) * valid


# --------------------------------------------------------------------------------------
# model  (models/decoder.py, models/pos_encoding.py)
# --------------------------------------------------------------------------------------
_MLP_CFG = {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None",
            "n_neurons": 32, "n_hidden_layers": 1}


def _mlp(n_in, n_out, width, seed):
    cfg = dict(_MLP_CFG, n_neurons=width)
    return tcnn.Network(n_in, n_out, cfg, seed=seed)


class PosEncoding(nn.Module):
    """decoder.py:30-48 + pos_encoding.py:31-46,61-71."""

    def __init__(self, cfg, bound, seed=0):
        super().__init__()
        self.pe_fn = tcnn.Encoding(3, {"otype": "OneBlob", "n_bins": cfg["pos"]["n_bins"]})
        self.pe_dim = self.pe_fn.n_output_dims
        dim_max = (bound[:, 1] - bound[:, 0]).max()
        self.resolution = int(dim_max / cfg["grid"]["voxel_size"])
        pls = np.exp2(np.log2(self.resolution / 16) / (16 - 1))
        self.grid_fn = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": 16,
                                         "n_features_per_level": 2,
                                         "log2_hashmap_size": cfg["grid"]["hash_size"],
                                         "base_resolution": 16, "per_level_scale": pls},
                                     seed=seed + 1)
        self.grid_dim = self.grid_fn.n_output_dims

    def forward(self, pts):
        return self.pe_fn(pts), self.grid_fn(pts)


class Merge(nn.Module):
    """decoder.py:51-77."""

    def __init__(self, cfg, hidden_dim, feature_dim, bound, seed=0):
        super().__init__()
        self.bound = bound
        self.pe_fn = tcnn.Encoding(3, {"otype": "OneBlob", "n_bins": cfg["pos"]["n_bins"]})
        self.pe_dim = self.pe_fn.n_output_dims
        self.decoder = _mlp(self.pe_dim + feature_dim, hidden_dim, hidden_dim, seed + 5)

    def forward(self, p, o, features=None):
        n_refer, n_points, _ = features.shape
        p = (p - self.bound[:, 0]) / (self.bound[:, 1] - self.bound[:, 0])
        pe = self.pe_fn(p.flatten(0, 1))
        lat = self.decoder(torch.cat((pe, features.flatten(0, 1)), -1))
        return torch.mean(lat.reshape(n_refer, n_points, -1), 0)


class Coarse(nn.Module):
    """decoder.py:80-94."""

    def __init__(self, pts_dim, hidden_dim, feature_dim, seed=0):
        super().__init__()
        self.decoder = _mlp(pts_dim + feature_dim, hidden_dim + 1, hidden_dim, seed + 2)

    def forward(self, pe, features=None):
        return self.decoder(torch.cat((pe, features), -1)).float()


class Out(nn.Module):
    """decoder.py:97-125."""

    def __init__(self, pts_dim, feature_dim, hidden_dim, n_class, seed=0):
        super().__init__()
        self.color_decoder = _mlp(pts_dim + feature_dim, 3, hidden_dim, seed + 3)
        self.logit_decoder = _mlp(pts_dim + feature_dim, n_class, hidden_dim, seed + 4)

    def forward(self, pe, features):
        x = torch.cat((pe, features), -1)
        return torch.sigmoid(self.color_decoder(x)), self.logit_decoder(x)


class Decoder(nn.Module):
    """decoder.py:7-27; attribute names (hence state_dict keys) follow the reference."""

    def __init__(self, cfg, bound, n_class=40, seed=0):
        super().__init__()
        self.pe_fn = PosEncoding(cfg, bound, seed)
        self.pe_dim, self.grid_dim = self.pe_fn.pe_dim, self.pe_fn.grid_dim
        self.hidden_dim, self.pixel_dim, self.n_class = cfg["hidden_dim"], cfg["pixel_dim"], n_class
        self.coarse_fn = Coarse(self.pe_dim, self.hidden_dim, self.grid_dim, seed)
        self.out_fn = Out(self.pe_dim, self.hidden_dim * 2, self.hidden_dim, n_class, seed)
        self.merge = Merge(cfg, self.hidden_dim, self.pixel_dim, bound, seed)


def new_expert(pe_dim=48, grid_dim=32, hidden_dim=32, seed=100):
    """mapping.py:737-744: one class-wise fine MLP 80 -> 32 -> 33."""
    return _mlp(pe_dim + grid_dim, hidden_dim + 1, hidden_dim, seed)


# --------------------------------------------------------------------------------------
# pixel-feature branch  (utils/common.py:632-679)
# --------------------------------------------------------------------------------------
def feature_matching(H, W, K, pts_, refer_w2c, features, merge_fn):
    features = F.interpolate(features, size=[H, W], mode="bilinear", align_corners=True)
    ones = torch.ones(pts_.shape[0], 1, device=pts_.device)
    pts = torch.cat((pts_, ones), -1)
    cam = torch.matmul(refer_w2c, pts.permute(1, 0))           # [R,4,P]
    cam = torch.cat((cam[:, 0:1], -cam[:, 1:2], -cam[:, 2:3], cam[:, 3:4]), 1)
    proj_depth = cam[:, 2, :]
    img = torch.matmul(K[None], cam[:, :3, :])
    uv = img[:, :2, :] / (img[:, 2:3, :] + 1e-5)
    uv = torch.round(uv.permute(0, 2, 1))
    mask = ((uv[:, :, 0] > 0) * (uv[:, :, 0] < W - 1) * (uv[:, :, 1] > 0) * (uv[:, :, 1] < H - 1)
            * (proj_depth > 0))
    uv_i = (uv * mask[:, :, None]).to(torch.int64)
    code = []
    for r in range(features.shape[0]):
        h = uv_i[r, :, 1].clamp(0, H - 1)
        w = uv_i[r, :, 0].clamp(0, W - 1)
        code.append(features[r][:, h, w])
    code = torch.stack(code, 0).permute(0, 2, 1)                # [R,P,C]
    refer_c2w = torch.inverse(refer_w2c)
    refer_o = refer_c2w[:, :3, 3]
    refer_p = pts_[None] - refer_o[:, None, :]
    code = code * mask[:, :, None]
    return merge_fn(refer_p, refer_o, code), uv_i, mask


# --------------------------------------------------------------------------------------
# compositing + losses  (utils/common.py:506-537,764-802; tracking.py:85-96; mapping.py:110-126)
# --------------------------------------------------------------------------------------
def raw2nerf_color(raw, z_vals):
    rgb = raw[..., :3]
    alpha = torch.sigmoid(10 * raw[..., -1])
    ones = torch.ones((alpha.shape[0], 1), device=z_vals.device)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
    w = alpha * trans
    w = w / w.sum(dim=-1)[:, None]
    rgb_map = torch.sum(w[..., None] * rgb, -2)
    depth_map = torch.sum(w * z_vals, -1)
    tmp = z_vals - depth_map.unsqueeze(-1)
    depth_var = torch.sum(w * tmp * tmp, dim=-1)
    return depth_map, depth_var, rgb_map, w


def get_opacity_loss(z_vals, depth, occ, truncation=0.2, sigma=0.05):
    bs, n_sample = z_vals.shape
    depth = depth.unsqueeze(-1)
    occ = torch.sigmoid(10 * occ).reshape(bs, n_sample)
    front = (z_vals < depth - truncation).to(z_vals.dtype)
    back = (z_vals > depth + truncation).to(z_vals.dtype)
    valid = (depth > 0.0).to(z_vals.dtype)
    band = (1.0 - front) * (1.0 - back) * valid
    if torch.count_nonzero(front) > 0 and torch.count_nonzero(band) > 0:
        fs = ((occ * front * valid) ** 2).mean()
        pseudo = 0.5 * torch.exp(-0.5 * ((z_vals - depth) / sigma) ** 2)
        op = ((occ * band - pseudo * band) ** 2).mean()
    else:
        fs = torch.tensor(0.0)
        op = torch.tensor(0.0)
    return fs, op


def tracking_losses(samples, pred_color, pred_depth, pred_var, pred_logits):
    m = samples["mask"]
    p = ((samples["gt_color"][m, :] - pred_color[m, :]) ** 2).mean()
    d = (torch.abs(samples["gt_depth"] - pred_depth) / torch.sqrt(pred_var + 1e-10))[m].mean()
    l = F.cross_entropy(pred_logits[m, :], samples["gt_label"][m])
    return p, d, l


def mapping_losses(samples, pred_color, pred_depth, pred_logits, fine, coarse, opacity_sigma):
    gd = samples["gt_depth"]
    m = gd > 0
    d = torch.abs(gd[m] - pred_depth[m]).mean()
    p = ((samples["gt_color"] - pred_color) ** 2).mean()
    l = F.cross_entropy(pred_logits, samples["gt_label"])
    lt = ((coarse - fine) ** 2).mean()
    # quirk (mapping.py:896): opacity_sigma lands in the ``truncation`` slot; channel 32 is read
    fs, op = get_opacity_loss(samples["z_vals"], gd, fine[..., -1], opacity_sigma)
    return p, d, l, lt, fs, op


# --------------------------------------------------------------------------------------
# renderers  (tracking.py:188-214, mapping.py:590-635)
# --------------------------------------------------------------------------------------
def normalise(pts, bound):
    return (pts - bound[:, 0]) / (bound[:, 1] - bound[:, 0])


def tracker_renderer(decoder, bound, samples):
    pts = normalise(samples["pts"].flatten(0, 1), bound)
    z_vals = samples["z_vals"]
    n, s = z_vals.shape
    pix = samples["features"].flatten(0, 1)
    pe, grid = decoder.pe_fn(pts)
    lat = decoder.coarse_fn(pe, features=grid)
    color, logits = decoder.out_fn(pe, torch.cat((lat[:, 1:], pix), -1))
    values = torch.cat((color, lat[:, 0:1]), -1).reshape(n, s, -1)
    logits = logits.reshape(n, s, -1)
    depth, var, rgb, w = raw2nerf_color(values, z_vals)
    return rgb, depth, var, torch.sum(w[..., None] * logits, -2)


def fine_fn(experts, hidden_dim, pes, classes, features):
    """mapping.py:590-601; experts is {class id: Network}."""
    present = torch.unique(classes, sorted=True).cpu().int().numpy()
    lat = torch.zeros(pes.shape[0], hidden_dim + 1, device=pes.device)
    for c in present:
        if int(c) not in experts:
            raise ValueError("Fine decoders does NOT have class", c)
        sel = classes == int(c)
        if int(sel.sum()) > 1:
            lat[sel, :] = experts[int(c)](torch.cat((pes[sel], features[sel]), -1)).float()
    return lat


def mapper_renderer(decoder, experts, bound, samples):
    pts = samples["pts"]
    n, s, _ = pts.shape
    x = normalise(pts.flatten(0, 1), bound)
    z_vals = samples["z_vals"]
    # quirk (mapping.py:612-613): 1-D repeat tiles the labels, so class(p) = label[p mod n]
    cls = samples["gt_label"].repeat(1, s).flatten(0, 1)
    pix = samples["features"].flatten(0, 1)
    pe, grid = decoder.pe_fn(x)
    coarse = decoder.coarse_fn(pe, features=grid)
    fine = fine_fn(experts, decoder.hidden_dim, pe, cls, grid)
    color, logits = decoder.out_fn(pe, torch.cat((fine[:, 1:], pix), -1))
    values = torch.cat((color, fine[:, 0:1]), -1).reshape(n, s, -1)
    logits = logits.reshape(n, s, -1)
    depth, var, rgb, w = raw2nerf_color(values, z_vals)
    return rgb, depth, var, torch.sum(w[..., None] * logits, -2), fine, coarse


# --------------------------------------------------------------------------------------
# TV smoothness  (mapping.py:129-159, patch P1)
# --------------------------------------------------------------------------------------
def smoothness_points(bound, sample_points, tape, voxel_size=0.1, margin=0.05):
    volume = bound[:, 1] - bound[:, 0]
    grid_size = (sample_points - 1) * voxel_size
    offset_max = volume - grid_size - 2 * margin
    offset = tape.rand((3,)).to(offset_max) * offset_max + margin
    n = sample_points - 1
    ar = torch.arange(0, n, dtype=torch.long)
    cx, cy, cz = torch.meshgrid(ar, ar, ar, indexing="ij")
    coords = torch.stack([cx, cy, cz], -1).float().to(volume)
    pts = (coords + tape.rand((1, 1, 1, 3)).to(volume)) * voxel_size + bound[:, 0] + offset
    return normalise(pts, bound)                                 # float64 [n,n,n,3]


def smoothness(decoder, bound, sample_points, tape):
    pts = smoothness_points(bound, sample_points, tape)
    shp = pts.shape
    pe, grid = decoder.pe_fn(pts.reshape(-1, 3))
    occ = decoder.coarse_fn(pe, features=grid)[:, 0:1].reshape(*shp[:3], 1)
    tv = ((occ[1:] - occ[:-1]) ** 2).sum() + ((occ[:, 1:] - occ[:, :-1]) ** 2).sum() \
        + ((occ[:, :, 1:] - occ[:, :, :-1]) ** 2).sum()
    return tv / (sample_points ** 3)


# --------------------------------------------------------------------------------------
# sample stages  (tracking.py:128-186, mapping.py:471-588)
# --------------------------------------------------------------------------------------
def tracker_get_target_samples(cam, bound, decoder, frame, quad, T, refer_w2c, feats,
                               n_pixels, n_samples, n_surface, tape):
    """cam = dict(H,W,fx,fy,cx,cy,K). frame = dict(color,depth,label)."""
    H, W = cam["H"], cam["W"]
    R = get_rotation_from_quad(quad)
    img = torch.cat((frame["color"], frame["depth"].unsqueeze(-1), frame["label"].unsqueeze(-1)), -1)
    H0, H1, W0, W1 = 20, H - 20, 20, W - 20
    idx = uniform_indices(H0, H1, W0, W1, n_pixels, tape)
    i, j = uv_from_flat(idx, H0, W0, W1 - W0)
    rays_o, rays_d = rays_from_uv(i, j, R, T, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    smp = gather_window(img, H0, H1, W0, W1, idx)
    gc, gd, gl = smp[:, :3], smp[:, 3], smp[:, 4]
    far_bb, inside = far_plane(rays_o, rays_d, bound, gd)
    z = sample_along_rays(gd, n_samples, n_surface, far_bb, tape)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
    code, uv_i, fm_mask = feature_matching(H, W, cam["K"], pts.flatten(0, 1), refer_w2c.detach(),
                                           feats, decoder.merge)
    code = code.reshape(pts.shape[0], pts.shape[1], -1) * trunc_mask(z, gd)[..., None]
    mask = (gd > 0.01) * inside
    return {"gt_color": gc.float(), "gt_depth": gd.float(), "gt_label": gl.to(torch.int64),
            "rays_o": rays_o.float(), "rays_d": rays_d.float(), "pts": pts.float(),
            "z_vals": z.float(), "mask": mask, "features": code,
            "_idx": idx, "_uv": uv_i, "_fm_mask": fm_mask}


def mapper_get_target_samples(cam, bound, decoder, frames, quad_list, T_list, refer_idx,
                              target_idx, refer_c2w_fixed, feats, n_pixels_total, n_samples,
                              n_surface, tape):
    """frames: list of dict(color,depth,label); refer_idx[i]: list of keyframe ids (-1 = self);
    refer_c2w_fixed[i][k]: pose used when the id is not a target frame; feats[i]: [R,64,h,w]."""
    H, W = cam["H"], cam["W"]
    n_t = len(frames)
    n_pixels = n_pixels_total // n_t
    acc = {k: [] for k in ("gt_color", "gt_depth", "gt_label", "rays_o", "rays_d", "pts",
                           "z_vals", "mask", "features", "_idx")}
    for f in range(n_t):
        fr = frames[f]
        R = get_rotation_from_quad(quad_list[f])
        T = T_list[f]
        cur_c2w = c2w_from_quad_T(quad_list[f], T)
        img = torch.cat((fr["color"], fr["depth"].unsqueeze(-1), fr["label"].unsqueeze(-1)), -1)
        idx1 = uniform_indices(0, H, 0, W, n_pixels // 3 * 2, tape)
        idx2 = class_balanced_indices(img[:, :, -1], n_pixels // 3, tape)
        idx = torch.cat((idx1, idx2), 0)
        i, j = uv_from_flat(idx, 0, 0, W)
        rays_o, rays_d = rays_from_uv(i, j, R, T, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        smp = gather_window(img, 0, H, 0, W, idx)
        gc, gd, gl = smp[:, :3], smp[:, 3], smp[:, 4]
        far_bb, inside = far_plane(rays_o, rays_d, bound, gd)
        z = sample_along_rays(gd, n_samples, n_surface, far_bb, tape)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
        w2c = []
        for k, rid in enumerate(refer_idx[f]):
            if rid == -1:
                c2w = cur_c2w.detach()
            elif rid in target_idx:
                t = target_idx.index(rid)
                c2w = c2w_from_quad_T(quad_list[t], T_list[t]).detach()
            else:
                c2w = refer_c2w_fixed[f][k].detach()
            w2c.append(torch.inverse(c2w))
        code, _, _ = feature_matching(H, W, cam["K"], pts.flatten(0, 1), torch.stack(w2c, 0),
                                      feats[f], decoder.merge)
        code = code.reshape(pts.shape[0], pts.shape[1], -1) * trunc_mask(z, gd)[..., None]
        for k, v in (("gt_color", gc.float()), ("gt_depth", gd.float()),
                     ("gt_label", gl.to(torch.int64)), ("rays_o", rays_o.float()),
                     ("rays_d", rays_d.float()), ("pts", pts.float()), ("z_vals", z.float()),
                     ("mask", inside), ("features", code), ("_idx", idx)):
            acc[k].append(v)
    cat = {k: torch.cat(v, 0) for k, v in acc.items()}
    m = cat.pop("mask")
    out = {k: v[m] for k, v in cat.items()}
    out["_inside"] = m
    return out


# --------------------------------------------------------------------------------------
# inference path: full-frame render and point queries
# --------------------------------------------------------------------------------------
def eval_points(decoder, experts, bound, pts, pixel_pts, gt_label_pts=None, stage="fine"):
    """slams/meshing.py:461-498: occupancy / colour / label of free points.  Returns
    (values [P,4] = rgb | occupancy with -100 outside the bound, labels [P] int64 with -1 outside, or None)."""
    mask = ((pts[:, 0] < bound[0, 1]) & (pts[:, 0] > bound[0, 0]) & (pts[:, 1] < bound[1, 1]) & (pts[:, 1] > bound[1, 0])
            & (pts[:, 2] < bound[2, 1]) & (pts[:, 2] > bound[2, 0]))
    p = (pts.clone() - bound[:, 0]) / (bound[:, 1] - bound[:, 0])
    pe, grid = decoder.pe_fn(p)
    if stage == "coarse":
        lat = decoder.coarse_fn(pe, features=grid)
    else:
        lat = fine_fn(experts, decoder.hidden_dim, pe, gt_label_pts, grid)
    color, logits = decoder.out_fn(pe, torch.cat((lat[:, 1:], pixel_pts), -1))
    values = torch.cat((color, lat[:, 0:1]), -1)
    values[~mask, 3] = -100
    if stage == "coarse":
        return values, None
    labels = torch.argmax(logits, -1)
    labels[~mask] = -1
    return values, labels


def get_2d_feature(cam, decoder, points, keyframes, hidden_dim=32):
    """slams/meshing.py:294-377 (Mesher.get_2d_feature): pixel features and labels of free points from the key frames
    that see them.  keyframes: list of dict(est_c2w [4,4], gt_label [H,W], gt_depth [H,W], features [1,64,h,w] = the
    encoder output of the key frame's colour image).  Per key frame: project (x flipped, z < 0 in front), mask to the
    image, round + clamp, truncation mask against the key frame's depth, Merge over that ONE view; the codes are averaged
    over the key frames whose truncation mask holds, the label is the one of the LAST key frame that sees the point."""
    H, W = cam["H"], cam["W"]
    K = cam["K"].float()
    P = points.shape[0]
    pixel_pts = torch.zeros(P, hidden_dim)
    label_pts = torch.zeros(P)
    count_pts = torch.zeros(P)
    for kf in keyframes:
        c2w = kf["est_c2w"]
        w2c = torch.inverse(c2w).float()
        homo = torch.cat([points, torch.ones_like(points[:, :1])], dim=1).reshape(-1, 4, 1).float()
        cam_cord = (w2c @ homo)[:, :3]
        cam_cord[:, 0] *= -1
        uv = K @ cam_cord.float()
        z = uv[:, -1:] + 1e-8
        uv = (uv[:, :2] / z).float()
        seen = (uv[:, 0] < W) & (uv[:, 0] > 0) & (uv[:, 1] < H) & (uv[:, 1] > 0)
        seen = (seen & (z[:, :, 0] < 0)).reshape(-1)
        uv_ = uv[seen, :, 0]
        p = points[seen, :]
        if uv_.numel() == 0:
            continue
        uv_ = torch.round(uv_).to(torch.int64)
        uv_[:, 0] = uv_[:, 0].clamp(0, W - 1)
        uv_[:, 1] = uv_[:, 1].clamp(0, H - 1)
        label_seen = kf["gt_label"][uv_[:, 1], uv_[:, 0]]
        depth_seen = kf["gt_depth"][uv_[:, 1], uv_[:, 0]]
        depth_proj = -z[seen].squeeze()
        front = torch.where(depth_proj < depth_seen * 0.95, torch.ones_like(depth_seen), torch.zeros_like(depth_seen))
        back = torch.where(depth_proj > depth_seen * 1.05, torch.ones_like(depth_seen), torch.zeros_like(depth_seen))
        trunc = (1.0 - front) * (1.0 - back)
        feats = F.interpolate(kf["features"], size=[H, W], mode="bilinear", align_corners=True)
        ft = feats[0, :, uv_[:, 1], uv_[:, 0]].permute(1, 0).unsqueeze(0)
        refer_o = c2w[:3, 3].unsqueeze(0)
        refer_p = p[None, :, :].clone() - refer_o[:, None, :]
        code = decoder.merge(refer_p, refer_o, ft) * trunc[..., None]
        count_pts[seen] += trunc
        pixel_pts[seen, :] += code.float()
        label_pts[seen] = label_seen.float()
    ok = count_pts > 0
    pixel_pts[ok, :] = pixel_pts[ok, :] / count_pts[ok, None]
    return pixel_pts, label_pts


def frame_vis_render(cam, bound, decoder, experts, frame, c2w, refer_w2c, feats, n_samples, n_surface, tape,
                     n_pts_batch):
    """slams/mapping.py:636-690 (frame_vis without the plotting): every pixel of the frame, rays from ``c2w``,
    depth-guided samples, ONE reference view for the pixel features, the mapper renderer per chunk of
    ``n_pts_batch`` rays (the class rule class(p) = label[p mod n] therefore runs per chunk).
    refer_w2c [1,4,4], feats [1,64,h,w].  Returns (color [H,W,3], depth [H,W], label [H,W] int64)."""
    H, W = cam["H"], cam["W"]
    idx = torch.arange(H * W)
    i, j = uv_from_flat(idx, 0, 0, W)
    rays_o, rays_d = rays_from_uv(i, j, c2w[:3, :3], c2w[:3, 3], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    gd = frame["depth"].flatten(0, 1)
    gl = frame["label"].flatten(0, 1)
    t = (bound.unsqueeze(0) - rays_o.detach().unsqueeze(-1)) / rays_d.detach().unsqueeze(-1)
    far_bb = torch.min(torch.max(t, dim=2)[0], dim=1)[0].unsqueeze(-1) + 0.01
    z = sample_along_rays(gd, n_samples, n_surface, far_bb, tape)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]
    cols, deps, labs = [], [], []
    for a in range(0, H * W, n_pts_batch):
        b = min(a + n_pts_batch, H * W)
        code, _, _ = feature_matching(H, W, cam["K"], pts[a:b].flatten(0, 1), refer_w2c, feats, decoder.merge)
        smp = {"rays_o": rays_o[a:b].float(), "rays_d": rays_d[a:b].float(), "gt_label": gl[a:b].to(torch.int64),
               "pts": pts[a:b].float(), "z_vals": z[a:b].float(), "features": code.reshape(b - a, z.shape[1], -1)}
        rgb, depth, _, logits, _, _ = mapper_renderer(decoder, experts, bound, smp)
        cols.append(rgb)
        deps.append(depth)
        labs.append(torch.argmax(logits, -1))
    return torch.cat(cols, 0).reshape(H, W, 3), torch.cat(deps, 0).reshape(H, W), torch.cat(labs, 0).reshape(H, W)


# --------------------------------------------------------------------------------------
# key-frame selection by overlap  (slams/mapping.py:171-236), SURVEY 8 f2
# --------------------------------------------------------------------------------------
def keyframe_overlap(cam, gt_depth, c2w, keyframe_c2w, tape, n_samples=16, pixels=100):
    """percent_inside per key frame, as the reference computes it with numpy: ``pixels`` random pixels of the current
    frame, 16 depths in [0.8 d, d + 0.5] each, projected into every key frame (x flipped, z must be negative, 10 px
    border).  Returns a float64 numpy array [n_keyframes]."""
    import numpy as np
    H, W = cam["H"], cam["W"]
    fx, fy, cx, cy = cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    idx = uniform_indices(0, H, 0, W, pixels, tape)
    i, j = uv_from_flat(idx, 0, 0, W)
    rays_o, rays_d = rays_from_uv(i, j, c2w[:3, :3], c2w[:3, -1], fx, fy, cx, cy)
    d = gt_depth.reshape(-1)[idx].reshape(-1, 1).repeat(1, n_samples)
    t_vals = torch.linspace(0.0, 1.0, steps=n_samples)
    z_vals = d * 0.8 * (1.0 - t_vals) + (d + 0.5) * t_vals
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    vertices = pts.reshape(-1, 3).cpu().numpy()
    out = []
    for kf in keyframe_c2w:
        w2c = np.linalg.inv(kf.cpu().numpy())
        ones = np.ones_like(vertices[:, 0]).reshape(-1, 1)
        homo = np.concatenate([vertices, ones], axis=1).reshape(-1, 4, 1)
        cam_cord = (w2c @ homo)[:, :3]
        K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]]).reshape(3, 3)
        cam_cord[:, 0] *= -1
        uv = K @ cam_cord
        z = uv[:, -1:] + 1e-5
        uv = (uv[:, :2] / z).astype(np.float32)
        edge = 10
        mask = (uv[:, 0] < W - edge) * (uv[:, 0] > edge) * (uv[:, 1] < H - edge) * (uv[:, 1] > edge)
        mask = (mask & (z[:, :, 0] < 0)).reshape(-1)
        out.append(mask.sum() / uv.shape[0])
    return np.asarray(out, dtype=np.float64)


# --------------------------------------------------------------------------------------
# ResNet stem of the pixel-feature branch
# --------------------------------------------------------------------------------------
def stem_forward(images, conv_w, bn_weight, bn_bias, running_mean=None, running_var=None, training=True,
                 momentum=0.1, eps=1e-5):
    """models/encoder.py:9-17 over models/layers.py:97-100: images [B,N,H,W,3] -> [B,N,64,h,w].
    conv1 (7x7, stride 2, pad 3, no bias) -> bn1 -> ReLU.  The reference leaves the encoder in training mode
    (slams/tracking.py:30, slams/mapping.py:33: no .eval()), so bn1 uses the statistics of this batch and updates
    ``running_mean`` / ``running_var`` IN PLACE (torch semantics: momentum, unbiased variance)."""
    B, N = images.shape[:2]
    x = images.flatten(0, 1).permute(0, 3, 1, 2)
    x = F.conv2d(x, conv_w, None, stride=2, padding=3)
    x = F.batch_norm(x, running_mean, running_var, bn_weight, bn_bias, training, momentum, eps)
    x = F.relu(x)
    _, Cc, h, w = x.shape
    return x.reshape(B, N, Cc, h, w)
