"""Seeded synthetic Replica / ScanNet-shaped RGB-D + semantic inputs (no dataset on disk).

Tensor contract follows the reference readers (``datas/slam_datasets.py:64-149,153-228``):
``color [H,W,3] f32 in [0,1)``, ``depth [H,W] f32`` (metres along the camera -z axis, a few
zeros for missing returns), ``label [H,W] int64`` (class ids, block constant; np.vectorize over a dict gives int64), ``c2w [4,4] f32``.
Intrinsics and bounds are the reference's (``configs/replica/replica.yaml:3-12``,
``configs/replica/room_0.yaml:3-4``, ``configs/scannet/scannet.yaml:3-11``,
``configs/scannet/scene0000.yaml:3-4``); the bound enlargement is ``slams/dns_slam.py:100-107``.
Everything is generated on the CPU with an explicit ``torch.Generator`` so that the CUDA path,
the oracle and the golden fixtures see bit-identical inputs.
"""
import math

import numpy as np
import torch

SHAPES = {
    # name: H, W, fx, fy, cx, cy, bound, hash_size, voxel_size, training/tracking/mapping knobs
    "replica": dict(H=680, W=1200, fx=600.0, fy=600.0, cx=599.5, cy=339.5,
                    bound=[[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]], hash_size=16, voxel_size=0.02,
                    lr=0.005, lambda_color=5.0, lambda_depth=5.0, lambda_label=0.1,
                    lambda_smooth=1e-5, lambda_fs=10.0, lambda_opacity=10.0, smooth_pts=64,
                    opacity_sigma=0.05, cam_lr=1e-3, BA_cam_lr=5e-4,
                    tracking_pixels=500, tracking_iters=50, mapping_pixels=2000, mapping_iters=100),
    # 640x480 cropped by 10 px on each edge (slams/dns_slam.py:126-131)
    "scannet": dict(H=460, W=620, fx=577.590698, fy=578.729797, cx=308.905426, cy=232.683609,
                    bound=[[-0.1, 8.6], [-0.1, 8.9], [-0.3, 3.3]], hash_size=20, voxel_size=0.04,
                    lr=0.001, lambda_color=5.0, lambda_depth=1.0, lambda_label=0.1,
                    lambda_smooth=1e-3, lambda_fs=10.0, lambda_opacity=10.0, smooth_pts=128,
                    opacity_sigma=0.1, cam_lr=1e-3, BA_cam_lr=5e-4,
                    tracking_pixels=1000, tracking_iters=30, mapping_pixels=2000, mapping_iters=100),
    # tiny shape used by golden fixtures and CPU tests
    "tiny": dict(H=60, W=80, fx=60.0, fy=60.0, cx=39.5, cy=29.5,
                 bound=[[-1.0, 2.1], [-1.2, 1.7], [-1.1, 1.3]], hash_size=13, voxel_size=0.025,
                 lr=0.005, lambda_color=5.0, lambda_depth=5.0, lambda_label=0.1,
                 lambda_smooth=1e-5, lambda_fs=10.0, lambda_opacity=10.0, smooth_pts=8,
                 opacity_sigma=0.05, cam_lr=1e-3, BA_cam_lr=5e-4,
                 tracking_pixels=24, tracking_iters=3, mapping_pixels=48, mapping_iters=3),
}

MODEL_CFG = {"pts_dim": 3, "pixel_dim": 64, "hidden_dim": 32,
             "pos": {"method": "OneBlob", "n_bins": 16},
             "grid": {"method": "HashGrid", "hash_size": 16, "voxel_size": 0.02}}


def model_cfg(shape):
    s = SHAPES[shape]
    cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in MODEL_CFG.items()}
    cfg["grid"]["hash_size"] = s["hash_size"]
    cfg["grid"]["voxel_size"] = s["voxel_size"]
    return cfg


def load_bound(bound, scale=1.0, bound_divisible=0.32):
    """float64 [3,2]; upper corner enlarged to a multiple of ``bound_divisible``."""
    b = torch.from_numpy(np.array(bound, dtype=np.float64) * scale)
    b[:, 1] = (((b[:, 1] - b[:, 0]) / bound_divisible).int() + 1) * bound_divisible + b[:, 0]
    return b


def camera(shape):
    s = SHAPES[shape]
    K = torch.tensor([[s["fx"], 0, s["cx"]], [0, s["fy"], s["cy"]], [0, 0, 1]])
    return dict(H=s["H"], W=s["W"], fx=s["fx"], fy=s["fy"], cx=s["cx"], cy=s["cy"], K=K)


def _rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def trajectory(shape, n_frames):
    """Smooth camera path inside the room; frame 0 is close to identity rotation."""
    b = np.array(SHAPES[shape]["bound"], dtype=np.float64)
    ctr, ext = b.mean(1), (b[:, 1] - b[:, 0])
    out = []
    for f in range(n_frames):
        t = f / max(n_frames - 1, 1)
        R = _rot(0.15 * math.sin(2 * math.pi * t), 0.6 * t, 0.05 * math.sin(4 * math.pi * t))
        pos = ctr + ext * np.array([0.12 * math.sin(2 * math.pi * t), 0.10 * math.cos(2 * math.pi * t) - 0.10, 0.05 * t])
        c2w = np.eye(4)
        c2w[:3, :3], c2w[:3, 3] = R, pos
        out.append(torch.from_numpy(c2w).float())
    return out


def frame(shape, c2w, gen, n_class=40, zero_frac=0.02):
    """One RGB-D + label frame: depth is the exit distance of the pixel ray from a room box
    0.4 m inside the scene bound, so ``inside`` masks are mostly true."""
    s = SHAPES[shape]
    H, W = s["H"], s["W"]
    b = torch.tensor(s["bound"], dtype=torch.float64)
    shrink = 0.15 * (b[:, 1] - b[:, 0]).min()
    lo, hi = b[:, 0] + shrink, b[:, 1] - shrink
    jj, ii = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    dirs = torch.stack([(ii - s["cx"]) / s["fx"], -(jj - s["cy"]) / s["fy"], -torch.ones_like(ii)], -1)
    R, o = c2w[:3, :3].double(), c2w[:3, 3].double()
    d = dirs @ R.t()
    t = torch.stack(((lo - o) / d, (hi - o) / d), -1)
    depth = t.max(-1)[0].min(-1)[0].clamp(0.3, 20.0)
    depth = depth * (1.0 + 0.05 * torch.sin(ii / 37.0) * torch.cos(jj / 23.0))
    depth = depth.float()
    holes = torch.rand(H, W, generator=gen) < zero_frac
    depth[holes] = 0.0
    color = torch.rand(H, W, 3, generator=gen)
    bh, bw = max(H // 10, 1), max(W // 12, 1)
    label = ((torch.div(jj, bh, rounding_mode="floor") * 7 + torch.div(ii, bw, rounding_mode="floor") * 3) % n_class)
    return dict(color=color, depth=depth, label=label.long(), c2w=c2w)


def pixel_features(shape, n_views, gen, channels=64):
    """Stand-in for the frozen ResNet stem output ``[R,64,H/2,W/2]`` (patch P4)."""
    s = SHAPES[shape]
    return torch.randn(n_views, channels, s["H"] // 2, s["W"] // 2, generator=gen)
