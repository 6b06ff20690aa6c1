"""Synthetic Replica / ScanNet-shaped ray batches for the benchmarks, the smoke test and the
full-size property tests (no dataset, no checkpoint: seeded synthetic frames and random-init
weights of the reference architecture)."""
import torch

from . import decoder as _decoder
from . import fused, slam
from . import synthetic as syn


def split_samples(S):
    """n_samples_ray / n_surface_ray for a total of S samples (47 -> 32+15 as replica.yaml:28-29,
    96 -> 64+32 for BASELINE config 2)."""
    if S == 47:
        return 32, 15
    n_surface = max(S // 3, 1)
    return S - n_surface, n_surface


def make_decoder(shape, n_class, device, seed=0, table_scale=1000.0, all_experts=True):
    bound = syn.load_bound(syn.SHAPES[shape]["bound"])
    dec = _decoder.Decoder(syn.model_cfg(shape), bound, n_class=n_class, seed=seed, device=device)
    with torch.no_grad():
        dec.pe_fn.grid_fn.params.mul_(table_scale)
    if all_experts:
        for c in range(n_class):
            dec.activate_expert(c)
    return dec


def synthetic_batch(shape, mode, N, S, C, device, seed=0, n_frames=None, dec=None):
    """Returns (decoder, samples) with exactly N rays drawn from ``n_frames`` synthetic frames
    (mapping: 4 target frames as replica.yaml:41; tracking: 1).  ``samples`` holds the fields of
    tracking.py:177-185 / mapping.py:579-586 as contiguous CUDA tensors."""
    cam = syn.camera(shape)
    if dec is None:
        dec = make_decoder(shape, C, device, seed)
    n_frames = n_frames or (1 if mode == "track" else 4)
    n_s, n_f = split_samples(S)
    gen = torch.Generator().manual_seed(seed)
    poses = syn.trajectory(shape, max(n_frames, 2) * 2)
    H, W = cam["H"], cam["W"]
    window = (20, H - 20, 20, W - 20) if mode == "track" else (0, H, 0, W)
    per = [N // n_frames + (1 if f < N % n_frames else 0) for f in range(n_frames)]
    parts = []
    for f in range(n_frames):
        if per[f] == 0:
            continue
        c2w = poses[f * 2 + 1]
        fr = syn.frame(shape, c2w, gen, n_class=C)
        fr = {k: v.to(device).contiguous() for k, v in fr.items()}
        n_win = (window[1] - window[0]) * (window[3] - window[2])
        idx = torch.randint(n_win, (per[f],), generator=gen).to(device)
        s = fused.sample_rays(cam, dec.bound, fr, idx, window, c2w[:3, :3].to(device), c2w[:3, 3].to(device), n_s, n_f,
                              fused.fix_surface_draw(torch.rand(n_f, generator=gen), n_f), torch.rand(n_f, generator=gen))
        parts.append(s)
    cat = {k: torch.cat([p[k] for p in parts], 0).contiguous() for k in parts[0]}
    feats = torch.randn(N, S, 32, generator=gen).to(device) * 0.3
    feats = (feats * slam.trunc_mask(cat["z_vals"], cat["gt_depth"])[..., None]).contiguous()
    samples = dict(gt_color=cat["gt_color"], gt_depth=cat["gt_depth"], gt_label=cat["gt_label"],
                   rays_o=cat["rays_o"], rays_d=cat["rays_d"], z_vals=cat["z_vals"], features=feats,
                   mask=(cat["gt_depth"] > 0.01) * cat["inside"])
    return dec, samples


def slam_scene(shape, n_class, device, seed=0, n_target=4, n_refer=3):
    """Synthetic keyframes for whole-iteration timings: target frames with class tables, reference
    poses and channels-last pixel features (one [R,h,w,64] block per target frame)."""
    cam = syn.camera(shape)
    gen = torch.Generator().manual_seed(seed)
    poses = syn.trajectory(shape, n_target * 2 + 2)
    frames, feats, refer_c2w, refer_idx = [], [], [], []
    for f in range(n_target):
        fr = syn.frame(shape, poses[2 * f + 1], gen, n_class=n_class)
        frames.append({k: v.to(device).contiguous() for k, v in fr.items()})
        feats.append(fused.channels_last(syn.pixel_features(shape, n_refer, gen).to(device)))
        refer_idx.append([100 + 2 * f, 101 + 2 * f, -1][:n_refer])          # ids that are not target frames
        refer_c2w.append([poses[2 * f].to(device), poses[2 * f + 2].to(device), poses[2 * f + 1].to(device)][:n_refer])
    tables = [slam.class_tables(fr["label"], n_ids=n_class if fr["label"].is_cuda else None) for fr in frames]
    return dict(cam=cam, poses=poses, frames=frames, feats=feats, refer_idx=refer_idx, refer_c2w=refer_c2w,
                class_tables=tables, kf_idx=list(range(n_target)))


def tracking_draws(cam, n_pixels, n_iters, seed=0, n_surface=15):
    g = torch.Generator().manual_seed(seed)
    n_win = (cam["H"] - 40) * (cam["W"] - 40)
    return [dict(idx=torch.randint(n_win, (n_pixels,), generator=g), t_surface=torch.rand(n_surface, generator=g),
                 t_zero=torch.rand(n_surface, generator=g)) for _ in range(n_iters)]


def mapping_draws(scene, n_pixels, n_iters, seed=0, n_surface=15):
    """Draw tapes of mapping iterations in the reference's order, pre-generated on the host."""
    g = torch.Generator().manual_seed(seed)
    cam = scene["cam"]
    n_t = len(scene["frames"])
    npf = n_pixels // n_t
    out, tv = [], []
    for _ in range(n_iters):
        per = []
        for tab in scene["class_tables"]:
            counts = tab[3].tolist()
            n_c = len(counts)
            n_k = (npf // 3) // n_c
            cd = []
            for c in range(n_c):
                m = (npf // 3) - n_k * (n_c - 1) if c == 0 else n_k
                if counts[c] != 1:
                    cd.append(torch.randint(counts[c], (m,), generator=g))
            per.append(dict(idx_uniform=torch.randint(cam["H"] * cam["W"], (npf // 3 * 2,), generator=g), class_draws=cd,
                            t_surface=torch.rand(n_surface, generator=g), t_zero=torch.rand(n_surface, generator=g)))
        out.append(per)
        tv.append((torch.rand(3, generator=g), torch.rand(1, 1, 1, 3, generator=g)))
    return out, tv
