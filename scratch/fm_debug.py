import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_featmerge as T
from gpu_util import rel_err
from dns_slam_b200 import fused, slam
for rep in range(3):
    c = T._case("replica", 2, 3, 1500, seed=21, n_class=40)
    dev, dec, cat = c["dev"], c["dec"], c["cat"]
    w2c = torch.cat(c["w2c"], 0).to(dev)
    c2w = torch.inverse(torch.cat(c["w2c"], 0)).to(dev)
    feats = [fused.channels_last(f.to(dev)) for f in c["feats"]]
    views = fused.Views(w2c, c2w[:, :3, 3].contiguous(), feats, c["ray_start"])
    mp = dec.merge.decoder.params.detach()
    f1, ws = fused.featmerge_raw(c["cam"], dec.merge.bound, views, cat["rays_o"], cat["rays_d"], cat["z_vals"], cat["gt_depth"], mp)
    f1b, _ = fused.featmerge_raw(c["cam"], dec.merge.bound, views, cat["rays_o"], cat["rays_d"], cat["z_vals"], cat["gt_depth"], mp)
    parts = []
    for f in range(2):
        r0, r1 = c["ray_start"][f], c["ray_start"][f + 1]
        pts = cat["rays_o"][r0:r1, None, :] + cat["rays_d"][r0:r1, None, :] * cat["z_vals"][r0:r1, :, None]
        code = fused.feature_matching(c["cam"]["H"], c["cam"]["W"], c["cam"]["K"].to(dev), pts.flatten(0, 1), w2c[3*f:3*f+3], feats[f], dec.merge, refer_c2w=c2w[3*f:3*f+3])
        parts.append(code.reshape(r1 - r0, -1, 32) * slam.trunc_mask(cat["z_vals"][r0:r1], cat["gt_depth"][r0:r1])[..., None])
    f0 = torch.cat(parts, 0).detach()
    n_band = int(ws[:4].view(torch.int32)[0])
    nz1, nz0 = (f1 != 0).any(-1), (f0 != 0).any(-1)
    rows = (f1 - f0).flatten(0, 1).norm(dim=-1)
    print(rep, "band", n_band, "nz1", int(nz1.sum()), "nz0", int(nz0.sum()), "rel", rel_err(f1, f0), "twice", rel_err(f1, f1b),
          "bad rows", int((rows > 1e-4 * f0.abs().max()).sum()), "nan", bool(torch.isnan(f1).any()), bool(torch.isnan(f0).any()),
          "max f0", float(f0.abs().max()), "max f1", float(f1.abs().max()))
    bad = torch.nonzero(rows > 1e-4 * f0.abs().max()).reshape(-1)[:5].tolist()
    for i in bad:
        print("   row", i, "ray", i // 15, "s", i % 15, f1.flatten(0,1)[i][:4].tolist(), f0.flatten(0,1)[i][:4].tolist())
