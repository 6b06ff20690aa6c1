"""A/B of the Jacobian image (ablate build: DNS_NO_JIMG=1 re-reads the corners) + gradient agreement of the two paths."""
import os, sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod, fused
dev = torch.device("cuda:0")
R, S, C = 131072, 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
def rays_grad():
    cfg = ms._config(samples); p = ms._views(dec.flat)
    ms.grad.zero_()
    out = fused.render_raw(cfg, p["table"], p["coarse"], p["color"], p["logit"], p["experts"], samples["rays_o"], samples["rays_d"],
                           samples.get("features"), ms._views(ms.grad), True, True)
    return out[2].clone(), out[3].clone(), ms.grad.clone()
os.environ["DNS_NO_JIMG"] = "1"; g_ref = rays_grad()
os.environ.pop("DNS_NO_JIMG"); g_j = rays_grad()
rowerr = ((g_j[1] - g_ref[1]).norm(dim=1) / g_ref[1].norm(dim=1).clamp_min(1e-20))
print("same weights: d_rays_o / d_rays_d / params, jacobian image vs re-read (norm-wise):",
      ["%.1e" % float((a - b).norm() / b.norm()) for a, b in zip(g_j, g_ref)], "ray rows: worst %.1e median %.1e" % (float(rowerr.max()), float(rowerr.median())), flush=True)
ref = None
for tag, env in (("re-read corners", {"DNS_NO_JIMG": "1"}), ("jacobian image", {}), ("re-read corners", {"DNS_NO_JIMG": "1"}), ("jacobian image", {})):
    os.environ.pop("DNS_NO_JIMG", None); os.environ.update(env)
    g = rays_grad()
    if ref is None: ref = g
    errs = [float((a - b).norm() / b.norm()) for a, b in zip(g, ref)]
    rowerr = ((g[1] - ref[1]).norm(dim=1) / ref[1].norm(dim=1).clamp_min(1e-20))
    for _ in range(2): ms.step(samples)
    torch.cuda.synchronize()
    _lib.profile_read(True); _lib.profile_enable(True)
    n = 6
    for _ in range(n): ms.step(samples)
    torch.cuda.synchronize(); _lib.profile_enable(False)
    ph, _ = _lib.profile_read(True)
    print(f"{tag:16s}", {k: round(v / n, 3) for k, v in ph.items() if k in ("point_fwd", "ray", "point_bwd", "dw_gemm")},
          "d_rays_o/d_rays_d/params vs first:", ["%.1e" % e for e in errs], "worst ray row %.1e median %.1e" % (float(rowerr.max()), float(rowerr.median())), flush=True)
