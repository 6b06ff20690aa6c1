"""TV smoothness term (fwd + bwd + dW) at the Replica (63^3) and ScanNet (127^3) lattices, with the per-cell
warp pre-reduction of the backward scatter up to different level limits (DNS_TV_AGG, ablate build)."""
import os, sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, fused, synthetic as syn
dev = torch.device("cuda:0")
for shape in ("replica", "scannet"):
    s = syn.SHAPES[shape]
    dec = bench_util.make_decoder(shape, 40, dev, seed=0)
    g = torch.Generator().manual_seed(1)
    off, jit = fused.tv_offsets(dec.bound, s["smooth_pts"], torch.rand(3, generator=g), torch.rand(1, 1, 1, 3, generator=g))
    d_t, d_c = torch.zeros_like(dec.view("table")), torch.zeros_like(dec.view("coarse"))
    def tv():
        return fused.tv_raw(dec.pe_fn.grid_fn.gstruct, dec.bound, dec.view("table"), dec.view("coarse"), s["smooth_pts"], off, jit,
                            s["lambda_smooth"], d_t, d_c)
    ref = None
    for tag, env in (("agg 0", {"DNS_TV_AGG": "0"}), ("agg 0.6", {"DNS_TV_AGG": "0.6"}), ("agg 1.0", {}), ("agg 1.5", {"DNS_TV_AGG": "1.5"}),
                     ("agg all", {"DNS_TV_AGG": "1e9"}), ("agg 0", {"DNS_TV_AGG": "0"}), ("agg 1.0", {}), ("agg 1 nopriv", {"DNS_NO_PRIV": "1"})):
        os.environ.pop("DNS_NO_PRIV", None); os.environ.pop("DNS_TV_AGG", None); os.environ.update(env)
        for _ in range(3): tv()
        d_t.zero_(); d_c.zero_(); tv(); gt = d_t.clone()
        if ref is None: ref = gt
        err = float((gt - ref).norm() / ref.norm())
        torch.cuda.synchronize()
        _lib.profile_read(True); _lib.profile_enable(True)
        for _ in range(10): tv()
        torch.cuda.synchronize(); _lib.profile_enable(False)
        ph, _ = _lib.profile_read(True)
        print(shape, f"{tag:12s}", {k: round(v / 10, 3) for k, v in ph.items() if v}, "grad diff vs first", f"{err:.1e}", flush=True)
