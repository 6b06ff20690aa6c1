"""Ablation series of the point kernels at the default size (ablate build): DNS_DBG bits 2 = no stash stores, 4 = no table
reductions, 8 = no corner re-read."""
import os, sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
R, S, C = 131072, 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
for dbg in (0, 2, 4, 8, 6, 12, 14, 0):
    os.environ["DNS_DBG"] = str(dbg)
    for _ in range(2): ms.step(samples)
    torch.cuda.synchronize()
    _lib.profile_read(True); _lib.profile_enable(True)
    for _ in range(5): ms.step(samples)
    torch.cuda.synchronize(); _lib.profile_enable(False)
    ph, _ = _lib.profile_read(True)
    print(f"DBG={dbg:2d}", {k: round(v / 5, 3) for k, v in ph.items() if k in ("point_fwd", "ray", "point_bwd", "dw_gemm")}, flush=True)
