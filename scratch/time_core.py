"""Phase timers of the core step (point_fwd -> ray -> point_bwd -> dW) at the default size, for A/B of library variants
(DNS_SLAM_B200_LIB selects the build)."""
import os, sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
R, S, C = int(os.environ.get("AB_RAYS", 131072)), 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
for rep in range(2):
    for _ in range(3): out = ms.step(samples)
    torch.cuda.synchronize()
    _lib.profile_read(True); _lib.profile_enable(True)
    n = 8
    for _ in range(n): out = ms.step(samples)
    torch.cuda.synchronize(); _lib.profile_enable(False)
    ph, _ = _lib.profile_read(True)
    print(os.path.basename(_lib.LIB_PATH), {k: round(v / n, 3) for k, v in ph.items() if k in ("point_fwd", "ray", "point_bwd", "dw_gemm")},
          "loss", [round(float(x), 6) for x in out[0][:7]] if isinstance(out, (tuple, list)) else "", flush=True)
