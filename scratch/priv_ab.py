"""A/B of the privatised gradient copies (ablate build: DNS_SLAM_B200_LIB=dns_slam_b200/libdns_slam_b200_ablate.so).
Core step only (point_fwd -> ray -> point_bwd -> dW) at the default size, phase timers per setting."""
import os, sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
R, S, C = int(os.environ.get("AB_RAYS", 131072)), 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
def run(tag, env):
    for k in ("DNS_NO_PRIV", "DNS_PRIV_COPIES", "DNS_PRIV_LEVELS", "DNS_DBG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(2): ms.step(samples)
    torch.cuda.synchronize()
    _lib.profile_read(True); _lib.profile_enable(True)
    n = 5
    for _ in range(n): out = ms.step(samples)
    torch.cuda.synchronize(); _lib.profile_enable(False)
    ph, _ = _lib.profile_read(True)
    print(f"{tag:28s}", {k: round(v / n, 3) for k, v in ph.items() if k in ("point_fwd", "ray", "point_bwd", "dw_gemm")}, flush=True)
    return ms.grad.clone()
g0 = None
for tag, env in [("no priv", {"DNS_NO_PRIV": "1"}), ("priv 8 copies x 4 levels", {}), ("priv 4 copies", {"DNS_PRIV_COPIES": "4"}),
                 ("priv 2 copies", {"DNS_PRIV_COPIES": "2"}), ("priv 8 x 1 level", {"DNS_PRIV_LEVELS": "1"}),
                 ("priv 8 x 2 levels", {"DNS_PRIV_LEVELS": "2"}), ("priv 8 x 3 levels", {"DNS_PRIV_LEVELS": "3"}),
                 ("no atomics (floor)", {"DNS_DBG": "4"}), ("no priv again", {"DNS_NO_PRIV": "1"})]:
    run(tag, env)
