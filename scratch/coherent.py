"""Upper bound of a locality-sorted slot order: all rays in ONE class and the generic (order-preserving) class sort make
the slots follow the ray / sample order, i.e. warps hold consecutive samples of a ray."""
import torch, sys, os
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
R, S, C = 131072, 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
if len(sys.argv) > 1 and sys.argv[1] == "oneclass":
    samples["gt_label"] = torch.zeros_like(samples["gt_label"])
ms = stepmod.MappingStep(dec, 5e-3)
for _ in range(3): ms.step(samples)
torch.cuda.synchronize()
_lib.profile_read(True); _lib.profile_enable(True)
for _ in range(5): ms.step(samples)
torch.cuda.synchronize(); _lib.profile_enable(False)
phase, ln = _lib.profile_read(True)
print(sys.argv[1:], os.environ.get("DNS_GENERIC_PREP"), {k: round(v/5,3) for k,v in phase.items() if v})
