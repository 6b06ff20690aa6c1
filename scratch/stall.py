"""Per-step CUDA-event times of the default mapping step: looks for one-off stalls inside a timed loop.
usage: python scratch/stall.py [rays] [profile 0/1]"""
import sys, time, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
prof = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dec = bench_util.make_decoder("replica", 40, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, 47, 40, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
for _ in range(3): ms.step(samples)
torch.cuda.synchronize()
if prof:
    _lib.profile_read(True); _lib.profile_enable(True)
n = 20
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host = []
ev[0].record()
for i in range(n):
    t0 = time.perf_counter()
    ms.step(samples)
    host.append((time.perf_counter() - t0) * 1e3)
    ev[i + 1].record()
torch.cuda.synchronize()
print("rays", R, "profile", prof)
print("gpu ms :", " ".join(f"{ev[i].elapsed_time(ev[i+1]):.1f}" for i in range(n)))
print("host ms:", " ".join(f"{h:.1f}" for h in host))
