import sys, torch
sys.path.insert(0, '.')
import bench
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
dec = bench_util.make_decoder("replica", 40, dev, seed=0)
for (R, S) in ((1024, 96), (500, 47), (4096, 47)):
    mode = "track" if S == 96 or R == 500 else "map"
    dec1, smp1 = bench_util.synthetic_batch("replica", mode, R, S, 40, dev, seed=9, dec=dec)
    if mode == "track":
        ts = stepmod.TrackingStep(dec1, dict(p=5.0, d=5.0, l=0.1)); fn = lambda: ts.forward_backward(smp1)
    else:
        ms = stepmod.MappingStep(dec1, 5e-3); fn = lambda: ms.step(smp1)
    t = bench._time_cuda(fn, 50, 10)
    _lib.profile_read(True); _lib.profile_enable(True)
    for _ in range(20): fn()
    torch.cuda.synchronize(); _lib.profile_enable(False)
    ph, ln = _lib.profile_read(True)
    print(mode, R, S, "ms:", round(t, 4), {k: round(v / 20, 4) for k, v in ph.items() if v}, {k: v // 20 for k, v in ln.items() if v}, flush=True)
