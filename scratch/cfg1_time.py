import sys, torch
sys.path.insert(0, '.')
import bench
from dns_slam_b200 import bench_util, step as stepmod
dev = torch.device("cuda:0")
dec = bench_util.make_decoder("replica", 40, dev, seed=0)
dec1, smp1 = bench_util.synthetic_batch("replica", "track", 1024, 96, 40, dev, seed=9, dec=dec)
ts = stepmod.TrackingStep(dec1, dict(p=5.0, d=5.0, l=0.1))
for _ in range(3):
    print("config1 tracking 1024x96 ms:", bench._time_cuda(lambda: ts.forward_backward(smp1), 50, 10), flush=True)
