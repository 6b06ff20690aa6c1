import torch, sys
sys.path.insert(0, '.')
from dns_slam_b200 import bench_util, slam, synthetic as syn
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
shape = "replica"; s = syn.SHAPES[shape]
dec = bench_util.make_decoder(shape, 40, dev, seed=1)
sc = bench_util.slam_scene(shape, 40, dev, seed=2)
cam = sc["cam"]
trk = slam.TrackerCore(cam, dec, s["tracking_pixels"], 32, 15, 5.0, 5.0, 0.1, freeze_decoder=True)
td = bench_util.tracking_draws(cam, s["tracking_pixels"], 10)
est = sc["poses"][3].clone(); est[:3, 3] += 0.01
refer_w2c = torch.inverse(sc["poses"][2]); feats2 = sc["feats"][1][:2].contiguous()
def run(n): slam.track_frame(trk, sc["frames"][1], refer_w2c, feats2, est, n, 1e-3, lambda it: td[it % 10], use_graph=True)
run(10); torch.cuda.synchronize()
N = 53
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(N); torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev)
print("total device us per iteration ~", tot / N, "kernels/it", sum(e.count for e in ev) / N)
for e in sorted(ev, key=lambda e: -e.device_time_total)[:28]:
    print(f"{e.device_time_total/N:8.1f} us/it x{e.count/N:5.1f}  {e.key[:100]}")
