import torch, sys, time
sys.path.insert(0, '.')
from dns_slam_b200 import bench_util, slam, synthetic as syn
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
shape = "replica"; s = syn.SHAPES[shape]
dec = bench_util.make_decoder(shape, 40, dev, seed=1)
sc = bench_util.slam_scene(shape, 40, dev, seed=2)
cam = sc["cam"]
trk = slam.TrackerCore(cam, dec, s["tracking_pixels"], 32, 15, 5.0, 5.0, 0.1)
td = bench_util.tracking_draws(cam, s["tracking_pixels"], 10)
est = sc["poses"][3].clone(); est[:3, 3] += 0.01
refer_w2c = torch.inverse(sc["poses"][2]); feats2 = sc["feats"][1][:2].contiguous()
def track(): slam.track_frame(trk, sc["frames"][1], refer_w2c, feats2, est, 10, 1e-3, lambda it: td[it])
track(); torch.cuda.synchronize()
t0 = time.perf_counter(); track(); torch.cuda.synchronize(); print("track 10 it wall ms", (time.perf_counter()-t0)*1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    track(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=50))
