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
def run(n): slam.track_frame(trk, sc["frames"][1], refer_w2c, feats2, est, n, 1e-3, lambda it: td[it % 10])
run(3); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    run(1); torch.cuda.synchronize()
ev = prof.key_averages(group_by_input_shape=True)
rows = sorted([e for e in ev if e.key.startswith("aten::") and e.device_time_total > 0], key=lambda e: -e.count)[:40]
for e in rows:
    print(f"x{e.count:4d} {e.device_time_total:7.1f}us  {e.key:28s} {str(e.input_shapes)[:110]}")
