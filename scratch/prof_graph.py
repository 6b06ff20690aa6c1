import torch, sys
sys.path.insert(0, '.')
from dns_slam_b200 import bench_util, slam, synthetic as syn
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
shape = sys.argv[1] if len(sys.argv) > 1 else "replica"; s = syn.SHAPES[shape]
dec = bench_util.make_decoder(shape, 40, dev, seed=1)
sc = bench_util.slam_scene(shape, 40, dev, seed=2)
cam = sc["cam"]
mp = slam.MapperCore(cam, dec, s["mapping_pixels"], 32, 15,
                     lambdas=dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=10.0, fs=s["lambda_fs"], op=s["lambda_opacity"]),
                     opacity_sigma=s["opacity_sigma"], smooth_pts=s["smooth_pts"], lambda_sm=s["lambda_smooth"])
mdg, tvg = bench_util.mapping_draws(sc, s["mapping_pixels"], 8)
target = dict(kf_idx=sc["kf_idx"], frames=sc["frames"], class_tables=sc["class_tables"])
refer = dict(kf_idx=sc["refer_idx"], est_c2w=sc["refer_c2w"])
est_list = [sc["poses"][2 * f + 1].clone() for f in range(len(sc["frames"]))]
def run(n): slam.map_optimize(mp, target, refer, sc["feats"], est_list, n, s["lr"], s["BA_cam_lr"], True, [], lambda it: mdg[it % 8], lambda it: tvg[it % 8], use_graph=True)
run(8); torch.cuda.synchronize()
N = 24
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(N); torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(ev, key=lambda e: -e.device_time_total)[:40]
tot = sum(e.device_time_total for e in ev)
print("total device us per iteration ~", tot / N)
for e in rows:
    print(f"{e.device_time_total/N:9.1f} us/it  x{e.count/N:6.1f}  {e.key[:110]}")
