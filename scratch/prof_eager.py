import torch, sys
sys.path.insert(0, '.')
from dns_slam_b200 import bench_util, slam, synthetic as syn
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
shape = "replica"; s = syn.SHAPES[shape]
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
def run(n): slam.map_optimize(mp, target, refer, sc["feats"], est_list, n, s["lr"], s["BA_cam_lr"], True, [], lambda it: mdg[it % 8], lambda it: tvg[it % 8])
run(3); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    run(1); torch.cuda.synchronize()
ev = prof.key_averages(group_by_input_shape=True)
rows = sorted([e for e in ev if e.key.startswith("aten::") and e.count >= 4], key=lambda e: -e.count)[:45]
for e in rows:
    print(f"x{e.count:4d}  {e.key:32s} {str(e.input_shapes)[:120]}")
