"""A few MappingFrameStep iterations at the SLAM batch size (Replica, 4 x 500 rays) -- for an ncu launch list:
python scratch/native_iter.py [n_iterations]"""
import sys, copy, torch
sys.path.insert(0, '.')
import bench
from dns_slam_b200 import encoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sys.argv = sys.argv[:1]
args = bench.parse()
dev = torch.device("cuda:0")
scene = bench.host_scene(args.shape, args.n_class)
a2 = copy.copy(args); a2.rays_per_gpu = 2000
stem = encoder.ResNet().to(dev)
hp = {"frames": scene["frames"], "refer_img": scene["refer_img"]}
frames_dev, feats, tables = bench.upload_scene(scene, hp, dev, stem, args.n_class)
dec = bench.build_decoder(a2, scene, dev)
st = bench.build_gpu_step(a2, scene, dec, 0, 1, None, frames_dev, feats, tables)
gen = torch.Generator().manual_seed(5)
draws = [st.make_host_draws(gen).to(dev) for _ in range(4)]
for i in range(n):
    st.step(draws[i % 4])
torch.cuda.synchronize()
print("MARK last iteration done", st.n_total)
