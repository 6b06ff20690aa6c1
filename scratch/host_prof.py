"""Host-side cost of one native mapping step at the SLAM batch size (2000 rays, 4 key frames): cProfile of step()."""
import sys, time, torch, argparse, cProfile, pstats, io
sys.path.insert(0, '.')
import bench
from dns_slam_b200 import encoder
args = argparse.Namespace(shape="replica", n_class=40, rays_per_gpu=2000, samples=47, gpus=1, steps=20, warmup=5)
dev = torch.device("cuda:0")
scene = bench.host_scene("replica", 40)
hp = {"frames": scene["frames"], "refer_img": scene["refer_img"]}
stem = encoder.ResNet().to(dev)
dec = bench.build_decoder(args, scene, dev)
fd, feats, tables = bench.upload_scene(scene, hp, dev, stem, 40)
st = bench.build_gpu_step(args, scene, dec, 0, 1, None, fd, feats, tables)
gen = torch.Generator().manual_seed(1)
draws = [st.make_host_draws(gen) for _ in range(4)]
for i in range(10):
    st.upload(draws[i % 4]); st.step()
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for i in range(n):
    st.upload(draws[i % 4]); st.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue {1e3 * (t1 - t0) / n:.3f} ms per step, wall {1e3 * (t2 - t0) / n:.3f} ms per step")
pr = cProfile.Profile(); pr.enable()
for i in range(100):
    st.upload(draws[i % 4]); st.step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
