"""Where the one-off part of bench.py's e2e region goes: upload_scene / build_gpu_step / first step, three rounds."""
import sys, time, torch, argparse
sys.path.insert(0, '.')
import bench
from dns_slam_b200 import encoder
args = argparse.Namespace(shape="replica", n_class=40, rays_per_gpu=131072, samples=47, gpus=1, steps=20, warmup=5)
dev = torch.device("cuda:0")
scene = bench.host_scene("replica", 40)
hp = {"frames": [{k: fr[k].contiguous().pin_memory() for k in ("color", "depth", "label")} for fr in scene["frames"]],
      "refer_img": [x.contiguous().pin_memory() for x in scene["refer_img"]]}
stem = encoder.ResNet().to(dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rnd in range(4):
    dec = bench.build_decoder(args, scene, dev)
    t0 = T()
    fd, feats, tables = bench.upload_scene(scene, hp, dev, stem, 40)
    t1 = T()
    st = bench.build_gpu_step(args, scene, dec, 0, 1, None, fd, feats, tables)
    t2 = T()
    if rnd == 0:
        gen = torch.Generator().manual_seed(1)
        draws = [st.make_host_draws(gen) for _ in range(4)]
    st.upload(draws[0]); st.step(); st.read_result()
    t3 = T()
    for i in range(5):
        st.upload(draws[i % 4]); st.step(); st.read_result()
    t4 = T()
    print(f"round {rnd}: upload_scene {1e3*(t1-t0):.1f} ms, build step {1e3*(t2-t1):.1f} ms, first step {1e3*(t3-t2):.1f} ms, next 5 steps {1e3*(t4-t3)/5:.1f} ms each", flush=True)
    del st, dec, fd, feats, tables
import cProfile, pstats, io
dec = bench.build_decoder(args, scene, dev)
pr = cProfile.Profile(); pr.enable()
fd, feats, tables = bench.upload_scene(scene, hp, dev, stem, 40)
st = bench.build_gpu_step(args, scene, dec, 0, 1, None, fd, feats, tables)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:5000])
