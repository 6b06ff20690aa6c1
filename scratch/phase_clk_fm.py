"""Average clock cycles of thread 0 per phase of k_featmerge_fwd / _bwd (ablate build, DNS_PHASE_CLK_FM) in the bench step."""
import os, sys, torch, argparse
sys.path.insert(0, '.')
import bench
from dns_slam_b200 import encoder
args = argparse.Namespace(shape="replica", n_class=40, rays_per_gpu=131072, samples=47, gpus=1, steps=20, warmup=5)
dev = torch.device("cuda:0")
scene = bench.host_scene("replica", 40)
hp = {"frames": scene["frames"], "refer_img": scene["refer_img"]}
stem = encoder.ResNet().to(dev)
dec = bench.build_decoder(args, scene, dev)
fd, feats, tables = bench.upload_scene(scene, hp, dev, stem, 40)
st = bench.build_gpu_step(args, scene, dec, 0, 1, None, fd, feats, tables)
gen = torch.Generator().manual_seed(1)
draws = [st.make_host_draws(gen).to(dev) for _ in range(2)]
for i in range(2): st.step(draws[i % 2])
torch.cuda.synchronize()
clk = torch.zeros(32, dtype=torch.int64, device=dev)
os.environ["DNS_PHASE_CLK_FM"] = hex(clk.data_ptr())
n = 3
for i in range(n): st.step(draws[i % 2])
torch.cuda.synchronize()
rows = int(st.fm_ws[:4].view(torch.int32)[0])
tiles = n * ((rows + 41) // 42)
c = clk.tolist()
fw = {1: "row geometry (z, rays, projection), tap descriptors, barrier", 2: "OneBlob + cooperative tap gather -> X tile, barrier", 3: "GEMM H = X W1 (+ bulk store of the tile image) + wait",
      4: "H epilogue, barrier", 5: "GEMM O = H W2 + wait", 6: "O -> staging, barrier", 7: "mean over the views + store, barrier"}
bw = {17: "X tile bulk load issued, row geometry, dO tile, barrier", 18: "GEMM H (waits for the X tile) + wait", 19: "H epilogue, barrier", 20: "GEMMs dH, dW2 + wait",
      21: "dH epilogue, barrier", 22: "GEMMs dX, dW1 + wait", 23: "OneBlob backward, per-sample ray gradients (atomics), barrier"}
for name, d in (("k_featmerge_fwd", fw), ("k_featmerge_bwd", bw)):
    tot = sum(c[i] for i in d)
    print(f"{name}: {tot / tiles:.0f} cycles per tile of 42 band samples x 3 views = {tot / tiles / 1.965e3:.2f} us")
    for i, nm in d.items():
        print(f"  {c[i] / tiles:8.0f} cyc  {100 * c[i] / tot:5.1f} %  {nm}")
