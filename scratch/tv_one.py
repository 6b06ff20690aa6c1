"""A few TV smoothness calls at one shape (for ncu): python scratch/tv_one.py scannet 4"""
import sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import bench_util, fused, synthetic as syn
dev = torch.device("cuda:0")
shape, reps = sys.argv[1], int(sys.argv[2])
s = syn.SHAPES[shape]
dec = bench_util.make_decoder(shape, 40, dev, seed=0)
g = torch.Generator().manual_seed(1)
off, jit = fused.tv_offsets(dec.bound, s["smooth_pts"], torch.rand(3, generator=g), torch.rand(1, 1, 1, 3, generator=g))
d_t, d_c = torch.zeros_like(dec.view("table")), torch.zeros_like(dec.view("coarse"))
for _ in range(reps):
    fused.tv_raw(dec.pe_fn.grid_fn.gstruct, dec.bound, dec.view("table"), dec.view("coarse"), s["smooth_pts"], off, jit,
                 s["lambda_smooth"], d_t, d_c)
torch.cuda.synchronize()
print("ok", float(d_t.abs().sum()))
