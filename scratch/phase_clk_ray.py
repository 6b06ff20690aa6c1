"""Average clock cycles of thread 0 per phase of k_ray_tc2 (ablate build, DNS_PHASE_CLK_RAY)."""
import os, sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
R, S, C = 131072, 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
clk = torch.zeros(24, dtype=torch.int64, device=dev)
names = {1: "X staging: latent / feature / ray loads, OneBlob, tile + image stores", 2: "barrier", 3: "forward GEMM (+ compositing scans of group 1) + wait",
         4: "hidden read back (TMEM), barrier", 5: "colour head, staging of w*h / w*rgb, barrier", 6: "per-ray sums (36 columns x S), barrier",
         7: "logit layer 2 on the composited hidden state, barrier", 8: "per-ray losses (warp per ray), barrier", 9: "QV = dlogit W2, colour part of d_w, barrier",
         10: "d_w, barrier", 11: "d_u / d_alpha scans, image stores (H colour, dpre), dH tile + image, barrier", 12: "backward GEMM + wait",
         13: "dX epilogue: d_features store, dfine scatter (slot image), OneBlob backward, barrier", 14: "ray-gradient staging, barrier", 15: "ray-gradient sums + stores"}
for _ in range(2): ms.step(samples)
torch.cuda.synchronize()
os.environ["DNS_PHASE_CLK_RAY"] = hex(clk.data_ptr())
n = 3
for _ in range(n): ms.step(samples)
torch.cuda.synchronize()
ctas = n * ((R + 4) // 5)
c = clk.tolist()
tot = sum(c)
print(f"{tot / ctas:.0f} cycles per CTA (5 rays x 47 samples) = {tot / ctas / 1.965e3:.2f} us")
for i in range(1, 16):
    print(f"  {c[i] / ctas:8.0f} cyc  {100 * c[i] / tot:5.1f} %  {names[i]}")
