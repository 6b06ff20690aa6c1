"""Average clock cycles of thread 0 per phase of k_point_bwd_tc2 (ablate build, DNS_PHASE_CLK)."""
import os, sys, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
R, S, C = 131072, 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
clk = torch.zeros(16, dtype=torch.int64, device=dev)
names = ["W2 loads + tmem alloc", "perm -> ray -> point", "dOut rows (slot-order row loads)", "barrier 1", "GEMM dH + W1 prefetch + wait",
         "dH epilogue (H image read, stores)", "barrier 2", "GEMM dX + wait", "dX read, OneBlob bwd, barrier 3, dealloc",
         "hash-grid backward (thread 0)", "ray gradients"]
for dbg in (0, 14):
    os.environ["DNS_DBG"] = str(dbg)
    os.environ.pop("DNS_PHASE_CLK", None)
    for _ in range(2): ms.step(samples)
    torch.cuda.synchronize()
    clk.zero_()
    os.environ["DNS_PHASE_CLK"] = hex(clk.data_ptr())
    n = 3
    for _ in range(n): ms.step(samples)
    torch.cuda.synchronize()
    tiles = n * ((R * S + 127) // 128 + 40)
    c = clk.tolist()
    tot = sum(c[:11])
    print(f"DNS_DBG={dbg}: {tot / tiles:.0f} cycles per tile = {tot / tiles / 1.965e3:.2f} us")
    for i, nm in enumerate(names):
        print(f"  {c[i] / tiles:8.0f} cyc  {100 * c[i] / tot:5.1f} %  {nm}")
