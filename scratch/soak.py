"""Soak: many steps at several batch sizes (mbarrier pipelines must never hang; losses must stay finite)."""
import sys, time, torch
sys.path.insert(0, '.')
from dns_slam_b200 import bench_util, step as stepmod
dev = torch.device("cuda:0")
t0 = time.time()
for R, S, C, steps in [(131072, 47, 40, 150), (2000, 47, 40, 400), (500, 47, 40, 400), (33333, 96, 12, 60), (777, 13, 3, 300)]:
    dec = bench_util.make_decoder("replica", C, dev, seed=R)
    _, smp = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=R + 1, dec=dec)
    smp = {k: v for k, v in smp.items() if k != "mask"}
    ms = stepmod.MappingStep(dec, 5e-3)
    first = last = None
    for it in range(steps):
        out = ms.step(smp)
        if it % 50 == 0 or it == steps - 1:
            l = out[0].tolist()
            assert all(x == x for x in l[:7]), (R, S, C, it, l)
            first = first or l
            last = l
    torch.cuda.synchronize()
    print(f"R={R} S={S} C={C} steps={steps}: total-loss {first[6]:.4f} -> {last[6]:.4f}  ({time.time()-t0:.1f}s)", flush=True)
print("soak ok")
