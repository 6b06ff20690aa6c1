import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device("cuda:0")
for rep in range(2):
    r = bench._iteration_timings("replica", 40, dev)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if "ms" in k})
