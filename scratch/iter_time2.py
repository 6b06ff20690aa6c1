import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device("cuda:0")
for shape in ("replica", "scannet"):
    print(shape, json.dumps(bench._iteration_timings(shape, 40, dev)), flush=True)
