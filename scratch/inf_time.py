import sys, torch
sys.path.insert(0, '.')
import bench
print(bench._inference_timings("replica", 40, torch.device("cuda:0")))
