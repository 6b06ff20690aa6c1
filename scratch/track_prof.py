import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import synthetic_slam
from dns_slam_b200 import slam
orig = slam.TrackLoop.run
def timed_run(self, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    self._reset(*a[:4]); torch.cuda.synchronize(); t1 = time.perf_counter()
    r = orig(self, *a, **k); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"reset {1e3*(t1-t0):.1f} ms, run(incl reset) {1e3*(t2-t1):.1f} ms", flush=True)
    return r
slam.TrackLoop.run = timed_run
out = synthetic_slam.run("scannet", 12, n_class=40, track_iters=30, map_iters=20, use_graph=True, verbose=False, map_every=5)
print(out["timings"])
import cProfile, pstats
