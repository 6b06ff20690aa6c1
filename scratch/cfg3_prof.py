"""Host profile of the config-3 example loop (ScanNet shape, reference iteration counts) over a few frames."""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, '.'); sys.path.insert(0, 'examples')
import torch
import synthetic_slam as ex
from dns_slam_b200 import synthetic as syn
shape, n = sys.argv[1] if len(sys.argv) > 1 else "scannet", int(sys.argv[2]) if len(sys.argv) > 2 else 26
s = syn.SHAPES[shape]
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
out = ex.run(shape, n, n_class=40, track_iters=s["tracking_iters"], map_iters=s["mapping_iters"], map_every=5, verbose=False)
pr.disable()
torch.cuda.synchronize()
print("wall", time.perf_counter() - t0, out["timings"])
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(45); print(st.getvalue()[:9000])
