import torch, sys, time
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod, fused, synthetic as syn
dev = torch.device("cuda:0")
R, S, C = 131072, 47, 40
dec = bench_util.make_decoder("replica", C, dev, seed=0)
_, samples = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=100, dec=dec)
ms = stepmod.MappingStep(dec, 5e-3)
for _ in range(3): ms.step(samples)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = [0.0, 0.0, 0.0]
cpu = 0.0
for it in range(5):
    t0 = time.perf_counter()
    ev[0].record()
    ms.grad.zero_()
    ev[1].record()
    out = ms.forward_backward.__wrapped__(ms, samples) if hasattr(ms.forward_backward, "__wrapped__") else None
    if out is None:
        cfg = ms._config(samples); p = ms._views(dec.flat)
        out = fused.render_raw(cfg, p["table"], p["coarse"], p["color"], p["logit"], p["experts"], samples["rays_o"], samples["rays_d"], samples.get("features"), ms._views(ms.grad), True, True)
    ev[2].record()
    ms.t += 1
    fused.adam_step(dec.flat, ms.grad, ms.m, ms.v, ms.lr, ms.t)
    ev[3].record()
    cpu += time.perf_counter() - t0
    torch.cuda.synchronize()
    for k in range(3): tot[k] += ev[k].elapsed_time(ev[k+1])
print("per step ms: zero %.3f render_raw %.3f adam %.3f | cpu issue %.3f" % (tot[0]/5, tot[1]/5, tot[2]/5, cpu/5*1e3))
# C call only, with phase timers
_lib.profile_read(True); _lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ms.step(samples)
e1.record(); torch.cuda.synchronize()
_lib.profile_enable(False)
phase, ln = _lib.profile_read(True)
print("loop ms/step %.3f, phase sum %.3f" % (e0.elapsed_time(e1)/5, sum(phase.values())/5), {k: round(v/5,3) for k,v in phase.items() if v})
