"""Random shapes: tcgen05 path vs fp32 SIMT path of dns_render_fwd_bwd (mapping and tracking), losses, predictions,
ray / feature gradients and the flat parameter gradient.  usage: python scratch/fuzz.py [n_cases] [seed]"""
import sys, random, torch
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
dev = torch.device("cuda:0")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
L = _lib.lib()
worst = {}
def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
bad = 0
for case in range(n_cases):
    mode = rng.choice(["map", "map", "track"])
    S = rng.choice([1, 2, 5, 13, 31, 32, 33, 47, 64, 65, 96, 127, 128, 129, 200, 256])
    N = rng.choice([1, 2, 3, 7, 64, 100, 257, 700, 1500, 4096])
    if N * S > 300000: N = max(1, 300000 // S)
    C = rng.choice([1, 2, 5, 9, 40, 101])
    dec, samples = bench_util.synthetic_batch("tiny", mode, N, S, C, dev, seed=case, n_frames=1)
    out = {}
    for tc in (0, 1):
        L.dns_set_tensor_cores(tc)
        if mode == "map":
            smp = {k: v for k, v in samples.items() if k != "mask"}
            ms = stepmod.MappingStep(dec, 5e-3)
            o = ms.forward_backward(smp)
            out[tc] = (o, ms.grad.clone())
        else:
            ts = stepmod.TrackingStep(dec)
            o = ts.forward_backward(samples)
            out[tc] = (o, None)
    L.dns_set_tensor_cores(1)
    o0, g0 = out[0]; o1, g1 = out[1]
    errs = {"loss": rel(o1[0][:7], o0[0][:7])}
    for k in ("color", "depth", "var", "logits"):
        errs[k] = rel(o1[1][k], o0[1][k])
    errs["d_rays_o"], errs["d_rays_d"] = rel(o1[2], o0[2]), rel(o1[3], o0[3])
    if o0[4] is not None:
        a, b = o0[4].reshape(-1, 32).double(), o1[4].reshape(-1, 32).double()
        pp = (a - b).norm(dim=1) / (a.norm(dim=1) + 1e-30)
        errs["d_feat_q99"] = float(torch.quantile(pp, 0.99)) if pp.numel() else 0.0
    if g0 is not None:
        for k in ("table", "coarse", "color", "logit", "experts"):
            a, n = dec.layout[k]
            errs["g_" + k] = rel(g1[a:a + n], g0[a:a + n])
    finite = all(torch.isfinite(t).all().item() for t in (o1[0][:7], o1[1]["color"], o1[2], o1[3]))
    flag = [k for k, v in errs.items() if not (v < 2e-3)]
    # NaN in both paths alike (e.g. a fully masked tracking batch) is agreement, not a failure
    if flag and not finite and not all(torch.isfinite(t).all().item() for t in (o0[0][:7],)):
        flag = []
    for k, v in errs.items():
        worst[k] = max(worst.get(k, 0.0), v if v == v else 0.0)
    print(f"case {case:3d} {mode:5s} N={N:5d} S={S:3d} C={C:3d}  max err {max(v for v in errs.values() if v == v):.2e}  {'FLAG ' + str(flag) if flag else 'ok'}", flush=True)
    bad += bool(flag)
print("worst per quantity:", {k: f"{v:.1e}" for k, v in worst.items()})
print("flagged cases:", bad)
