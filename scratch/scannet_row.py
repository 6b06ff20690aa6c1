import os, sys, contextlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity_configs as T
from dns_slam_b200 import bench_util, fused, step as stepmod, synthetic as syn
dev = torch.device("cuda:0")
s = syn.SHAPES["scannet"]
dec, samples = bench_util.synthetic_batch("scannet", "map", 2048, 47, 40, dev, seed=11)
samples = {k: v for k, v in samples.items() if k != "mask"}
lam = dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=10.0, fs=s["lambda_fs"], op=s["lambda_opacity"])
o = T._oracle_mapping("scannet", dec, samples, 40, lam, s["opacity_sigma"])
ms = stepmod.MappingStep(dec, s["lr"], lam, s["opacity_sigma"])
out_tc = ms.forward_backward(samples)
with fused.simt_path():
    out_si = ms.forward_backward(samples)
want = o["grads"]["features"].flatten(0, 1)
tc, si = out_tc[4].flatten(0, 1).cpu(), out_si[4].flatten(0, 1).cpu()
err = (tc - want).norm(dim=-1)
i = int(err.argmax()); r, sidx = divmod(i, 47)
print("worst row", i, "ray", r, "sample", sidx, "abs err", float(err[i]), "row norm", float(want[i].norm()), "total norm", float(want.norm()))
print("gt_depth", float(samples["gt_depth"][r]), "z", samples["z_vals"][r].cpu()[max(0, sidx-2):sidx+3].tolist())
print("want", want[i][:8].tolist()); print("tc  ", tc[i][:8].tolist()); print("simt", si[i][:8].tolist())
print("ratio tc/want", (tc[i] / want[i])[:12].tolist())
print("pred color tc/oracle", out_tc[1]["color"][r].cpu().tolist(), o["pred"]["color"][r].tolist())
print("pred depth tc/oracle", float(out_tc[1]["depth"][r]), float(o["pred"]["depth"][r]), "var", float(out_tc[1]["var"][r]))
print("features row", samples["features"][r, sidx, :6].cpu().tolist())
ro, rd = samples["rays_o"][r].cpu(), samples["rays_d"][r].cpu()
pt = ro + rd * samples["z_vals"][r, sidx].cpu()
b = dec.bound.cpu()
print("x normalised", ((pt.double() - b[:, 0]) / (b[:, 1] - b[:, 0])).tolist())
# neighbours of the row
for k in range(max(0, sidx - 2), min(47, sidx + 3)):
    j = r * 47 + k
    print("  s", k, "err", float((tc[j] - want[j]).norm() / (want[j].norm() + 1e-30)), "norm", float(want[j].norm()))
dro = (out_tc[2].cpu() - o["grads"]["rays_o"]).norm(dim=-1)
print("worst d_rays_o ray", int(dro.argmax()), float(dro.max()), float(o["grads"]["rays_o"].norm()))
# latents of the bad sample and the layer-1 pre-activations of the colour / logit nets in float64
cfg = ms._config(samples); cfg.want_latents = True
p = ms._views(dec.flat)
_, preds, *_ = fused.render_raw(cfg, p["table"], p["coarse"], p["color"], p["logit"], p["experts"], samples["rays_o"], samples["rays_d"],
                                samples["features"], None, False, False, forward_only=1)
lat = preds["fine"][i].cpu().double()
print("latent row: occ", float(lat[0]), "abs max", float(lat[1:].abs().max()), "rms", float(lat[1:].pow(2).mean().sqrt()))
from oracle import tcnn_standin as otc
xn = ((pt.double() - b[:, 0]) / (b[:, 1] - b[:, 0])).float()[None]
pe = otc.Encoding(3, {"otype": "OneBlob", "n_bins": 16})(xn)[0].double()
X = torch.cat((pe, lat[1:], samples["features"][r, sidx].cpu().double()))
for name in ("color", "logit"):
    W1 = p[name][:32 * 112].view(32, 112).cpu().double()
    h = W1 @ X
    terms = (W1.abs() @ X.abs())
    order = h.abs().argsort()[:4]
    print(name, "smallest |h|:", [(int(j), float(h[j]), float(terms[j])) for j in order])
