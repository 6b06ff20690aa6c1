"""Round 2 debugging aid: relative errors of every gradient of the fused call against the CPU oracle, for the tcgen05
path and the fp32 SIMT path, at the BASELINE shapes.  python scratch/parity_table.py"""
import os, sys, contextlib
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import rel_err
import test_gpu_parity_configs as T
from dns_slam_b200 import bench_util, fused, step as stepmod, synthetic as syn

dev = torch.device("cuda:0")
LAM = T.LAM


def run(shape, N, C, seed, hash_note=""):
    dec, samples = bench_util.synthetic_batch(shape, "map", N, 47, C, dev, seed=seed, n_frames=4 if N >= 64 else 2)
    samples = {k: v for k, v in samples.items() if k != "mask"}
    o = T._oracle_mapping(shape, dec, samples, C, LAM, 0.05)
    for name, ctx in (("tc", contextlib.nullcontext()), ("simt", fused.simt_path())):
        ms = stepmod.MappingStep(dec, 5e-3, LAM, 0.05)
        with ctx:
            losses, preds, d_o, d_d, d_f = ms.forward_backward(samples)
        gv = ms._views(ms.grad)
        row = {k: rel_err(gv[k], o["grads"][k]) for k in ("table", "coarse", "color", "logit")}
        row["experts"] = rel_err(gv["experts"][:, :3616], o["grads"]["experts"][:, :3616])
        row["d_feat"] = rel_err(d_f, o["grads"]["features"])
        row["d_o"], row["d_d"] = rel_err(d_o, o["grads"]["rays_o"]), rel_err(d_d, o["grads"]["rays_d"])
        q = T._row_quantiles(d_o, o["grads"]["rays_o"])
        row["pred"] = max(rel_err(preds[k], o["pred"][k]) for k in ("color", "depth", "logits"))
        row["loss"] = rel_err(losses[:7], torch.stack([x.detach() for x in o["losses"]]).float())
        print(f"{shape:8s} N={N:5d} {name:5s} " + " ".join(f"{k}={v:.1e}" for k, v in row.items()) +
              f" | d_o per ray: median {q[0]:.1e} q99.9 {q[1]:.1e} max {q[2]:.1e}", flush=True)
        # per-level error of the table gradient
        if name == "tc":
            g = dec.pe_fn.grid_fn
            offs = list(g.tables["offset"])
            lv = [rel_err(gv["table"][2 * offs[l]:2 * offs[l + 1]], o["grads"]["table"][2 * offs[l]:2 * offs[l + 1]]) for l in range(16)]
            print("      table per level:", " ".join(f"{v:.0e}" for v in lv), flush=True)


for args in (("tiny", 45, 6, 5), ("tiny", 300, 6, 6), ("tiny", 2000, 6, 7), ("replica", 4096, 40, 7), ("scannet", 2048, 40, 11)):
    run(*args)
# determinism / sharding: full batch twice, and two shards summed
dec, samples = bench_util.synthetic_batch("tiny", "map", 300, 47, 6, dev, seed=5, n_frames=2)
samples = {k: v for k, v in samples.items() if k != "mask"}
ms = stepmod.MappingStep(dec, 5e-3, LAM, 0.05)
ms.forward_backward(samples); g1 = ms.grad.clone()
ms.forward_backward(samples); g2 = ms.grad.clone()
print("full twice:", rel_err(g1, g2), "table", rel_err(ms._views(g1)["table"], ms._views(g2)["table"]))
N = 300
shards = [stepmod.shard_bounds(N, 2, r) for r in range(2)]
local = [{k: v[lo:hi].contiguous() for k, v in samples.items()} for lo, hi in shards]
counts = sum(fused.render_counts(ms._config(s)) for s in local)
gs = torch.zeros_like(g1)
for (lo, hi), s in zip(shards, local):
    ms.forward_backward(s, cfg=ms._config(s).shard(N, lo, samples["gt_label"], counts))
    gs += ms.grad
va, vb = ms._views(gs), ms._views(g1)
print("2 shards vs full:", {k: f"{rel_err(va[k], vb[k]):.1e}" for k in va})
d = (va["table"] - vb["table"]).abs()
print("table entries differing by > 1e-3 of max:", int((d > 1e-3 * vb["table"].abs().max()).sum()), "of", int((vb["table"] != 0).sum()))
