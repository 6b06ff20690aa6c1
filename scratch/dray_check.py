import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from dns_slam_b200 import _lib, bench_util, step as stepmod
from oracle import reference_path as rp
dev = torch.device("cuda:0")
L = _lib.lib()
LAM = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
for (N, S, C, seed) in [(257, 31, 40, 39), (700, 13, 101, 43), (700, 47, 9, 8)]:
    dec, samples = bench_util.synthetic_batch("tiny", "map", N, S, C, dev, seed=seed, n_frames=1)
    smp = {k: v for k, v in samples.items() if k != "mask"}
    out = {}
    for tc in (0, 1):
        L.dns_set_tensor_cores(tc)
        ms = stepmod.MappingStep(dec, 5e-3, lambdas=LAM)
        o = ms.forward_backward(smp)
        out[tc] = (o[2].clone(), o[3].clone(), o[0].clone())
    L.dns_set_tensor_cores(1)
    # operator composition with autograd (independent of the fused kernels)
    ro = smp["rays_o"].clone().requires_grad_(True); rd = smp["rays_d"].clone().requires_grad_(True)
    s2 = dict(smp, rays_o=ro, rays_d=rd)
    s2["pts"] = ro[:, None, :] + rd[:, None, :] * smp["z_vals"][:, :, None]
    pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(dec, dec.fine_decoders, dec.bound, s2)
    p, d, l, lt, fs, op = rp.mapping_losses(s2, pc, pd, pl, fine, coarse, 0.05)
    total = LAM["p"] * p + LAM["d"] * d + LAM["l"] * l + LAM["lt"] * lt + LAM["fs"] * fs + LAM["op"] * op
    total.backward()
    print(f"N={N} S={S} C={C}: d_rays_d  TC vs ops {rel(out[1][1], rd.grad):.2e}  SIMT vs ops {rel(out[0][1], rd.grad):.2e}  TC vs SIMT {rel(out[1][1], out[0][1]):.2e}"
          f" | d_rays_o TC vs ops {rel(out[1][0], ro.grad):.2e} SIMT vs ops {rel(out[0][0], ro.grad):.2e}")
    # per-ray view: how concentrated is the difference?
    diff = (out[1][1] - rd.grad).norm(dim=1); ref = rd.grad.norm(dim=1)
    k = torch.topk(diff, 3)
    print("   worst rays (TC vs ops):", [(int(i), float(diff[i]), float(ref[i])) for i in k.indices], " total norm", float(rd.grad.norm()))
