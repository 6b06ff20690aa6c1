import torch, sys
sys.path.insert(0, '.')
from dns_slam_b200 import _lib, bench_util, step as stepmod
from oracle import reference_path as rp
dev = torch.device("cuda:0")
dec, samples = bench_util.synthetic_batch("tiny", "map", 700, 47, 9, dev, seed=2, n_frames=2)
samples = {k: v for k, v in samples.items() if k != "mask"}
lam = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
ms = stepmod.MappingStep(dec, 5e-3, lam)
L = _lib.lib()
L.dns_set_tensor_cores(0); o0 = ms.forward_backward(samples)
L.dns_set_tensor_cores(1); o1 = ms.forward_backward(samples)
# third opinion: operator kernels (fp32 SIMT) + torch autograd, reference-style composition
L.dns_set_tensor_cores(0)
smp = dict(samples)
smp["features"] = samples["features"].clone().requires_grad_(True)
smp["pts"] = smp["rays_o"][:, None, :] + smp["rays_d"][:, None, :] * smp["z_vals"][:, :, None]
experts = {c: dec.fine_decoders[c] for c in dec.fine_decoders}
pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(dec, experts, dec.bound, smp)
p, d, l, lt, fs, op = rp.mapping_losses(smp, pc, pd, pl, fine, coarse, 0.05)
loss = 5*p + 5*d + 0.1*l + 10*lt + 10*fs + 10*op
loss.backward()
L.dns_set_tensor_cores(1)
t = smp["features"].grad.double()
a, b = o0[4].double(), o1[4].double()
def rel(x, y): return float((x-y).norm()/y.norm())
print("simt vs truth", rel(a, t), " tc vs truth", rel(b, t), " simt vs tc", rel(a, b))
pa = (a-t).reshape(-1,32).norm(dim=1)/(t.reshape(-1,32).norm(dim=1)+1e-30)
pb = (b-t).reshape(-1,32).norm(dim=1)/(t.reshape(-1,32).norm(dim=1)+1e-30)
print("per-point rel err quantiles simt:", [float(q) for q in torch.quantile(pa, torch.tensor([0.5,0.9,0.99,1.0], dtype=torch.double, device=dev))])
print("per-point rel err quantiles tc  :", [float(q) for q in torch.quantile(pb, torch.tensor([0.5,0.9,0.99,1.0], dtype=torch.double, device=dev))])
print("losses", o0[0][:7].tolist(), float(loss))
bad = torch.nonzero(pb > 1e-3).reshape(-1)
print("n bad", bad.numel(), "of", pb.numel())
print("tile positions of bad points:", sorted(set(int(i) % 141 for i in bad.tolist())))
print("CTAs of bad points (first 30):", [int(i)//141 for i in bad.tolist()][:30])
print("rel errs:", [round(float(pb[i]),4) for i in bad.tolist()][:30])
o2 = ms.forward_backward(samples)
b2 = o2[4].double()
pb2 = (b2-t).reshape(-1,32).norm(dim=1)/(t.reshape(-1,32).norm(dim=1)+1e-30)
bad2 = torch.nonzero(pb2 > 1e-3).reshape(-1)
print("second TC run bad:", bad2.tolist(), " first:", bad.tolist())
i = int(bad[1]); 
print("cols rel err at worst point:", [round(float(abs(b.reshape(-1,32)[i,k]-t.reshape(-1,32)[i,k])/(abs(t.reshape(-1,32)[i,k])+1e-30)),3) for k in range(32)])
print("truth row:", [float('%.3g'%v) for v in t.reshape(-1,32)[i].tolist()])
print("tc row   :", [float('%.3g'%v) for v in b.reshape(-1,32)[i].tolist()])
r = i // 47
print("d_rays_o simt/tc at that ray:", o0[2][r].tolist(), o1[2][r].tolist())
print("z at point, depth:", float(samples['z_vals'].reshape(-1)[i]), float(samples['gt_depth'][r]))
