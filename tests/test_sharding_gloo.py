"""Ray sharding across ranks, host logic on CPU: two gloo processes run
``ShardedMappingStep.step_sharded`` with the CPU oracle plugged in as the local compute.  The
all-reduced flat gradient and losses must equal the single-process full-batch oracle -- this pins
the global denominators, the global ``class(p) = label[p mod N]`` rule, the shard bounds, the
label all-gather, the packed all-reduce and the replicated Adam."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_TOTAL, S, C, SHAPE = 22, 9, 5, "tiny"


def _batch():
    g = torch.Generator().manual_seed(3)
    z = torch.sort(torch.rand(N_TOTAL, S, generator=g) * 1.5 + 0.2, -1)[0]
    d = torch.rand(N_TOTAL, generator=g) * 1.5 + 0.2
    d[4] = 0.0
    return dict(rays_o=torch.rand(N_TOTAL, 3, generator=g) * 0.2 + 0.3,
                rays_d=torch.randn(N_TOTAL, 3, generator=g) * 0.3, z_vals=z, gt_depth=d,
                gt_color=torch.rand(N_TOTAL, 3, generator=g), gt_label=torch.randint(C, (N_TOTAL,), generator=g),
                features=torch.randn(N_TOTAL, S, 32, generator=g) * 0.3)


def _models():
    from oracle.make_golden import build_models
    return build_models(SHAPE, C, 5, expert_classes=range(C))


LAM = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)


def _oracle_shard_loss(dec, experts, bound, smp, lo, n_total, labels_all, counts):
    """Partial loss of rays [lo, lo+n) with GLOBAL denominators; sums over shards to mapping_losses."""
    from oracle import reference_path as rp
    n = smp["z_vals"].shape[0]
    P, Ptot = n * S, n_total * S
    pts = smp["rays_o"][:, None, :] + smp["rays_d"][:, None, :] * smp["z_vals"][:, :, None]
    x = rp.normalise(pts.flatten(0, 1), bound)
    cls = labels_all[(lo * S + torch.arange(P)) % n_total]
    pe, grid = dec.pe_fn(x)
    coarse = dec.coarse_fn(pe, features=grid)
    fine = rp.fine_fn(experts, 32, pe, cls, grid)
    color, logits = dec.out_fn(pe, torch.cat((fine[:, 1:], smp["features"].flatten(0, 1)), -1))
    values = torch.cat((color, fine[:, 0:1]), -1).reshape(n, S, -1)
    depth, var, rgb, w = rp.raw2nerf_color(values, smp["z_vals"])
    plog = torch.sum(w[..., None] * logits.reshape(n, S, -1), -2)
    n_dpos, n_front, n_band = int(counts[1]), int(counts[2]), int(counts[3])
    gd = smp["gt_depth"]
    p = ((smp["gt_color"] - rgb) ** 2).sum() / (3 * n_total)
    m = gd > 0
    d = torch.abs(gd[m] - depth[m]).sum() / n_dpos
    l = F.cross_entropy(plog, smp["gt_label"], reduction="sum") / n_total
    lt = ((coarse - fine) ** 2).sum() / (33 * Ptot)
    occ = torch.sigmoid(10 * fine[:, 32]).reshape(n, S)
    dd = gd.unsqueeze(-1)
    front = (smp["z_vals"] < dd - 0.05).float()
    back = (smp["z_vals"] > dd + 0.05).float()
    valid = (dd > 0).float()
    band = (1 - front) * (1 - back) * valid
    fs = op = torch.zeros(())
    if n_front > 0 and n_band > 0:
        fs = ((occ * front * valid) ** 2).sum() / Ptot
        pseudo = 0.5 * torch.exp(-0.5 * ((smp["z_vals"] - dd) / 0.05) ** 2)
        op = ((occ * band - pseudo * band) ** 2).sum() / Ptot
    total = LAM["p"] * p + LAM["d"] * d + LAM["l"] * l + LAM["lt"] * lt + LAM["fs"] * fs + LAM["op"] * op
    return total, torch.stack([p, d, l, lt, fs, op, total, torch.tensor(float(n_total))]).detach()


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from dns_slam_b200 import decoder as pdec, step as stepmod, synthetic as syn
    from dns_slam_b200.decoder import EXPERT_PARAMS
    bound, odec, oexp = _models()
    dec = pdec.Decoder(syn.model_cfg(SHAPE), bound, n_class=C, device="cpu")      # flat layout only
    names = {"table": odec.pe_fn.grid_fn.params, "coarse": odec.coarse_fn.decoder.params,
             "color": odec.out_fn.color_decoder.params, "logit": odec.out_fn.logit_decoder.params}

    class CpuStep(stepmod.ShardedMappingStep):
        def _local_counts(self, cfg):
            z, d = cfg.z_vals, cfg.gt_depth.unsqueeze(-1)
            front, back = z < d - 0.05, z > d + 0.05
            band = (~front) & (~back) & (d > 0)
            return torch.tensor([z.shape[0], int((cfg.gt_depth > 0).sum()), int(front.sum()), int(band.sum())],
                                dtype=torch.int32)

        def forward_backward(self, samples, need_drays=True, need_dfeat=True, cfg=None):
            self.grad.zero_()
            for prm in list(odec.parameters()) + [e.params for e in oexp.values()]:
                prm.grad = None
            total, losses = _oracle_shard_loss(odec, oexp, bound, samples, cfg.ray_offset, cfg.n_rays_total,
                                               cfg.gt_label_all, cfg.global_counts)
            total.backward()
            v = self._views(self.grad)
            for k, prm in names.items():
                v[k].copy_(prm.grad)
            for c, e in oexp.items():
                if e.params.grad is not None:
                    v["experts"][c].copy_(e.params.grad)
            return (losses, None, None, None, None)

        def _adam(self):
            pass

    st = CpuStep(dec, 5e-3, stepmod.TorchComm(), rank, world, lambdas=LAM)
    batch = _batch()
    lo, hi = stepmod.shard_bounds(N_TOTAL, world, rank)
    local = {k: v[lo:hi].contiguous() for k, v in batch.items()}
    out = st.step_sharded(local, N_TOTAL)
    if rank == 0:
        v = st._views(st.grad)
        ret["losses"] = out[0].clone()
        ret["grads"] = {k: v[k].clone() for k in v}
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_bounds():
    from dns_slam_b200.step import shard_bounds
    for n, w in ((22, 2), (1000003, 8), (5, 8), (4096, 4)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharded_step_equals_full_batch():
    from oracle import reference_path as rp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    # single-process reference: the unsharded oracle (mapper_renderer + mapping_losses)
    bound, odec, oexp = _models()
    b = _batch()
    b["pts"] = b["rays_o"][:, None, :] + b["rays_d"][:, None, :] * b["z_vals"][:, :, None]
    pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(odec, oexp, bound, b)
    p, d, l, lt, fs, op = rp.mapping_losses(b, pc, pd, pl, fine, coarse, 0.05)
    total = LAM["p"] * p + LAM["d"] * d + LAM["l"] * l + LAM["lt"] * lt + LAM["fs"] * fs + LAM["op"] * op
    total.backward()
    want = torch.stack([p, d, l, lt, fs, op, total]).detach()
    torch.testing.assert_close(ret["losses"][:7], want, rtol=1e-5, atol=1e-7)
    g = ret["grads"]
    torch.testing.assert_close(g["table"], odec.pe_fn.grid_fn.params.grad, rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(g["coarse"], odec.coarse_fn.decoder.params.grad, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(g["color"], odec.out_fn.color_decoder.params.grad, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(g["logit"], odec.out_fn.logit_decoder.params.grad, rtol=1e-4, atol=1e-8)
    for c, e in oexp.items():
        if e.params.grad is not None:
            torch.testing.assert_close(g["experts"][c], e.params.grad, rtol=1e-4, atol=1e-8)
