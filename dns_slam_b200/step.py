"""Autograd-free optimisation steps over the decoder's flat parameter buffer.

``MappingStep`` is the native body of one mapping iteration's core (``slams/mapping.py:888-910``
without sampling): zero the flat gradient, ONE fused render + loss + backward call that scatters
every gradient straight into the flat buffer, an optional NCCL all-reduce of that buffer when the
ray batch is sharded across GPUs, and ONE fused Adam kernel over all parameters
(``torch.optim.Adam`` defaults; a fresh state per ``optimize()`` call as ``mapping.py:438-468``).
``TrackingStep`` is the pose-only counterpart (``slams/tracking.py:313-340``).
"""
import torch

from . import _lib, fused
from .decoder import EXPERT_PARAMS


class MappingStep:
    def __init__(self, decoder, lr, lambdas=None, opacity_sigma=0.05, process_group=None, world_size=1):
        self.dec = decoder
        self.lr = lr
        self.lambdas = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
        self.lambdas.update(lambdas or {})
        self.opacity_sigma = opacity_sigma
        self.pg, self.world = process_group, world_size
        self.grad = torch.zeros_like(decoder.flat)
        self.reset()

    def reset(self):
        """Fresh Adam state (the reference builds a new optimiser for every optimize() call)."""
        self.m = torch.zeros_like(self.dec.flat)
        self.v = torch.zeros_like(self.dec.flat)
        self.t = 0

    def _views(self, buf):
        lay = self.dec.layout
        out = {k: buf[lay[k][0]:lay[k][0] + lay[k][1]] for k in ("table", "coarse", "color", "logit")}
        a, n = lay["experts"]
        out["experts"] = buf[a:a + n].view(-1, EXPERT_PARAMS)
        return out

    def _config(self, samples):
        dec = self.dec
        return fused.RenderConfig(_lib.MODE_MAP, dec.bound, dec.pe_fn.grid_fn.gstruct, samples["z_vals"],
                                  samples["gt_color"], samples["gt_depth"], samples["gt_label"], None,
                                  dec.class_to_expert, dec.n_class, self.lambdas, opacity_trunc=self.opacity_sigma)

    def forward_backward(self, samples, need_drays=True, need_dfeat=True, cfg=None):
        """Fills ``self.grad`` (flat) and returns (losses[8], preds, d_rays_o, d_rays_d, d_features)."""
        dec = self.dec
        self.grad.zero_()
        cfg = cfg or self._config(samples)
        p = self._views(dec.flat)
        return fused.render_raw(cfg, p["table"], p["coarse"], p["color"], p["logit"], p["experts"],
                                samples["rays_o"], samples["rays_d"], samples.get("features"),
                                self._views(self.grad), need_drays, need_dfeat)

    def step(self, samples, need_drays=True, need_dfeat=True):
        out = self.forward_backward(samples, need_drays, need_dfeat)
        if self.world > 1:
            torch.distributed.all_reduce(self.grad, group=self.pg)
        self.t += 1
        fused.adam_step(self.dec.flat, self.grad, self.m, self.v, self.lr, self.t)
        return out


def shard_bounds(n_total, world, rank):
    """Contiguous, balanced ray ranges: rank r owns [lo, hi)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedMappingStep(MappingStep):
    """One mapping batch sharded by rays over the ranks of ``process_group``.

    Every rank holds the full decoder (replicated) and ITS slice of the ray batch.  Per step:
    (1) labels of the whole batch are all-gathered once (the class rule ``class(p) = label[p mod N]``
    of mapping.py:612-613 reaches across shards) and the four batch-global counts are all-reduced
    (tiny); (2) each rank runs the fused kernels on its rays with GLOBAL denominators, so partial
    losses and gradients SUM to the single-GPU values; (3) ONE all-reduce of the flat gradient
    buffer (+ 8 loss partials appended to it); (4) identical fused Adam on every rank -- no
    parameter broadcast.  ``comm`` is any object with all_reduce_sum(tensor) / all_gather(tensor)
    (torch.distributed over NCCL on GPUs; the CPU tests plug gloo)."""

    def __init__(self, decoder, lr, comm, rank, world, **kw):
        super().__init__(decoder, lr, **kw)
        self.comm, self.rank, self.world_n = comm, rank, world
        self.packed = torch.zeros(self.grad.numel() + 8, device=self.grad.device)
        self.grad = self.packed[:-8]                 # gradients and loss partials travel together

    def step_sharded(self, local_samples, n_total, need_drays=True, need_dfeat=True):
        lo, hi = shard_bounds(n_total, self.world_n, self.rank)
        assert local_samples["z_vals"].shape[0] == hi - lo, "local batch does not match the shard bounds"
        sizes = [b - a for a, b in (shard_bounds(n_total, self.world_n, r) for r in range(self.world_n))]
        labels_all = self.comm.all_gather(local_samples["gt_label"].contiguous(), sizes)
        cfg = self._config(local_samples)
        counts = self._local_counts(cfg)
        self.comm.all_reduce_sum(counts)
        cfg.shard(n_total, lo, labels_all, counts)
        out = self.forward_backward(local_samples, need_drays, need_dfeat, cfg=cfg)
        self.packed[-8:] = out[0]
        self.comm.all_reduce_sum(self.packed)
        losses = self.packed[-8:].clone()
        losses[7] = losses[7] / self.world_n          # n_valid is a batch constant, not a partial sum
        self.t += 1
        self._adam()
        return (losses,) + tuple(out[1:])

    # the two device calls besides forward_backward (overridden by the CPU host-logic tests)
    def _local_counts(self, cfg):
        return fused.render_counts(cfg)

    def _adam(self):
        fused.adam_step(self.dec.flat, self.grad, self.m, self.v, self.lr, self.t)


class TorchComm:
    """torch.distributed adapter (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        self.group = group

    def all_reduce_sum(self, t):
        torch.distributed.all_reduce(t, group=self.group)
        return t

    def all_gather(self, t, sizes=None):
        """Concatenation over ranks of possibly different-length 1-D tensors.  ``sizes`` (per-rank lengths known to the
        caller, e.g. from shard_bounds) skips the size exchange and its device->host read."""
        world = torch.distributed.get_world_size(self.group)
        if sizes is not None and len(set(sizes)) == 1:
            out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
            torch.distributed.all_gather_into_tensor(out, t, group=self.group)
            return out
        if sizes is None:
            n = torch.tensor([t.numel()], device=t.device)
            sizes = [torch.zeros_like(n) for _ in range(world)]
            torch.distributed.all_gather(sizes, n, group=self.group)
        mx = int(max(int(s) for s in sizes))
        pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
        pad[:t.numel()] = t
        outs = [torch.zeros_like(pad) for _ in range(world)]
        torch.distributed.all_gather(outs, pad, group=self.group)
        return torch.cat([o[:int(s)] for o, s in zip(outs, sizes)])


class TrackingStep:
    """Fused tracking forward/backward: gradients w.r.t. the rays only (the decoder is frozen in
    tracking), to be chained to quaternion / translation by the caller."""

    def __init__(self, decoder, lambdas=None):
        self.dec = decoder
        self.lambdas = dict(p=5.0, d=5.0, l=0.1)
        self.lambdas.update(lambdas or {})

    def forward_backward(self, samples, need_dfeat=True):
        dec = self.dec
        cfg = fused.RenderConfig(_lib.MODE_TRACK, dec.bound, dec.pe_fn.grid_fn.gstruct, samples["z_vals"],
                                 samples["gt_color"], samples["gt_depth"], samples["gt_label"], samples.get("mask"),
                                 None, dec.n_class, self.lambdas)
        return fused.render_raw(cfg, dec.view("table"), dec.view("coarse"), dec.view("color"), dec.view("logit"), None,
                                samples["rays_o"], samples["rays_d"], samples.get("features"), None, True, need_dfeat)


class HostBatchPipeline:
    """Feeds ray batches that live in PINNED HOST memory to a step function, double buffered: the H2D copy
    of batch k+1 runs on a copy stream while the kernels of batch k run on the compute stream, and the loss
    vector of every step is read back into pinned memory.  Used for the end-to-end number of bench.py (every
    step's inputs really cross PCIe inside the timed region) and for hosts that sample on the CPU."""

    def __init__(self, template_host_batch, device):
        self.dev = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.bufs = [{k: torch.empty_like(v, device=device) for k, v in template_host_batch.items()} for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]      # H2D of buffer i finished
        self.free = [torch.cuda.Event() for _ in range(2)]       # compute on buffer i finished
        self.result_host = torch.empty(8, pin_memory=True)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in template_host_batch.values())
        for e in self.free:
            e.record(torch.cuda.current_stream(device))

    def _upload(self, host_batch, i):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[i])
            for k, v in host_batch.items():
                self.bufs[i][k].copy_(v, non_blocking=True)
            self.ready[i].record(self.copy_stream)

    def run(self, host_batches, step_fn):
        """host_batches: sequence of dicts of pinned host tensors (may repeat the same object);
        step_fn(device_batch) -> tuple whose first element is the loss vector [8]."""
        cur = torch.cuda.current_stream(self.dev)
        n = len(host_batches)
        if n == 0:
            return self.result_host
        self._upload(host_batches[0], 0)
        for k in range(n):
            i = k & 1
            if k + 1 < n:
                self._upload(host_batches[k + 1], (k + 1) & 1)
            cur.wait_event(self.ready[i])
            out = step_fn(self.bufs[i])
            self.free[i].record(cur)
            self.result_host.copy_(out[0], non_blocking=True)
        return self.result_host
