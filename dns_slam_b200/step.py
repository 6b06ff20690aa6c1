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

    def forward_backward(self, samples, need_drays=True, need_dfeat=True):
        """Fills ``self.grad`` (flat) and returns (losses[8], preds, d_rays_o, d_rays_d, d_features)."""
        dec = self.dec
        self.grad.zero_()
        cfg = fused.RenderConfig(_lib.MODE_MAP, dec.bound, dec.pe_fn.grid_fn.gstruct, samples["z_vals"],
                                 samples["gt_color"], samples["gt_depth"], samples["gt_label"], None,
                                 dec.class_to_expert, dec.n_class, self.lambdas, opacity_trunc=self.opacity_sigma)
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
