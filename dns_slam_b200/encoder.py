"""ResNet stem of the pixel-feature branch -- drop-in for ``models/encoder.py:4-17`` over ``models/layers.py:52-114``.

What the reference keeps of ResNet18 is ``conv1`` (7x7 / 2, 3 -> 64, no bias) -> ``bn1`` -> ReLU, evaluated once per
frame under ``no_grad`` (``slams/tracking.py:295-296``) or followed by ``.clone().detach()``
(``slams/mapping.py:768,846``).  It never calls ``.eval()`` on it, so ``bn1`` runs in TRAINING mode: batch statistics
of the views passed in one call, running statistics updated.  ``ResNet`` below keeps the module tree
(``conv_blocks.conv1`` / ``conv_blocks.bn1``: same ``state_dict`` keys and initialisation) and the call signature;
``forward_cl`` returns the channels-last tensor ``dns_feature_gather`` reads, ``forward`` a permuted VIEW of it in the
reference's ``[B, N, C, h, w]`` shape.  The arithmetic runs in ``dns_stem_fwd`` (csrc/stem.cu); there is no torch
fallback.  The feature maps carry no gradient (as in the reference).
"""
import math

import torch
import torch.nn as nn

from . import _lib


class Stem(nn.Module):
    """``models/layers.py:52-72``: the parameters of ``ResNet(BasicBlock, [2, 2, 2, 2])`` that are left."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        n = 7 * 7 * 64
        self.conv1.weight.data.normal_(0, math.sqrt(2.0 / n))
        self.bn1.weight.data.fill_(1)
        self.bn1.bias.data.zero_()

    def load_pretrained(self, state_dict):
        """``ResNet18(pretrained=True)`` (``models/layers.py:119-133``) with the download replaced by a state dict the
        caller has read (torchvision's ``resnet18`` file): keys this module does not have are ignored."""
        own = self.state_dict()
        for k, v in state_dict.items():
            if k in own:
                own[k] = v
        self.load_state_dict(own)
        return self

    @torch.no_grad()
    def forward_cl(self, x_hwc):
        """x_hwc [n, H, W, 3] float32 (cuda) -> [n, h, w, 64] channels-last."""
        bn = self.bn1
        x = x_hwc.detach().float().contiguous()
        n, H, W, _ = x.shape
        h, w = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        out = torch.empty(n, h, w, 64, device=x.device)
        ws_bytes = _lib.lib().dns_stem_workspace_bytes()
        ws = getattr(self, "_ws", None)
        if ws is None or ws.device != x.device or ws.numel() < ws_bytes:
            ws = self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        training = self.training or bn.running_mean is None
        momentum = 0.1 if bn.momentum is None else bn.momentum
        if training and bn.running_mean is not None and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
            if bn.momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        _lib.check(_lib.lib().dns_stem_fwd(
            _lib.ptr(x, torch.float32), n, H, W, _lib.ptr(self.conv1.weight.detach(), torch.float32),
            _lib.ptr(bn.weight.detach(), torch.float32), _lib.ptr(bn.bias.detach(), torch.float32),
            float(bn.eps), float(momentum), int(training),
            _lib.ptr(bn.running_mean, torch.float32, allow_none=True),
            _lib.ptr(bn.running_var, torch.float32, allow_none=True), _lib.ptr(out), _lib.ptr(ws), ws_bytes,
            _lib.stream()))
        return out

    def forward(self, x_nchw):
        """``models/layers.py:97-119``: [n, 3, H, W] -> [n, 64, h, w] (a view of the channels-last result)."""
        return self.forward_cl(x_nchw.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


def ResNet18(pretrained=False, state_dict=None):
    """``models/layers.py:119-133``; there is no network here, so ``pretrained`` needs the weights as ``state_dict``."""
    model = Stem()
    if pretrained:
        if state_dict is None:
            raise RuntimeError("dns_slam_b200.encoder.ResNet18(pretrained=True) needs state_dict= (no download here)")
        model.load_pretrained(state_dict)
    return model


class ResNet(nn.Module):
    """``models/encoder.py:4-17``."""

    def __init__(self, state_dict=None):
        super().__init__()
        self.conv_blocks = ResNet18(pretrained=state_dict is not None, state_dict=state_dict)

    def forward_cl(self, images):
        """images [B, N, H, W, 3] -> [B * N, h, w, 64] channels-last: pass this to ``fused.feature_matching``."""
        return self.conv_blocks.forward_cl(images.flatten(0, 1))

    def forward(self, images):
        """images [B, N, H, W, 3] -> [B, N, 64, h, w], as ``models/encoder.py:9-17``."""
        B, N = images.shape[:2]
        f = self.forward_cl(images)
        return f.permute(0, 3, 1, 2).reshape(B, N, 64, f.shape[1], f.shape[2])
