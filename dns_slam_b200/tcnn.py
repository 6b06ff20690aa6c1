"""Operator surface: a drop-in for the ``tinycudann`` torch binding as the reference uses it.

``import dns_slam_b200.tcnn as tcnn`` provides ``tcnn.Encoding`` (``HashGrid`` and ``OneBlob``,
``models/pos_encoding.py:31-46,61-71``) and ``tcnn.Network`` (``CutlassMLP``, one hidden layer,
ReLU, no bias; ``models/decoder.py:58-65,84-91,101-117``, ``slams/mapping.py:737-744``) with the
same constructor signatures, the same single flat fp32 ``params`` Parameter
(``[W1 (n_neurons x in_pad) | W2 (out_pad x n_neurons)]`` row-major; hash table ``[entry][feature]``
with levels concatenated) and autograd support.  Every forward/backward is a hand-written
sm_100a kernel reached through the C ABI; CPU tensors are rejected (no fallback).
Arithmetic is fp32 throughout (tinycudann itself computes the MLPs in fp16).
"""
import ctypes as C
import math

import torch
from torch import nn

from . import _lib, grid as _grid


def _f32c(x):
    return x.to(torch.float32).contiguous()


class _OneBlobFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_bins):
        P, D = x.shape
        out = torch.empty(P, D * n_bins, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().dns_oneblob_fwd(_lib.ptr(x, torch.float32), P, D, n_bins, _lib.ptr(out), _lib.stream()))
        ctx.save_for_backward(x)
        ctx.n_bins = n_bins
        return out

    @staticmethod
    def backward(ctx, d_out):
        (x,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        P, D = x.shape
        d_x = torch.empty_like(x)
        _lib.check(_lib.lib().dns_oneblob_bwd(_lib.ptr(x), _lib.ptr(_f32c(d_out)), P, D, ctx.n_bins,
                                              _lib.ptr(d_x), _lib.stream()))
        return d_x, None


class _HashGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, gstruct, width):
        P = x.shape[0]
        out = torch.empty(P, width, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().dns_hashgrid_fwd(C.byref(gstruct), _lib.ptr(x, torch.float32),
                                               _lib.ptr(params, torch.float32), P, _lib.ptr(out), _lib.stream()))
        ctx.save_for_backward(x, params)
        ctx.gstruct = gstruct
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, params = ctx.saved_tensors
        P = x.shape[0]
        need_x, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_params = torch.zeros_like(params) if need_p else None
        d_x = torch.empty_like(x) if need_x else None
        if need_x or need_p:
            _lib.check(_lib.lib().dns_hashgrid_bwd(C.byref(ctx.gstruct), _lib.ptr(x), _lib.ptr(params),
                                                   _lib.ptr(_f32c(d_out)), P, _lib.ptr(d_params, allow_none=True),
                                                   _lib.ptr(d_x, allow_none=True), _lib.stream()))
        return d_x, d_params, None, None


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, n_in, n_out):
        P = x.shape[0]
        out = torch.empty(P, n_out, device=x.device, dtype=torch.float32)
        hidden = torch.empty(P, 32, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().dns_mlp_fwd(_lib.ptr(x, torch.float32), _lib.ptr(params, torch.float32), P, n_in,
                                          n_out, _lib.ptr(out), _lib.ptr(hidden), _lib.stream()))
        ctx.save_for_backward(x, params, hidden)
        ctx.dims = (n_in, n_out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, params, hidden = ctx.saved_tensors
        n_in, n_out = ctx.dims
        P = x.shape[0]
        need_x, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_hidden = torch.empty_like(hidden)
        d_x = torch.empty_like(x) if need_x else None
        d_params = torch.zeros_like(params) if need_p else None
        _lib.check(_lib.lib().dns_mlp_bwd(_lib.ptr(x), _lib.ptr(params), _lib.ptr(hidden), _lib.ptr(_f32c(d_out)),
                                          P, n_in, n_out, _lib.ptr(d_hidden), _lib.ptr(d_x, allow_none=True),
                                          _lib.ptr(d_params, allow_none=True), _lib.stream()))
        return d_x, d_params, None, None


class Encoding(nn.Module):
    """``tcnn.Encoding(n_input_dims, encoding_config, seed=1337, dtype=None)``."""

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None, device="cuda"):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.encoding_config = dict(encoding_config)
        otype = encoding_config["otype"].lower()
        if otype in ("hashgrid", "grid"):
            if n_input_dims != 3:
                raise ValueError("dns_slam_b200 HashGrid supports 3-D inputs")
            if encoding_config.get("type", "Hash") != "Hash" and otype == "grid":
                raise ValueError("dns_slam_b200 supports the Hash grid type only")
            self.otype = "hashgrid"
            self.tables = _grid.level_tables(int(encoding_config.get("n_levels", 16)),
                                             int(encoding_config.get("base_resolution", 16)),
                                             float(encoding_config.get("per_level_scale", 2.0)),
                                             int(encoding_config.get("log2_hashmap_size", 19)),
                                             int(encoding_config.get("n_features_per_level", 2)))
            if self.tables["n_features"] != 2 or self.tables["n_levels"] > _lib.MAX_LEVELS:
                raise ValueError("dns_slam_b200 HashGrid supports F=2 and at most 16 levels")
            self.gstruct = _grid.to_struct(self.tables)
            self.n_output_dims = self.tables["n_levels"] * 2
            g = torch.Generator().manual_seed(seed)
            init = (torch.rand(self.tables["n_entries"] * 2, generator=g) * 2 - 1) * 1e-4
            self.params = nn.Parameter(init.to(device))
        elif otype == "oneblob":
            self.otype = "oneblob"
            self.n_bins = int(encoding_config.get("n_bins", 16))
            self.n_output_dims = n_input_dims * self.n_bins
            self.params = nn.Parameter(torch.zeros(0, device=device))
        else:
            raise ValueError(f"dns_slam_b200.tcnn.Encoding: otype {encoding_config['otype']!r} is not on the "
                             "DNS-SLAM hot path (HashGrid and OneBlob are)")

    def forward(self, x):
        x = _f32c(x)
        if self.otype == "oneblob":
            return _OneBlobFn.apply(x, self.n_bins)
        return _HashGridFn.apply(x, self.params, self.gstruct, self.n_output_dims)


def xavier_uniform(rows, cols, gen):
    a = math.sqrt(6.0 / (rows + cols))
    return (torch.rand(rows, cols, generator=gen) * 2 - 1) * a


def init_network_params(n_in_pad, out_pad, width, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.cat([xavier_uniform(width, n_in_pad, g).reshape(-1), xavier_uniform(out_pad, width, g).reshape(-1)])


class Network(nn.Module):
    """``tcnn.Network(n_input_dims, n_output_dims, network_config, seed=1337)``."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337, device="cuda"):
        super().__init__()
        if int(network_config.get("n_hidden_layers", 1)) != 1 or int(network_config.get("n_neurons", 32)) != 32:
            raise ValueError("dns_slam_b200.tcnn.Network implements the reference's MLPs: 1 hidden layer of 32")
        if network_config.get("activation", "ReLU") != "ReLU" or network_config.get("output_activation", "None") != "None":
            raise ValueError("dns_slam_b200.tcnn.Network: ReLU hidden / linear output only")
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.network_config = dict(network_config)
        self.in_pad = _grid.next_multiple(n_input_dims, 16)
        self.out_pad = _grid.next_multiple(n_output_dims, 16)
        self.params = nn.Parameter(init_network_params(self.in_pad, self.out_pad, 32, seed).to(device))

    def forward(self, x):
        x = x.to(torch.float32)
        if self.in_pad != self.n_input_dims:  # tcnn pads the input with ones
            x = torch.nn.functional.pad(x, (0, self.in_pad - self.n_input_dims), value=1.0)
        return _MlpFn.apply(x.contiguous(), self.params, self.in_pad, self.n_output_dims)
