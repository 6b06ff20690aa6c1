"""TEST INFRASTRUCTURE ONLY -- fp32 PyTorch stand-in for the ``tinycudann`` torch bindings.

PARITY UNPINNED: tiny-cuda-nn (NVlabs, ``bindings/torch``; reference ``requirements.txt:35``
has no commit pin, ``README.md:56-57`` suggests 91ee479d275d322a65726435040fc20b56b9c991) is
not installed here and its source is not on this machine.  This file restates the published
algorithm of the three operators the reference reaches:

* ``Encoding(otype="HashGrid")``  -- ``models/pos_encoding.py:31-46``
  (tiny-cuda-nn ``encodings/grid.h``: ``grid_scale``, ``grid_resolution``, ``grid_index``,
  ``grid_hash`` coherent-prime, ``pos_fract`` linear interpolation, params per level rounded
  up to 8 and capped at ``2**log2_hashmap_size``)
* ``Encoding(otype="OneBlob")``   -- ``models/pos_encoding.py:61-71``
  (``encodings/oneblob.h``: periodic quartic-kernel CDF differences)
* ``Network(otype="CutlassMLP")`` -- ``models/decoder.py:58-65,84-91,101-117``,
  ``slams/mapping.py:737-744`` (bias-free, ReLU hidden, inputs padded to x16 with ones,
  outputs padded to x16 and sliced; fp32 here although tcnn computes in fp16)

It can be installed as ``sys.modules['tinycudann']`` so that the reference's own
``models/decoder.py`` runs on CPU (see ``oracle/make_golden.py``).

Bit-exact contract for hash indices: ``pos = float32(float64(x) * float64(scale_l) + 0.5)``
(an ``fmaf`` emulated through double), ``g = uint32(int(floor(pos)))``; per-level
``scale / resolution / size / offset`` tables are computed ONCE on the host
(``grid_level_tables``) and the same numbers are handed to the CUDA kernels.
"""
import math

import numpy as np
import torch
from torch import nn

PRIME_Y = 2654435761
PRIME_Z = 805459861
_M32 = 0xFFFFFFFF


def next_multiple(v, m):
    return ((int(v) + m - 1) // m) * m


def grid_level_tables(n_levels, base_resolution, per_level_scale, log2_hashmap_size,
                      n_features=2):
    """Per-level (scale f32, resolution u32, size u32 entries, offset u32 entries, hashed)."""
    log2_pls = np.float32(np.log2(np.float32(per_level_scale)))
    scale = np.zeros(n_levels, np.float32)
    res = np.zeros(n_levels, np.int64)
    size = np.zeros(n_levels, np.int64)
    offset = np.zeros(n_levels + 1, np.int64)
    hashed = np.zeros(n_levels, np.int64)
    cap = 1 << int(log2_hashmap_size)
    max_params = 0xFFFFFFFF // 2
    for l in range(n_levels):
        s = np.float32(np.exp2(np.float32(l) * log2_pls)) * np.float32(base_resolution) - np.float32(1.0)
        s = np.float32(s)
        r = int(np.ceil(s)) + 1
        n = min(r ** 3, max_params)
        n = next_multiple(n, 8)
        n = min(n, cap)
        scale[l], res[l], size[l] = s, r, n
        hashed[l] = 1 if r ** 3 > n else 0
        offset[l + 1] = offset[l] + n
    return dict(scale=scale, res=res, size=size, offset=offset, hashed=hashed,
                n_levels=n_levels, n_features=n_features, n_entries=int(offset[-1]))


class _HashGrid(nn.Module):
    def __init__(self, cfg, seed=1337):
        super().__init__()
        self.L = int(cfg.get("n_levels", 16))
        self.F = int(cfg.get("n_features_per_level", 2))
        self.tables = grid_level_tables(self.L, int(cfg.get("base_resolution", 16)),
                                        float(cfg.get("per_level_scale", 2.0)),
                                        int(cfg.get("log2_hashmap_size", 19)), self.F)
        self.n_output_dims = self.L * self.F
        g = torch.Generator().manual_seed(seed)
        n = self.tables["n_entries"] * self.F
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * 1e-4)

    def corner_indices(self, x):
        """uint32 table indices [P, L, 8] (int64 storage), weights [P, L, 8], frac [P, L, 3]."""
        t = self.tables
        xd = x.to(torch.float32).double()
        idx_l, w_l = [], []
        for l in range(self.L):
            pos = (xd * float(t["scale"][l]) + 0.5).float()
            fl = torch.floor(pos)
            w = pos - fl
            g = fl.to(torch.int64) & _M32
            res, n = int(t["res"][l]), int(t["size"][l])
            ids, ws = [], []
            for c in range(8):
                b = [(c >> d) & 1 for d in range(3)]
                cx = (g[:, 0] + b[0]) & _M32
                cy = (g[:, 1] + b[1]) & _M32
                cz = (g[:, 2] + b[2]) & _M32
                if t["hashed"][l]:
                    i = cx ^ ((cy * PRIME_Y) & _M32) ^ ((cz * PRIME_Z) & _M32)
                else:
                    i = (cx + ((cy * res) & _M32) + ((cz * ((res * res) & _M32)) & _M32)) & _M32
                i = i % n + int(t["offset"][l])
                wt = torch.ones_like(w[:, 0])
                for d in range(3):
                    wt = wt * (w[:, d] if b[d] else (1.0 - w[:, d]))
                ids.append(i)
                ws.append(wt)
            idx_l.append(torch.stack(ids, -1))
            w_l.append(torch.stack(ws, -1))
        return torch.stack(idx_l, 1), torch.stack(w_l, 1)

    def forward(self, x):
        idx, w = self.corner_indices(x)                       # [P,L,8]
        tab = self.params.view(-1, self.F)
        vals = tab[idx.reshape(-1)].view(*idx.shape, self.F)  # [P,L,8,F]
        out = (w.unsqueeze(-1) * vals).sum(2)                 # [P,L,F]
        return out.reshape(x.shape[0], self.L * self.F)


def _quartic_cdf(u):
    u2 = u * u
    u4 = u2 * u2
    return torch.clamp((15.0 / 16.0) * u * (1.0 - (2.0 / 3.0) * u2 + (1.0 / 5.0) * u4) + 0.5, 0.0, 1.0)


class _OneBlob(nn.Module):
    def __init__(self, n_input_dims, cfg):
        super().__init__()
        self.n_bins = int(cfg.get("n_bins", 16))
        self.n_in = n_input_dims
        self.n_output_dims = n_input_dims * self.n_bins
        self.params = nn.Parameter(torch.zeros(0))

    def forward(self, x):
        x = x.to(torch.float32)
        nb = self.n_bins
        b = torch.arange(nb + 1, dtype=torch.float32, device=x.device) / nb   # boundaries
        d = b.view(1, 1, -1) - x.unsqueeze(-1)                                 # [P,D,nb+1]
        cdf = _quartic_cdf(d * nb) + _quartic_cdf((d - 1.0) * nb) + _quartic_cdf((d + 1.0) * nb)
        out = cdf[..., 1:] - cdf[..., :-1]
        return out.reshape(x.shape[0], self.n_in * nb)


class Encoding(nn.Module):
    """Signature of ``tcnn.Encoding`` as called at ``models/pos_encoding.py:34-45,63-70``."""

    def __init__(self, n_input_dims, encoding_config, dtype=torch.float, seed=1337):
        super().__init__()
        otype = encoding_config["otype"].lower()
        if otype in ("hashgrid", "grid"):
            self.impl = _HashGrid(encoding_config, seed)
        elif otype == "oneblob":
            self.impl = _OneBlob(n_input_dims, encoding_config)
        else:
            raise ValueError("oracle stand-in supports HashGrid and OneBlob only, got " + otype)
        self.n_input_dims = n_input_dims
        self.n_output_dims = self.impl.n_output_dims
        self.params = self.impl.params

    def forward(self, x):
        return self.impl(x)


def xavier_uniform_(w, gen):
    fan_out, fan_in = w.shape
    a = math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        w.copy_((torch.rand(w.shape, generator=gen) * 2 - 1) * a)
    return w


class Network(nn.Module):
    """Signature of ``tcnn.Network`` (``models/decoder.py:58-65``): one flat fp32 ``params``
    vector laid out ``[W1 (n_neurons x in_pad) | W2 (out_pad x n_neurons)]`` row-major."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        assert int(network_config.get("n_hidden_layers", 1)) == 1
        assert network_config.get("activation", "ReLU") == "ReLU"
        assert network_config.get("output_activation", "None") == "None"
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.width = int(network_config["n_neurons"])
        self.in_pad = next_multiple(n_input_dims, 16)
        self.out_pad = next_multiple(n_output_dims, 16)
        g = torch.Generator().manual_seed(seed)
        w1 = xavier_uniform_(torch.empty(self.width, self.in_pad), g)
        w2 = xavier_uniform_(torch.empty(self.out_pad, self.width), g)
        self.params = nn.Parameter(torch.cat([w1.reshape(-1), w2.reshape(-1)]))

    def weights(self):
        n1 = self.width * self.in_pad
        return (self.params[:n1].view(self.width, self.in_pad),
                self.params[n1:].view(self.out_pad, self.width))

    def forward(self, x):
        x = x.to(torch.float32)
        if self.in_pad != self.n_input_dims:
            x = torch.nn.functional.pad(x, (0, self.in_pad - self.n_input_dims), value=1.0)
        w1, w2 = self.weights()
        h = torch.relu(x @ w1.t())
        return (h @ w2.t())[:, :self.n_output_dims]
