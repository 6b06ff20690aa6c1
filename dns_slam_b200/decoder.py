"""Model surface: the duck-typed ``Decoder`` that ``slams/tracking.py`` / ``slams/mapping.py`` use.

Mirrors ``models/decoder.py`` attribute for attribute (so ``state_dict`` keys are the
reference's: ``pe_fn.grid_fn.params``, ``coarse_fn.decoder.params``,
``out_fn.{color,logit}_decoder.params``, ``merge.decoder.params``):

    decoder.pe_fn(x[P,3])              -> (pe[P,48], grid[P,32])      decoder.py:45-48
    decoder.coarse_fn(pe, features=g)  -> [P,33]                      decoder.py:93-94
    decoder.out_fn(pe, feat[P,64])     -> (rgb[P,3], logits[P,C])     decoder.py:122-125
    decoder.merge(p[R,P,3], o, f)      -> [P,32]                      decoder.py:67-77

All parameters live in ONE contiguous fp32 buffer ``decoder.flat`` laid out
``[hash table | coarse | colour | logit | merge | class experts]`` (the module Parameters are views
into it), so that the fused Adam step and the multi-GPU gradient all-reduce are single flat
operations.  The class-wise fine MLPs of ``Mapper.set_decoder``
(``slams/mapping.py:727-761``) are rows of a pre-allocated expert bank.
"""
import torch
from torch import nn

from . import tcnn

_MLP = {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None",
        "n_neurons": 32, "n_hidden_layers": 1}
EXPERT_PARAMS = 32 * 80 + 48 * 32  # 4096


class Pos_Encoding(nn.Module):
    def __init__(self, cfg, bound, seed=0, device="cuda"):
        super().__init__()
        self.pe_fn = tcnn.Encoding(3, {"otype": "OneBlob", "n_bins": cfg["pos"]["n_bins"]}, device=device)
        self.pe_dim = self.pe_fn.n_output_dims
        dim_max = (bound[:, 1] - bound[:, 0]).max()
        self.resolution = int(dim_max / cfg["grid"]["voxel_size"])
        from .grid import per_level_scale
        self.grid_fn = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2,
                                         "log2_hashmap_size": cfg["grid"]["hash_size"], "base_resolution": 16,
                                         "per_level_scale": per_level_scale(self.resolution)},
                                     seed=seed + 1, device=device)
        self.grid_dim = self.grid_fn.n_output_dims

    def forward(self, pts):
        return self.pe_fn(pts), self.grid_fn(pts)


class Merge(nn.Module):
    # True: the drop-in operator chain (OneBlob -> concat -> MLP -> mean) instead of the fused kernels; A/B hook of the
    # parity tests
    use_operator_chain = False

    def __init__(self, cfg, hidden_dim=32, feature_dim=64, bound=None, seed=0, device="cuda"):
        super().__init__()
        self.bound = bound
        self.pe_fn = tcnn.Encoding(3, {"otype": "OneBlob", "n_bins": cfg["pos"]["n_bins"]}, device=device)
        self.pe_dim = self.pe_fn.n_output_dims
        self.decoder = tcnn.Network(self.pe_dim + feature_dim, hidden_dim, _MLP, seed=seed + 5, device=device)

    def forward(self, p, o, features=None):
        n_refer, n_points, _ = features.shape
        if (features.is_cuda and features.shape[-1] == 64 and self.decoder.n_output_dims == 32
                and not self.use_operator_chain):
            # fused tcgen05 path: OneBlob + concat + MLP + mean over the views in one kernel each way
            from . import fused
            return fused.merge_fused(p, features, self.decoder.params, self.bound)
        p = (p - self.bound[:, 0]) / (self.bound[:, 1] - self.bound[:, 0])
        pe = self.pe_fn(p.flatten(0, 1))
        lat = self.decoder(torch.cat((pe, features.flatten(0, 1)), -1))
        return torch.mean(lat.reshape(n_refer, n_points, -1), 0)


class Coarse(nn.Module):
    def __init__(self, pts_dim, hidden_dim, feature_dim, seed=0, device="cuda"):
        super().__init__()
        self.decoder = tcnn.Network(pts_dim + feature_dim, hidden_dim + 1, _MLP, seed=seed + 2, device=device)

    def forward(self, pe, features=None):
        return self.decoder(torch.cat((pe, features), -1)).float()


class Out(nn.Module):
    def __init__(self, pts_dim, feature_dim, hidden_dim, n_class, seed=0, device="cuda"):
        super().__init__()
        self.color_decoder = tcnn.Network(pts_dim + feature_dim, 3, _MLP, seed=seed + 3, device=device)
        self.logit_decoder = tcnn.Network(pts_dim + feature_dim, n_class, _MLP, seed=seed + 4, device=device)
        self.sigmoid = nn.Sigmoid()

    def forward(self, pe, features):
        x = torch.cat((pe, features), -1)
        return self.sigmoid(self.color_decoder(x)), self.logit_decoder(x)


class Expert(nn.Module):
    """One class-wise fine MLP 80 -> 32 -> 33 (``slams/mapping.py:737-744``); ``params`` is a row
    of the decoder's expert bank."""

    def __init__(self, params_view):
        super().__init__()
        self.n_input_dims, self.n_output_dims = 80, 33
        self.params = nn.Parameter(params_view)

    def forward(self, x, features=None):
        """``fine_decoders[c](torch.cat((pe, grid), -1))`` (slams/mapping.py:600) or ``fine_decoders[c](pe,
        features=grid)`` (eval_2d.py:143)."""
        if features is not None:
            x = torch.cat((x, features), -1)
        return tcnn._MlpFn.apply(x.to(torch.float32).contiguous(), self.params, 80, 33)


class Decoder(nn.Module):
    def __init__(self, cfg, bound, n_class=40, seed=0, device="cuda", n_class_ids=None):
        super().__init__()
        bound = bound.to(device)
        self.bound = bound
        self.pe_fn = Pos_Encoding(cfg, bound, seed, device)
        self.pe_dim, self.grid_dim = self.pe_fn.pe_dim, self.pe_fn.grid_dim
        self.pts_dim, self.hidden_dim, self.pixel_dim = cfg["pts_dim"], cfg["hidden_dim"], cfg["pixel_dim"]
        if self.hidden_dim != 32 or self.pe_dim != 48 or self.grid_dim != 32:
            raise ValueError("dns_slam_b200 is built for hidden_dim=32, OneBlob 16 bins, 16x2 hash grid")
        self.n_class = n_class
        self.coarse_fn = Coarse(self.pe_dim, self.hidden_dim, self.grid_dim, seed, device)
        self.out_fn = Out(self.pe_dim, self.hidden_dim * 2, self.hidden_dim, n_class, seed, device)
        self.merge = Merge(cfg, self.hidden_dim, self.pixel_dim, bound, seed, device)
        # ---- one flat buffer; module Parameters become views into it
        self.n_class_ids = int(n_class_ids if n_class_ids is not None else n_class)
        named = [("table", self.pe_fn.grid_fn), ("coarse", self.coarse_fn.decoder),
                 ("color", self.out_fn.color_decoder), ("logit", self.out_fn.logit_decoder),
                 ("merge", self.merge.decoder)]
        self.layout, off = {}, 0
        for name, mod in named:
            n = mod.params.numel()
            self.layout[name] = (off, n)
            off += (n + 63) // 64 * 64
        self.layout["experts"] = (off, self.n_class_ids * EXPERT_PARAMS)
        off += self.n_class_ids * EXPERT_PARAMS
        flat = torch.zeros(off, device=device, dtype=torch.float32)
        for name, mod in named:
            a, n = self.layout[name]
            flat[a:a + n] = mod.params.detach()
            mod.params = nn.Parameter(flat[a:a + n])
        a, n = self.layout["experts"]
        for c in range(self.n_class_ids):
            flat[a + c * EXPERT_PARAMS:a + (c + 1) * EXPERT_PARAMS] = tcnn.init_network_params(
                80, 48, 32, seed + 100 + c).to(device)
        self.flat = flat
        self.expert_params = nn.Parameter(flat[a:a + n].view(self.n_class_ids, EXPERT_PARAMS))
        # class id -> expert row; -1 until the class is activated (mapping.py:736-749).  A persistent buffer: the
        # activation state travels with state_dict(), load_state_dict() rebuilds the expert modules from it.
        self.register_buffer("class_to_expert", torch.full((self.n_class_ids,), -1, dtype=torch.int32, device=device))
        self._experts = {}

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Reference-shaped entries (no ``class_to_expert`` / ``expert_params`` keys) load too; the experts named by the
        loaded ``class_to_expert`` are re-created (their modules are plain Python state)."""
        own = self.state_dict()
        merged = {k: state_dict.get(k, v) for k, v in own.items()}
        extra = [k for k in state_dict if k not in own]
        if strict and extra:
            raise KeyError(f"unexpected keys in the decoder state: {extra}")
        out = super().load_state_dict(merged, strict=True)
        self._experts = {}
        for c in torch.nonzero(self.class_to_expert >= 0).reshape(-1).tolist():
            self._experts[c] = Expert(self.expert_params.detach()[c])
        return out

    def _apply(self, fn, recurse=True):
        """``.to(device)`` / ``.cuda()``: the module Parameters must stay views of ONE flat buffer (fused Adam and the
        gradient all-reduce run over it), so the buffer is moved and the views are rebuilt; dtype changes are refused."""
        flat = fn(self.flat)
        if flat.dtype != torch.float32:
            raise RuntimeError("dns_slam_b200.Decoder keeps fp32 parameters (one flat buffer shared with the C ABI)")
        if flat.data_ptr() == self.flat.data_ptr():
            return self
        self.flat = flat.detach()
        for name, mod in (("table", self.pe_fn.grid_fn), ("coarse", self.coarse_fn.decoder), ("color", self.out_fn.color_decoder),
                          ("logit", self.out_fn.logit_decoder), ("merge", self.merge.decoder)):
            a, n = self.layout[name]
            mod.params = nn.Parameter(self.flat[a:a + n], requires_grad=mod.params.requires_grad)
        a, n = self.layout["experts"]
        self.expert_params = nn.Parameter(self.flat[a:a + n].view(self.n_class_ids, EXPERT_PARAMS))
        self.bound = fn(self.bound)
        self.merge.bound = self.bound
        self._buffers["class_to_expert"] = fn(self.class_to_expert)
        self._experts = {c: Expert(self.expert_params.detach()[c]) for c in self._experts}
        return self

    # ---- class-wise experts -------------------------------------------------------------
    def activate_expert(self, class_id):
        c = int(class_id)
        if c < 0 or c >= self.n_class_ids:
            raise ValueError("Unknown semantic class", c)
        if c not in self._experts:
            self._experts[c] = Expert(self.expert_params.detach()[c])
            self.class_to_expert[c] = c
        return self._experts[c]

    def copy_weights_from(self, other):
        """Weight hand-off mapper -> tracker (the reference deep-copies the shared decoder for every frame,
        slams/tracking.py:81, 296-302): ONE device copy of the flat buffer plus the expert table."""
        if other.flat.numel() != self.flat.numel():
            raise ValueError("decoders of different shape")
        with torch.no_grad():
            self.flat.copy_(other.flat)
            self.class_to_expert.copy_(other.class_to_expert)
        for c in other.fine_decoders:
            self.activate_expert(c)
        return self

    @property
    def fine_decoders(self):
        """``{class id: module}`` like ``Mapper.fine_decoders``."""
        return self._experts

    def view(self, name):
        a, n = self.layout[name]
        return self.flat[a:a + n]

    def expert_rows(self):
        """(offset, rows, row length) of the class-expert bank inside ``flat``: independent parameter tensors for Adam."""
        a, n = self.layout["experts"]
        return a, self.n_class_ids, EXPERT_PARAMS
