"""Checkpoint reader / writer with the reference's wire format.

``Checkpoint`` mirrors ``models/checkpoint.py:5-66``: modules registered by keyword are stored as their
``state_dict()`` under that keyword (the reference registers ``decoder=...``, ``slams/dns_slam.py:51-52``),
extra keyword arguments of ``save`` are stored verbatim (``slams/mapping.py:1119-1127``: scene, idx,
fine_decoders, keyframe_dict, keyframe_list, estimate_c2w_list, gt_c2w_list), ``load`` copies the keys
both sides know and returns the rest.

The decoder of this package has the reference's parameter names (``pe_fn.grid_fn.params``,
``coarse_fn.decoder.params``, ``out_fn.color_decoder.params``, ``out_fn.logit_decoder.params``,
``merge.decoder.params``: one flat fp32 ``params`` vector per tcnn module, weights row-major
``[out][in]`` with the output width padded to 16), so a reference ``decoder`` entry loads as is.
``fine_decoders``: the reference pickles ``{class id: tcnn.Network}`` MODULE objects and its consumers CALL them
(``extract_mesh.py:146-157``, ``eval_2d.py:385-386,143``).  The writer therefore stores ``{class id: Expert}`` --
stand-alone ``dns_slam_b200.decoder.Expert`` modules (own copy of the 4096 weights, callable as
``m(torch.cat((pe, grid), -1))`` and ``m(pe, features=grid)``, with ``parameters()`` / ``state_dict()['params']``), which
a reference-side reader can use as is once this package is importable; ``as_modules=False`` stores plain
``{"params": tensor}`` dicts.  The reader accepts every form (a module with ``state_dict()``, a ``params`` attribute, a
dict with ``params`` or a bare tensor).  When ``save`` gets no ``fine_decoders`` argument it takes them from the registered
decoder, so the activation state is never lost.  Parity of the tcnn layout itself is unpinned (no tinycudann in this
build).
"""
import os

import torch

from .decoder import EXPERT_PARAMS


class Checkpoint:
    def __init__(self, checkpoint_dir="./chkpts", device=None, **kwargs):
        self.module_dict = kwargs
        self.device = device
        self.checkpoint_dir = checkpoint_dir
        os.makedirs(checkpoint_dir, exist_ok=True)

    def _path(self, filename):
        return filename if os.path.isabs(filename) else os.path.join(self.checkpoint_dir, filename)

    def save(self, filename, as_modules=True, **kwargs):
        """checkpoint.py:21-35."""
        out = dict(kwargs)
        if "fine_decoders" not in out:
            for v in self.module_dict.values():
                if hasattr(v, "fine_decoders"):
                    out["fine_decoders"] = v.fine_decoders
        if "fine_decoders" in out:
            out["fine_decoders"] = fine_decoders_state(out["fine_decoders"], as_modules)
        for k, v in self.module_dict.items():
            out[k] = v.state_dict()
        torch.save(out, self._path(filename))

    def load(self, filename, verbose=False):
        """checkpoint.py:37-66: keys present on both sides are copied; returns the non-module entries."""
        state = torch.load(self._path(filename), map_location=self.device, weights_only=False)
        for k, mod in self.module_dict.items():
            if k not in state:
                print(f'Warning: Could not find "{k}" in checkpoint!')
                continue
            own = dict(mod.state_dict())
            for kk, vv in state[k].items():
                if kk in own:
                    if tuple(own[kk].shape) != tuple(vv.shape):
                        raise ValueError(f"checkpoint entry {k}.{kk} has shape {tuple(vv.shape)}, "
                                         f"the module expects {tuple(own[kk].shape)}")
                    own[kk] = vv
                    if verbose:
                        print(kk)
            mod.load_state_dict(own)
            if "fine_decoders" in state and hasattr(mod, "activate_expert"):
                load_fine_decoders(mod, state["fine_decoders"])
        return {k: v for k, v in state.items() if k not in self.module_dict}


def _expert_vector(obj):
    if isinstance(obj, torch.Tensor):
        return obj
    if isinstance(obj, dict):
        return obj["params"]
    if hasattr(obj, "state_dict"):
        return obj.state_dict()["params"]
    return obj.params


def fine_decoders_state(fine_decoders, as_modules=True):
    """{class id: module | state | tensor} -> {class id: stand-alone Expert module on the CPU} (or, with
    ``as_modules=False``, {class id: {"params": fp32 vector}})."""
    from .decoder import Expert
    out = {}
    for c, m in fine_decoders.items():
        vec = _expert_vector(m).detach().float().cpu().clone()
        out[int(c)] = Expert(vec) if as_modules else {"params": vec}
    return out


def load_fine_decoders(decoder, fine_decoders):
    """Creates the class experts named in the entry (mapping.py:727-761) and fills their weights."""
    for c, m in fine_decoders.items():
        vec = _expert_vector(m).detach().float().reshape(-1)
        if vec.numel() != EXPERT_PARAMS:
            raise ValueError(f"class expert {c}: {vec.numel()} parameters, expected {EXPERT_PARAMS}")
        decoder.activate_expert(int(c))
        with torch.no_grad():
            decoder.expert_params[int(c)].copy_(vec.to(decoder.expert_params.device))
    return decoder.fine_decoders
