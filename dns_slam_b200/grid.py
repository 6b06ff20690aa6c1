"""Host-side geometry of the multi-resolution hash grid.

Computes, once and in fp32 as tiny-cuda-nn does (``grid_scale`` / ``grid_resolution`` of
``encodings/grid.h``; reached from the reference through ``models/pos_encoding.py:31-46``),
the per-level ``scale``, ``resolution``, table ``size`` and ``offset`` and packs them into the
``dns_grid`` struct of the C ABI.  The numbers are produced on the host so that the CUDA
kernels and any CPU checker use the very same tables (bit-exact hash indices).
"""
import numpy as np

from . import _lib


def next_multiple(v, m):
    return ((int(v) + m - 1) // m) * m


def level_tables(n_levels=16, base_resolution=16, per_level_scale=2.0, log2_hashmap_size=19,
                 n_features=2):
    log2_pls = np.float32(np.log2(np.float32(per_level_scale)))
    cap = 1 << int(log2_hashmap_size)
    max_params = 0xFFFFFFFF // 2
    scale, res, size, hashed, offset = [], [], [], [], [0]
    for l in range(n_levels):
        s = np.float32(np.float32(np.exp2(np.float32(l) * log2_pls)) * np.float32(base_resolution)
                       - np.float32(1.0))
        r = int(np.ceil(s)) + 1
        n = min(next_multiple(min(r ** 3, max_params), 8), cap)
        scale.append(float(s))
        res.append(r)
        size.append(n)
        hashed.append(1 if r ** 3 > n else 0)
        offset.append(offset[-1] + n)
    return dict(n_levels=n_levels, n_features=n_features, scale=scale, res=res, size=size,
                offset=offset, hashed=hashed, n_entries=offset[-1])


def per_level_scale(desired_resolution, base_resolution=16, n_levels=16):
    """models/pos_encoding.py:33."""
    return float(np.exp2(np.log2(desired_resolution / base_resolution) / (n_levels - 1)))


def to_struct(t):
    g = _lib.Grid()
    g.n_levels, g.n_features = t["n_levels"], t["n_features"]
    for l in range(t["n_levels"]):
        g.scale[l] = t["scale"][l]
        g.res[l] = t["res"][l]
        g.size[l] = t["size"][l]
        g.hashed[l] = t["hashed"][l]
    for l in range(t["n_levels"] + 1):
        g.offset[l] = t["offset"][l]
    return g
