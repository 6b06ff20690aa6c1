"""The per-cell corner index arithmetic of the fused kernels (csrc/common.cuh: corner_indices8) restated in Python and
compared with the per-corner uint32 form of the oracle's hash-grid stand-in (oracle/tcnn_standin.py: corner_indices):
dense levels wrap modulo their size, cells with negative coordinates wrap modulo 2^32 first.  The CUDA code itself is
checked bit for bit on the GPU (tests/test_gpu_parity_configs.py::test_scannet_grid_indices_bit_exact); this test pins the
ALGEBRA the three cases of the helper rest on, for every dense resolution of the Replica / ScanNet encoders."""
import numpy as np

M = 0xFFFFFFFF


def per_corner(g, res, size):
    out = []
    for c in range(8):
        cx, cy, cz = (g[0] + (c & 1)) & M, (g[1] + ((c >> 1) & 1)) & M, (g[2] + (c >> 2)) & M
        i = (cx + ((cy * res) & M) + ((cz * ((res * res) & M)) & M)) & M
        out.append(i % size)
    return out


def per_cell(g, res, size):
    r2 = (res * res) & M
    base = (g[0] + ((g[1] * res) & M) + ((g[2] * r2) & M)) & M
    dmax = 1 + res + r2
    out = []
    if base <= M - dmax and dmax < size:
        r0 = base % size if base + dmax >= size else base     # ONE modulo per cell, none for a cell inside the level
        for c in range(8):
            v = r0 + (c & 1) + (res if c & 2 else 0) + (r2 if c & 4 else 0)
            out.append(v - size if v >= size else v)            # delta < size: a conditional subtraction suffices
    else:                                                       # base within delta of 2^32: per-corner form
        for c in range(8):
            v = (base + (c & 1) + (res if c & 2 else 0) + (r2 if c & 4 else 0)) & M
            out.append(v % size)
    return out


def test_per_cell_form_equals_per_corner_form():
    from oracle.tcnn_standin import grid_level_tables
    rng = np.random.default_rng(0)
    levels = set()
    for base_res, scale, log2_t in ((16, 1.2599, 16), (16, 1.1924, 20), (2, 1.5, 13)):
        t = grid_level_tables(16, base_res, scale, log2_t)
        levels |= {(int(r), int(n)) for r, n, h in zip(t["res"], t["size"], t["hashed"]) if not h}
    assert len(levels) >= 10
    for res, size in sorted(levels):
        cells = [rng.integers(-3 * res, 4 * res, 3) for _ in range(3000)]           # far outside the bound both ways
        cells += [rng.integers(-3, 3, 3) for _ in range(600)]                        # around the 2^32 wrap
        cells += [np.array(v) for v in ((0, 0, 0), (res - 1, res - 1, res - 1), (res, res, res), (-1, 0, 0), (0, -1, 0),
                                        (0, 0, -1), (-1, -1, -1), (res - 2, res - 1, res - 1))]
        for cell in cells:
            g = [int(v) & M for v in cell]
            assert per_corner(g, res, size) == per_cell(g, res, size), (res, size, cell)
