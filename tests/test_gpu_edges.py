"""Edge cases of the fused call (through the C ABI): ragged / minimal / maximal shapes, ray chunking
inside one call, missing class experts, empty input, zero-depth rays, a fully masked tracking batch."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import close, rel_err  # noqa: E402


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _oracle_mapping(dec, samples, lam, sigma=0.05):
    """Reference-style composition on the CUDA operator kernels + torch autograd (independent of the fused path)."""
    from oracle import reference_path as rp
    smp = dict(samples)
    smp["pts"] = smp["rays_o"][:, None, :] + smp["rays_d"][:, None, :] * smp["z_vals"][:, :, None]
    pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(dec, dec.fine_decoders, dec.bound, smp)
    p, d, l, lt, fs, op = rp.mapping_losses(smp, pc, pd, pl, fine, coarse, sigma)
    total = lam["p"] * p + lam["d"] * d + lam["l"] * l + lam["lt"] * lt + lam["fs"] * fs + lam["op"] * op
    return total, (pc, pd, pv, pl)


LAM = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)


@pytest.mark.parametrize("N,S,C", [(1, 47, 3), (3, 5, 1), (130, 96, 101), (257, 33, 7), (64, 256, 4)])
def test_ragged_shapes_vs_operator_composition(N, S, C):
    from dns_slam_b200 import _lib, bench_util, fused
    dev = _dev()
    dec, samples = bench_util.synthetic_batch("tiny", "map", N, S, C, dev, seed=N + S, n_frames=1)
    samples = {k: v for k, v in samples.items() if k != "mask"}
    ld, preds = fused.render_and_loss(dec, samples, _lib.MODE_MAP, lambdas=LAM)
    dec.zero_grad()
    ld["total"].backward()
    g_table = dec.pe_fn.grid_fn.params.grad.clone()
    g_exp = dec.expert_params.grad.clone()
    dec.zero_grad()
    total, (pc, pd, pv, pl) = _oracle_mapping(dec, samples, LAM)
    total.backward()
    close(ld["total"], total, rtol=1e-3, atol=1e-6, name="total")
    close(preds["color"], pc, rtol=1e-3, atol=1e-5, name="color")
    close(preds["depth"], pd, rtol=1e-3, atol=1e-5, name="depth")
    close(preds["logits"], pl, rtol=1e-3, atol=1e-4, name="logits")
    assert rel_err(g_table, dec.pe_fn.grid_fn.params.grad) < 2e-3
    ge = torch.stack([dec.fine_decoders[c].params.grad if dec.fine_decoders[c].params.grad is not None
                      else torch.zeros(4096, device=dev) for c in range(C)])
    assert rel_err(g_exp, ge) < 2e-3


def test_chunked_call_equals_single_chunk():
    """A workspace that only fits a fraction of the rays makes dns_render_fwd_bwd process ray chunks internally
    (global denominators, global class rule): same losses and gradients as the one-chunk call."""
    import ctypes as C
    from dns_slam_b200 import _lib, bench_util, fused, step as stepmod
    dev = _dev()
    N, S, Cn = 1000, 47, 9
    dec, samples = bench_util.synthetic_batch("tiny", "map", N, S, Cn, dev, seed=4, n_frames=2)
    samples = {k: v for k, v in samples.items() if k != "mask"}
    ms = stepmod.MappingStep(dec, 5e-3, LAM)
    full = ms.forward_backward(samples)
    g_full, l_full = ms.grad.clone(), full[0].clone()
    need = _lib.lib().dns_render_workspace_bytes(_lib.MODE_MAP, N, S, Cn, Cn)
    orig = fused.workspace
    small = torch.empty(int(need * 0.3), dtype=torch.uint8, device=dev)
    fused.workspace = lambda nbytes, device: small
    try:
        part = ms.forward_backward(samples)
    finally:
        fused.workspace = orig
    close(part[0][:7], l_full[:7], rtol=1e-4, atol=1e-7, name="chunked losses")
    assert rel_err(ms.grad, g_full) < 1e-3
    close(part[1]["color"], full[1]["color"], rtol=1e-4, atol=1e-6, name="chunked colour")


def test_missing_expert_raises_value_error():
    from dns_slam_b200 import _lib, bench_util, fused
    dev = _dev()
    dec = bench_util.make_decoder("tiny", 6, dev, all_experts=False)
    for c in (0, 1, 2):
        dec.activate_expert(c)
    _, samples = bench_util.synthetic_batch("tiny", "map", 64, 47, 6, dev, seed=1, n_frames=1, dec=dec)
    with pytest.raises(ValueError):
        fused.render_and_loss(dec, samples, _lib.MODE_MAP, lambdas=LAM, strict=True)


def test_empty_and_oversized_inputs_fail_loudly():
    from dns_slam_b200 import _lib, bench_util, fused
    dev = _dev()
    dec, samples = bench_util.synthetic_batch("tiny", "map", 8, 47, 4, dev, seed=2, n_frames=1)
    empty = {k: v[:0].contiguous() for k, v in samples.items()}
    with pytest.raises(RuntimeError):
        fused.render_and_loss(dec, empty, _lib.MODE_MAP, lambdas=LAM)
    big = dict(samples)
    big["z_vals"] = torch.sort(torch.rand(8, 300, device=dev), -1)[0] + 0.1
    big["features"] = torch.zeros(8, 300, 32, device=dev)
    with pytest.raises(RuntimeError):
        fused.render_and_loss(dec, big, _lib.MODE_MAP, lambdas=LAM)


def test_tracking_all_masked_is_nan_like_reference():
    """mean over an empty selection is NaN in the reference (tracking.py:85-96); the fused call must not crash."""
    from dns_slam_b200 import _lib, bench_util, fused
    dev = _dev()
    dec, samples = bench_util.synthetic_batch("tiny", "track", 32, 47, 4, dev, seed=3)
    samples["mask"] = torch.zeros(32, dtype=torch.bool, device=dev)
    ld, _ = fused.render_and_loss(dec, samples, _lib.MODE_TRACK, freeze_decoder=True)
    assert not bool(torch.isfinite(ld["total"]))
    samples["mask"][5] = True
    ld, _ = fused.render_and_loss(dec, samples, _lib.MODE_TRACK, freeze_decoder=True)
    assert bool(torch.isfinite(ld["total"]))


@pytest.mark.parametrize("N,S,C,mode", [(1, 47, 3, "map"), (131, 47, 40, "map"), (257, 33, 7, "map"), (100, 96, 5, "track"),
                                        (37, 13, 4, "track")])
def test_workspace_guard_bands_stay_intact(N, S, C, mode, monkeypatch):
    """The kernels may only write inside [workspace, workspace + dns_render_workspace_bytes): the fused call runs on a
    buffer of exactly that size cut out of a larger allocation whose borders carry a byte pattern."""
    from dns_slam_b200 import _lib, bench_util, fused
    dev = _dev()
    dec, samples = bench_util.synthetic_batch("tiny", mode, N, S, C, dev, seed=3 * N + S, n_frames=1)
    if mode == "map":
        samples = {k: v for k, v in samples.items() if k != "mask"}
    guard = 1 << 16
    state = {}

    def guarded(nbytes, device):
        nbytes = int(nbytes)
        big = torch.full((nbytes + 2 * guard + 512,), 0xCD, dtype=torch.uint8, device=device)
        off = guard + (-(big.data_ptr() + guard)) % 256          # keep the 256-byte alignment of the carve
        state["big"], state["off"], state["n"] = big, off, nbytes
        return big[off:off + nbytes]

    monkeypatch.setattr(fused, "workspace", guarded)
    m = _lib.MODE_MAP if mode == "map" else _lib.MODE_TRACK
    ld, _ = fused.render_and_loss(dec, samples, m, lambdas=LAM if mode == "map" else dict(p=5.0, d=5.0, l=0.1))
    ld["total"].backward()
    torch.cuda.synchronize()
    big, off, n = state["big"], state["off"], state["n"]
    assert bool((big[:off] == 0xCD).all()), "write below the workspace"
    assert bool((big[off + n:] == 0xCD).all()), "write above the workspace"
    assert torch.isfinite(ld["total"]).item()
