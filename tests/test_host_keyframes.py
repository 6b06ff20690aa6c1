"""Key-frame selection by overlap (slams/mapping.py:171-236): the batched device formulation against the numpy loop of
the reference (oracle restatement), and the decoder weight hand-off.  Pure torch: runs without a GPU."""
import numpy as np
import torch

from dns_slam_b200 import decoder as D
from dns_slam_b200 import slam
from dns_slam_b200 import synthetic as syn
from oracle import reference_path as rp


def test_keyframe_overlap_matches_numpy_loop():
    cam = syn.camera("tiny")
    poses = syn.trajectory("tiny", 9)
    gen = torch.Generator().manual_seed(3)
    fr = syn.frame("tiny", poses[4], gen, n_class=4)
    kfs = torch.stack([poses[i] for i in (0, 2, 3, 5, 8)], 0)
    tape = rp.DrawTape(seed=11)
    want = rp.keyframe_overlap(cam, fr["depth"], poses[4], kfs, tape, pixels=100)
    idx = tape.items[0][1]
    got = slam.keyframe_overlap(cam, fr["depth"], poses[4], kfs, idx)
    assert np.abs(got.numpy().astype(np.float64) - want).max() <= 1.0 / 1600 + 1e-9     # at most one probe point apart
    sel = slam.keyframe_selection_overlap(cam, fr["depth"], poses[4], kfs, 3, idx)
    ref_order = [i for i in sorted(range(len(want)), key=lambda i: want[i], reverse=True) if want[i] > 0][:3]
    assert len(sel) == len(ref_order) and all(abs(want[a] - want[b]) <= 1.0 / 1600 + 1e-9 for a, b in zip(sel, ref_order))
    perm = [2, 0, 1, 4, 3]
    sel_p = slam.keyframe_selection_overlap(cam, fr["depth"], poses[4], kfs, 2, idx, perm=perm)
    assert len(sel_p) <= 2 and set(sel_p) <= set(range(5))


def test_decoder_weight_hand_off():
    mk = lambda seed: D.Decoder(syn.model_cfg("tiny"), syn.load_bound(syn.SHAPES["tiny"]["bound"]), n_class=5, seed=seed,  # noqa: E731
                                device="cpu")
    a, b = mk(1), mk(2)
    a.activate_expert(4)
    b.copy_weights_from(a)
    assert torch.equal(a.flat, b.flat) and list(b.fine_decoders) == [4]
    assert torch.equal(b.coarse_fn.decoder.params, a.coarse_fn.decoder.params)       # module views follow the flat buffer
