"""``slam.uniq_class_indices`` (the sampler of Mapper.decoder_init) against the golden vectors produced by the reference's
own ``utils/common.py:364-403`` (get_samples_by_uniq_class; oracle/make_golden.py uniq_class_case): rays spread over a
GIVEN class list -- class 0 of the list takes the remainder, a class with one pixel is repeated without a draw, an
absent class is skipped.  Bit exact (index work)."""
import os

import torch

from dns_slam_b200 import slam, synthetic as syn


def test_uniq_class_indices_match_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "uniq_class_tiny.pt"), weights_only=False)
    meta = g["meta"]
    gen = torch.Generator().manual_seed(meta["seed"])
    fr = syn.frame(meta["shape"], syn.trajectory(meta["shape"], 4)[meta["pose_index"]], gen, n_class=meta["n_class"])
    label = fr["label"].clone()
    r, c, v = meta["single"]
    label[r, c] = v
    tables = slam.class_tables(label)
    for case in g["cases"]:
        draws = [t for kind, t in case["tape"] if kind == "randint"]
        idx, used = slam.uniq_class_indices(tables, case["n"], case["class_list"], draws)
        assert used == len(draws)
        assert torch.equal(idx, case["indices"]), case["class_list"]
        assert torch.equal(label.reshape(-1)[idx], case["labels"])
