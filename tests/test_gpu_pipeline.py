"""End-to-end wiring on the GPU (examples/synthetic_slam.py): mapping window -> weight hand-off -> tracking -> key frames
by overlap -> checkpoint -> full-frame render, eager and CUDA-graph replayed.  Checks that everything stays finite,
that tracking lowers its loss, and that the checkpoint restores the decoder."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples"))


@pytest.mark.parametrize("use_graph", [False, True])
def test_synthetic_slam_runs_end_to_end(tmp_path, use_graph):
    import synthetic_slam
    from dns_slam_b200 import bench_util, checkpoint
    assert torch.cuda.is_available()
    out = synthetic_slam.run("tiny", n_frames=3, track_iters=6, map_iters=6, use_graph=use_graph, out_dir=str(tmp_path),
                             verbose=False)
    for kind, f, a, b, *_ in out["log"]:
        assert a == a and b == b, (kind, f, a, b)
        if kind == "track":
            assert b <= a + 1e-6                       # the best loss of the pose loop is no worse than its first
    color, depth, label = out["render"]
    assert torch.isfinite(color).all() and torch.isfinite(depth).all() and label.dtype == torch.int64
    dev = torch.device("cuda:0")
    fresh = bench_util.make_decoder("tiny", 5, dev, seed=77, all_experts=False)
    rest = checkpoint.Checkpoint(str(tmp_path), device=dev, decoder=fresh).load("model.pt")
    assert torch.equal(fresh.flat, out["decoder"].flat)
    assert sorted(fresh.fine_decoders) == sorted(out["decoder"].fine_decoders)
    assert rest["idx"] == 2 and len(rest["keyframe_list"]) >= 1
