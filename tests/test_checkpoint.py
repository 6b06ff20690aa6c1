"""Checkpoint wire format (models/checkpoint.py:21-66, slams/mapping.py:1119-1145): round trip through this
package's writer, and a reference-shaped file (module state under ``decoder``, class experts as pickled
objects exposing ``params``, extra bookkeeping entries) read back into a fresh decoder."""
import torch

from dns_slam_b200 import checkpoint as ck
from dns_slam_b200 import decoder as D
from dns_slam_b200 import synthetic as syn


def _decoder(seed):
    return D.Decoder(syn.model_cfg("tiny"), syn.load_bound(syn.SHAPES["tiny"]["bound"]), n_class=5, seed=seed, device="cpu")


class _FakeTcnnNetwork(torch.nn.Module):          # what the reference pickles per class: a module with ``params``
    def __init__(self, vec):
        super().__init__()
        self.params = torch.nn.Parameter(vec.clone())


def test_round_trip(tmp_path):
    a, b = _decoder(1), _decoder(2)
    a.activate_expert(3)
    a.activate_expert(0)
    c = ck.Checkpoint(str(tmp_path), device="cpu", decoder=a)
    c.save("model.pt", scene="room0", idx=torch.tensor([7]), fine_decoders=a.fine_decoders,
           keyframe_list=[0, 5], estimate_c2w_list=torch.eye(4)[None].repeat(3, 1, 1))
    rest = ck.Checkpoint(str(tmp_path), device="cpu", decoder=b).load("model.pt")
    assert torch.equal(a.flat, b.flat)
    assert sorted(b.fine_decoders) == [0, 3]
    assert b.class_to_expert.tolist() == a.class_to_expert.tolist()
    assert rest["scene"] == "room0" and int(rest["idx"]) == 7 and rest["keyframe_list"] == [0, 5]
    assert set(rest["fine_decoders"]) == {0, 3}
    # what a reference-side consumer does with the entry (extract_mesh.py:146-157, eval_2d.py:143): module objects
    m = rest["fine_decoders"][3]
    assert torch.equal(m.state_dict()["params"], a.expert_params[3].detach()) and len(list(m.parameters())) == 1
    assert callable(m)


def test_activation_state_survives_without_the_kwarg(tmp_path):
    """save() without fine_decoders= and a plain state_dict round trip keep the activated experts."""
    a, b, c2 = _decoder(1), _decoder(2), _decoder(3)
    a.activate_expert(4)
    c = ck.Checkpoint(str(tmp_path), device="cpu", decoder=a)
    c.save("m.pt", idx=1)
    ck.Checkpoint(str(tmp_path), device="cpu", decoder=b).load("m.pt")
    assert sorted(b.fine_decoders) == [4] and int(b.class_to_expert[4]) == 4
    c2.load_state_dict(a.state_dict())
    assert sorted(c2.fine_decoders) == [4] and torch.equal(c2.flat, a.flat)
    # the module Parameters are still views of the flat buffer after a (no-op) .to()
    c2.to("cpu")
    with torch.no_grad():
        c2.flat.add_(1.0)
    assert torch.equal(c2.coarse_fn.decoder.params, c2.view("coarse"))


def test_reads_reference_shaped_file(tmp_path):
    src, dst = _decoder(3), _decoder(4)
    ref_keys = ["pe_fn.grid_fn.params", "coarse_fn.decoder.params", "out_fn.color_decoder.params",
                "out_fn.logit_decoder.params", "merge.decoder.params"]
    sd = src.state_dict()
    vec = torch.randn(D.EXPERT_PARAMS)
    torch.save({"decoder": {k: sd[k] for k in ref_keys}, "fine_decoders": {2: _FakeTcnnNetwork(vec)}, "idx": 11},
               str(tmp_path / "ref.pt"))
    before = dst.expert_params.detach().clone()
    rest = ck.Checkpoint(str(tmp_path), device="cpu", decoder=dst).load("ref.pt")
    for k in ref_keys:
        assert torch.equal(dst.state_dict()[k], sd[k])
    assert torch.equal(dst.expert_params[2], vec) and list(dst.fine_decoders) == [2]
    assert torch.equal(dst.expert_params[1], before[1])       # untouched rows keep their initialisation
    assert rest["idx"] == 11


def test_shape_mismatch_is_loud(tmp_path):
    dst = _decoder(5)
    torch.save({"decoder": {"coarse_fn.decoder.params": torch.zeros(7)}}, str(tmp_path / "bad.pt"))
    try:
        ck.Checkpoint(str(tmp_path), device="cpu", decoder=dst).load("bad.pt")
    except ValueError as e:
        assert "coarse_fn.decoder.params" in str(e)
    else:
        raise AssertionError("a wrong-sized entry must not load silently")
