"""``FusedAdam`` / ``dns_adam_multi`` (the optimiser groups of slams/tracking.py:119-124 and slams/mapping.py:464-466)
against ``torch.optim.Adam`` on the same gradients: several groups with their own learning rates, a flat-buffer group,
eager and as CUDA-graph replays (device-side step counter).  Tolerance 1e-6 absolute on parameters of order one."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    return torch.device("cuda:0")


def _setup(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    flat = torch.randn(5000, generator=g).to(dev)
    a = torch.nn.Parameter(flat[:3000].view(30, 100))
    b = torch.nn.Parameter(flat[3000:])
    q = torch.randn(4, generator=g).to(dev).requires_grad_(True)
    t = torch.randn(3, generator=g).to(dev).requires_grad_(True)
    return flat, a, b, q, t


def _loss(a, b, q, t, x):
    return ((a * x[0]).sin().sum() + (b * b).sum() * x[1] + (q * q).sum() * (t * x[2]).cos().sum())


def test_fused_adam_matches_torch():
    from dns_slam_b200 import fused
    dev = _dev()
    flat1, a1, b1, q1, t1 = _setup(dev)
    flat2, a2, b2, q2, t2 = _setup(dev)
    ours = fused.FusedAdam([{"params": [a1, b1], "lr": 5e-3, "flat": flat1}, {"params": [q1], "lr": 1e-3},
                            {"params": [t1], "lr": 2e-4}])
    ref = torch.optim.Adam([{"params": [a2, b2], "lr": 5e-3}, {"params": [q2], "lr": 1e-3},
                            {"params": [t2], "lr": 2e-4}])
    assert ours.n_segs == 3                      # the flat group is ONE segment
    xs = torch.rand(7, 3, generator=torch.Generator().manual_seed(1)).to(dev)
    for x in xs:
        ours.zero_grad(); _loss(a1, b1, q1, t1, x).backward(); ours.step()
        ref.zero_grad(); _loss(a2, b2, q2, t2, x).backward(); ref.step()
    for u, v in ((flat1, flat2), (q1, q2), (t1, t2)):
        torch.testing.assert_close(u.detach(), v.detach(), rtol=1e-5, atol=1e-6)
    assert int(ours.step_dev) == 7


def test_fused_adam_graph_replay_advances_the_step():
    from dns_slam_b200 import fused
    dev = _dev()
    flat1, a1, b1, q1, t1 = _setup(dev, 3)
    flat2, a2, b2, q2, t2 = _setup(dev, 3)
    ours = fused.FusedAdam([{"params": [a1, b1], "lr": 5e-3}, {"params": [q1, t1], "lr": 1e-3}])   # per-tensor segments
    ref = torch.optim.Adam([{"params": [a2, b2], "lr": 5e-3}, {"params": [q2, t2], "lr": 1e-3}])
    assert ours.n_segs == 4
    x = torch.rand(3, device=dev)

    def one():
        ours.zero_grad(set_to_none=True)
        _loss(a1, b1, q1, t1, x).backward()
        ours.step()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            one()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one()
    for _ in range(4):
        graph.replay()
    n = 2 + 4               # capture only records
    for _ in range(n):
        ref.zero_grad(); _loss(a2, b2, q2, t2, x).backward(); ref.step()
    assert int(ours.step_dev) == n
    for u, v in ((flat1, flat2), (q1, q2), (t1, t2)):
        torch.testing.assert_close(u.detach(), v.detach(), rtol=1e-5, atol=1e-6)


def test_expert_rows_follow_torch_per_tensor_semantics():
    """the class experts are independent parameter tensors in the reference (slams/mapping.py:445-446); an
    expert whose class is absent from an iteration has ``.grad = None`` there and torch.optim.Adam skips it (no moment
    decay, no step count).  ``dns_adam_seg.row_len`` reproduces that per row of the expert bank."""
    from dns_slam_b200 import fused
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    n_rows, row_len, head = 5, 64, 100
    flat = torch.randn(head + n_rows * row_len, generator=g).to(dev)
    grad = torch.zeros_like(flat)
    ref_head = flat[:head].clone().requires_grad_(True)
    ref_rows = [flat[head + r * row_len:head + (r + 1) * row_len].clone().requires_grad_(True) for r in range(n_rows)]
    ref = torch.optim.Adam([ref_head] + ref_rows, lr=5e-3)
    ours = fused.AdamSegments(fused.split_expert_rows(flat, grad, 5e-3, (head, n_rows, row_len)))
    present = [[0, 1, 2, 3, 4], [0, 2], [], [1, 2, 4], [3], [0, 1, 2, 3, 4], [2]]      # classes seen per iteration
    for it, rows in enumerate(present):
        grad.zero_()
        gh = torch.randn(head, generator=g).to(dev)
        grad[:head] = gh
        ref_head.grad = gh.clone()
        for r in range(n_rows):
            if r in rows:
                gr = (torch.randn(row_len, generator=g) * (10.0 if it % 2 else 0.1)).to(dev)
                grad[head + r * row_len:head + (r + 1) * row_len] = gr
                ref_rows[r].grad = gr.clone()
            else:
                ref_rows[r].grad = None
        ref.step()
        ours.step()
    torch.testing.assert_close(flat[:head], ref_head.detach(), rtol=1e-5, atol=1e-6)
    for r in range(n_rows):
        torch.testing.assert_close(flat[head + r * row_len:head + (r + 1) * row_len], ref_rows[r].detach(), rtol=1e-5, atol=1e-6,
                                   msg=lambda m, r=r: f"expert row {r}: {m}")
