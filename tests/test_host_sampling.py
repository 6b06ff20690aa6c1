"""Host-side sampling logic that needs no GPU: the vectorised class-balanced draw against the per-class
loop of utils/common.py:315-330 (a class with a single pixel consumes no draw and is repeated)."""
import torch

from dns_slam_b200 import slam


def _loop_form(tab, n, draws):
    counts, starts = tab.counts_h, tab.starts_h
    order = tab[1]
    nc = len(counts)
    nk = n // nc
    out, di = [], 0
    for c in range(nc):
        m = n - nk * (nc - 1) if c == 0 else nk
        if counts[c] == 1:
            out.append(order[starts[c]].reshape(1).repeat(m))
        else:
            out.append(order[starts[c] + draws[di]])
            di += 1
    return torch.cat(out), di


def test_class_balanced_indices_matches_loop():
    g = torch.Generator().manual_seed(0)
    for single in (False, True):
        lab = torch.randint(0, 6, (40, 50), generator=g)
        if single:
            lab[3, 3] = 17                     # exactly one pixel of class 17
        tab = slam.class_tables(lab)
        for n in (100, 37):
            counts = tab.counts_h
            nc = len(counts)
            nk = n // nc
            draws = []
            for c in range(nc):
                m = n - nk * (nc - 1) if c == 0 else nk
                if counts[c] != 1:
                    draws.append(torch.randint(counts[c], (m,), generator=g))
            want, used_want = _loop_form(tab, n, draws)
            got, used = slam.class_balanced_indices(tab, n, draws)
            assert used == used_want
            assert torch.equal(got, want)


def test_quad2rotation_closed_form_backward():
    """The custom backward of quad2rotation (utils/common.py:406-429) against autograd of the formula."""
    torch.manual_seed(0)
    q = torch.randn(5, 4, requires_grad=True)
    G = torch.randn(5, 3, 3)
    R1 = slam.quad2rotation(q)
    (R1 * G).sum().backward()
    g1, q.grad = q.grad.clone(), None
    R2 = slam._quad2rotation_formula(q)
    (R2 * G).sum().backward()
    assert torch.equal(R1, R2)
    assert torch.allclose(g1, q.grad, rtol=1e-5, atol=1e-5)


def test_vectorised_quad2rotation_is_bit_identical():
    """The 9-vector evaluation used by the autograd wrapper against the element-wise formula of common.py:406-429."""
    torch.manual_seed(1)
    for _ in range(300):
        q = torch.randn(4, 4) * (torch.rand(1) * 3 + 0.05)
        assert torch.equal(slam._quad2rotation_vectorised(q), slam._quad2rotation_formula(q))
