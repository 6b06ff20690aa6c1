"""Host logic of the frames + draws boundary (dns_slam_b200.step.FrameBatchPlan): slot bookkeeping, the class-balanced
draw of utils/common.py:307-330 split into (offset, slot_base) and resolved as order[slot_base + offset] (what the
sampling kernel does), the draw tape in the reference's call order, and the rank slices of a sharded batch."""
import torch

from dns_slam_b200 import slam, step, synthetic as syn
from oracle import reference_path as rp


def _tables(n_frames=2, n_class=6, seed=3):
    gen = torch.Generator().manual_seed(seed)
    poses = syn.trajectory("tiny", 4)
    frames = [syn.frame("tiny", poses[i], gen, n_class=n_class) for i in range(n_frames)]
    # one class with a single pixel: consumes no draw (common.py:324-325)
    frames[0]["label"][0, 0] = n_class + 3
    return frames, [slam.class_tables(f["label"]) for f in frames]


def test_plan_resolves_to_oracle_indices():
    frames, tables = _tables()
    cam = syn.camera("tiny")
    bound = syn.load_bound(syn.SHAPES["tiny"]["bound"])
    plan = step.FrameBatchPlan(tables, 96, 15, (0, cam["H"], 0, cam["W"]), bound, 8)
    buf, tape = plan.make_host_draws(torch.Generator().manual_seed(1), pinned=False, return_tape=True)
    t = rp.DrawTape(tape)
    for f, fr in enumerate(frames):
        n_f = 96 // 2
        idx1 = rp.uniform_indices(0, cam["H"], 0, cam["W"], n_f // 3 * 2, t)
        idx2 = rp.class_balanced_indices(fr["label"], n_f // 3, t)
        t.rand((15,)), t.rand((15,))
        idx = plan.draw_view(buf, f"idx{f}", torch.int64)
        n_u = plan.slices[f][0][1]
        assert torch.equal(idx[:n_u], idx1)
        mine = tables[f][1][plan.slot_base[f].long() + idx[n_u:]]
        assert torch.equal(mine, idx2), f
    t.rand((3,)), t.rand((1, 1, 1, 3))
    assert t.pos == len(tape)


def test_rank_slices_tile_the_global_slots():
    _, tables = _tables()
    cam = syn.camera("tiny")
    bound = syn.load_bound(syn.SHAPES["tiny"]["bound"])
    win = (0, cam["H"], 0, cam["W"])
    whole = step.FrameBatchPlan(tables, 100, 15, win, bound, 8)
    parts = [step.FrameBatchPlan(tables, 100, 15, win, bound, 8, rank=r, world=3) for r in range(3)]
    assert sum(p.n_local for p in parts) == whole.n_total == whole.n_local
    assert [p.ray_offset for p in parts] == [0, parts[0].n_local, parts[0].n_local + parts[1].n_local]
    for f in range(2):
        assert torch.equal(torch.cat([p.slot_base[f] for p in parts]), whole.slot_base[f])
        us = [p.slices[f][0] for p in parts]
        assert us[0][0] == 0 and us[-1][1] == whole.n_u and all(us[i][1] == us[i + 1][0] for i in range(2))


def test_pack_draws_equals_the_generated_buffer():
    """``FrameBatchPlan.pack_draws`` (the draw dictionaries of slam.map_optimize -> the step's byte buffer) must produce the
    very bytes ``make_host_draws`` writes for the same raw draws: uniform indices, per-class offsets (a single-pixel class
    consumes no draw and gets offset 0), the forced 0.5 of common.py:572-573, the float64 TV offsets."""
    frames, tables = _tables()
    cam = syn.camera("tiny")
    bound = syn.load_bound(syn.SHAPES["tiny"]["bound"])
    plan = step.FrameBatchPlan(tables, 96, 15, (0, cam["H"], 0, cam["W"]), bound, 8)
    buf, tape = plan.make_host_draws(torch.Generator().manual_seed(7), pinned=False, return_tape=True)
    # rebuild the dictionaries from the tape: per frame [uniform, one randint per class with > 1 pixel, rand, rand], then TV
    it = iter(tape)
    per = []
    for f in range(plan.F):
        u = next(it)[1]
        cd = [next(it)[1] for _, _, m, count, _ in plan.class_slot_ranges(f) if count != 1]
        ts, tz = next(it)[1], next(it)[1]
        per.append(dict(idx_uniform=u, class_draws=cd, t_surface=ts, t_zero=tz))
    tv = (next(it)[1], next(it)[1])
    packed = plan.pack_draws(per, tv)
    assert torch.equal(packed, buf)
    # wrong sizes are refused instead of silently shifting slots
    bad = [dict(d) for d in per]
    bad[0]["idx_uniform"] = bad[0]["idx_uniform"][:-1]
    try:
        plan.pack_draws(bad, tv)
        raise AssertionError("short uniform draw accepted")
    except ValueError:
        pass
