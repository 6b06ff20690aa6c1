"""Autograd-free optimisation steps over the decoder's flat parameter buffer.

``MappingStep`` is the native body of one mapping iteration's core (``slams/mapping.py:888-910``
without sampling): zero the flat gradient, ONE fused render + loss + backward call that scatters
every gradient straight into the flat buffer, an optional NCCL all-reduce of that buffer when the
ray batch is sharded across GPUs, and ONE fused Adam kernel over all parameters
(``torch.optim.Adam`` defaults; a fresh state per ``optimize()`` call as ``mapping.py:438-468``).
``TrackingStep`` is the pose-only counterpart (``slams/tracking.py:313-340``).
"""
import torch

from . import _lib, fused
from .decoder import EXPERT_PARAMS


class MappingStep:
    def __init__(self, decoder, lr, lambdas=None, opacity_sigma=0.05, process_group=None, world_size=1):
        self.dec = decoder
        self.lr = lr
        self.lambdas = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
        self.lambdas.update(lambdas or {})
        self.opacity_sigma = opacity_sigma
        self.pg, self.world = process_group, world_size
        self.grad = torch.zeros_like(decoder.flat)
        self.reset()

    def reset(self):
        """Fresh Adam state (the reference builds a new optimiser for every optimize() call)."""
        self.adam = None
        self.t = 0

    def _adam(self):
        if self.adam is None:     # class experts are independent tensors: rows without gradient are skipped (fused.FusedAdam)
            self.adam = fused.AdamSegments(fused.split_expert_rows(self.dec.flat, self.grad, self.lr, self.dec.expert_rows()))
        self.adam.step()

    def _views(self, buf):
        lay = self.dec.layout
        out = {k: buf[lay[k][0]:lay[k][0] + lay[k][1]] for k in ("table", "coarse", "color", "logit")}
        a, n = lay["experts"]
        out["experts"] = buf[a:a + n].view(-1, EXPERT_PARAMS)
        return out

    def _config(self, samples):
        dec = self.dec
        return fused.RenderConfig(_lib.MODE_MAP, dec.bound, dec.pe_fn.grid_fn.gstruct, samples["z_vals"],
                                  samples["gt_color"], samples["gt_depth"], samples["gt_label"], None,
                                  dec.class_to_expert, dec.n_class, self.lambdas, opacity_trunc=self.opacity_sigma)

    def forward_backward(self, samples, need_drays=True, need_dfeat=True, cfg=None):
        """Fills ``self.grad`` (flat) and returns (losses[8], preds, d_rays_o, d_rays_d, d_features)."""
        dec = self.dec
        self.grad.zero_()
        cfg = cfg or self._config(samples)
        p = self._views(dec.flat)
        return fused.render_raw(cfg, p["table"], p["coarse"], p["color"], p["logit"], p["experts"],
                                samples["rays_o"], samples["rays_d"], samples.get("features"),
                                self._views(self.grad), need_drays, need_dfeat)

    def step(self, samples, need_drays=True, need_dfeat=True):
        out = self.forward_backward(samples, need_drays, need_dfeat)
        if self.world > 1:
            torch.distributed.all_reduce(self.grad, group=self.pg)
        self.t += 1
        self._adam()
        return out


def shard_bounds(n_total, world, rank):
    """Contiguous, balanced ray ranges: rank r owns [lo, hi)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedMappingStep(MappingStep):
    """One mapping batch sharded by rays over the ranks of ``process_group``.

    Every rank holds the full decoder (replicated) and ITS slice of the ray batch.  Per step:
    (1) labels of the whole batch are all-gathered once (the class rule ``class(p) = label[p mod N]``
    of mapping.py:612-613 reaches across shards) and the four batch-global counts are all-reduced
    (tiny); (2) each rank runs the fused kernels on its rays with GLOBAL denominators, so partial
    losses and gradients SUM to the single-GPU values; (3) ONE all-reduce of the flat gradient
    buffer (+ 8 loss partials appended to it); (4) identical fused Adam on every rank -- no
    parameter broadcast.  ``comm`` is any object with all_reduce_sum(tensor) / all_gather(tensor)
    (torch.distributed over NCCL on GPUs; the CPU tests plug gloo)."""

    def __init__(self, decoder, lr, comm, rank, world, **kw):
        super().__init__(decoder, lr, **kw)
        self.comm, self.rank, self.world_n = comm, rank, world
        self.packed = torch.zeros(self.grad.numel() + 8, device=self.grad.device)
        self.grad = self.packed[:-8]                 # gradients and loss partials travel together

    def step_sharded(self, local_samples, n_total, need_drays=True, need_dfeat=True):
        lo, hi = shard_bounds(n_total, self.world_n, self.rank)
        assert local_samples["z_vals"].shape[0] == hi - lo, "local batch does not match the shard bounds"
        sizes = [b - a for a, b in (shard_bounds(n_total, self.world_n, r) for r in range(self.world_n))]
        labels_all = self.comm.all_gather(local_samples["gt_label"].contiguous(), sizes)
        cfg = self._config(local_samples)
        counts = self._local_counts(cfg)
        self.comm.all_reduce_sum(counts)
        cfg.shard(n_total, lo, labels_all, counts)
        out = self.forward_backward(local_samples, need_drays, need_dfeat, cfg=cfg)
        self.packed[-8:] = out[0]
        self.comm.all_reduce_sum(self.packed)
        losses = self.packed[-8:].clone()
        losses[7] = losses[7] / self.world_n          # n_valid is a batch constant, not a partial sum
        self.t += 1
        self._adam()
        return (losses,) + tuple(out[1:])

    # the two device calls besides forward_backward (overridden by the CPU host-logic tests)
    def _local_counts(self, cfg):
        return fused.render_counts(cfg)



class TorchComm:
    """torch.distributed adapter (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        self.group = group

    def all_reduce_sum(self, t):
        torch.distributed.all_reduce(t, group=self.group)
        return t

    def all_reduce_max(self, t):
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=self.group)
        return t

    def all_gather(self, t, sizes=None):
        """Concatenation over ranks of possibly different-length 1-D tensors.  ``sizes`` (per-rank lengths known to the
        caller, e.g. from shard_bounds) skips the size exchange and its device->host read."""
        world = torch.distributed.get_world_size(self.group)
        if sizes is not None and len(set(sizes)) == 1:
            out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
            torch.distributed.all_gather_into_tensor(out, t, group=self.group)
            return out
        if sizes is None:
            n = torch.tensor([t.numel()], device=t.device)
            sizes = [torch.zeros_like(n) for _ in range(world)]
            torch.distributed.all_gather(sizes, n, group=self.group)
        mx = int(max(int(s) for s in sizes))
        pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
        pad[:t.numel()] = t
        outs = [torch.zeros_like(pad) for _ in range(world)]
        torch.distributed.all_gather(outs, pad, group=self.group)
        return torch.cat([o[:int(s)] for o, s in zip(outs, sizes)])


class TrackingStep:
    """Fused tracking forward/backward: gradients w.r.t. the rays only (the decoder is frozen in
    tracking), to be chained to quaternion / translation by the caller."""

    def __init__(self, decoder, lambdas=None):
        self.dec = decoder
        self.lambdas = dict(p=5.0, d=5.0, l=0.1)
        self.lambdas.update(lambdas or {})

    def forward_backward(self, samples, need_dfeat=True):
        dec = self.dec
        cfg = fused.RenderConfig(_lib.MODE_TRACK, dec.bound, dec.pe_fn.grid_fn.gstruct, samples["z_vals"],
                                 samples["gt_color"], samples["gt_depth"], samples["gt_label"], samples.get("mask"),
                                 None, dec.n_class, self.lambdas)
        return fused.render_raw(cfg, dec.view("table"), dec.view("coarse"), dec.view("color"), dec.view("logit"), None,
                                samples["rays_o"], samples["rays_d"], samples.get("features"), None, True, need_dfeat)


class HostBatchPipeline:
    """Feeds ray batches that live in PINNED HOST memory to a step function, double buffered: the H2D copy
    of batch k+1 runs on a copy stream while the kernels of batch k run on the compute stream, and the loss
    vector of every step is read back into pinned memory.  Used for the end-to-end number of bench.py (every
    step's inputs really cross PCIe inside the timed region) and for hosts that sample on the CPU."""

    def __init__(self, template_host_batch, device):
        self.dev = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.bufs = [{k: torch.empty_like(v, device=device) for k, v in template_host_batch.items()} for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]      # H2D of buffer i finished
        self.free = [torch.cuda.Event() for _ in range(2)]       # compute on buffer i finished
        self.result_host = torch.empty(8, pin_memory=True)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in template_host_batch.values())
        for e in self.free:
            e.record(torch.cuda.current_stream(device))

    def _upload(self, host_batch, i):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[i])
            for k, v in host_batch.items():
                self.bufs[i][k].copy_(v, non_blocking=True)
            self.ready[i].record(self.copy_stream)

    def run(self, host_batches, step_fn):
        """host_batches: sequence of dicts of pinned host tensors (may repeat the same object);
        step_fn(device_batch) -> tuple whose first element is the loss vector [8]."""
        cur = torch.cuda.current_stream(self.dev)
        n = len(host_batches)
        if n == 0:
            return self.result_host
        self._upload(host_batches[0], 0)
        for k in range(n):
            i = k & 1
            if k + 1 < n:
                self._upload(host_batches[k + 1], (k + 1) & 1)
            cur.wait_event(self.ready[i])
            out = step_fn(self.bufs[i])
            self.free[i].record(cur)
            self.result_host.copy_(out[0], non_blocking=True)
        return self.result_host


class FrameBatchPlan:
    """Ray bookkeeping and draw layout of a mapping batch over F target frames (host logic, device agnostic).

    Per frame the GLOBAL slot list is ``[n_u uniform | n_c class balanced]`` with ``n_f = n_rays // F``,
    ``n_u = n_f // 3 * 2``, ``n_c = n_f // 3`` (slams/mapping.py:504-508); rank r owns slice r of both parts.  The
    iteration's draws travel as ONE byte buffer: per frame ``idx`` int64 [n_local] (uniform window indices, then the
    class-balanced offsets into the pixels of each slot's class), ``t_surface`` / ``t_zero`` f32 [n_surface]
    (common.py:572-573 already applied), then the TV ``offset | jitter`` f64 [6] (mapping.py:133-140)."""

    def __init__(self, class_tables, n_rays, n_surface, window, bound, smooth_pts, rank=0, world=1):
        self.tables, self.nf, self.window, self.smooth_pts = class_tables, n_surface, window, smooth_pts
        self.bound = bound.detach().double().cpu()      # host copy, made ONCE: the TV offsets are host arithmetic per iteration
        self.rank, self.world = rank, world
        F = self.F = len(class_tables)
        n_f = n_rays // F
        n_u, n_c = n_f // 3 * 2, n_f // 3
        self.n_u, self.n_c = n_u, n_c
        self.n_total = F * (n_u + n_c)
        self.slices, self.ray_start = [], [0]
        per_rank = [0] * world
        for f in range(F):
            for r in range(world):
                (a, b), (c, d) = shard_bounds(n_u, world, r), shard_bounds(n_c, world, r)
                per_rank[r] += (b - a) + (d - c)
                if r == rank:
                    self.slices.append(((a, b), (c, d)))
                    self.ray_start.append(self.ray_start[-1] + (b - a) + (d - c))
        self.n_local = self.ray_start[-1]
        self.rank_sizes = per_rank
        self.ray_offset = sum(per_rank[:rank])
        # class-balanced slots: first pixel (in the label-sorted order) of the class each slot draws from
        self.slot_base = []
        for f in range(F):
            _, (c, d) = self.slices[f]
            base = []
            for _, _, m, _, start in self.class_slot_ranges(f):
                base += [start] * m
            self.slot_base.append(torch.tensor(base[c:d], dtype=torch.int32))
        self.draw_layout, off = {}, 0
        for f in range(F):
            n = self.ray_start[f + 1] - self.ray_start[f]
            for name, nbytes in ((f"idx{f}", 8 * n), (f"ts{f}", 4 * n_surface), (f"tz{f}", 4 * n_surface)):
                self.draw_layout[name] = (off, nbytes)
                off = (off + nbytes + 15) & ~15
        self.draw_layout["tv"] = (off, 48)
        self.draw_bytes = off + 48

    def _slot_ranges(self, f):
        """``class_slot_ranges`` cached per frame (the tables do not change during a mapping call)."""
        cache = self.__dict__.setdefault("_slot_range_cache", {})
        if f not in cache:
            cache[f] = self.class_slot_ranges(f)
        return cache[f]

    def class_slot_ranges(self, f):
        """[(class position, first slot, slots, pixels of the class, first pixel)] of frame f's GLOBAL class-balanced slot
        list (common.py:315-330: class 0 of the sorted list takes the remainder)."""
        tab = self.tables[f]
        counts, starts = tab.counts_h, tab.starts_h
        n_class = len(counts)
        n_k = self.n_c // n_class
        out, s0 = [], 0
        for c in range(n_class):
            m = self.n_c - n_k * (n_class - 1) if c == 0 else n_k
            out.append((c, s0, m, counts[c], starts[c]))
            s0 += m
        return out

    def draw_view(self, buf, name, dtype):
        off, n = self.draw_layout[name]
        return buf[off:off + n].view(dtype)

    def make_host_draws(self, gen, pinned=True, return_tape=False, shared_gen=None):
        """Seeded host-side draws of one iteration in the reference's order: per frame one uniform
        ``randint`` over the window, one ``randint`` per class with more than one pixel, ``rand(n_surface)`` twice; then
        the two TV draws.  With ``return_tape`` also the raw draws as an oracle ``DrawTape`` item list (the reference's
        call order; meaningful for a single rank).  ``shared_gen``: generator for the draws that belong to the whole
        batch rather than to a ray (the per-frame surface offsets and the TV lattice): every rank of a sharded batch must
        pass one with the SAME seed, ``gen`` then only feeds the rank's own pixel draws."""
        sg = shared_gen if shared_gen is not None else gen
        from . import fused as _fused
        buf = torch.zeros(self.draw_bytes, dtype=torch.uint8)
        if pinned:
            buf = buf.pin_memory()
        tape = []
        H0, H1, W0, W1 = self.window
        n_win = (H1 - H0) * (W1 - W0)
        for f in range(self.F):
            (a, b), (c, d) = self.slices[f]
            idx = self.draw_view(buf, f"idx{f}", torch.int64)
            u = torch.randint(n_win, (b - a,), generator=gen)
            idx[:b - a] = u
            tape.append(("randint", u))
            for _, s0, m, count, _ in self.class_slot_ranges(f):
                lo, hi = max(s0, c), min(s0 + m, d)       # this rank's part of the class's slots
                if count == 1:                            # a class with one pixel is repeated without a draw
                    continue
                if hi <= lo:
                    if m == 0:                            # the reference still issues the (empty) randint
                        tape.append(("randint", torch.zeros(0, dtype=torch.int64)))
                    continue
                dr = torch.randint(count, (hi - lo,), generator=gen)
                idx[(b - a) + lo - c:(b - a) + hi - c] = dr
                tape.append(("randint", dr))
            ts = torch.rand(self.nf, generator=sg)
            tz = torch.rand(self.nf, generator=sg)
            tape += [("rand", ts.clone()), ("rand", tz.clone())]
            if not bool((ts == 0.5).any()):
                ts[self.nf // 2 + 1] = 0.5                                # common.py:572-573
            self.draw_view(buf, f"ts{f}", torch.float32).copy_(ts)
            self.draw_view(buf, f"tz{f}", torch.float32).copy_(tz)
        r3, r113 = torch.rand(3, generator=sg), torch.rand(1, 1, 1, 3, generator=sg)
        tape += [("rand", r3), ("rand", r113)]
        off, jit = _fused.tv_offsets(self.bound, self.smooth_pts, r3, r113)
        self.draw_view(buf, "tv", torch.float64).copy_(torch.cat((off, jit)))
        return (buf, tape) if return_tape else buf

    def pack_draws(self, per_frame, tv_draws, buf=None):
        """The draw buffer of one iteration from the draw DICTIONARIES of ``slam.map_optimize`` (``per_frame[f]`` =
        dict(idx_uniform, class_draws=[one randint per class with more than one pixel], t_surface, t_zero),
        ``tv_draws`` = (rand(3), rand(1,1,1,3))); single rank.  ``buf``: a (pinned) byte buffer to fill."""
        from . import fused as _fused
        if self.world != 1:
            raise ValueError("pack_draws packs the draws of an unsharded batch")
        if buf is None:
            buf = torch.zeros(self.draw_bytes, dtype=torch.uint8)
        for f in range(self.F):
            d = per_frame[f]
            idx = self.draw_view(buf, f"idx{f}", torch.int64)
            u = d["idx_uniform"].reshape(-1)
            if u.numel() != self.n_u:
                raise ValueError(f"frame {f}: {u.numel()} uniform draws, the plan has {self.n_u} slots")
            pieces, k = [u], 0          # ONE concatenation + ONE copy per frame (a slice assignment per class was 0.5 ms / iteration)
            for _, s0, m, count, _ in self._slot_ranges(f):
                if count == 1:                  # repeated without a draw (common.py:321-323): offset 0
                    pieces.append(torch.zeros(m, dtype=torch.int64))
                    continue
                dr = d["class_draws"][k].reshape(-1)
                k += 1
                if dr.numel() != m:
                    raise ValueError(f"frame {f}: class draw of {dr.numel()} values for {m} slots")
                pieces.append(dr)
            idx.copy_(torch.cat(pieces))
            ts = d["t_surface"].detach().cpu().to(torch.float32).clone()
            if not bool((ts == 0.5).any()):
                ts[self.nf // 2 + 1] = 0.5                                # common.py:572-573
            self.draw_view(buf, f"ts{f}", torch.float32).copy_(ts)
            self.draw_view(buf, f"tz{f}", torch.float32).copy_(d["t_zero"].detach().cpu().to(torch.float32))
        off, jit = _fused.tv_offsets(self.bound, self.smooth_pts, tv_draws[0], tv_draws[1])
        self.draw_view(buf, "tv", torch.float64).copy_(torch.cat((off, jit)))
        return buf


class MappingFrameStep:
    """One mapping iteration (``slams/mapping.py:884-910``) from what a SLAM host actually holds: key frames and their
    feature maps resident on the device, camera poses, and the iteration's hoisted random draws (~8 B per ray).  No
    autograd and no library kernels in the loop (torch only zeroes the gradient buffer and copies the 9-float loss
    vector):

        dns_pose_prepare -> dns_sample_rays per target frame (uniform + class-balanced draws resolved in the kernel)
        -> dns_featmerge_fwd (feature matching + Merge + truncation mask, band samples only)
        -> dns_render_fwd_bwd (all six losses, every gradient into ONE flat buffer) -> dns_featmerge_bwd
        -> dns_tv_fwd_bwd (Mapper.smoothness) -> dns_pose_grad -> [NCCL all-reduce of the flat buffer] -> dns_adam_multi
        over decoder + quaternions + translations (frame 0 frozen, mapping.py:457).

    Ray sharding: rank r draws slice r of every frame's uniform and class-balanced slots; the global batch is
    the rank-major concatenation (a fixed permutation of the reference's frame-major order: the draws are i.i.d.).
    Batch-global scalars are exchanged first (max depth per frame, band / depth counts, labels for the class rule
    ``class(p) = label[p mod N]``); losses and gradients of the ranks SUM to the single-GPU values.

    ``draws``: one byte buffer (``draw_layout``): per frame ``idx`` int64 [n_f] (uniform window indices, then the
    class-balanced offsets), ``t_surface`` / ``t_zero`` f32 [n_surface] each (common.py:572-573 already applied), and the
    TV ``offset | jitter`` f64 [6] (mapping.py:133-140).  ``upload`` copies a pinned host image of it in ONE H2D copy.
    """

    def __init__(self, dec, cam, frames, class_tables, feats_cl, est_c2w, refer_idx, refer_c2w, kf_idx, n_rays,
                 n_samples_ray=32, n_surface_ray=15, lr=5e-3, BA_cam_lr=5e-4, is_BA=True, lambdas=None, opacity_sigma=0.05,
                 smooth_pts=64, lambda_sm=1e-5, with_tv=True, comm=None, rank=0, world=1):
        from . import slam
        self.dec, self.cam = dec, cam
        dev = self.dev = dec.bound.device
        self.frames, self.tables, self.feats = frames, class_tables, [f.contiguous() for f in feats_cl]
        F = self.F = len(frames)
        self.ns, self.nf = n_samples_ray, n_surface_ray
        self.S = n_samples_ray + n_surface_ray
        self.lambdas = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
        self.lambdas.update(lambdas or {})
        self.opacity_sigma, self.smooth_pts, self.lambda_sm, self.with_tv = opacity_sigma, smooth_pts, lambda_sm, with_tv
        self.comm, self.rank, self.world = comm, rank, world
        H, W = cam["H"], cam["W"]
        self.window = (0, H, 0, W)
        plan = self.plan = FrameBatchPlan(class_tables, n_rays, n_surface_ray, self.window, dec.bound, smooth_pts, rank, world)
        self.slices, self.ray_start, self.n_local, self.n_total = plan.slices, plan.ray_start, plan.n_local, plan.n_total
        self.rank_sizes, self.ray_offset = plan.rank_sizes, plan.ray_offset
        self.slot_base = [b.to(dev) for b in plan.slot_base]
        N, S = self.n_local, self.S
        # ---- poses: quaternion / translation device parameters; reference views
        self.quats = torch.stack([slam.quad_from_matrix(c[:3, :3]) for c in est_c2w], 0).to(dev).contiguous()
        self.trans = torch.stack([c[:3, 3].detach().clone().float() for c in est_c2w], 0).to(dev).contiguous()
        n_ref = {len(x) for x in refer_idx}
        if len(n_ref) != 1:
            raise ValueError("MappingFrameStep needs the same number of reference views for every target frame")
        self.R = n_ref.pop()
        src, fixed = [], []
        for i in range(F):
            for k, rid in enumerate(refer_idx[i]):
                if rid == -1:
                    src.append(i)
                elif rid in kf_idx:
                    src.append(kf_idx.index(rid))
                else:
                    src.append(-1)
                c = refer_c2w[i][k] if src[-1] == -1 else None      # views that follow a target frame need no fixed pose
                fixed.append(c.detach().to(dev).float() if c is not None else torch.eye(4, device=dev))
        fixed = torch.stack(fixed, 0)
        self.view_src = torch.tensor(src, dtype=torch.int32, device=dev)
        self.fixed_w2c = fused.rigid_inverse(fixed).contiguous()
        self.fixed_cam_o = fixed[:, :3, 3].contiguous()
        V = F * self.R
        self.R_all = torch.empty(F, 3, 3, device=dev)
        self.w2c, self.cam_o = torch.empty(V, 4, 4, device=dev), torch.empty(V, 3, device=dev)
        # ---- batch buffers
        f32 = torch.float32
        self.batch = dict(gt_color=torch.empty(N, 3, device=dev), gt_depth=torch.empty(N, device=dev),
                          gt_label=torch.empty(N, dtype=torch.int64, device=dev), rays_o=torch.empty(N, 3, device=dev),
                          rays_d=torch.empty(N, 3, device=dev), z_vals=torch.empty(N, S, device=dev),
                          inside=torch.empty(N, dtype=torch.uint8, device=dev),
                          pixel=torch.empty(N, dtype=torch.int64, device=dev))
        self.scratch = torch.zeros(F, 2, device=dev)          # per frame: max depth, rays outside the bound
        self.features = torch.zeros(N, S, 32, device=dev)      # band rows are rewritten every step, the others never read
        self.fm_ws = torch.empty(int(_lib.lib().dns_featmerge_workspace_bytes(N, S)), dtype=torch.uint8, device=dev)
        self.fm_stash = fused.featmerge_stash(N, S, self.R, dev)      # forward operand tiles kept for the backward
        self.t_lin = torch.linspace(0.0, 1.0, steps=n_samples_ray).to(dev)
        # ---- draws: ONE device buffer, refreshed with one H2D copy
        self.draw_bytes = plan.draw_bytes
        self.draws_dev = torch.zeros(self.draw_bytes, dtype=torch.uint8, device=dev)
        # ---- gradients: [decoder flat | d_quats | d_trans | losses 8 | smooth 1 | pad], one all-reduce
        nflat = dec.flat.numel()
        self.packed = torch.zeros(nflat + 7 * F + 12, device=dev)
        self.grad = self.packed[:nflat]
        self.d_quats = self.packed[nflat:nflat + 4 * F].view(F, 4)
        self.d_trans = self.packed[nflat + 4 * F:nflat + 7 * F].view(F, 3)
        self.loss_vec = self.packed[nflat + 7 * F:nflat + 7 * F + 9]
        self.pose_scratch = torch.empty(12 * F, device=dev)
        self.d_features = torch.empty(N, S, 32, device=dev)
        segs = fused.split_expert_rows(dec.flat, self.grad, lr, dec.expert_rows())
        self.opt_poses = bool(is_BA)
        if self.opt_poses:
            f0 = 0 if F == 1 else 1      # the oldest target frame stays fixed (mapping.py:457)
            segs += [(self.quats[f0:].view(-1), self.d_quats[f0:].view(-1), BA_cam_lr),
                     (self.trans[f0:].view(-1), self.d_trans[f0:].view(-1), BA_cam_lr)]
        self.adam = fused.AdamSegments(segs)
        # result vector: 9 losses | per frame (max depth, rays outside) | over ALL steps so far: rays outside, min n_valid
        self.result_host = torch.empty(11 + 2 * F, pin_memory=True)
        self.result_dev = torch.zeros(11 + 2 * F, device=dev)

    def draw_view(self, buf, name, dtype):
        return self.plan.draw_view(buf, name, dtype)

    def make_host_draws(self, gen, pinned=True, return_tape=False, shared_gen=None):
        return self.plan.make_host_draws(gen, pinned, return_tape, shared_gen)

    def upload(self, host_buf, stream=None):
        """ONE host-to-device copy of an iteration's draws (pinned ``host_buf`` -> asynchronous)."""
        if stream is None:
            self.draws_dev.copy_(host_buf, non_blocking=True)
        else:
            with torch.cuda.stream(stream):
                self.draws_dev.copy_(host_buf, non_blocking=True)

    # ------------------------------------------------------------------ stages
    def _sample(self, phase, draws):
        b = self.batch
        pend = []          # all target frames in one pair of launches (dns_sample_rays_batch)
        for f in range(self.F):
            r0, r1 = self.ray_start[f], self.ray_start[f + 1]
            if r1 == r0:
                continue
            (a, bb), _ = self.slices[f]
            out = {k: v[r0:r1] for k, v in b.items()}
            out["scratch"] = self.scratch[f]
            fused.sample_rays(self.cam, self.dec.bound, self.frames[f], self.draw_view(draws, f"idx{f}", torch.int64),
                              self.window, self.R_all[f], self.trans[f], self.ns, self.nf,
                              self.draw_view(draws, f"ts{f}", torch.float32), self.draw_view(draws, f"tz{f}", torch.float32),
                              t_lin=self.t_lin, class_order=self.tables[f][1], slot_base=self.slot_base[f],
                              n_direct=bb - a, out=out, phase=phase, defer=pend)
        fused.sample_rays_flush(pend)

    def _views(self):
        return fused.Views(self.w2c, self.cam_o, self.feats, self.ray_start)

    def _flat_views(self, buf):
        lay = self.dec.layout
        out = {k: buf[lay[k][0]:lay[k][0] + lay[k][1]] for k in ("table", "coarse", "color", "logit", "merge")}
        a, n = lay["experts"]
        out["experts"] = buf[a:a + n].view(-1, EXPERT_PARAMS)
        return out

    def step(self, draws=None):
        """One iteration on the draws in ``self.draws_dev`` (or the given device byte buffer).  Returns the device vector
        [p, d, l, lt, fs, op, total incl. smoothness, n_valid (< 0: error flag), smooth | per frame: max depth, rays outside |
        rays outside and min n_valid over ALL steps of this object]."""
        draws = self.draws_dev if draws is None else draws
        dec, b, F = self.dec, self.batch, self.F
        L = _lib.lib()
        f32 = torch.float32
        _lib.check(L.dns_pose_prepare(_lib.ptr(self.quats, f32), _lib.ptr(self.trans, f32), F, _lib.ptr(self.view_src, torch.int32),
                                      _lib.ptr(self.fixed_w2c, f32), _lib.ptr(self.fixed_cam_o, f32), F * self.R,
                                      _lib.ptr(self.R_all), _lib.ptr(self.w2c), _lib.ptr(self.cam_o), _lib.stream()))
        if self.world > 1:     # a frame's rays are spread over the ranks: its max depth is a batch-global scalar
            self._sample(1, draws)
            self.comm.all_reduce_max(self.scratch)
            self._sample(2, draws)
        else:
            self._sample(0, draws)
        self.packed.zero_()
        views = self._views()
        merge_p = dec.view("merge")
        fused.featmerge_raw(self.cam, dec.merge.bound, views, b["rays_o"], b["rays_d"], b["z_vals"], b["gt_depth"], merge_p,
                            True, ws=self.fm_ws, out=self.features, stash=self.fm_stash, zero_fill=fused._use_simt)
        cfg = fused.RenderConfig(_lib.MODE_MAP, dec.bound, dec.pe_fn.grid_fn.gstruct, b["z_vals"], b["gt_color"], b["gt_depth"],
                                 b["gt_label"], None, dec.class_to_expert, dec.n_class, self.lambdas,
                                 opacity_trunc=self.opacity_sigma)
        if self.world > 1:
            labels_all = self.comm.all_gather(b["gt_label"], self.rank_sizes)
            counts = fused.render_counts(cfg)
            self.comm.all_reduce_sum(counts)
            cfg.shard(self.n_total, self.ray_offset, labels_all, counts)
        p, g = self._flat_views(dec.flat), self._flat_views(self.grad)
        losses, _, d_o, d_d, d_f = fused.render_raw(cfg, p["table"], p["coarse"], p["color"], p["logit"], p["experts"],
                                                    b["rays_o"], b["rays_d"], self.features, g, True, True,
                                                    features_band_only=True)
        fused.featmerge_bwd_raw(self.cam, dec.merge.bound, views, b["rays_o"], b["rays_d"], b["z_vals"], b["gt_depth"], merge_p,
                                d_f, self.fm_ws, g["merge"], d_o if self.opt_poses else None, d_d if self.opt_poses else None,
                                stash=self.fm_stash)
        sm, w = None, 1.0 / self.world
        if self.with_tv:       # parameter-only work, replicated: every rank contributes 1 / world of it
            sm = fused.tv_raw(dec.pe_fn.grid_fn.gstruct, dec.bound, p["table"], p["coarse"], self.smooth_pts, None, None,
                              self.lambda_sm * w, g["table"], g["coarse"], oj_dev=self.draw_view(draws, "tv", torch.float64))
        if self.opt_poses:
            fused.pose_grad_raw(self.cam, self.window, d_o, d_d, b["pixel"], self.ray_start, self.quats, self.d_quats,
                                self.d_trans, self.pose_scratch)
        # loss vector (it travels with the gradients through the all-reduce) and the result vector: dns_map_step_result
        def result(phase, nvalid_scale=1.0):
            _lib.check(L.dns_map_step_result(phase, _lib.ptr(losses, f32), _lib.ptr(sm, f32, allow_none=True), w,
                                             self.lambda_sm * w, nvalid_scale, _lib.ptr(self.scratch, f32), F,
                                             self.loss_vec.data_ptr(), _lib.ptr(self.result_dev, f32), _lib.stream()))
        if self.world > 1:
            result(1)
            self.comm.all_reduce_sum(self.packed)
            result(2, 1.0 / self.world)               # n_valid is a batch constant, not a partial sum
        else:
            result(3)
        self.adam.step()
        return self.result_dev

    def read_result(self):
        """Device -> pinned host read of the last step's result vector (asynchronous); ``check`` raises as the
        reference would (mapping.py:594-595) once the copy has landed."""
        self.result_host.copy_(self.result_dev, non_blocking=True)
        return self.result_host

    def check(self, host_vec=None):
        v = self.result_host if host_vec is None else host_vec
        F2 = 9 + 2 * self.F
        if float(v[7]) < 0 or float(v[F2 + 1]) < 0:          # this step, or any step since construction
            fused.raise_on_flag(torch.cat((v[:7], torch.minimum(v[7:8], v[F2 + 1:F2 + 2]))))
        outside = float(v[9:F2].reshape(self.F, 2)[:, 1].sum()) + float(v[F2])
        if outside > 0:
            raise RuntimeError(f"{int(outside)} sampled rays leave the scene bound before their depth (mapping.py:525 "
                               "drops them): static-shape step not applicable, use slam.map_optimize")


class TrackingFrameStep:
    """The pose loop of ``Tracker.run`` (slams/tracking.py:304-346) from what the host holds: the new RGB-D-label frame, the
    previous frame's pose, the two feature maps and the iteration's draws.  Per iteration, back to back on the device and
    without autograd or PyTorch kernels on the data path:

        dns_pose_prepare (R(q); the current view follows the pose estimate, tracking.py:317-319)
        -> dns_sample_rays (window [20, H-20) x [20, W-20), tracking.py:140-147) -> dns_featmerge_fwd (band samples only)
        -> dns_render_fwd_bwd in TRACK mode (decoder frozen: ray and pixel-feature gradients only)
        -> dns_featmerge_bwd (into the ray gradients) -> dns_pose_grad -> dns_track_best (best pose / loss history on the
        device instead of the reference's per-iteration host comparison) -> dns_adam_multi over (translation, quaternion).

    One object serves every frame of a sequence (``reset``): buffers, workspaces and optimiser state are allocated once.
    ``draws``: one byte buffer ``idx`` int64 [n] | ``t_surface`` f32 [n_surface] (common.py:572-573 applied) |
    ``t_zero`` f32 [n_surface]; ``pack_draws`` fills a pinned host image from the draw dictionary of ``slam.track_frame``."""

    def __init__(self, dec, cam, n_pixels, n_iters, n_samples_ray=32, n_surface_ray=15, lambdas=None, cam_lr=1e-3,
                 seperate_LR=False, feat_shape=None):
        self.dec, self.cam = dec, cam
        dev = self.dev = dec.bound.device
        self.ns, self.nf, self.S = n_samples_ray, n_surface_ray, n_samples_ray + n_surface_ray
        self.lambdas = dict(p=5.0, d=5.0, l=0.1)
        self.lambdas.update(lambdas or {})
        N, S = int(n_pixels), self.S
        self.N, self.n_iters = N, int(n_iters)
        H, W = cam["H"], cam["W"]
        self.window = (20, H - 20, 20, W - 20)
        self.quats, self.trans = torch.zeros(1, 4, device=dev), torch.zeros(1, 3, device=dev)
        self.view_src = torch.tensor([-1, 0], dtype=torch.int32, device=dev)     # view 0: previous frame, view 1: this frame
        self.fixed_w2c, self.fixed_cam_o = torch.zeros(2, 4, 4, device=dev), torch.zeros(2, 3, device=dev)
        self.R_all = torch.empty(1, 3, 3, device=dev)
        self.w2c, self.cam_o = torch.empty(2, 4, 4, device=dev), torch.empty(2, 3, device=dev)
        self.batch = dict(gt_color=torch.empty(N, 3, device=dev), gt_depth=torch.empty(N, device=dev),
                          gt_label=torch.empty(N, dtype=torch.int64, device=dev), rays_o=torch.empty(N, 3, device=dev),
                          rays_d=torch.empty(N, 3, device=dev), z_vals=torch.empty(N, S, device=dev),
                          inside=torch.empty(N, dtype=torch.uint8, device=dev),
                          pixel=torch.empty(N, dtype=torch.int64, device=dev), scratch=torch.zeros(2, device=dev))
        self.features = torch.zeros(N, S, 32, device=dev)      # band rows are rewritten every step, the others never read
        self.fm_ws = torch.empty(int(_lib.lib().dns_featmerge_workspace_bytes(N, S)), dtype=torch.uint8, device=dev)
        self.fm_stash = fused.featmerge_stash(N, S, 2, dev)
        self.t_lin = torch.linspace(0.0, 1.0, steps=n_samples_ray).to(dev)
        self.off_ts = (8 * N + 15) & ~15
        self.off_tz = (self.off_ts + 4 * self.nf + 15) & ~15
        self.draw_bytes = self.off_tz + 4 * self.nf
        self.draws_dev = torch.zeros(self.draw_bytes, dtype=torch.uint8, device=dev)
        self.grads = torch.zeros(7, device=dev)                    # d_quat [4] | d_trans [3]
        self.d_quats, self.d_trans = self.grads[:4].view(1, 4), self.grads[4:].view(1, 3)
        self.pose_scratch = torch.empty(12, device=dev)
        self.adam = fused.AdamSegments([(self.trans.view(-1), self.d_trans.view(-1), cam_lr * (0.2 if seperate_LR else 1.0)),
                                        (self.quats.view(-1), self.d_quats.view(-1), cam_lr)])      # tracking.py:119-124
        # state vector: best [quad | T] (7) | best loss | running min of n_valid (error flag) | loss history
        self.state = torch.zeros(9 + self.n_iters, device=dev)
        self.best, self.best_loss, self.err_min = self.state[:7], self.state[7:8], self.state[8:9]
        self.hist = self.state[9:]
        self.slot = torch.zeros(1, dtype=torch.int32, device=dev)
        self.frame = self.feats = None      # static copies of the frame images / feature maps (a captured graph reads them)
        self.ws = torch.empty(fused.render_workspace_bytes(_lib.MODE_TRACK, N, S, dec.n_class) + 4096, dtype=torch.uint8,
                              device=dev)       # own workspace: the shared scratch may be re-allocated by later calls
        self.graph = None
        self.ring = [torch.zeros(self.draw_bytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
        self.ring_done = [None] * len(self.ring)
        self.turn = 0

    def reset(self, frame, refer_w2c, feats_cl, est_c2w):
        """A new frame: images (device tensors, used in place), the previous frame's world-to-camera matrix, the
        [2,h,w,64] channels-last feature maps of (previous, new) frame, the start pose; fresh Adam state."""
        from . import slam
        dev = self.dev
        if feats_cl.shape[0] != 2:
            raise ValueError("tracking uses two views: the previous frame and the new one (tracking.py:293-296)")
        if self.frame is None or any(self.frame[k].shape != frame[k].shape for k in self.frame) \
                or self.feats[0].shape != feats_cl.shape:
            self.frame = {k: torch.empty_like(frame[k], device=dev).contiguous() for k in ("color", "depth", "label")}
            self.feats = [torch.empty_like(feats_cl, device=dev).contiguous()]
            self.graph = None
        w2c = refer_w2c.detach().to(dev, torch.float32)
        with torch.no_grad():
            for k, v in self.frame.items():
                v.copy_(frame[k], non_blocking=True)
            self.feats[0].copy_(feats_cl, non_blocking=True)
            self.fixed_w2c[0].copy_(w2c)
            self.fixed_cam_o[0].copy_(fused.rigid_inverse(w2c)[:3, 3])
            self.quats[0].copy_(slam.quad_from_matrix(est_c2w[:3, :3]), non_blocking=True)
            self.trans[0].copy_(est_c2w[:3, 3].detach().float(), non_blocking=True)
            self.state.zero_()
            self.best[:4].copy_(self.quats[0])
            self.best[4:].copy_(self.trans[0])
            self.best_loss.fill_(1e10)
            self.slot.zero_()
        self.adam.reset_state()

    def pack_draws(self, d, buf):
        idx = d["idx"].reshape(-1)
        if idx.numel() != self.N:
            raise ValueError(f"{idx.numel()} pixel draws for a step of {self.N} rays")
        buf[:8 * self.N].view(torch.int64).copy_(idx)
        ts = d["t_surface"].detach().cpu().to(torch.float32).clone()
        if not bool((ts == 0.5).any()):
            ts[self.nf // 2 + 1] = 0.5                                     # common.py:572-573
        buf[self.off_ts:self.off_ts + 4 * self.nf].view(torch.float32).copy_(ts)
        buf[self.off_tz:self.off_tz + 4 * self.nf].view(torch.float32).copy_(d["t_zero"].detach().cpu().to(torch.float32))
        return buf

    def upload_draws(self, d):
        """Draw dictionary -> pinned staging buffer (ring of 4) -> ONE asynchronous H2D copy."""
        k = self.turn % len(self.ring)
        self.turn += 1
        if self.ring_done[k] is not None:
            self.ring_done[k].synchronize()
        self.draws_dev.copy_(self.pack_draws(d, self.ring[k]), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.ring_done[k] = ev

    def step(self):
        dec, b, N = self.dec, self.batch, self.N
        L = _lib.lib()
        f32 = torch.float32
        _lib.check(L.dns_pose_prepare(_lib.ptr(self.quats, f32), _lib.ptr(self.trans, f32), 1, _lib.ptr(self.view_src, torch.int32),
                                      _lib.ptr(self.fixed_w2c, f32), _lib.ptr(self.fixed_cam_o, f32), 2,
                                      _lib.ptr(self.R_all), _lib.ptr(self.w2c), _lib.ptr(self.cam_o), _lib.stream()))
        dr = self.draws_dev
        fused.sample_rays(self.cam, dec.bound, self.frame, dr[:8 * N].view(torch.int64), self.window, self.R_all[0], self.trans[0],
                          self.ns, self.nf, dr[self.off_ts:self.off_ts + 4 * self.nf].view(f32),
                          dr[self.off_tz:self.off_tz + 4 * self.nf].view(f32), t_lin=self.t_lin, out=b)
        mask = (b["gt_depth"] > 0.01).to(torch.uint8).mul_(b["inside"])          # tracking.py:172-173
        views = fused.Views(self.w2c, self.cam_o, self.feats, (0, N))
        merge_p = dec.view("merge")
        fused.featmerge_raw(self.cam, dec.merge.bound, views, b["rays_o"], b["rays_d"], b["z_vals"], b["gt_depth"], merge_p,
                            True, ws=self.fm_ws, out=self.features, stash=self.fm_stash, zero_fill=fused._use_simt)
        cfg = fused.RenderConfig(_lib.MODE_TRACK, dec.bound, dec.pe_fn.grid_fn.gstruct, b["z_vals"], b["gt_color"], b["gt_depth"],
                                 b["gt_label"], mask, None, dec.n_class, self.lambdas)
        losses, _, d_o, d_d, d_f = fused.render_raw(cfg, dec.view("table"), dec.view("coarse"), dec.view("color"),
                                                    dec.view("logit"), None, b["rays_o"], b["rays_d"], self.features, None,
                                                    True, True, ws=self.ws, features_band_only=True)
        fused.featmerge_bwd_raw(self.cam, dec.merge.bound, views, b["rays_o"], b["rays_d"], b["z_vals"], b["gt_depth"], merge_p,
                                d_f, self.fm_ws, None, d_o, d_d, stash=self.fm_stash)
        fused.pose_grad_raw(self.cam, self.window, d_o, d_d, b["pixel"], (0, N), self.quats, self.d_quats, self.d_trans,
                            self.pose_scratch)
        _lib.check(L.dns_track_best(_lib.ptr(losses, f32), _lib.ptr(self.quats, f32), _lib.ptr(self.trans, f32),
                                    _lib.ptr(self.best), _lib.ptr(self.best_loss), _lib.ptr(self.hist, allow_none=True),
                                    _lib.ptr(self.slot, torch.int32), self.n_iters, _lib.ptr(self.err_min), _lib.stream()))
        self.adam.step()
        return losses

    def run(self, frame, refer_w2c, feats_cl, est_c2w, draws_fn, n_iters=None, use_graph=True):
        """The whole loop of one frame; ONE host read at the end.  Returns (best [quad|T], best loss, loss history).
        ``use_graph``: the ~25 launches of ``step`` are captured ONCE (after two eager iterations of the first frame) and
        replayed for every later iteration of every later frame -- the iteration is launch bound at tracking sizes."""
        n = self.n_iters if n_iters is None else int(n_iters)
        if n > self.n_iters:
            raise ValueError("more iterations than the step was built for")
        self.reset(frame, refer_w2c, feats_cl, est_c2w)
        it = 0
        if use_graph and self.graph is None and n > 2:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for it in range(2):
                    self.upload_draws(draws_fn(it))
                    self.step()
            torch.cuda.current_stream().wait_stream(side)
            it = 2
            self.upload_draws(draws_fn(it))
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):     # capture only records: the captured iteration is replayed below
                self.step()
            self.graph.replay()
            it = 3
        for it in range(it, n):
            self.upload_draws(draws_fn(it))
            if use_graph and self.graph is not None:
                self.graph.replay()
            else:
                self.step()
        v = self.state.clone()
        if n > 0 and float(v[8]) < 0:     # a label outside the semantic head raises (torch's cross_entropy would)
            fused.raise_on_flag(torch.cat((v[:7] * 0, v[8:9])))
        return v[:7], v[7], v[9:9 + n]
