"""Host-side mirror of the reference's iteration bodies, on the native kernels.

    TrackerCore.get_target_samples / iteration   <->  slams/tracking.py:128-186, 313-340
    MapperCore.get_target_samples / iteration     <->  slams/mapping.py:471-588, 881-910

Same names, argument meaning and loss dictionary as the reference; every random draw is an
explicit input (``draws``), so that the CPU oracle and the GPU consume identical pixel indices
and surface offsets (oracle patch P5).  Orchestration (keyframes, process plumbing, logging,
meshing) is out of scope.
"""
import torch

from . import _lib, fused

BOTTOM = [0.0, 0.0, 0.0, 1.0]


def _quad2rotation_formula(quad):
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
            two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr),
            two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)]
    return torch.stack(rows, -1).reshape(quad.shape[0], 3, 3)


_QUAD_VEC = {}


def _quad2rotation_vectorised(quad):
    """The nine entries of ``_quad2rotation_formula`` evaluated as 9-vectors: R_flat = base + coef * two_s * (A*B + sgn*C*D)
    with the SAME fp32 operations per entry (x**2 == x*x, a - b == a + (-1)*b, 1 - v == 1 + (-1)*v are exact
    identities), so the result is bit-identical, in ~10 launches instead of ~45."""
    key = str(quad.device)
    if key not in _QUAD_VEC:
        # entry:        00      01      02      10      11      12      20      21      22      (r,i,j,k = 0,1,2,3)
        idx = torch.tensor([[2, 1, 1, 1, 1, 2, 1, 2, 1],      # A
                            [2, 2, 3, 2, 1, 3, 3, 3, 1],      # B
                            [3, 3, 2, 3, 3, 1, 2, 1, 2],      # C
                            [3, 0, 0, 0, 3, 0, 0, 0, 2]])     # D
        sgn = torch.tensor([1., -1., 1., 1., 1., -1., -1., 1., 1.])
        coef = torch.tensor([-1., 1., 1., 1., -1., 1., 1., 1., -1.])
        base = torch.tensor([1., 0., 0., 0., 1., 0., 0., 0., 1.])
        _QUAD_VEC[key] = tuple(t.to(quad.device) for t in (idx.reshape(-1), sgn, coef, base))
    idx, sgn, coef, base = _QUAD_VEC[key]
    g = quad[:, idx].reshape(quad.shape[0], 4, 9)
    two_s = 2.0 / (quad * quad).sum(-1)
    inner = g[:, 0] * g[:, 1] + sgn * (g[:, 2] * g[:, 3])
    return (base + coef * (two_s[:, None] * inner)).reshape(quad.shape[0], 3, 3)


_QUAD_HESS = {}


def _quad_hessian(device):
    """C[m,a,b,p] = d^2 A_ab / dq_m dq_p of the quadratic form A(q) with R = I + (2/|q|^2) A(q): a constant."""
    key = str(device)
    if key not in _QUAD_HESS:
        def A(q):
            r, i, j, k = q[0], q[1], q[2], q[3]
            return torch.stack([-(j * j + k * k), i * j - k * r, i * k + j * r,
                                i * j + k * r, -(i * i + k * k), j * k - i * r,
                                i * k - j * r, j * k + i * r, -(i * i + j * j)]).reshape(3, 3)
        q0 = torch.zeros(4, dtype=torch.float64)
        H = torch.stack([torch.autograd.functional.hessian(lambda q, a=a, b=b: A(q)[a, b], q0)
                         for a in range(3) for b in range(3)]).reshape(3, 3, 4, 4)       # [a,b,m,p]
        _QUAD_HESS[key] = H.permute(2, 0, 1, 3).contiguous().to(torch.float32).to(device)
    return _QUAD_HESS[key]


class _Quad2Rot(torch.autograd.Function):
    """Forward: the reference's element-wise formula, evaluated without a graph (same values).  Backward: the
    closed form  dL/dq = s (G : dA/dq) - s^2 q (G : A),  s = 2/|q|^2, A = (R - I)/s, in three small tensor ops
    instead of the ~150 scalar-sized autograd kernels the formula leaves behind (they dominated a graph-replayed
    iteration)."""

    @staticmethod
    def forward(ctx, quad):
        with torch.no_grad():
            R = _quad2rotation_vectorised(quad)
        ctx.save_for_backward(quad, R)
        return R

    @staticmethod
    def backward(ctx, G):
        quad, R = ctx.saved_tensors
        C = _quad_hessian(quad.device)
        s = 2.0 / (quad * quad).sum(-1)                                    # [B]
        dA = torch.einsum("mabp,np->nmab", C, quad)                        # [B,4,3,3]
        term1 = s[:, None] * (dA * G[:, None]).sum((-1, -2))               # [B,4]
        eye = torch.eye(3, device=quad.device, dtype=quad.dtype)
        GA = ((R - eye) * G).sum((-1, -2)) / s                             # G : A
        return term1 - (s * s * GA)[:, None] * quad


def quad2rotation(quad):
    """utils/common.py:406-429 (device-safe)."""
    if quad.requires_grad and torch.is_grad_enabled():
        return _Quad2Rot.apply(quad)
    return _quad2rotation_formula(quad)


def get_rotation_from_quad(quad):
    return quad2rotation(quad.unsqueeze(0))[0] if quad.dim() == 1 else quad2rotation(quad)


def c2w_from_quad_T(quad, T):
    R = get_rotation_from_quad(quad)
    return torch.cat([torch.cat((R, T[:, None]), -1), fused.bottom_row(quad.device)], 0)


def trunc_mask(z, gt_depth):
    """tracking.py:167-170."""
    d = gt_depth[:, None]
    front = (z < d * 0.95).to(z.dtype)
    back = (z > d * 1.05).to(z.dtype)
    return (1.0 - front) * (1.0 - back) * (d > 0.0).to(z.dtype)


class ClassTables(tuple):
    """(classes, order, starts, counts) with host copies of the small per-class lists made ONCE per frame,
    so that the per-iteration index draw needs no device->host read."""

    def __new__(cls, classes, order, starts, counts, host=None):
        self = super().__new__(cls, (classes, order, starts, counts))
        self.classes_h, self.starts_h, self.counts_h = host or (classes.tolist(), starts.tolist(), counts.tolist())
        return self


def class_tables(label_win, n_ids=None):
    """Per-frame tables for the class-balanced draw (common.py:312-322): labels do not change
    between iterations, so ``unique`` / ``nonzero`` are done once per frame, not per iteration.
    Returns (classes ascending, sorted pixel indices, start offsets, counts).  On the device with ``n_ids`` (the
    decoder's number of class ids) the tables come from ``dns_class_tables`` (a stable counting sort, three small
    kernels); host tensors (tests, tools) take the torch formulation."""
    flat = label_win.reshape(-1)
    if flat.is_cuda and n_ids is not None:
        import ctypes as C
        flat = flat.contiguous()
        dev, n = flat.device, flat.numel()
        L = _lib.lib()
        order = torch.empty(n, dtype=torch.int64, device=dev)
        cs = torch.empty(2, n_ids, dtype=torch.int32, device=dev)
        err = torch.empty(1, dtype=torch.int32, device=dev)
        ws = torch.empty(int(L.dns_class_tables_workspace_bytes(n, n_ids)), dtype=torch.uint8, device=dev)
        _lib.check(L.dns_class_tables(_lib.ptr(flat, torch.int64), n, n_ids, _lib.ptr(order), cs[0].data_ptr(),
                                      cs[1].data_ptr(), _lib.ptr(err), ws.data_ptr(), ws.numel(), _lib.stream()))
        host = torch.cat((cs.reshape(-1), err)).tolist()          # ONE small read per frame (the host lists below)
        if host[-1]:
            raise ValueError("label outside [0, n_class_ids)")
        counts_all, starts_all = host[:n_ids], host[n_ids:2 * n_ids]
        present = [c for c in range(n_ids) if counts_all[c] > 0]
        idx = torch.tensor(present, dtype=torch.int64, device=dev)
        return ClassTables(idx, order, cs[1].long()[idx], cs[0].long()[idx],
                           host=(present, [starts_all[c] for c in present], [counts_all[c] for c in present]))
    order = torch.sort(flat, stable=True)[1]
    classes, counts = torch.unique_consecutive(flat[order], return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    return ClassTables(classes, order, starts, counts)


def class_balanced_offsets(tables, n, draws):
    """common.py:315-330 with the per-class randint draws supplied in order (a class with one pixel consumes no
    draw): returns (offsets [n] int64, slot_base [n] int32, draws used) such that slot j samples window pixel
    ``order[slot_base[j] + offsets[j]]`` -- the lookup itself happens in the sampling kernel (``dns_sample_rays``)."""
    classes, order, starts, counts = tables
    counts_h, starts_h = tables.counts_h, tables.starts_h
    n_class = len(counts_h)
    n_k = n // n_class
    # slot -> first pixel of its class: fixed per frame and n (cached on the tables object)
    cache = tables.__dict__.setdefault("_slot_cache", {})
    if n not in cache:
        base, keep = [], []
        for c in range(n_class):
            m = n - n_k * (n_class - 1) if c == 0 else n_k
            base += [starts_h[c]] * m
            keep += [counts_h[c] != 1] * m
        base_t = torch.tensor(base, dtype=torch.int32, device=order.device)
        pos = None if all(keep) else torch.nonzero(torch.tensor(keep, device=order.device)).reshape(-1)
        cache[n] = (base_t, pos, sum(1 for c in counts_h if c != 1))
    base_t, pos, n_used = cache[n]
    flat = torch.cat([d.reshape(-1) for d in draws[:n_used]], 0).to(order.device) if n_used else None
    if pos is None:
        off = flat
    else:
        off = torch.zeros(base_t.numel(), dtype=torch.int64, device=order.device)
        if flat is not None:
            off[pos] = flat
    return off, base_t, n_used


def class_balanced_indices(tables, n, draws):
    """The resolved window indices of ``class_balanced_offsets`` (host-side helper of tests / tools)."""
    off, base_t, n_used = class_balanced_offsets(tables, n, draws)
    return tables[1][base_t.long() + off], n_used


class TrackerCore:
    """The part of ``Tracker`` that sits on the hot path."""

    def __init__(self, cam, decoder, n_pixels, n_samples_ray=32, n_surface_ray=15, lambda_p=5.0,
                 lambda_d=5.0, lambda_l=0.1, freeze_decoder=False):
        self.cam, self.decoder, self.freeze_decoder = cam, decoder, freeze_decoder
        self.H, self.W = cam["H"], cam["W"]
        self.K = cam["K"].to(decoder.bound.device)
        self.n_pixels, self.n_samples_ray, self.n_surface_ray = n_pixels, n_samples_ray, n_surface_ray
        self.lambda_p, self.lambda_d, self.lambda_l = lambda_p, lambda_d, lambda_l

    def get_target_samples(self, cur_frames, refer_frames, features_cl, draws):
        """draws = dict(idx [n] int64, t_surface [15], t_zero [15]).  ``refer_frames['est_w2c']`` [R,4,4]; optional
        ``refer_frames['est_c2w']`` (the poses it was inverted from; the reference inverts back, common.py:672)."""
        quad, T = cur_frames["est_quad"], cur_frames["est_T"]
        R = get_rotation_from_quad(quad.detach())
        window = (20, self.H - 20, 20, self.W - 20)
        idx = draws["idx"].to(quad.device)
        want_pose = quad.requires_grad or T.requires_grad
        s = fused.sample_rays(self.cam, self.decoder.bound, cur_frames, idx, window, R, T, self.n_samples_ray,
                              self.n_surface_ray, fused.fix_surface_draw(draws["t_surface"], self.n_surface_ray),
                              draws["t_zero"], want_pixel=want_pose)
        n = idx.numel()
        rays_o, rays_d = s["rays_o"], s["rays_d"]
        if want_pose:
            rays_o, rays_d = fused.pose_rays(quad.reshape(1, 4), T.reshape(1, 3), rays_o, rays_d, s["pixel"], (0, n),
                                             self.cam, window)
        z = s["z_vals"]
        w2c = refer_frames["est_w2c"].detach()
        c2w = refer_frames.get("est_c2w")
        cam_o = (fused.rigid_inverse(w2c) if c2w is None else c2w.detach())[:, :3, 3]
        views = fused.Views(w2c, cam_o, [features_cl], (0, n))
        mp = self.decoder.merge.decoder.params
        if self.freeze_decoder:   # tracking never updates the Merge weights: skip their gradient GEMMs
            mp = mp.detach()
        code = fused.feature_merge(self.cam, self.decoder.merge.bound, views, rays_o, rays_d, z, s["gt_depth"], mp)
        mask = (s["gt_depth"] > 0.01) * s["inside"]
        return {"gt_color": s["gt_color"], "gt_depth": s["gt_depth"], "gt_label": s["gt_label"],
                "rays_o": rays_o, "rays_d": rays_d, "z_vals": z, "mask": mask, "features": code}

    def iteration(self, cur_frames, refer_frames, features_cl, draws):
        """Body of tracking.py:322-329; returns (loss dict, preds, samples)."""
        samples = self.get_target_samples(cur_frames, refer_frames, features_cl, draws)
        ld, preds = fused.render_and_loss(self.decoder, samples, _lib.MODE_TRACK,
                                          lambdas=dict(p=self.lambda_p, d=self.lambda_d, l=self.lambda_l),
                                          freeze_decoder=self.freeze_decoder)
        return ld, preds, samples


class MapperCore:
    """The part of ``Mapper`` that sits on the hot path."""

    def __init__(self, cam, decoder, n_pixels, n_samples_ray=32, n_surface_ray=15, lambdas=None,
                 opacity_sigma=0.05, smooth_pts=64, lambda_sm=1e-5):
        self.cam, self.decoder = cam, decoder
        self.H, self.W = cam["H"], cam["W"]
        self.K = cam["K"].to(decoder.bound.device)
        self.n_pixels, self.n_samples_ray, self.n_surface_ray = n_pixels, n_samples_ray, n_surface_ray
        self.lambdas = dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0)
        self.lambdas.update(lambdas or {})
        self.opacity_sigma, self.smooth_pts, self.lambda_sm = opacity_sigma, smooth_pts, lambda_sm

    def get_target_samples(self, target_frames, quad_list, T_list, refer_frames, features_cl, draws,
                           assume_inside=False):
        """target_frames: dict(kf_idx, frames=[dict(color,depth,label)], class_tables=[...]);
        refer_frames: dict(kf_idx=[[...]], est_c2w=[[...]]); features_cl[i]: [R,h,w,64];
        draws[i] = dict(idx_uniform, class_draws=[...], t_surface, t_zero)."""
        n_t = len(target_frames["frames"])
        n_pixels = self.n_pixels // n_t
        window = (0, self.H, 0, self.W)
        target_idx = target_frames["kf_idx"]
        # all target poses in ONE evaluation of the quaternion formula (same per-element arithmetic, no graph: the
        # backward is the closed form of fused.pose_rays)
        quats = torch.stack(list(quad_list), 0)
        T_all = torch.stack(list(T_list), 0)
        dev = quats.device
        with torch.no_grad():
            R_all = _quad2rotation_formula(quats)
            c2w_all = torch.cat((torch.cat((R_all, T_all[:, :, None]), -1),
                                 fused.bottom_row(dev)[None].expand(n_t, 1, 4)), 1)
            # reference-view poses of every target frame, inverted in ONE batched call (mapping.py:534-551)
            refer_c2w = []
            for i in range(n_t):
                for k, rid in enumerate(refer_frames["kf_idx"][i]):
                    if rid == -1:
                        refer_c2w.append(c2w_all[i])
                    elif rid in target_idx:
                        refer_c2w.append(c2w_all[target_idx.index(rid)])
                    else:
                        refer_c2w.append(refer_frames["est_c2w"][i][k].detach().to(dev))
            n_ref = [len(refer_frames["kf_idx"][i]) for i in range(n_t)]
            refer_c2w = torch.stack(refer_c2w, 0)
            refer_w2c = fused.rigid_inverse(refer_c2w)
        ref_at = [sum(n_ref[:i]) for i in range(n_t + 1)]
        # pixel draws and the sampling kernel per frame (each frame has its own images); the class-balanced third of
        # the draws is resolved to pixels inside the kernel
        want_pose = quats.requires_grad or T_all.requires_grad
        s_l = []
        for i in range(n_t):
            idx1 = draws[i]["idx_uniform"].to(dev)
            tab = target_frames["class_tables"][i]
            off, base, _ = class_balanced_offsets(tab, n_pixels // 3, draws[i]["class_draws"])
            s_l.append(fused.sample_rays(self.cam, self.decoder.bound, target_frames["frames"][i], torch.cat((idx1, off), 0),
                                         window, R_all[i], T_all[i].detach(), self.n_samples_ray, self.n_surface_ray,
                                         fused.fix_surface_draw(draws[i]["t_surface"], self.n_surface_ray),
                                         draws[i]["t_zero"], class_order=tab[1], slot_base=base,
                                         n_direct=idx1.numel(), want_pixel=want_pose))
        sizes = [int(x["z_vals"].shape[0]) for x in s_l]
        row_at = [sum(sizes[:i]) for i in range(n_t + 1)]
        cat = {k: torch.cat([x[k] for x in s_l], 0) for k in ("gt_color", "gt_depth", "gt_label", "rays_o", "rays_d",
                                                               "z_vals", "inside")}
        rays_o, rays_d = cat["rays_o"], cat["rays_d"]
        if want_pose:   # pose gradients for the rays of ALL frames at once (values stay the kernel's)
            rays_o, rays_d = fused.pose_rays(quats, T_all, rays_o, rays_d, torch.cat([x["pixel"] for x in s_l], 0), row_at,
                                             self.cam, window)
        mp = self.decoder.merge.decoder.params
        bound = self.decoder.merge.bound
        if len(set(n_ref)) == 1:
            # the whole pixel-feature branch (projection, gather, Merge MLP, mean over the views, truncation mask) of
            # all target frames in ONE fused call, band samples only
            views = fused.Views(refer_w2c, refer_c2w[:, :3, 3], features_cl, row_at)
            code = fused.feature_merge(self.cam, bound, views, rays_o, rays_d, cat["z_vals"], cat["gt_depth"], mp)
        else:
            parts = []
            for i in range(n_t):
                a, b, r0, r1 = ref_at[i], ref_at[i + 1], row_at[i], row_at[i + 1]
                views = fused.Views(refer_w2c[a:b], refer_c2w[a:b, :3, 3], [features_cl[i]], (0, r1 - r0))
                parts.append(fused.feature_merge(self.cam, bound, views, rays_o[r0:r1], rays_d[r0:r1],
                                                 cat["z_vals"][r0:r1], cat["gt_depth"][r0:r1], mp))
            code = torch.cat(parts, 0)
        m = cat.pop("inside")
        cat["rays_o"], cat["rays_d"], cat["features"] = rays_o, rays_d, code
        if assume_inside:          # graph replay: no compaction; the flag accumulates IN PLACE in a persistent device
            if getattr(self, "inside_ok", None) is None:   # bool (allocated before warm-up), read once after the loop
                self.inside_ok = torch.ones((), dtype=torch.bool, device=dev)
            self.inside_ok.logical_and_(m.all())
            return cat
        if bool(m.all()):          # one small D2H read; the reference syncs here too (mapping.py:576)
            return cat
        return {k: v[m] for k, v in cat.items()}

    def iteration(self, target_frames, quad_list, T_list, refer_frames, features_cl, draws, tv_draws,
                  lambda_lt=None, want_latents=False, assume_inside=False):
        """Body of mapping.py:884-907; returns (loss dict incl. smooth_loss and total, preds, samples)."""
        samples = self.get_target_samples(target_frames, quad_list, T_list, refer_frames, features_cl, draws,
                                          assume_inside=assume_inside)
        lam = dict(self.lambdas)
        if lambda_lt is not None:
            lam["lt"] = lambda_lt
        ld, preds = fused.render_and_loss(self.decoder, samples, _lib.MODE_MAP, lambdas=lam,
                                          opacity_sigma=self.opacity_sigma, want_latents=want_latents)
        sm = fused.tv_loss(self.decoder, self.smooth_pts, tv_draws[0], tv_draws[1])
        ld["smooth_loss"] = sm.detach()
        ld["total"] = ld["total"] + self.lambda_sm * sm
        return ld, preds, samples


# ----------------------------------------------------------------------------------------
# whole-loop drop-ins: the optimisation loops around the iteration bodies
# ----------------------------------------------------------------------------------------
def quad_from_matrix(R):
    """Rotation matrix -> quaternion (w,x,y,z) (utils/common.py:485-504 without ``mathutils``; the sign
    is immaterial because quad2rotation uses 2/|q|^2)."""
    import numpy as np
    R = np.asarray(R.detach().cpu() if isinstance(R, torch.Tensor) else R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return torch.tensor(q, dtype=torch.float32)


def uniq_class_indices(tables, n, class_list, draws):
    """common.py:375-393 (get_samples_by_uniq_class): n rays spread over the GIVEN classes; class 0 of
    the list takes the remainder, a class with one pixel is repeated, an absent class is skipped."""
    classes, order, starts, counts = tables
    lookup = {int(c): k for k, c in enumerate(tables.classes_h)}
    counts_h, starts_h = tables.counts_h, tables.starts_h
    n_class = len(class_list)
    n_k = n // n_class
    out, di = [], 0
    for pos, cid in enumerate(class_list):
        m = n - n_k * (n_class - 1) if pos == 0 else n_k
        k = lookup.get(int(cid))
        if k is None:
            continue
        if counts_h[k] == 1:
            out.append(order[starts_h[k]].reshape(1).repeat(m))
        else:
            out.append(order[starts_h[k] + draws[di].to(order.device)])
            di += 1
    return torch.cat(out, -1), di


# Benchmarks set this to a dict: the graph loops then bracket their replays with CUDA events and store
# ``replay_ms_per_iteration`` (device time of one replayed iteration, draw uploads included).
graph_timing = None


def _timed_replays(n, body):
    if graph_timing is None or n <= 0:
        for it in range(n):
            body(it)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(n):
        body(it)
    e1.record()
    e1.synchronize()
    graph_timing["replay_ms_per_iteration"] = e0.elapsed_time(e1) / n


def track_frame(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR=False,
                use_graph=False, native=None):
    """The pose-optimisation loop of ``Tracker.run`` (slams/tracking.py:304-346): Adam over
    (translation, quaternion), the best-loss pose is kept.  The ``loss < current_min_loss`` test that
    costs the reference a host sync per iteration (tracking.py:331) stays on the device.
    ``draws_fn(it)`` -> dict(idx, t_surface, t_zero).  Returns (best [quad|T] 7-vector, best loss, losses).
    ``use_graph``: capture ONE iteration (sampling, feature matching, fused render + backward, Adam, best
    pose) in a CUDA graph after three eager warm-up iterations and replay it -- the loop is launch bound
    (hundreds of tiny launches per iteration), not GPU bound.
    ``native`` (default: = ``use_graph`` for a tracker with a frozen decoder, i.e. the fast path): the loop runs through
    ``step.TrackingFrameStep`` -- no autograd, no PyTorch kernels on the data path, about 25 launches of this library per
    iteration, nothing to capture or warm up; one step object is cached on the tracker and serves every frame."""
    if native is None:
        native = use_graph and tracker.freeze_decoder
    if native:
        if not tracker.freeze_decoder:
            raise ValueError("the native tracking loop moves the pose only (TrackerCore(freeze_decoder=True), tracking.py:108-126)")
        return _track_frame_native(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR)
    if use_graph:
        return _track_frame_graph(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR)
    dev = tracker.decoder.bound.device
    quad = quad_from_matrix(est_c2w[:3, :3]).to(dev).requires_grad_(True)
    T = est_c2w[:3, 3].detach().clone().to(dev).requires_grad_(True)
    opt = fused.make_adam([{"params": [T], "lr": cam_lr * (0.2 if seperate_LR else 1.0)},
                           {"params": [quad], "lr": cam_lr}])
    best_loss = torch.full((), 1e10, device=dev)
    best = torch.cat((quad, T), 0).detach().clone()
    cur = dict(frame, est_quad=quad, est_T=T)
    history = []
    for it in range(n_iters):
        opt.zero_grad()
        cur_w2c = fused.rigid_inverse(c2w_from_quad_T(quad, T))
        est_w2c = torch.stack((refer_w2c.to(dev), cur_w2c), 0)
        ld, _, _ = tracker.iteration(cur, {"est_w2c": est_w2c}, features_cl, draws_fn(it))
        loss = ld["total"]
        with torch.no_grad():
            better = loss < best_loss
            best_loss = torch.where(better, loss, best_loss)
            best = torch.where(better, torch.cat((quad, T), 0), best)
        history.append(loss.detach())
        err = ld["n_valid"].detach() if it == 0 else torch.minimum(err, ld["n_valid"].detach())
        loss.backward()
        opt.step()
    if n_iters > 0:   # ONE host read per frame: a label outside the semantic head raises (torch's cross_entropy would)
        fused.raise_on_flag(torch.stack([err] * 8))
    return best, best_loss, torch.stack(history)


class TrackLoop:
    """The pose loop of one frame (slams/tracking.py:304-346) as ONE captured CUDA graph that is REUSED for every later
    frame: frame images, reference pose, feature maps, pose leaves, Adam moments, best pose and loss history live in
    static device buffers that are refreshed per frame (a few device copies), so a new frame costs no capture, no
    warm-up and no allocation -- tracking is launch bound (hundreds of tiny launches per iteration), not GPU bound.
    Cached on the tracker per (frame size, feature size, iteration count, learning rates)."""

    def __init__(self, tracker, frame, refer_w2c, features_cl, n_iters, cam_lr, seperate_LR, draws0):
        dev = tracker.decoder.bound.device
        self.tracker, self.n_iters, self.dev = tracker, n_iters, dev
        self.frame = {k: torch.empty_like(v) for k, v in frame.items() if isinstance(v, torch.Tensor)}
        self.refer = torch.empty(4, 4, device=dev)
        self.feats = torch.empty_like(features_cl)
        self.quad = torch.zeros(4, device=dev, requires_grad=True)
        self.T = torch.zeros(3, device=dev, requires_grad=True)
        self.opt = fused.make_adam([{"params": [self.T], "lr": cam_lr * (0.2 if seperate_LR else 1.0)},
                                    {"params": [self.quad], "lr": cam_lr}], capturable=True)
        self.packed = _PackedStatic(draws0, dev)
        self.best_loss = torch.full((), 1e10, device=dev)
        self.best = torch.zeros(7, device=dev)
        self.hist = torch.zeros(n_iters, device=dev)
        self.slot = torch.zeros((), dtype=torch.int64, device=dev)
        self.err_flag = torch.zeros((), device=dev)
        self.cur = dict(self.frame, est_quad=self.quad, est_T=self.T)
        self.graph = None
        self._draws0_loaded = True           # _PackedStatic has consumed draws0: do not fetch draws_fn(0) twice

    def _one(self):
        self.opt.zero_grad(set_to_none=True)
        quad, T = self.quad, self.T
        est_w2c = torch.stack((self.refer, fused.rigid_inverse(c2w_from_quad_T(quad, T))), 0)
        ld, _, _ = self.tracker.iteration(self.cur, {"est_w2c": est_w2c}, self.feats, self.packed.static)
        loss = ld["total"]
        with torch.no_grad():
            better = loss < self.best_loss
            self.best.copy_(torch.where(better, torch.cat((quad, T), 0), self.best))
            self.best_loss.copy_(torch.where(better, loss, self.best_loss))
            self.hist.index_copy_(0, self.slot.reshape(1), loss.detach().reshape(1))
            self.slot.add_(1)
            self.err_flag.copy_(torch.minimum(self.err_flag, ld["n_valid"].detach()))
        loss.backward()
        self.opt.step()

    def _reset(self, frame, refer_w2c, features_cl, est_c2w):
        with torch.no_grad():
            for k, v in self.frame.items():
                v.copy_(frame[k], non_blocking=True)
            self.refer.copy_(refer_w2c, non_blocking=True)
            self.feats.copy_(features_cl, non_blocking=True)
            self.quad.copy_(quad_from_matrix(est_c2w[:3, :3]), non_blocking=True)
            self.T.copy_(est_c2w[:3, 3].detach(), non_blocking=True)
            self.best.copy_(torch.cat((self.quad, self.T), 0))
            self.best_loss.fill_(1e10)
            self.hist.zero_()
            self.slot.zero_()
            self.err_flag.zero_()
        if hasattr(self.opt, "reset_state"):
            self.opt.reset_state()           # fresh optimiser per frame (tracking.py:119-124)

    def run(self, frame, refer_w2c, features_cl, est_c2w, draws_fn):
        self._reset(frame, refer_w2c, features_cl, est_c2w)
        n_iters = self.n_iters
        start = 0
        if self.graph is None:               # first frame: three eager iterations (they count), then the capture
            n_warm = min(3, n_iters)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for it in range(n_warm):
                    if it > 0 or not self._draws0_loaded:
                        self.packed.update(draws_fn(it))
                    self._one()
            self._draws0_loaded = False
            torch.cuda.current_stream().wait_stream(side)
            start = n_warm
            if n_iters > n_warm:
                self.packed.update(draws_fn(n_warm))
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._one()
                # the captured launches hold pointers into the shared scratch of fused.workspace(): keep THIS buffer
                # alive (a later, larger call re-allocates the shared one; the graph keeps using the one it captured)
                self._ws_ref = fused._ws_cache.get((self.dev.type, self.dev.index))

        def replay(j):                       # capture only records: the captured iteration is replayed too
            self.packed.update(draws_fn(start + j))
            self.graph.replay()
        if self.graph is not None:
            _timed_replays(n_iters - start, replay)
        fused.raise_on_flag(torch.stack([self.err_flag] * 8))      # the single host read of the loop
        return self.best.clone(), self.best_loss.clone(), self.hist.clone()


def _track_frame_native(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR):
    from . import step as stepmod
    key = (int(n_iters), float(cam_lr), bool(seperate_LR), tracker.n_pixels)
    cache = tracker.__dict__.setdefault("_track_steps", {})
    st = cache.get(key)
    if st is None:
        if len(cache) >= 2:
            cache.clear()
        st = cache[key] = stepmod.TrackingFrameStep(
            tracker.decoder, tracker.cam, tracker.n_pixels, n_iters, tracker.n_samples_ray, tracker.n_surface_ray,
            dict(p=tracker.lambda_p, d=tracker.lambda_d, l=tracker.lambda_l), cam_lr, seperate_LR)
    return st.run(frame, refer_w2c, features_cl, est_c2w, draws_fn)


def _track_frame_graph(tracker, frame, refer_w2c, features_cl, est_c2w, n_iters, cam_lr, draws_fn, seperate_LR):
    dev = tracker.decoder.bound.device
    key = (tuple(frame["color"].shape), tuple(features_cl.shape), int(n_iters), float(cam_lr), bool(seperate_LR),
           tracker.n_pixels)
    cache = tracker.__dict__.setdefault("_track_loops", {})
    loop = cache.get(key)
    if loop is None:
        if len(cache) >= 2:                  # a captured graph pins its buffers: keep the cache small
            cache.clear()
        loop = cache[key] = TrackLoop(tracker, frame, refer_w2c.to(dev), features_cl, n_iters, cam_lr, seperate_LR, draws_fn(0))
    return loop.run(frame, refer_w2c.to(dev), features_cl, est_c2w, draws_fn)


def map_optimize(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr, is_BA,
                 new_decoders, draws_fn, tv_draws_fn, use_graph=False, native=None, history=None):
    """The optimisation loop of ``Mapper.optimize`` (slams/mapping.py:868-910): one Adam over the decoder
    (hash grid + MLPs), the class experts present and -- when ``is_BA`` -- quaternion / translation of every
    target frame but the oldest (mapping.py:457); lambda_lt follows the schedule of mapping.py:898-904.
    Returns (quad_list, T_list, loss dict of the last iteration).
    ``use_graph``: after three eager iterations ONE iteration (4 frames of sampling + feature matching, fused
    render/backward, TV, Adam) is captured in a CUDA graph and replayed with fresh draws.  Replay needs static
    shapes, so it assumes every sampled ray passes the inside test of mapping.py:525 (checked once at the end;
    the call falls back to the eager loop from the saved state otherwise) and a constant lambda_lt.
    ``native`` (default: = ``use_graph``, i.e. the fast path): the whole loop runs through ``step.MappingFrameStep`` --
    no autograd, no PyTorch kernels on the data path, nothing to capture (about 60 launches of this library per
    iteration, so eager is already launch-cheap); same static-shape assumption and the same fallback.  The captured
    autograd loop remains for windows whose target frames have different numbers of reference views.
    ``history``: optional list that receives the total loss of every iteration (device scalars).
    ``mapper.last_path`` tells which loop produced the result ("native", "graph" or "eager")."""
    if native is None:
        native = use_graph
    mapper.last_path = "eager"
    if native and n_iters > 0 and len({len(x) for x in refer_frames["kf_idx"]}) == 1:
        out = _map_optimize_native(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr,
                                   is_BA, new_decoders, draws_fn, tv_draws_fn, history)
        if out is not None:
            return out
    elif use_graph and len(new_decoders) == 0 and n_iters > 4:
        out = _map_optimize_graph(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr,
                                  is_BA, draws_fn, tv_draws_fn)
        if out is not None:
            mapper.last_path = "graph"
            return out
    dec = mapper.decoder
    dev = dec.bound.device
    n_t = len(target_frames["frames"])
    quad_list, T_list = [], []
    for f in range(n_t):
        c2w = est_c2w_list[f]
        q = quad_from_matrix(c2w[:3, :3]).to(dev)
        t = c2w[:3, 3].detach().clone().to(dev)
        if (n_t == 1 or f != 0) and is_BA:
            q.requires_grad_(True)
            t.requires_grad_(True)
        quad_list.append(q)
        T_list.append(t)
    net = [p for p in dec.parameters() if p.requires_grad and p.numel() > 0]
    cam_lr = BA_cam_lr * float(is_BA)
    opt = fused.make_adam([{"params": net, "lr": lr, "flat": dec.flat, "rows": dec.expert_rows()},
                           {"params": [q for q in quad_list if q.requires_grad], "lr": cam_lr},
                           {"params": [t for t in T_list if t.requires_grad], "lr": cam_lr}])
    ld = None
    err_flag = torch.zeros((), device=dev)
    for it in range(n_iters):
        opt.zero_grad()
        lam_lt = (10.0 if it > n_iters // 2 else 0.0) if len(new_decoders) > 0 else 10.0
        ld, _, _ = mapper.iteration(target_frames, quad_list, T_list, refer_frames, features_cl, draws_fn(it),
                                    tv_draws_fn(it), lambda_lt=lam_lt)
        ld["total"].backward()
        opt.step()
        err_flag = torch.minimum(err_flag, ld["n_valid"].detach())
    if n_iters > 0:   # ONE host read per optimize call: the reference raises inside the iteration (mapping.py:594-595)
        fused.raise_on_flag(torch.stack([err_flag] * 8))
    # detached: a loss dictionary that still holds its autograd graph keeps the AccumulateGrad nodes of this (default)
    # stream alive, and the next CUDA-graph capture of a loop then fails ("legacy stream depends on a capturing stream")
    return quad_list, T_list, (None if ld is None else {k: v.detach() for k, v in ld.items()})


def keyframe_overlap(cam, gt_depth, c2w, keyframe_c2w, idx, n_samples=16):
    """``Mapper.keyframe_selection_overlap`` (slams/mapping.py:171-231) without the per-key-frame numpy loop: the
    ``pixels`` x 16 probe points of the current frame are projected into ALL key frames with one batched matmul on the
    device.  ``idx`` are the drawn pixel indices (flat, whole image), ``keyframe_c2w`` [K,4,4].  Returns
    percent_inside [K] (float32, on the device)."""
    dev = keyframe_c2w.device
    H, W = cam["H"], cam["W"]
    idx = idx.to(dev)
    dirs = fused.pixel_dirs(cam, idx, (0, H, 0, W))
    c2w = c2w.to(dev)
    p = dirs.reshape(-1, 1, 3) * c2w[:3, :3]
    rays_d = (p[..., 0] + p[..., 1]) + p[..., 2]
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    d = gt_depth.to(dev).reshape(-1)[idx].reshape(-1, 1)
    t_vals = torch.linspace(0.0, 1.0, steps=n_samples, device=dev)
    z_vals = d * 0.8 * (1.0 - t_vals) + (d + 0.5) * t_vals
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., None]).reshape(-1, 3)          # [V,3]
    w2c = torch.linalg.inv(keyframe_c2w.double())                                               # [K,4,4]
    homo = torch.cat((pts.double(), torch.ones(pts.shape[0], 1, dtype=torch.float64, device=dev)), -1)
    cam_cord = torch.einsum("kab,vb->kva", w2c, homo)[..., :3]                                   # [K,V,3]
    cam_cord = cam_cord * torch.tensor([-1.0, 1.0, 1.0], dtype=torch.float64, device=dev)
    K = cam["K"].to(dev).double()
    uv = torch.einsum("ab,kvb->kva", K, cam_cord)
    z = uv[..., 2:3] + 1e-5
    uv = (uv[..., :2] / z).float()
    edge = 10
    mask = (uv[..., 0] < W - edge) & (uv[..., 0] > edge) & (uv[..., 1] < H - edge) & (uv[..., 1] > edge) & (z[..., 0] < 0)
    return mask.float().mean(-1)


def keyframe_selection_overlap(cam, gt_depth, c2w, keyframe_c2w, k, idx, perm=None, th=0.0):
    """mapping.py:226-236: key frames with percent_inside > th, in descending order, shuffled by ``perm`` (the
    reference's np.random.permutation, passed in like every other draw; None keeps the sorted order), first k."""
    pct = keyframe_overlap(cam, gt_depth, c2w, keyframe_c2w, idx).cpu()
    order = sorted(range(pct.numel()), key=lambda i: float(pct[i]), reverse=True)
    sel = [i for i in order if float(pct[i]) > th]
    if perm is not None:
        sel = [sel[int(j)] for j in perm if int(j) < len(sel)]
    return sel[:k]


def decoder_init(mapper, decoder_idx, frame, class_table, cur_c2w, features_cl, lr, draws_fn, tv_draws_fn,
                 n_iters=100, n_rays=300):
    """``Mapper.decoder_init`` (slams/mapping.py:764-836): warm-up of freshly created class experts on the
    current frame: rays spread over the new classes (get_samples_by_uniq_class), losses p + d + l + fs +
    opacity + TV (no latent term), Adam over the decoder and the new experts."""
    dec = mapper.decoder
    dev = dec.bound.device
    for c in decoder_idx:
        dec.activate_expert(c)
    net = [p for p in dec.parameters() if p.requires_grad and p.numel() > 0]
    opt = fused.make_adam([{"params": net, "lr": lr, "flat": dec.flat, "rows": dec.expert_rows()}])
    R, T = cur_c2w[:3, :3].to(dev), cur_c2w[:3, 3].to(dev)
    window = (0, mapper.H, 0, mapper.W)
    c2w = cur_c2w.to(dev).unsqueeze(0)
    w2c = fused.rigid_inverse(c2w)
    lam = dict(mapper.lambdas, lt=0.0)
    ld, err = None, None
    for it in range(n_iters):
        opt.zero_grad()
        d = draws_fn(it)
        idx, _ = uniq_class_indices(class_table, n_rays, decoder_idx, d["class_draws"])
        s = fused.sample_rays(mapper.cam, dec.bound, frame, idx, window, R, T, mapper.n_samples_ray,
                              mapper.n_surface_ray, fused.fix_surface_draw(d["t_surface"], mapper.n_surface_ray),
                              d["t_zero"])
        n = idx.numel()
        views = fused.Views(w2c, c2w[:, :3, 3], [features_cl], (0, n))
        code = fused.feature_merge(mapper.cam, dec.merge.bound, views, s["rays_o"], s["rays_d"], s["z_vals"], s["gt_depth"],
                                   dec.merge.decoder.params, apply_trunc=False)   # no trunc mask here (mapping.py:807-809)
        samples = {"gt_color": s["gt_color"], "gt_depth": s["gt_depth"], "gt_label": s["gt_label"],
                   "rays_o": s["rays_o"], "rays_d": s["rays_d"], "z_vals": s["z_vals"], "features": code}
        ld, _ = fused.render_and_loss(dec, samples, _lib.MODE_MAP, lambdas=lam, opacity_sigma=mapper.opacity_sigma)
        tv = tv_draws_fn(it)
        sm = fused.tv_loss(dec, mapper.smooth_pts, tv[0], tv[1])
        (ld["total"] + mapper.lambda_sm * sm).backward()
        opt.step()
        err = ld["n_valid"].detach() if err is None else torch.minimum(err, ld["n_valid"].detach())
    if err is not None:   # ONE host read per call: 'Fine decoders does NOT have class' (mapping.py:594-595)
        fused.raise_on_flag(torch.stack([err] * 8))
    return None if ld is None else {k: v.detach() for k, v in ld.items()}


class _PackedStatic:
    """Static device copies of a nested draw structure (dict / list of host tensors) for graph replay.
    All leaves live in ONE device buffer and are refreshed with ONE host-to-device copy per iteration from a
    small ring of pinned staging buffers (a mapping iteration has ~170 draw tensors; one copy each was
    1-2 ms of launch overhead per replayed iteration)."""

    RING = 4

    def __init__(self, template, dev):
        self.leaves = []                      # (offset, nbytes, dtype, shape)
        off = 0

        def plan(d):
            nonlocal off
            if isinstance(d, torch.Tensor):
                n = d.numel() * d.element_size()
                self.leaves.append((off, n, d.dtype, tuple(d.shape)))
                off = (off + n + 15) & ~15
                return len(self.leaves) - 1
            if isinstance(d, dict):
                return {k: plan(v) for k, v in d.items()}
            return [plan(v) for v in d]

        self.index = plan(template)
        self.nbytes = max(off, 16)
        self.dev_buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=dev)
        self.host = [torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory() for _ in range(self.RING)]
        self.done = [None] * self.RING
        self.turn = 0
        self.static = self._views(self.dev_buf, self.index)
        self.update(template)

    def _leaf(self, buf, i):
        off, n, dt, shape = self.leaves[i]
        return buf[off:off + n].view(dt).view(shape)

    def _views(self, buf, idx):
        if isinstance(idx, int):
            return self._leaf(buf, idx)
        if isinstance(idx, dict):
            return {k: self._views(buf, v) for k, v in idx.items()}
        return [self._views(buf, v) for v in idx]

    def _fill(self, buf, idx, src):
        if isinstance(idx, int):
            self._leaf(buf, idx).copy_(src)
        elif isinstance(idx, dict):
            for k, v in idx.items():
                self._fill(buf, v, src[k])
        else:
            for v, s in zip(idx, src):
                self._fill(buf, v, s)

    def update(self, src):
        k = self.turn
        self.turn = (k + 1) % self.RING
        if self.done[k] is not None:
            self.done[k].synchronize()        # the copy that last read this staging buffer has finished
        self._fill(self.host[k], self.index, src)
        self.dev_buf.copy_(self.host[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.done[k] = ev


def _map_optimize_native(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr, is_BA,
                         new_decoders, draws_fn, tv_draws_fn, history=None):
    """``Mapper.optimize`` (slams/mapping.py:868-910) through ``step.MappingFrameStep``: per iteration ONE pinned draw
    buffer goes up, ``dns_pose_prepare -> dns_sample_rays -> dns_featmerge_fwd -> dns_render_fwd_bwd -> dns_featmerge_bwd
    -> dns_tv_fwd_bwd -> dns_pose_grad -> dns_adam_multi`` run back to back, and the loop ends with ONE host read (losses,
    error flag, rays that left the bound).  Returns None -- with the decoder restored -- when a sampled ray failed the
    inside test of mapping.py:525 (the reference drops such rays; the static-shape step cannot), so that the caller
    repeats the call with the compacting eager loop."""
    from . import step as stepmod
    dec = mapper.decoder
    dev = dec.bound.device
    saved = dec.flat.detach().clone()
    n_t = len(target_frames["frames"])
    st = stepmod.MappingFrameStep(dec, mapper.cam, target_frames["frames"], target_frames["class_tables"], features_cl,
                                  est_c2w_list, refer_frames["kf_idx"], refer_frames["est_c2w"], target_frames["kf_idx"],
                                  mapper.n_pixels, mapper.n_samples_ray, mapper.n_surface_ray, lr=lr, BA_cam_lr=BA_cam_lr,
                                  is_BA=is_BA, lambdas=mapper.lambdas, opacity_sigma=mapper.opacity_sigma,
                                  smooth_pts=mapper.smooth_pts, lambda_sm=mapper.lambda_sm, with_tv=True)
    rings = mapper.__dict__.setdefault("_draw_rings", {})     # pinned staging buffers, kept across calls (pinning is slow)
    ring = rings.get(st.draw_bytes)
    if ring is None:
        rings.clear()
        ring = rings[st.draw_bytes] = [torch.zeros(st.draw_bytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
    done = [None] * len(ring)
    F2 = 9 + 2 * n_t
    for it in range(n_iters):
        k = it % len(ring)
        if done[k] is not None:
            done[k].synchronize()           # the staging buffer's previous copy has landed
        st.plan.pack_draws(draws_fn(it), tv_draws_fn(it), ring[k])
        st.upload(ring[k])
        done[k] = torch.cuda.Event()
        done[k].record()
        st.lambdas["lt"] = (10.0 if it > n_iters // 2 else 0.0) if len(new_decoders) > 0 else 10.0   # mapping.py:898-904
        res = st.step()
        if history is not None:
            history.append(res[6].clone())
        if it == 0 and n_iters > 8 and float(res[F2]) > 0:
            break        # early look (one extra host read): a frame that pokes out of the bound shows in the first batch
    v = st.result_dev.tolist()               # the single host read of the loop
    outside, err = v[F2], min(v[7], v[F2 + 1])
    mapper.last_graph_ok = outside == 0
    if err < 0:                              # the reference raises inside the iteration (mapping.py:594-595)
        with torch.no_grad():
            dec.flat.copy_(saved)
        fused.raise_on_flag(torch.tensor([0.0] * 7 + [err]))
    if outside > 0:
        with torch.no_grad():
            dec.flat.copy_(saved)
        return None
    mapper.last_path = "native"
    ld = {k: torch.tensor(v[i], device=dev) for i, k in enumerate(fused.LOSS_KEYS)}
    ld["smooth_loss"] = torch.tensor(v[8], device=dev)
    return [st.quats[f].clone() for f in range(n_t)], [st.trans[f].clone() for f in range(n_t)], ld


def _map_optimize_graph(mapper, target_frames, refer_frames, features_cl, est_c2w_list, n_iters, lr, BA_cam_lr, is_BA,
                        draws_fn, tv_draws_fn):
    dec = mapper.decoder
    dev = dec.bound.device
    n_t = len(target_frames["frames"])
    saved = dec.flat.detach().clone()
    quad_list, T_list = [], []
    for f in range(n_t):
        c2w = est_c2w_list[f]
        q = quad_from_matrix(c2w[:3, :3]).to(dev)
        t = c2w[:3, 3].detach().clone().to(dev)
        if (n_t == 1 or f != 0) and is_BA:
            q.requires_grad_(True)
            t.requires_grad_(True)
        quad_list.append(q)
        T_list.append(t)
    net = [p for p in dec.parameters() if p.requires_grad and p.numel() > 0]
    cam_lr = BA_cam_lr * float(is_BA)
    groups = [{"params": net, "lr": lr, "flat": dec.flat, "rows": dec.expert_rows()}]
    if any(q.requires_grad for q in quad_list):
        groups += [{"params": [q for q in quad_list if q.requires_grad], "lr": cam_lr},
                   {"params": [t for t in T_list if t.requires_grad], "lr": cam_lr}]
    opt = fused.make_adam(groups, capturable=True)
    packed = _PackedStatic([draws_fn(0), list(tv_draws_fn(0))], dev)
    static_d, static_tv = packed.static
    # persistent device flags, updated IN PLACE by every (eager or replayed) iteration and read ONCE after the loop
    mapper.inside_ok = torch.ones((), dtype=torch.bool, device=dev)
    err_flag = torch.zeros((), device=dev)
    last = {}

    def one():
        opt.zero_grad(set_to_none=True)
        ld, _, _ = mapper.iteration(target_frames, quad_list, T_list, refer_frames, features_cl, static_d, static_tv,
                                    lambda_lt=10.0, assume_inside=True)
        ld["total"].backward()
        opt.step()
        err_flag.copy_(torch.minimum(err_flag, ld["n_valid"].detach()))
        for k, v in ld.items():
            if k not in last:
                last[k] = torch.zeros_like(v.detach())
            last[k].copy_(v.detach())

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for it in range(3):
            if it > 0:
                packed.update([draws_fn(it), list(tv_draws_fn(it))])
            one()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    packed.update([draws_fn(3), list(tv_draws_fn(3))])
    with torch.cuda.graph(graph):
        one()
    def replay(j):
        packed.update([draws_fn(3 + j), list(tv_draws_fn(3 + j))])
        graph.replay()
    _timed_replays(n_iters - 3, replay)
    flags = torch.stack((mapper.inside_ok.float(), err_flag)).tolist()   # the single host read of the loop
    ok = flags[0] != 0.0
    mapper.inside_ok = None
    mapper.last_graph_ok = ok
    if flags[1] < 0:                     # the reference raises inside the iteration (mapping.py:594-595)
        with torch.no_grad():
            dec.flat.copy_(saved)
        fused.raise_on_flag(torch.tensor([0.0] * 7 + [flags[1]]))
    if not ok:                           # some ray left the bound: static shapes were wrong, redo eagerly
        with torch.no_grad():
            dec.flat.copy_(saved)
        return None
    return quad_list, T_list, dict(last)
