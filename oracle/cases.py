"""TEST INFRASTRUCTURE ONLY -- rebuilds the seeded inputs of the golden cases and runs the
oracle restatement (oracle/reference_path.py) over them.  Shared by tests/ (oracle-vs-golden
on CPU, CUDA-vs-oracle on the GPU) so both sides see identical tensors.
"""
import torch

from dns_slam_b200 import synthetic as syn

from . import reference_path as rp
from .make_golden import build_models


def tracking_inputs(meta):
    shape, n_class, seed = meta["shape"], meta["n_class"], meta["seed"]
    gen = torch.Generator().manual_seed(seed)
    bound, dec, _ = build_models(shape, n_class, seed)
    cam = syn.camera(shape)
    poses = syn.trajectory(shape, 6)
    fr = syn.frame(shape, poses[meta["pose_index"]], gen, n_class=n_class)
    feats = syn.pixel_features(shape, 2, gen)
    return dict(bound=bound, decoder=dec, cam=cam, poses=poses, frame=fr, feats=feats)


def run_tracking(meta, quad, T, tape_items, inp=None):
    """One tracking iteration body (tracking.py:316-338) through the oracle."""
    inp = inp or tracking_inputs(meta)
    s = syn.SHAPES[meta["shape"]]
    dec, bound, cam = inp["decoder"], inp["bound"], inp["cam"]
    quad = quad.clone().requires_grad_(True)
    T = T.clone().requires_grad_(True)
    cur_c2w = rp.c2w_from_quad_T(quad, T)
    refer_w2c = torch.stack((torch.inverse(inp["poses"][meta["refer_index"]]), torch.inverse(cur_c2w)), 0)
    tape = rp.DrawTape(tape_items)
    samples = rp.tracker_get_target_samples(cam, bound, dec, inp["frame"], quad, T, refer_w2c,
                                            inp["feats"], s["tracking_pixels"], meta["n_samples"],
                                            meta["n_surface"], tape)
    pc, pd, pv, pl = rp.tracker_renderer(dec, bound, samples)
    p, d, l = rp.tracking_losses(samples, pc, pd, pv, pl)
    loss = s["lambda_color"] * p + s["lambda_depth"] * d + s["lambda_label"] * l
    for prm in dec.parameters():
        prm.grad = None
    loss.backward()
    named = dict(dec.named_parameters())
    return dict(samples=samples, pred=dict(color=pc, depth=pd, var=pv, logits=pl),
                loss=dict(p=p, d=d, l=l, total=loss),
                grad=dict(quad=quad.grad, T=T.grad,
                          coarse=named["coarse_fn.decoder.params"].grad,
                          color=named["out_fn.color_decoder.params"].grad,
                          logit=named["out_fn.logit_decoder.params"].grad,
                          merge=named["merge.decoder.params"].grad,
                          table=named["pe_fn.grid_fn.params"].grad),
                inputs=inp)


def mapping_inputs(meta):
    shape, n_class, seed = meta["shape"], meta["n_class"], meta["seed"]
    gen = torch.Generator().manual_seed(seed)
    bound, dec, experts = build_models(shape, n_class, seed, expert_classes=range(n_class))
    cam = syn.camera(shape)
    poses = syn.trajectory(shape, 8)
    frames = [syn.frame(shape, poses[i], gen, n_class=n_class) for i in meta["tgt_ids"]]
    feats = [syn.pixel_features(shape, 3, gen) for _ in meta["tgt_ids"]]
    refer_c2w = [[poses[k if k >= 0 else meta["tgt_ids"][f]].clone() for k in ids]
                 for f, ids in enumerate(meta["refer_idx"])]
    return dict(bound=bound, decoder=dec, experts=experts, cam=cam, poses=poses, frames=frames,
                feats=feats, refer_c2w=refer_c2w)


def run_mapping(meta, quad_list, T_list, tape_items, inp=None):
    """One mapping iteration body (mapping.py:882-909) through the oracle."""
    inp = inp or mapping_inputs(meta)
    s = syn.SHAPES[meta["shape"]]
    dec, experts, bound, cam = inp["decoder"], inp["experts"], inp["bound"], inp["cam"]
    quad_list = [q.clone().requires_grad_(i != 0) for i, q in enumerate(quad_list)]
    T_list = [t.clone().requires_grad_(i != 0) for i, t in enumerate(T_list)]
    tape = rp.DrawTape(tape_items)
    samples = rp.mapper_get_target_samples(cam, bound, dec, inp["frames"], quad_list, T_list,
                                           meta["refer_idx"], meta["tgt_ids"], inp["refer_c2w"],
                                           inp["feats"], s["mapping_pixels"], meta["n_samples"],
                                           meta["n_surface"], tape)
    pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(dec, experts, bound, samples)
    p, d, l, lt, fs, op = rp.mapping_losses(samples, pc, pd, pl, fine, coarse, s["opacity_sigma"])
    sm = rp.smoothness(dec, bound, s["smooth_pts"], tape)
    loss = s["lambda_color"] * p + s["lambda_depth"] * d + s["lambda_label"] * l \
        + meta["lambda_lt"] * lt + meta["lambda_sm"] * sm + s["lambda_fs"] * fs + s["lambda_opacity"] * op
    for prm in list(dec.parameters()) + [e.params for e in experts.values()]:
        prm.grad = None
    loss.backward()
    named = dict(dec.named_parameters())
    return dict(samples=samples,
                pred=dict(color=pc, depth=pd, var=pv, logits=pl, fine=fine, coarse=coarse),
                loss=dict(p=p, d=d, l=l, lt=lt, sm=sm, fs=fs, op=op, total=loss),
                grad=dict(quad=[q.grad for q in quad_list], T=[t.grad for t in T_list],
                          coarse=named["coarse_fn.decoder.params"].grad,
                          color=named["out_fn.color_decoder.params"].grad,
                          logit=named["out_fn.logit_decoder.params"].grad,
                          merge=named["merge.decoder.params"].grad,
                          experts={c: e.params.grad for c, e in experts.items()},
                          table=named["pe_fn.grid_fn.params"].grad),
                inputs=inp)
