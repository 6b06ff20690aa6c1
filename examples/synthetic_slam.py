"""A minimal tracking + mapping loop on a synthetic RGB-D + label sequence, wired the way slams/dns_slam.py wires the
reference: the mapper optimises the shared decoder on a window of key frames, the tracker takes a copy of the weights
and optimises the pose of every new frame, key frames are chosen by overlap, the result is check-pointed and one frame
is rendered.  Everything numerical goes through the C ABI; this file only orchestrates (orchestration is not part of
the scoped path -- it is here to show the drop-in surface end to end and is exercised by tests/test_gpu_pipeline.py).

    python examples/synthetic_slam.py [shape] [n_frames] [map_every]

``map_every`` = the reference's ``mapping.every_frame`` (5 in configs/replica/replica.yaml and scannet.yaml): frames in
between are only tracked.  ``python examples/synthetic_slam.py scannet 200 5`` is BASELINE.json's configuration 3.
"""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dns_slam_b200 import bench_util, checkpoint, encoder, inference, slam  # noqa: E402
from dns_slam_b200 import synthetic as syn  # noqa: E402


def run(shape="tiny", n_frames=4, n_class=5, track_iters=8, map_iters=8, use_graph=True, seed=0, out_dir=None, verbose=True,
        map_every=1):
    dev = torch.device("cuda:0")
    s = syn.SHAPES[shape]
    cam = syn.camera(shape)
    gen = torch.Generator().manual_seed(seed)
    poses = syn.trajectory(shape, n_frames + 1)
    frames = [{k: v.to(dev).contiguous() for k, v in syn.frame(shape, poses[i], gen, n_class=n_class).items()}
              for i in range(n_frames)]
    torch.manual_seed(seed)
    stem = encoder.ResNet().to(dev)     # random-initialised stem (no pretrained weights without a network), training-mode bn1
    feats = [stem.forward_cl(fr["color"][None, None]) for fr in frames]      # [1, h, w, 64] channels-last per frame

    def pair_feats(f):
        """Tracking sees two views, the previous frame and the new one, encoded in ONE call as in
        slams/tracking.py:293-296 (bn1 statistics over both): [2, h, w, 64]."""
        return stem.forward_cl(torch.stack((frames[f]["color"], frames[f + 1]["color"]), 0)[None])
    shared = bench_util.make_decoder(shape, n_class, dev, seed=seed, all_experts=False)
    tracker_dec = bench_util.make_decoder(shape, n_class, dev, seed=seed + 1, all_experts=False)
    mapper = slam.MapperCore(cam, shared, s["mapping_pixels"], 32, 15,
                             lambdas=dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=10.0,
                                          fs=s["lambda_fs"], op=s["lambda_opacity"]),
                             opacity_sigma=s["opacity_sigma"], smooth_pts=s["smooth_pts"], lambda_sm=s["lambda_smooth"])
    tracker = slam.TrackerCore(cam, tracker_dec, s["tracking_pixels"], 32, 15, s["lambda_color"], s["lambda_depth"],
                               s["lambda_label"], freeze_decoder=True)
    est = [poses[0].clone()]
    keyframes = [0]
    log = []
    import time
    spent = {"track": [], "map": []}        # wall seconds per tracked frame / mapping call (device work included)

    def timed(kind, fn, *a, **kw):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **kw)
        torch.cuda.synchronize()
        spent[kind].append(time.perf_counter() - t0)
        return r
    for f in range(n_frames):
        if f % map_every != 0 and f + 1 < n_frames:      # tracked-only frame (mapping.every_frame)
            tracker_dec.copy_weights_from(shared)
            td = bench_util.tracking_draws(cam, s["tracking_pixels"], track_iters, seed=seed + 100 + f)
            best, best_loss, hist = timed("track", slam.track_frame, tracker, frames[f + 1], torch.inverse(est[f]), pair_feats(f),
                                          est[f].clone(), track_iters, s["cam_lr"], lambda it: td[it], use_graph=use_graph)
            est.append(slam.c2w_from_quad_T(best[:4], best[4:]).cpu())
            log.append(("track", f + 1, float(hist[0]), float(best_loss)))
            continue
        # ---- mapping on the window [selected key frames ..., current frame] (mapping.py:839-949)
        if f > 0:
            kf_c2w = torch.stack([est[k] for k in keyframes], 0).to(dev)
            idx = torch.randint(cam["H"] * cam["W"], (100,), generator=gen)
            picked = slam.keyframe_selection_overlap(cam, frames[f]["depth"], est[f], kf_c2w, 2, idx)
            window = sorted({keyframes[i] for i in picked} | {f})
        else:
            window = [0]
        for c in torch.unique(torch.cat([frames[k]["label"].reshape(-1) for k in window])).tolist():
            shared.activate_expert(int(c))                       # experts appear with their classes (mapping.py:727-761)
        target = dict(kf_idx=window, frames=[frames[k] for k in window],
                      class_tables=[slam.class_tables(frames[k]["label"], n_ids=n_class) for k in window])
        refer = dict(kf_idx=[[-1]] * len(window), est_c2w=[[est[k].to(dev)] for k in window])
        scene = dict(cam=cam, frames=target["frames"], class_tables=target["class_tables"])
        md, tv = bench_util.mapping_draws(scene, s["mapping_pixels"], map_iters, seed=seed + 10 * f)
        quads, Ts, losses = timed("map", slam.map_optimize, mapper, target, refer, [feats[k] for k in window],
                                  [est[k] for k in window], map_iters, s["lr"], s["BA_cam_lr"], len(window) > 1, [],
                                  lambda it: md[it], lambda it: tv[it], use_graph=use_graph)
        for k, q, t in zip(window, quads, Ts):
            est[k] = slam.c2w_from_quad_T(q.detach(), t.detach()).cpu()
        log.append(("map", f, float(losses["p_loss"]), float(losses["d_loss"]),
                    bool(use_graph and getattr(mapper, "last_path", "eager") != "eager")))
        if f not in keyframes:
            keyframes.append(f)
        if f + 1 == n_frames:
            break
        # ---- tracking of the next frame with a copy of the weights (tracking.py:296-346)
        tracker_dec.copy_weights_from(shared)
        guess = est[f].clone()
        td = bench_util.tracking_draws(cam, s["tracking_pixels"], track_iters, seed=seed + 100 + f)
        best, best_loss, hist = timed("track", slam.track_frame, tracker, frames[f + 1], torch.inverse(est[f]), pair_feats(f), guess,
                                      track_iters, s["cam_lr"], lambda it: td[it], use_graph=use_graph)
        est.append(slam.c2w_from_quad_T(best[:4], best[4:]).cpu())
        log.append(("track", f + 1, float(hist[0]), float(best_loss)))
    # ---- checkpoint (mapping.py:1119-1145) and one rendered frame (mapping.py:636-690)
    out_dir = out_dir or tempfile.mkdtemp(prefix="dns_slam_b200_")
    ck = checkpoint.Checkpoint(out_dir, device=dev, decoder=shared)
    ck.save("model.pt", idx=n_frames - 1, fine_decoders=shared.fine_decoders, keyframe_list=keyframes,
            estimate_c2w_list=torch.stack(est, 0))
    g = torch.Generator().manual_seed(seed + 999)
    color, depth, label = inference.render_frame(cam, shared, frames[-1], est[-1], torch.inverse(est[-1]), feats[-1], 32, 15,
                                                 torch.rand(15, generator=g), torch.rand(15, generator=g), n_pts_batch=2048)
    if verbose:
        for kind, f, u, v, *_ in log:
            if kind == "map":
                print("map   frame %d  p_loss %.4f  d_loss %.4f (last iteration)" % (f, u, v))
            else:
                print("track frame %d  loss %.4f (first iteration) -> %.4f (best)" % (f, u, v))
        print("rendered", tuple(color.shape), "checkpoint", os.path.join(out_dir, "model.pt"))
    timings = {k: {"calls": len(v), "ms_per_call": 1e3 * sum(v) / max(len(v), 1)} for k, v in spent.items()}
    timings["track"]["ms_per_iteration"] = timings["track"]["ms_per_call"] / max(track_iters, 1)
    timings["map"]["ms_per_iteration"] = timings["map"]["ms_per_call"] / max(map_iters, 1)
    return dict(log=log, est=est, gt=poses, render=(color, depth, label), out_dir=out_dir, decoder=shared, timings=timings)


if __name__ == "__main__":
    import time
    shape_ = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    n_ = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    every_ = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    if shape_ == "tiny":
        run(shape_, n_, map_every=every_)
    else:   # the reference's iteration counts for the shape (configs/*/*.yaml), 40 classes
        s_ = syn.SHAPES[shape_]
        t0 = time.perf_counter()
        out = run(shape_, n_, n_class=40, track_iters=s_["tracking_iters"], map_iters=s_["mapping_iters"], map_every=every_,
                  verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tr = [e for e in out["log"] if e[0] == "track"]
        mp_ = [e for e in out["log"] if e[0] == "map"]
        err = torch.stack([(a[:3, 3] - b[:3, 3]).norm() for a, b in zip(out["est"], out["gt"])])
        print(f"{shape_}: {n_} frames, mapping every {every_}: {dt:.1f} s wall incl. synthetic data generation "
              f"({len(tr)} tracked frames x {s_['tracking_iters']} iterations, {len(mp_)} mapping calls x {s_['mapping_iters']} "
              f"iterations); translation drift vs the synthetic trajectory (random-colour frames, not an accuracy "
              f"figure): mean {float(err.mean()):.4f} m, max {float(err.max()):.4f} m; last map p_loss {mp_[-1][2]:.4f} d_loss {mp_[-1][3]:.4f}; "
              f"{sum(1 for e in mp_ if e[4])} of {len(mp_)} mapping calls on the native loop (the others had rays "
              f"leaving the bound and ran the compacting eager loop); per call: {out['timings']}")
