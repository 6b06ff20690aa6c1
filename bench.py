#!/usr/bin/env python
"""Benchmark of the DNS-SLAM render-and-optimise hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is ONE MAPPING ITERATION (``slams/mapping.py:884-910``) over a Replica-shaped ray batch drawn from four
key frames that live on the device: ray / pixel sampling (uniform + class-balanced), far plane and depth-guided
z values, feature matching + Merge + truncation mask for the band samples, fused encode + MLP + render + all six
ray losses + TV smoothness, the backward down to hash table, every MLP, the class experts, Merge and the camera
poses, [NCCL exchanges when N > 1] and Adam over decoder + poses.  What crosses the host boundary per step is what a
SLAM host holds per iteration: its hoisted random draws (8 B per ray) in, the loss vector out.

``value``  rays/s of whole-job steps with the step's draws already resident in HBM.
``e2e``    the same steps through the public step API from HOST memory: the key frames (colour / depth / label), the
           reference views' images, their upload, the ResNet stem and the per-frame class tables happen INSIDE the timed
           region (once, as at the start of a mapping call), then every step uploads its draws from pinned memory and
           reads the loss vector back.
Each rank owns ``--rays-per-gpu`` rays of the global batch (weak scaling: 131072 x 8 = the 1M-ray mapping batch of
BASELINE.json config 4).  ``extra`` holds BASELINE configs 0-3, the core-only step of round 1 and the inference / stem
timings.  ``--impl reference`` times the oracle port of the SAME iteration on the host cores (the reference is Python
on tinycudann, which is CUDA-only: hash grid / MLPs run as the fp32 PyTorch stand-in), on the same frames, poses,
weights and draw generator, over a bounded number of rays per step.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays/sec fused encode+MLP+render+loss fwd/bwd"
BYTES_PT = dict(point_fwd=1024 + 4, ray=256, point_bwd=2048)      # SURVEY 8d, per sample point
BYTES_RAY = dict(point_fwd=24, ray=228, point_bwd=0)              # per ray
TV_BYTES_PT = 1024 + 2048 + 4                                     # SURVEY 8d, per lattice point (upper bound)
FEAT_BYTES_ROW = 4 * 256                                          # 4 bilinear taps x 64 channels per (band sample, view)
N_TARGET, N_REFER = 4, 3                                          # replica.yaml:41 (mapping window), refer views


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays-per-gpu", type=int, default=131072)
    ap.add_argument("--samples", type=int, default=47)
    ap.add_argument("--n-class", type=int, default=40)
    ap.add_argument("--shape", default="replica")
    ap.add_argument("--cpu-rays", type=int, default=16384)
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML library calls
    (spawning nvidia-smi every 100 ms stalls kernel launches by ~20 %; the query set is the one of
    B200_PROFILING.md: clocks.sm, clocks.max.sm, hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown,
    sw_power_cap)."""

    def __init__(self, index):
        self.index, self.sm, self.max_sm, self.bits, self.stop, self.err = index, [], None, 0, False, None
        self.nv = self.h = self.get_reasons = None
        # NVML is initialised HERE, before the timed region: nvmlInit takes driver-wide locks for 100-200 ms
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) \
                or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception as e:      # noqa: BLE001 - clocks are evidence, not a dependency of the measurement
            self.err = repr(e)
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        if self.nv is None:
            return
        try:
            while not self.stop:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.bits |= int(self.get_reasons(self.h))
                time.sleep(0.02)
        except Exception as e:      # noqa: BLE001
            self.err = repr(e)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=5)

    def summary(self):
        sm = sorted(self.sm)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted(n for b, n in names.items() if self.bits & b)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
               "samples": len(sm), "source": "nvml"}
        if self.err:
            out["error"] = self.err
        return out


# ----------------------------------------------------------------------------------------------------------------
# the workload as the HOST holds it (CPU tensors, seeded): identical for the GPU arm and the CPU reference arm
# ----------------------------------------------------------------------------------------------------------------
def host_scene(shape, n_class, seed=100):
    """Four target key frames, three reference views each (two neighbouring poses with their own images, and the frame
    itself), poses, and every weight of the model -- all generated here with torch on the CPU."""
    from dns_slam_b200 import synthetic as syn
    s = syn.SHAPES[shape]
    gen = torch.Generator().manual_seed(seed)
    poses = syn.trajectory(shape, 2 * N_TARGET + 2)
    H, W = s["H"], s["W"]
    frames, refer_img, refer_c2w, refer_idx = [], [], [], []
    for f in range(N_TARGET):
        fr = syn.frame(shape, poses[2 * f + 1], gen, n_class=n_class)
        frames.append(fr)
        refer_img.append(torch.stack((torch.rand(H, W, 3, generator=gen), torch.rand(H, W, 3, generator=gen), fr["color"]), 0))
        refer_idx.append([100 + 2 * f, 101 + 2 * f, -1])
        refer_c2w.append([poses[2 * f], poses[2 * f + 2], poses[2 * f + 1]])
    est = [poses[2 * f + 1].clone() for f in range(N_TARGET)]

    def uni(n, a):
        return (torch.rand(n, generator=gen) * 2 - 1) * a

    bound = syn.load_bound(s["bound"])
    weights = {"table_scale": 0.1, "coarse": uni(4096, (6.0 / (80 + 32)) ** 0.5), "color": uni(32 * 112 + 16 * 32, 0.2),
               "logit": uni(32 * 112 + ((n_class + 15) // 16 * 16) * 32, 0.2), "merge": uni(32 * 112 + 32 * 32, 0.2),
               "experts": uni(n_class * 4096, (6.0 / (80 + 32)) ** 0.5).view(n_class, 4096),
               "stem_conv": torch.randn(64, 3, 7, 7, generator=gen) * (2.0 / (7 * 7 * 64)) ** 0.5}
    weights["table_seed"] = seed + 1
    return dict(shape=shape, cam=syn.camera(shape), bound=bound, frames=frames, refer_img=refer_img, refer_c2w=refer_c2w,
                refer_idx=refer_idx, est=est, kf_idx=list(range(N_TARGET)), weights=weights, s=s)


def table_values(n, seed, scale):
    return (torch.rand(n, generator=torch.Generator().manual_seed(seed)) * 2 - 1) * scale


# ----------------------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the same iteration
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference(args, scene, n_rays, steps, warmup, seed=1):
    """mapping.py:884-910 through oracle/reference_path.py on the host cores: get_target_samples (incl. the 209 MB/view
    up-sample of common.py:646), renderer, 7 losses, backward, torch Adam over decoder + experts + poses.  Test
    infrastructure used here ONLY as the CPU baseline.  Returns (rays/s, s/step, steps timed)."""
    from oracle import reference_path as rp
    from dns_slam_b200 import slam, step as stepmod, synthetic as syn
    torch.set_num_threads(os.cpu_count())
    s, cam, bound, w = scene["s"], scene["cam"], scene["bound"], scene["weights"]
    C = args.n_class
    dec = rp.Decoder(syn.model_cfg(scene["shape"]), bound, n_class=C)
    with torch.no_grad():
        dec.pe_fn.grid_fn.params.copy_(table_values(dec.pe_fn.grid_fn.params.numel(), w["table_seed"], w["table_scale"]))
        dec.coarse_fn.decoder.params.copy_(w["coarse"])
        dec.out_fn.color_decoder.params.copy_(w["color"])
        dec.out_fn.logit_decoder.params.copy_(w["logit"])
        dec.merge.decoder.params.copy_(w["merge"])
    experts = {}
    for c in range(C):
        e = rp.new_expert(seed=c)
        with torch.no_grad():
            e.params.copy_(w["experts"][c])
        experts[c] = e
    # once per mapping call: the stem over the reference views of every target frame (mapping.py:846)
    with torch.no_grad():
        feats = [rp.stem_forward(img[None], w["stem_conv"], torch.ones(64), torch.zeros(64))[0] for img in scene["refer_img"]]
    quad = [slam.quad_from_matrix(c[:3, :3]).requires_grad_(f != 0) for f, c in enumerate(scene["est"])]
    T = [c[:3, 3].clone().requires_grad_(f != 0) for f, c in enumerate(scene["est"])]
    opt = torch.optim.Adam([{"params": list(dec.parameters()) + [e.params for e in experts.values()], "lr": s["lr"]},
                            {"params": quad[1:], "lr": s["BA_cam_lr"]}, {"params": T[1:], "lr": s["BA_cam_lr"]}])
    tables = [slam.class_tables(f["label"]) for f in scene["frames"]]
    plan = stepmod.FrameBatchPlan(tables, n_rays, 15, (0, cam["H"], 0, cam["W"]), bound, s["smooth_pts"])
    gen = torch.Generator().manual_seed(seed)
    times = []
    for it in range(warmup + steps):
        _, tape = plan.make_host_draws(gen, pinned=False, return_tape=True)
        t0 = time.perf_counter()
        opt.zero_grad()
        tp = rp.DrawTape(tape)
        smp = rp.mapper_get_target_samples(cam, bound, dec, scene["frames"], quad, T, scene["refer_idx"], scene["kf_idx"],
                                           scene["refer_c2w"], feats, n_rays, 32, 15, tp)
        pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(dec, experts, bound, smp)
        p, d, l, lt, fs, op = rp.mapping_losses(smp, pc, pd, pl, fine, coarse, s["opacity_sigma"])
        sm = rp.smoothness(dec, bound, s["smooth_pts"], tp)
        loss = s["lambda_color"] * p + s["lambda_depth"] * d + s["lambda_label"] * l + 10 * lt \
            + s["lambda_fs"] * fs + s["lambda_opacity"] * op + s["lambda_smooth"] * sm
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return plan.n_total / sec, sec, len(times)


# ----------------------------------------------------------------------------------------------------------------
def build_decoder(args, scene, dev):
    """The product decoder carrying the scene's weights (all class experts active)."""
    from dns_slam_b200 import bench_util
    w = scene["weights"]
    dec = bench_util.make_decoder(scene["shape"], args.n_class, dev, seed=0, table_scale=1.0)
    with torch.no_grad():
        dec.view("table").copy_(table_values(dec.view("table").numel(), w["table_seed"], w["table_scale"]))
        for k in ("coarse", "color", "logit", "merge"):
            dec.view(k).copy_(w[k])
        dec.expert_params.copy_(w["experts"])
    return dec


def build_gpu_step(args, scene, dec, rank, world, comm, frames_dev, feats, tables):
    """MappingFrameStep over key frames that are on the device (fresh Adam state, as per optimize() call)."""
    from dns_slam_b200 import step as stepmod
    s = scene["s"]
    lam = dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=10.0, fs=s["lambda_fs"], op=s["lambda_opacity"])
    st = stepmod.MappingFrameStep(dec, scene["cam"], frames_dev, tables, feats, scene["est"], scene["refer_idx"],
                                  scene["refer_c2w"], scene["kf_idx"], args.rays_per_gpu * world, 32, 15, lr=s["lr"],
                                  BA_cam_lr=s["BA_cam_lr"], is_BA=True, lambdas=lam, opacity_sigma=s["opacity_sigma"],
                                  smooth_pts=s["smooth_pts"], lambda_sm=s["lambda_smooth"], with_tv=True, comm=comm,
                                  rank=rank, world=world)
    return st


def upload_scene(scene, host_pinned, dev, stem, n_class=None):
    """What happens once per mapping call: key frames and reference images host -> device, the stem over the reference
    views (mapping.py:846), the per-frame class tables (labels do not change between iterations)."""
    from dns_slam_b200 import slam
    frames_dev, feats, tables = [], [], []
    for f in range(N_TARGET):
        fr = {k: host_pinned["frames"][f][k].to(dev, non_blocking=True) for k in ("color", "depth", "label")}
        frames_dev.append(fr)
        feats.append(stem.forward_cl(host_pinned["refer_img"][f].to(dev, non_blocking=True)[None]))
        tables.append(slam.class_tables(fr["label"], n_ids=n_class))
    return frames_dev, feats, tables


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    S, C, R = args.samples, args.n_class, args.rays_per_gpu
    from dns_slam_b200 import synthetic as syn
    workload = (f"{args.shape}_mapping_iteration_{R}rays_per_gpu_x{S}samples_{C}classes_"
                f"hash2^{syn.SHAPES[args.shape]['hash_size']}_4keyframes_3views_sampling+featurematching+render+7losses+backward+adam")
    config = {"workload": workload, "rays_per_gpu": R, "n_samples": S, "n_class": C,
              "boundary": "host holds key frames + poses + per-iteration draws; everything else on the device",
              "cache": "key-frame feature maps (627 MB), per-step latents and activation stash exceed the 126 MB L2",
              "parallelism": f"rays sharded x{args.gpus} (key frames and parameters replicated), max-depth / counts / label "
                             f"exchange + ONE flat-gradient all-reduce (NCCL)" if args.gpus > 1 else "single GPU"}
    if S != 47:
        raise SystemExit("the benchmarked iteration uses the reference's 32 + 15 samples per ray")
    scene = host_scene(args.shape, C)

    if args.impl == "reference":
        if rank != 0:
            return
        n_cpu = min(args.cpu_rays, R)
        # bounded sample: one probe step sizes the rays per step so that K steps end within a few minutes
        _, probe, _ = cpu_reference(args, scene, n_cpu, 1, 0)
        while probe * (args.steps + 1) > 240.0 and n_cpu > 2048:
            n_cpu //= 2
            _, probe, _ = cpu_reference(args, scene, n_cpu, 1, 0)
        rps, sec, k = cpu_reference(args, scene, n_cpu, max(1, args.steps), 1)
        line = {"impl": "reference", "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"{n_cpu} of {R} rays per step; the whole iteration (sampling, feature matching "
                                           f"incl. up-sample, render, 7 losses, backward, Adam), {k} timed steps"},
                "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU: dns_slam_b200 has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        from dns_slam_b200 import step as stepmod
        comm = stepmod.TorchComm(torch.distributed.group.WORLD)
    from dns_slam_b200 import _lib, encoder

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # host side of the boundary: pinned key frames / reference images, the stem's weights
    host_pinned = {"frames": [{k: fr[k].contiguous().pin_memory() for k in ("color", "depth", "label")} for fr in scene["frames"]],
                   "refer_img": [x.contiguous().pin_memory() for x in scene["refer_img"]]}
    frame_bytes = sum(v.numel() * v.element_size() for fr in host_pinned["frames"] for v in fr.values()) \
        + sum(x.numel() * x.element_size() for x in host_pinned["refer_img"])
    stem = encoder.ResNet().to(dev)
    with torch.no_grad():
        stem.conv_blocks.conv1.weight.copy_(scene["weights"]["stem_conv"])

    # ---------------- device-resident timing
    frames_dev, feats, tables = upload_scene(scene, host_pinned, dev, stem, args.n_class)
    dec = build_decoder(args, scene, dev)
    st = build_gpu_step(args, scene, dec, rank, world, comm, frames_dev, feats, tables)
    gen = torch.Generator().manual_seed(1000 + rank)           # the rank's own pixel draws
    shared_gen = torch.Generator().manual_seed(999) if world > 1 else None   # batch-level draws: same on every rank
    n_bufs = max(args.steps, args.warmup)
    host_draws = [st.make_host_draws(gen, shared_gen=shared_gen) for _ in range(min(n_bufs, 8))]
    dev_draws = [h.to(dev) for h in host_draws]
    for i in range(args.warmup):
        st.step(dev_draws[i % len(dev_draws)])
    res = st.step(dev_draws[0]).cpu()
    st.check(res)                     # error flags / rays outside the bound would void the static-shape step
    barrier()
    _lib.profile_read(reset=True)
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            out = st.step(dev_draws[i % len(dev_draws)])
        e1.record()
        barrier()
    _lib.profile_enable(False)
    phase_ms, launches = _lib.profile_read(reset=True)
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
    step_ms = float(t_ms) / args.steps
    value = R * world / (step_ms * 1e-3)
    losses = out.cpu().tolist()
    band_rows = int(st.fm_ws[:4].view(torch.int32)[0])

    # ---------------- end to end: scene from HOST memory inside the timed region, draws from pinned memory every step,
    # the loss vector read back every step
    del st, dec, frames_dev, feats, tables         # (the caching allocator keeps the blocks: steady state of a SLAM run)
    # one untimed rehearsal of the whole call (upload, stem, class tables, step construction, two steps): the first call after
    # a free re-shapes the caching allocator's blocks (20-60 ms, once); a SLAM run is in that steady state from its second
    # mapping call on
    dec = build_decoder(args, scene, dev)
    frames_dev, feats, tables = upload_scene(scene, host_pinned, dev, stem, args.n_class)
    st = build_gpu_step(args, scene, dec, rank, world, comm, frames_dev, feats, tables)
    for i in range(2):
        st.upload(host_draws[i % len(host_draws)])
        st.step()
        st.read_result()
    barrier()
    del st, dec, frames_dev, feats, tables
    dec = build_decoder(args, scene, dev)          # weights are device state like in the reference (not per-call input)
    barrier()
    e0.record()
    frames_dev, feats, tables = upload_scene(scene, host_pinned, dev, stem, args.n_class)
    st = build_gpu_step(args, scene, dec, rank, world, comm, frames_dev, feats, tables)
    for i in range(args.steps):
        st.upload(host_draws[i % len(host_draws)])
        st.step()
        st.read_result()
    e1.record()
    barrier()
    st.check()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
    e2e_val = R * world / (float(t2) / args.steps * 1e-3)
    h2d = st.draw_bytes + frame_bytes / args.steps
    d2h = st.result_host.numel() * 4

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (CUDA events inside the timed region)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern = {k: phase_ms.get(k, 0.0) / args.steps for k in ("point_fwd", "ray", "point_bwd", "dw_gemm", "feature", "sample",
                                                           "tv_fwd", "tv_bwd", "adam", "class_prep", "prep", "finalize")}
    dom = max(("point_fwd", "ray", "point_bwd"), key=lambda k: kern[k])
    alg = R * (S * BYTES_PT[dom] + BYTES_RAY[dom])
    achieved = alg / (kern[dom] * 1e-3) / 1e9
    tv_pts = (scene["s"]["smooth_pts"] - 1) ** 3
    step_alg = R * (S * 3332 + 252) + tv_pts * TV_BYTES_PT + band_rows * N_REFER * FEAT_BYTES_ROW
    # DRAM traffic of the dominant kernel: bytes per sample point from the round's `ncu --set full` capture of this build
    # (profiles/ncu_traffic.json, written by scratch/ncu_summary.py from profiles/run_ncu.sh), scaled to this launch; a
    # run cannot measure it itself (numbers taken under a profiler are not bench values)
    traffic, traffic_src = None, "no ncu capture on file (profiles/ncu_traffic.json missing)"
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tj):
        tr = json.load(open(tj))
        kname = {"point_fwd": "k_point_fwd_tc2<1>", "ray": "k_ray_tc2", "point_bwd": "k_point_bwd_tc2<1>"}[dom]
        bpp = tr.get("bytes_per_point", {}).get(kname)
        if bpp is not None:
            traffic = bpp * R * S
            traffic_src = (f"profiles/{tr.get('capture')}_ncu_summary.md: {bpp:.0f} B per sample point (dram__bytes_read.sum + "
                           f"dram__bytes_write.sum of one `ncu --set full` launch at {tr.get('points_per_launch')} points) x "
                           f"{R * S} points of this launch")
    roof = {"bound": "hbm", "kernel": {"point_fwd": "k_point_fwd_tc2", "ray": "k_ray_tc2", "point_bwd": "k_point_bwd_tc2"}[dom],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_launch": alg, "kernel_ms": kern[dom],
            "step": {"algorithmic_bytes": step_alg, "achieved": step_alg / (step_ms * 1e-3) / 1e9,
                     "frac": step_alg / (step_ms * 1e-3) / 1e9 / peak,
                     "terms": {"rays": R * (S * 3332 + 252), "tv_lattice": tv_pts * TV_BYTES_PT,
                               "feature_rows": band_rows * N_REFER * FEAT_BYTES_ROW}},
            "phase_ms_per_step": kern}

    extra = {}
    if not args.no_extra and world == 1:
        del st, dec, frames_dev, feats, tables
        torch.cuda.empty_cache()
        extra = _extra_configs(args, dev, scene)

    cpu = None
    if world == 1 and not args.no_cpu:
        n_cpu = min(args.cpu_rays, R)
        rps, sec, k = cpu_reference(args, scene, n_cpu, 2, 1)
        cpu = {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{n_cpu} of {R} rays per step on the same key frames / poses / weights / draw generator (the whole "
                         f"iteration: sampling, feature matching incl. up-sample, render, 7 losses, backward, Adam), "
                         f"{k} timed steps, {sec * 1e3:.0f} ms each"}

    line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": f"key frames + reference images ({frame_bytes} B) uploaded, stem and class tables run once "
                            f"inside the timed region (the call was rehearsed once, untimed); {R * 8} B of draws per step"},
            "gpu_launches": int(sum(launches.values())), "launches_per_step": {k: v // args.steps for k, v in launches.items() if v},
            "roofline": roof, "cpu_baseline": cpu, "losses_last_step": losses, "band_samples_per_ray": band_rows / R,
            "extra": extra}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def _time_cuda(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _extra_configs(args, dev, scene):
    """BASELINE configs 0-3 and the component timings of earlier rounds (device-resident, CUDA events)."""
    from dns_slam_b200 import bench_util, fused, step as stepmod, synthetic as syn
    s = syn.SHAPES[args.shape]
    out = {}
    # the round-1 step: core only (features given), 131072 rays -- for continuity with BENCH_r01
    dec, smp = bench_util.synthetic_batch(args.shape, "map", args.rays_per_gpu, 47, args.n_class, dev, seed=100)
    ms = stepmod.MappingStep(dec, s["lr"], dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0), s["opacity_sigma"])
    t = _time_cuda(lambda: ms.step(smp), 10, 3)
    out["core_only_step_round1_definition"] = {"ms_per_step": t, "rays_per_s": args.rays_per_gpu / (t * 1e-3)}
    del dec, smp, ms
    torch.cuda.empty_cache()
    # config 2: mapping 4096 rays x 47, semantic head, hash-grid + MLP Adam step
    dec, smp = bench_util.synthetic_batch(args.shape, "map", 4096, 47, args.n_class, dev, seed=7)
    ms = stepmod.MappingStep(dec, s["lr"], dict(p=5.0, d=5.0, l=0.1, lt=10.0, fs=10.0, op=10.0), s["opacity_sigma"])
    t = _time_cuda(lambda: ms.step(smp), 20, 5)
    out["config2_mapping_4096x47_adam"] = {"ms_per_step": t, "rays_per_s": 4096 / (t * 1e-3)}
    g = torch.Generator().manual_seed(1)
    r3, r113 = torch.rand(3, generator=g), torch.rand(1, 1, 1, 3, generator=g)
    off, jit = fused.tv_offsets(dec.bound, s["smooth_pts"], r3, r113)
    gv = ms._views(ms.grad)

    def tv():
        fused.tv_raw(dec.pe_fn.grid_fn.gstruct, dec.bound, dec.view("table"), dec.view("coarse"), s["smooth_pts"],
                     off, jit, s["lambda_smooth"], gv["table"], gv["coarse"])
    t_tv = _time_cuda(tv, 20, 5)
    out["tv_smoothness_63^3"] = {"ms": t_tv, "lattice_points_per_s": 63 ** 3 / (t_tv * 1e-3)}
    # config 1: tracking 1024 rays x 96 samples, pose gradients only
    dec1, smp1 = bench_util.synthetic_batch(args.shape, "track", 1024, 96, args.n_class, dev, seed=9, dec=dec)
    ts = stepmod.TrackingStep(dec1, dict(p=5.0, d=5.0, l=0.1))
    t1 = _time_cuda(lambda: ts.forward_backward(smp1), 50, 10)
    out["config1_tracking_1024x96"] = {"ms_per_iteration": t1, "rays_per_s": 1024 / (t1 * 1e-3),
                                       "ms_per_10_iterations": 10 * t1}
    # the frames + draws step at the SLAM batch size (4 x 500 rays): whole iteration, eager launches and CUDA graph
    out["iteration_native_step_replica_2000"] = _native_iteration(args, dev, scene, 2000)
    # T_iter: whole iterations through the autograd drop-in host mirror
    for shape in ("replica", "scannet"):
        out["iteration_" + shape] = _iteration_timings(shape, args.n_class, dev)
    out["inference_replica"] = _inference_timings("replica", args.n_class, dev)
    out["stem_replica"] = _stem_timings("replica", dev)
    # config 0: one full mapping iteration at 1200x680 with the reference's 2000 rays on the host cores (oracle port)
    if not args.no_cpu:
        rps, sec, k = cpu_reference(args, scene, 2000, 2, 1)
        out["config0_cpu_full_iteration_1200x680_2000rays"] = {
            "ms_per_iteration": sec * 1e3, "rays_per_s": rps, "cores": os.cpu_count(), "timed_iterations": k,
            "gpu_ms_per_iteration": out["iteration_native_step_replica_2000"]["ms_per_iteration_cuda_graph"],
            "note": "oracle port of slams/mapping.py:884-910 incl. get_target_samples, smoothness, backward, Adam"}
    # config 3: ScanNet-shaped 200-frame tracking + mapping loop (examples/synthetic_slam.py)
    out["config3_scannet_200"] = _config3(dev)
    return out


def _native_iteration(args, dev, scene, n_rays):
    """MappingFrameStep at the reference's mapping batch size: ms per iteration, eager and as a CUDA-graph replay."""
    import copy
    from dns_slam_b200 import encoder
    a2 = copy.copy(args)
    a2.rays_per_gpu = n_rays
    stem = encoder.ResNet().to(dev)
    hp = {"frames": scene["frames"], "refer_img": scene["refer_img"]}
    frames_dev, feats, tables = upload_scene(scene, hp, dev, stem, args.n_class)
    dec = build_decoder(a2, scene, dev)
    st = build_gpu_step(a2, scene, dec, 0, 1, None, frames_dev, feats, tables)
    gen = torch.Generator().manual_seed(5)
    draws = [st.make_host_draws(gen).to(dev) for _ in range(4)]
    it = [0]

    def eager():
        st.step(draws[it[0] % 4])
        it[0] += 1
    t_eager = _time_cuda(eager, 30, 5)
    host = st.read_result()
    torch.cuda.synchronize()
    st.check(host)
    # CUDA graph: one captured iteration, replayed with fresh draws copied into the static buffer
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        st.step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        st.step()

    def replay():
        st.draws_dev.copy_(draws[it[0] % 4], non_blocking=True)
        graph.replay()
        it[0] += 1
    t_graph = _time_cuda(replay, 50, 5)
    return {"rays": st.n_total, "ms_per_iteration": t_eager, "ms_per_iteration_cuda_graph": t_graph,
            "rays_per_s_cuda_graph": st.n_total / (t_graph * 1e-3)}


def _config3(dev):
    """BASELINE config 3: 200 ScanNet-shaped frames, tracking every frame, mapping every 5th (scannet.yaml), through the
    example loop; wall seconds include the host-side synthetic data generation."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    try:
        import synthetic_slam
        t0 = time.perf_counter()
        log = synthetic_slam.run("scannet", 200, n_class=40, track_iters=30, map_iters=100, use_graph=True, verbose=False, map_every=5)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        mp_ = [x for x in log["log"] if x[0] == "map"]
        return {"wall_s": wall, "frames": 200, "timings": log["timings"],
                "mapping_calls_on_the_native_loop": sum(1 for x in mp_ if x[4]), "mapping_calls": len(mp_),
                "schedule": "scannet.yaml: 30 tracking iterations per frame, 100 mapping iterations every 5th frame"}
    except Exception as e:      # noqa: BLE001 - a failing extra must not void the headline line
        return {"error": repr(e)}


def _stem_timings(shape, dev):
    """ResNet stem (models/encoder.py:4-17): once per tracked frame (1 view) and once per mapping call
    (the refer views of the window; 3 here)."""
    from dns_slam_b200 import encoder, synthetic as syn
    s = syn.SHAPES[shape]
    H, W = s["H"], s["W"]
    h, w = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    torch.manual_seed(0)
    enc = encoder.ResNet().to(dev)
    res = {"frame": [H, W]}
    for n in (1, 3):
        x = torch.rand(1, n, H, W, 3, device=dev)
        t = _time_cuda(lambda: enc.forward_cl(x), 20, 3)
        res[f"ms_{n}_view" + ("s" if n > 1 else "")] = t
        res[f"gbs_{n}"] = n * (H * W * 3 + 3 * h * w * 64) * 4 / (t * 1e-3) / 1e9
    return res


def _inference_timings(shape, n_class, dev):
    """Inference path: one full-frame render (every pixel x 47 samples, frame_vis of
    slams/mapping.py:636-690) and free-point queries at the mesh-extraction batch size (meshing.py:640-655)."""
    from dns_slam_b200 import bench_util, inference
    dec = bench_util.make_decoder(shape, n_class, dev, seed=1)
    sc = bench_util.slam_scene(shape, n_class, dev, seed=2, n_target=1)
    cam = sc["cam"]
    g = torch.Generator().manual_seed(4)
    ts, tz = torch.rand(15, generator=g), torch.rand(15, generator=g)
    refer_w2c = torch.inverse(sc["poses"][0])

    def frame():
        inference.render_frame(cam, dec, sc["frames"][0], sc["poses"][1], refer_w2c, sc["feats"][0][:1].contiguous(),
                               32, 15, ts, tz, n_pts_batch=131072)
    t_frame = _time_cuda(frame, 3, 1)
    P = 1 << 21
    lo, hi = dec.bound[:, 0].float(), dec.bound[:, 1].float()
    pts = lo + (hi - lo) * torch.rand(P, 3, generator=g).to(dev)
    pix = (torch.randn(P, 32, generator=g) * 0.3).to(dev)
    lab = torch.randint(0, n_class, (P,), generator=g).to(dev)
    t_q = _time_cuda(lambda: inference.eval_points(dec, pts, pix, lab, "fine"), 5, 2)
    return {"full_frame_render_ms": t_frame, "frame_pixels": cam["H"] * cam["W"], "n_samples": 47,
            "frame_rays_per_s": cam["H"] * cam["W"] / (t_frame * 1e-3),
            "eval_points_ms_per_2M": t_q, "eval_points_per_s": P / (t_q * 1e-3),
            "grid_256^3_query_s_est": (256 ** 3 / P) * t_q * 1e-3}


def _iteration_timings(shape, n_class, dev):
    """ms per tracking / mapping ITERATION at the reference's default sizes (replica.yaml / scannet.yaml):
    the loops of slams/tracking.py:313-340 and slams/mapping.py:881-910 through the drop-in host mirror."""
    from dns_slam_b200 import bench_util, slam, synthetic as syn
    s = syn.SHAPES[shape]
    dec = bench_util.make_decoder(shape, n_class, dev, seed=1)
    sc = bench_util.slam_scene(shape, n_class, dev, seed=2)
    cam = sc["cam"]
    trk = slam.TrackerCore(cam, dec, s["tracking_pixels"], 32, 15, s["lambda_color"], s["lambda_depth"], s["lambda_label"],
                           freeze_decoder=True)
    n_it = 10
    td = bench_util.tracking_draws(cam, s["tracking_pixels"], n_it)
    est = sc["poses"][3].clone()
    est[:3, 3] += 0.01
    refer_w2c = torch.inverse(sc["poses"][2])
    feats2 = sc["feats"][1][:2].contiguous()

    def track():
        slam.track_frame(trk, sc["frames"][1], refer_w2c, feats2, est, n_it, s["cam_lr"], lambda it: td[it])
    t_track = _time_cuda(track, 3, 1) / n_it

    def graph_ms(fn):   # device time of one CUDA-graph-replayed iteration: events around the replays, best of 3
        best = None
        for _ in range(3):
            slam.graph_timing = {}
            fn()
            ms = slam.graph_timing.get("replay_ms_per_iteration")
            slam.graph_timing = None
            if ms is not None:
                best = ms if best is None else min(best, ms)
        return best
    t_track_graph = graph_ms(lambda: slam.track_frame(trk, sc["frames"][1], refer_w2c, feats2, est, 103, s["cam_lr"],
                                                      lambda it: td[it % n_it], use_graph=True, native=False))
    n_nat = s["tracking_iters"]           # whole calls of the native loop (per-frame reset and the final host read included)
    t_track_native = _time_cuda(lambda: slam.track_frame(trk, sc["frames"][1], refer_w2c, feats2, est, n_nat, s["cam_lr"],
                                                         lambda it: td[it % n_it], native=True), 3, 1) / n_nat
    mp = slam.MapperCore(cam, dec, s["mapping_pixels"], 32, 15,
                         lambdas=dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=10.0,
                                      fs=s["lambda_fs"], op=s["lambda_opacity"]),
                         opacity_sigma=s["opacity_sigma"], smooth_pts=s["smooth_pts"], lambda_sm=s["lambda_smooth"])
    m_it = 5
    md, tv = bench_util.mapping_draws(sc, s["mapping_pixels"], m_it)
    target = dict(kf_idx=sc["kf_idx"], frames=sc["frames"], class_tables=sc["class_tables"])
    refer = dict(kf_idx=sc["refer_idx"], est_c2w=sc["refer_c2w"])
    est_list = [sc["poses"][2 * f + 1].clone() for f in range(len(sc["frames"]))]

    def mapit():
        slam.map_optimize(mp, target, refer, sc["feats"], est_list, m_it, s["lr"], s["BA_cam_lr"], True, [],
                          lambda it: md[it], lambda it: tv[it])
    t_map = _time_cuda(mapit, 2, 1) / m_it
    mdg, tvg = bench_util.mapping_draws(sc, s["mapping_pixels"], 8)
    t_map_graph = graph_ms(lambda: slam.map_optimize(mp, target, refer, sc["feats"], est_list, 43, s["lr"], s["BA_cam_lr"], True, [],
                                                     lambda it: mdg[it % 8], lambda it: tvg[it % 8], use_graph=True, native=False))
    graph_ok = bool(getattr(mp, "last_graph_ok", False))
    m_nat = s["mapping_iters"]            # whole calls of the native loop (step construction and the final host read included)
    t_map_native = _time_cuda(lambda: slam.map_optimize(mp, target, refer, sc["feats"], est_list, m_nat, s["lr"], s["BA_cam_lr"],
                                                        True, [], lambda it: mdg[it % 8], lambda it: tvg[it % 8], native=True),
                              2, 1) / m_nat
    native_ok = mp.last_path == "native"
    return {"tracking_ms_per_iteration": t_track, "tracking_ms_per_iteration_cuda_graph": t_track_graph,
            "tracking_ms_per_iteration_native": t_track_native, "tracking_rays": s["tracking_pixels"],
            "mapping_ms_per_iteration": t_map, "mapping_ms_per_iteration_cuda_graph": t_map_graph,
            "mapping_ms_per_iteration_native": t_map_native, "mapping_graph_ok": graph_ok, "mapping_native_ok": native_ok,
            "mapping_rays": s["mapping_pixels"], "tv_lattice": (s["smooth_pts"] - 1) ** 3, "n_samples": 47,
            "note": "slam.track_frame / slam.map_optimize incl. sampling + feature matching + TV + Adam: the autograd drop-in "
                    "loop eager and as CUDA-graph replays (events around the replays), and the native loops "
                    "(step.TrackingFrameStep / step.MappingFrameStep, the default fast path; whole calls / iterations)"}


if __name__ == "__main__":
    main()
