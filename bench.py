#!/usr/bin/env python
"""Benchmark of the DNS-SLAM render-and-optimise hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the mapping core over one Replica-shaped ray batch already resident in
HBM: zero-grad -> fused encode + MLP + render + loss forward/backward (all six losses, hash-table,
MLP, class-expert, ray and pixel-feature gradients) -> [NCCL all-reduce of the flat gradient when
N > 1] -> fused Adam over the flat parameter buffer (``slams/mapping.py:888-910`` without the
sampling stage).  Each rank owns ``--rays-per-gpu`` rays (weak scaling: 131072 x 8 = the 1M-ray
mapping batch of BASELINE.json); ``value`` is whole-job rays/s.  ``e2e`` repeats the step with
the ray batch coming from pinned HOST memory every step and the loss dictionary read back.
``extra`` reports BASELINE configs 1 and 2 (tracking 1024 x 96, mapping 4096 x 47 + Adam) and the
full iteration with sampling / feature matching / TV.  ``--impl reference`` times the oracle port
of the reference's CPU path (the reference is Python on tinycudann, which is CUDA-only: the hash
grid / MLPs run as the fp32 PyTorch stand-in) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays/sec fused encode+MLP+render+loss fwd/bwd"
BYTES_PT = dict(point_fwd=1024 + 4, ray=256, point_bwd=2048)      # SURVEY 8d, per sample point
BYTES_RAY = dict(point_fwd=24, ray=228, point_bwd=0)              # per ray


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays-per-gpu", type=int, default=131072)
    ap.add_argument("--samples", type=int, default=47)
    ap.add_argument("--n-class", type=int, default=40)
    ap.add_argument("--shape", default="replica")
    ap.add_argument("--cpu-rays", type=int, default=2048)
    ap.add_argument("--no-extra", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML library calls
    (spawning nvidia-smi every 100 ms stalls kernel launches by ~20 %; the query set is the one of
    B200_PROFILING.md: clocks.sm, clocks.max.sm, hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown,
    sw_power_cap)."""

    def __init__(self, index):
        self.index, self.sm, self.max_sm, self.bits, self.stop, self.err = index, [], None, 0, False, None
        self.nv = self.h = self.get_reasons = None
        # NVML is initialised HERE, before the timed region: nvmlInit takes driver-wide locks for 100-200 ms, and
        # inside a short timed loop that showed up as an idle GPU (262 144 rays x 10 steps: 51.7 instead of 34.6 ms)
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
        # NVML clocks-event-reason bits
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted(n for b, n in names.items() if self.bits & b)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
               "samples": len(sm), "source": "nvml"}
        if self.err:
            out["error"] = self.err
        return out


def cpu_reference(args, dec_state, samples_cpu, steps, warmup):
    """The oracle port of mapping.py:888-910 (renderer + 6 losses + backward + torch Adam) on the
    host cores.  Test-infrastructure code used here ONLY as the CPU baseline."""
    from oracle import reference_path as rp
    from dns_slam_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count())
    bound = syn.load_bound(syn.SHAPES[args.shape]["bound"])
    dec = rp.Decoder(syn.model_cfg(args.shape), bound, n_class=args.n_class)
    with torch.no_grad():
        dec.pe_fn.grid_fn.params.copy_(dec_state["table"])
        dec.coarse_fn.decoder.params.copy_(dec_state["coarse"])
        dec.out_fn.color_decoder.params.copy_(dec_state["color"])
        dec.out_fn.logit_decoder.params.copy_(dec_state["logit"])
    experts = {}
    for c in range(args.n_class):
        e = rp.new_expert(seed=c)
        with torch.no_grad():
            e.params.copy_(dec_state["experts"][c])
        experts[c] = e
    s = syn.SHAPES[args.shape]
    params = list(dec.parameters()) + [e.params for e in experts.values()]
    opt = torch.optim.Adam(params, lr=s["lr"])
    smp = dict(samples_cpu)
    smp["pts"] = smp["rays_o"][:, None, :] + smp["rays_d"][:, None, :] * smp["z_vals"][:, :, None]
    n = smp["z_vals"].shape[0]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        pc, pd, pv, pl, fine, coarse = rp.mapper_renderer(dec, experts, bound, smp)
        p, d, l, lt, fs, op = rp.mapping_losses(smp, pc, pd, pl, fine, coarse, s["opacity_sigma"])
        loss = s["lambda_color"] * p + s["lambda_depth"] * d + s["lambda_label"] * l + 10 * lt \
            + s["lambda_fs"] * fs + s["lambda_opacity"] * op
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return n / sec, sec


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    S, C, R = args.samples, args.n_class, args.rays_per_gpu
    from dns_slam_b200 import synthetic as _syn
    workload = (f"{args.shape}_mapping_{R}rays_per_gpu_x{S}samples_{C}classes_"
                f"hash2^{_syn.SHAPES[args.shape]['hash_size']}_plus_adam")
    config = {"workload": workload, "rays_per_gpu": R, "n_samples": S, "n_class": C,
              "cache": "per-step inputs (~6.2 KB/ray) + activation stash exceed the 126 MB L2",
              "parallelism": f"rays sharded x{args.gpus}, flat-gradient all-reduce (NCCL)" if args.gpus > 1 else "single GPU"}

    from dns_slam_b200 import bench_util, synthetic as syn

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU path (oracle port) on a bounded sample of the same workload
        dev = torch.device("cpu")
        n_cpu = min(args.cpu_rays, R)
        dec_state, samples = _cpu_inputs(args, n_cpu)
        rps, sec = cpu_reference(args, dec_state, samples, max(1, args.steps), max(1, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"{n_cpu} of {R} rays per step, full step (render+6 losses+backward+Adam)"},
                "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU: dns_slam_b200 has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD
    from dns_slam_b200 import _lib, step as stepmod, fused

    s = syn.SHAPES[args.shape]
    dec = bench_util.make_decoder(args.shape, C, dev, seed=0)                       # identical on every rank
    _, samples = bench_util.synthetic_batch(args.shape, "map", R, S, C, dev, seed=100 + rank, dec=dec)
    lam = dict(p=s["lambda_color"], d=s["lambda_depth"], l=s["lambda_label"], lt=10.0, fs=s["lambda_fs"],
               op=s["lambda_opacity"])
    if world > 1:   # rays [rank*R, (rank+1)*R) of ONE global batch of R*world rays
        sh = stepmod.ShardedMappingStep(dec, s["lr"], stepmod.TorchComm(pg), rank, world, lambdas=lam,
                                        opacity_sigma=s["opacity_sigma"])

        class _Ms:          # same .step(samples) surface as MappingStep
            def step(self, smp):
                return sh.step_sharded(smp, R * world)
        ms = _Ms()
    else:
        ms = stepmod.MappingStep(dec, s["lr"], lam, s["opacity_sigma"])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    for _ in range(args.warmup):
        ms.step(samples)
    barrier()
    _lib.profile_read(reset=True)
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = ms.step(samples)
        e1.record()
        barrier()
    _lib.profile_enable(False)
    phase_ms, launches = _lib.profile_read(reset=True)
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
    step_ms = float(t_ms) / args.steps
    value = R * world / (step_ms * 1e-3)
    losses = out[0].cpu().tolist()

    # ---------------- end to end: inputs from pinned host memory every step, losses read back
    host = {k: v.detach().cpu().pin_memory() for k, v in samples.items() if k != "mask"}
    pipe = stepmod.HostBatchPipeline(host, dev)
    h2d = pipe.h2d_bytes
    pipe.run([host] * max(2, args.warmup // 2), ms.step)
    barrier()
    e0.record()
    pipe.run([host] * args.steps, ms.step)       # every step: H2D of its inputs (overlapped with the previous
    e1.record()                                  # step's kernels on a copy stream) + D2H of its loss vector
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
    e2e_val = R * world / (float(t2) / args.steps * 1e-3)

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
    kern = {k: phase_ms[k] / args.steps for k in ("point_fwd", "ray", "point_bwd", "dw_gemm", "adam", "class_prep", "prep")}
    dom = max(("point_fwd", "ray", "point_bwd"), key=lambda k: kern[k])
    alg = R * (S * BYTES_PT[dom] + BYTES_RAY[dom])
    achieved = alg / (kern[dom] * 1e-3) / 1e9
    traffic = None
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tj):
        bpp = json.load(open(tj)).get("bytes_per_point", {}).get(dom)   # ncu --set full, per sample point
        traffic = bpp * R * S if bpp is not None else None
    step_alg = R * (S * 3332 + 252)
    roof = {"bound": "hbm", "kernel": {"point_fwd": "k_point_fwd", "ray": "k_ray", "point_bwd": "k_point_bwd"}[dom],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_launch": alg, "kernel_ms": kern[dom],
            "step": {"algorithmic_bytes": step_alg, "achieved": step_alg / (step_ms * 1e-3) / 1e9,
                     "frac": step_alg / (step_ms * 1e-3) / 1e9 / peak},
            "phase_ms_per_step": kern}

    extra = {}
    if not args.no_extra:
        extra = _extra_configs(args, dev)

    cpu = None
    if world == 1:
        n_cpu = min(args.cpu_rays, R)
        dec_state = {"table": dec.view("table").cpu(), "coarse": dec.view("coarse").cpu(), "color": dec.view("color").cpu(),
                     "logit": dec.view("logit").cpu(), "experts": dec.expert_params.detach().cpu()}
        smp = {k: v[:n_cpu].detach().cpu() for k, v in samples.items()}
        rps, sec = cpu_reference(args, dec_state, smp, 3, 1)
        cpu = {"value": rps, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{n_cpu} of {R} rays per step (same step: render + 6 losses + backward + Adam), "
                         f"3 timed steps, {sec * 1e3:.0f} ms each"}

    line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32},
            "gpu_launches": int(sum(launches.values())), "launches_per_step": {k: v // args.steps for k, v in launches.items() if v},
            "roofline": roof, "cpu_baseline": cpu, "losses_last_step": losses, "extra": extra}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def _cpu_inputs(args, n_cpu):
    """Inputs of the reference arm, generated on the CPU only (no GPU needed for --impl reference)."""
    from oracle import reference_path as rp
    from dns_slam_b200 import synthetic as syn, bench_util
    S, C = args.samples, args.n_class
    bound = syn.load_bound(syn.SHAPES[args.shape]["bound"])
    odec = rp.Decoder(syn.model_cfg(args.shape), bound, n_class=C, seed=0)
    with torch.no_grad():
        odec.pe_fn.grid_fn.params.mul_(1000.0)
    dec_state = {"table": odec.pe_fn.grid_fn.params.detach(), "coarse": odec.coarse_fn.decoder.params.detach(),
                 "color": odec.out_fn.color_decoder.params.detach(), "logit": odec.out_fn.logit_decoder.params.detach(),
                 "experts": torch.stack([rp.new_expert(seed=100 + c).params.detach() for c in range(C)])}
    cam = syn.camera(args.shape)
    gen = torch.Generator().manual_seed(100)
    n_s, n_f = bench_util.split_samples(S)
    poses = syn.trajectory(args.shape, 8)
    parts = []
    per = n_cpu // 4
    for f in range(4):
        c2w = poses[2 * f + 1]
        fr = syn.frame(args.shape, c2w, gen, n_class=C)
        img = torch.cat((fr["color"], fr["depth"].unsqueeze(-1), fr["label"].unsqueeze(-1)), -1)
        tape = rp.DrawTape(seed=f)
        idx = rp.uniform_indices(0, cam["H"], 0, cam["W"], per, tape)
        i, j = rp.uv_from_flat(idx, 0, 0, cam["W"])
        ro, rd = rp.rays_from_uv(i, j, c2w[:3, :3], c2w[:3, 3], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        smp = rp.gather_window(img, 0, cam["H"], 0, cam["W"], idx)
        far, inside = rp.far_plane(ro, rd, bound, smp[:, 3])
        z = rp.sample_along_rays(smp[:, 3], n_s, n_f, far, tape)
        parts.append(dict(gt_color=smp[:, :3].float(), gt_depth=smp[:, 3].float(), gt_label=smp[:, 4].long(),
                          rays_o=ro.float(), rays_d=rd.float(), z_vals=z))
    cat = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
    cat["features"] = torch.randn(cat["z_vals"].shape[0], S, 32, generator=gen) * 0.3 * rp.trunc_mask(cat["z_vals"], cat["gt_depth"])[..., None]
    return dec_state, cat


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


def _extra_configs(args, dev):
    """BASELINE configs 1 and 2 on this GPU (device-resident, CUDA events)."""
    from dns_slam_b200 import bench_util, fused, step as stepmod, synthetic as syn
    s = syn.SHAPES[args.shape]
    out = {}
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
    # T_iter: whole iterations incl. sampling, feature matching + Merge, TV, Adam
    for shape in ("replica", "scannet"):
        out["iteration_" + shape] = _iteration_timings(shape, args.n_class, dev)
    out["inference_replica"] = _inference_timings("replica", args.n_class, dev)
    out["stem_replica"] = _stem_timings("replica", dev)
    return out


def _stem_timings(shape, dev):
    """ResNet stem (models/encoder.py:4-17): once per tracked frame (1 view) and once per mapping call
    (the refer views of the window; 3 here).  Bytes: the frames in, the 64-channel half-resolution map written by the
    convolution, read and written by the normalisation."""
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
    from dns_slam_b200 import bench_util, inference, synthetic as syn
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
    # device time of one CUDA-graph-replayed iteration: events around 100 replays (slam.graph_timing), best of 3
    def graph_ms(fn):
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
                                                      lambda it: td[it % n_it], use_graph=True))
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
                                                     lambda it: mdg[it % 8], lambda it: tvg[it % 8], use_graph=True))
    return {"tracking_ms_per_iteration": t_track, "tracking_ms_per_iteration_cuda_graph": t_track_graph,
            "tracking_rays": s["tracking_pixels"],
            "mapping_ms_per_iteration": t_map, "mapping_ms_per_iteration_cuda_graph": t_map_graph,
            "mapping_graph_ok": bool(getattr(mp, "last_graph_ok", False)), "mapping_rays": s["mapping_pixels"], "tv_lattice": (s["smooth_pts"] - 1) ** 3,
            "n_samples": 47, "note": "autograd drop-in path (render_and_loss + FusedAdam), sampling + feature matching + TV included"}


if __name__ == "__main__":
    main()
