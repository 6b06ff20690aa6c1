"""Soak: many steps at several batch sizes (mbarrier pipelines must never hang; losses must stay finite)."""
import sys, time, torch
sys.path.insert(0, '.')
from dns_slam_b200 import bench_util, step as stepmod
dev = torch.device("cuda:0")
t0 = time.time()
for R, S, C, steps in [(131072, 47, 40, 150), (2000, 47, 40, 400), (500, 47, 40, 400), (33333, 96, 12, 60), (777, 13, 3, 300)]:
    dec = bench_util.make_decoder("replica", C, dev, seed=R)
    _, smp = bench_util.synthetic_batch("replica", "map", R, S, C, dev, seed=R + 1, dec=dec)
    smp = {k: v for k, v in smp.items() if k != "mask"}
    ms = stepmod.MappingStep(dec, 5e-3)
    first = last = None
    for it in range(steps):
        out = ms.step(smp)
        if it % 50 == 0 or it == steps - 1:
            l = out[0].tolist()
            assert all(x == x for x in l[:7]), (R, S, C, it, l)
            first = first or l
            last = l
    torch.cuda.synchronize()
    print(f"R={R} S={S} C={C} steps={steps}: total-loss {first[6]:.4f} -> {last[6]:.4f}  ({time.time()-t0:.1f}s)", flush=True)
print("soak ok")

# ---- the native steps (frames + draws boundary): many steps with fresh draws, losses finite, no hang
import argparse
import bench
from dns_slam_b200 import encoder, slam, synthetic as syn
args = argparse.Namespace(shape="replica", n_class=40, rays_per_gpu=131072, samples=47, gpus=1, steps=20, warmup=5)
scene = bench.host_scene("replica", 40)
hp = {"frames": scene["frames"], "refer_img": scene["refer_img"]}
stem = encoder.ResNet().to(dev)
for R, steps in ((131072, 60), (2000, 300), (48, 300)):
    args.rays_per_gpu = R
    dec = bench.build_decoder(args, scene, dev)
    fd, feats, tables = bench.upload_scene(scene, hp, dev, stem, 40)
    st = bench.build_gpu_step(args, scene, dec, 0, 1, None, fd, feats, tables)
    gen = torch.Generator().manual_seed(R)
    draws = [st.make_host_draws(gen) for _ in range(8)]
    first = last = None
    for it in range(steps):
        st.upload(draws[it % 8]); out = st.step()
        if it % 50 == 0 or it == steps - 1:
            l = out.tolist()
            assert all(x == x for x in l[:9]), (R, it, l)
            first = first or l
            last = l
    st.read_result(); torch.cuda.synchronize(); st.check()
    print(f"native mapping step R={R} steps={steps}: total-loss {first[6]:.4f} -> {last[6]:.4f}  ({time.time()-t0:.1f}s)", flush=True)
    del st, dec, fd, feats, tables
sc = bench_util.slam_scene("replica", 40, dev, seed=2)
s = syn.SHAPES["replica"]
dec = bench_util.make_decoder("replica", 40, dev, seed=1)
trk = slam.TrackerCore(sc["cam"], dec, s["tracking_pixels"], 32, 15, s["lambda_color"], s["lambda_depth"], s["lambda_label"], freeze_decoder=True)
td = bench_util.tracking_draws(sc["cam"], s["tracking_pixels"], 50)
est = sc["poses"][3].clone(); est[:3, 3] += 0.01
for k in range(40):
    best, loss, hist = slam.track_frame(trk, sc["frames"][1], torch.inverse(sc["poses"][2]), sc["feats"][1][:2].contiguous(), est, 50, s["cam_lr"],
                                        lambda it: td[it], native=True)
    assert torch.isfinite(hist).all() and torch.isfinite(best).all()
print(f"native tracking loop: 40 frames x 50 iterations, best loss {float(loss):.4f}  ({time.time()-t0:.1f}s)")
print("soak ok")
