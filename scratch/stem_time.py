"""ResNet stem (dns_stem_fwd) at the Replica / ScanNet frame sizes; the cuDNN route of the reference
(conv2d + batch_norm + relu, then the channels-last copy the gather needs) is timed beside it."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from dns_slam_b200 import encoder, _lib
dev = torch.device("cuda:0")
enc = encoder.ResNet().to(dev)
bn = enc.conv_blocks.bn1


def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n, H, W in ((1, 680, 1200), (3, 680, 1200), (3, 460, 620)):
    x = torch.rand(1, n, H, W, 3, device=dev)
    ours = timed(lambda: enc.forward_cl(x))

    def ref():
        with torch.no_grad():
            y = F.conv2d(x[0].permute(0, 3, 1, 2), enc.conv_blocks.conv1.weight, None, 2, 3)
            y = F.relu(F.batch_norm(y, None, None, bn.weight, bn.bias, True, 0.1, 1e-5))
            return y.permute(0, 2, 3, 1).contiguous()
    lib = timed(ref)
    h, w = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    gb = n * (H * W * 3 + 3 * h * w * 64) * 4 / 1e9
    print(f"{n} x {H}x{W}: dns_stem_fwd {ours:.3f} ms ({gb / ours * 1e3:.0f} GB/s of its own traffic, "
          f"{n*h*w*64*147*2/ours/1e9:.1f} TFLOP/s), torch/cuDNN route {lib:.3f} ms")
