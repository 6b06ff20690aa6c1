"""Helpers shared by the GPU parity tests: weight transfer oracle -> product and tape splitting."""
import torch

from dns_slam_b200 import decoder as pdec
from dns_slam_b200 import synthetic as syn


def close(a, b, rtol=1e-3, atol=1e-5, name=""):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol, msg=lambda m: f"{name}: {m}")


def rel_err(a, b):
    """Norm-wise relative error: the 1e-3 bar of BASELINE.json is checked against this for gradients
    (element-wise for tensors without cancellation)."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def product_decoder_from_oracle(shape, odec, oexperts=None, n_class=None, device="cuda"):
    """Builds the CUDA Decoder and copies the oracle's weights into it."""
    bound = syn.load_bound(syn.SHAPES[shape]["bound"])
    n_class = n_class or odec.n_class
    dec = pdec.Decoder(syn.model_cfg(shape), bound, n_class=n_class, device=device)
    with torch.no_grad():
        dec.pe_fn.grid_fn.params.copy_(odec.pe_fn.grid_fn.params)
        dec.coarse_fn.decoder.params.copy_(odec.coarse_fn.decoder.params)
        dec.out_fn.color_decoder.params.copy_(odec.out_fn.color_decoder.params)
        dec.out_fn.logit_decoder.params.copy_(odec.out_fn.logit_decoder.params)
        dec.merge.decoder.params.copy_(odec.merge.decoder.params)
        if oexperts:
            for c, net in oexperts.items():
                dec.activate_expert(c)
                dec.expert_params[c].copy_(net.params)
    return dec


def frame_to(fr, device):
    return {k: (v.to(device).contiguous() if isinstance(v, torch.Tensor) else v) for k, v in fr.items()}


def split_mapping_tape(tape, frames, n_pixels_frame):
    """Recorded draw order of one mapping iteration: per target frame
    randint(uniform) -> one randint per class with more than one pixel -> rand(15) x2; then the two
    TV draws."""
    pos, out = 0, []
    for fr in frames:
        lab = fr["label"].reshape(-1)
        _, counts = torch.unique(lab, return_counts=True)
        n_draws = int((counts != 1).sum())
        d = {"idx_uniform": tape[pos][1]}
        pos += 1
        d["class_draws"] = [tape[pos + k][1] for k in range(n_draws)]
        pos += n_draws
        d["t_surface"], d["t_zero"] = tape[pos][1], tape[pos + 1][1]
        pos += 2
        out.append(d)
    tv = (tape[pos][1], tape[pos + 1][1])
    assert pos + 2 == len(tape)
    return out, tv
