import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import rel_err
import test_gpu_framestep as T
from oracle import cases
from dns_slam_b200 import synthetic as syn
gd = os.path.join(ROOT, "tests", "golden")
for n_rays in (48, 300):
    for with_tv in (True, False):
        st, g, meta, inp, s = T._build(gd, n_rays, with_tv=with_tv)
        buf, tape = st.make_host_draws(torch.Generator().manual_seed(4), return_tape=True)
        st.upload(buf)
        st.adam.step = lambda: None
        res = st.step().cpu()
        meta2 = dict(meta)
        if not with_tv:
            meta2["lambda_sm"] = 0.0
        syn.SHAPES["tiny"]["mapping_pixels"] = n_rays
        o = cases.run_mapping(meta2, g["quad"], g["T"], tape, inp)
        syn.SHAPES["tiny"]["mapping_pixels"] = 48
        gv = st._flat_views(st.grad)
        row = {k: rel_err(gv[k], o["grad"][k]) for k in ("table", "coarse", "color", "logit", "merge")}
        row["quad"] = rel_err(st.d_quats[1:], torch.stack(o["grad"]["quad"][1:]))
        row["T"] = rel_err(st.d_trans[1:], torch.stack(o["grad"]["T"][1:]))
        print(n_rays, with_tv, {k: f"{v:.1e}" for k, v in row.items()}, "sm", float(res[8]), float(o["loss"]["sm"]))
        t_g, t_o = gv["table"].cpu(), o["grad"]["table"]
        d = (t_g - t_o).abs()
        offs = list(st.dec.pe_fn.grid_fn.tables["offset"])
        print("   per level:", " ".join(f"{rel_err(t_g[2*offs[l]:2*offs[l+1]], t_o[2*offs[l]:2*offs[l+1]]):.0e}" for l in range(16)))
        i = int(d.argmax()); print("   worst entry", i, float(t_g[i]), float(t_o[i]), "nnz", int((t_g != 0).sum()), int((t_o != 0).sum()))
