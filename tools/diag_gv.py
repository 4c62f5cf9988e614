import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-groundingdino_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
from helpers import make_inputs
from oracle import oracle
from msda_b200 import functions as F

def run(g, variant):
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g[k])).cuda()
    gv, gl, ga = F.ms_deform_attn_backward(t("value"), t("shapes"), t("starts"), t("loc"), t("attn"), t("grad_out"), 64, variant=variant)
    return gv.cpu().numpy()

for name, shapes, N, M, Lq, dist in (("1lvl", [(9, 7)], 1, 1, 1, "U"), ("1lvl-q8", [(9, 7)], 1, 1, 8, "U"), ("2lvl", [(9, 7), (5, 4)], 1, 2, 3, "U"), ("4lvl", [(9, 7), (5, 4), (3, 2), (2, 1)], 2, 8, 5, "W")):
    g = make_inputs(shapes, N=N, M=M, D=32, Lq=Lq, P=4, dist=dist, seed=1)
    want = oracle.backward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], g["grad_out"], np.float64)[0]
    for variant in (1, 2):
        got = run(g, variant)
        err = np.abs(got - want)
        print(name, "variant", variant, "normwise", err.max() / np.abs(want).max())
        if variant == 2 and err.max() / np.abs(want).max() > 1e-4:
            bad = np.argwhere(err.max(-1) > 1e-4 * np.abs(want).max())
            print("  bad rows (b, s, m):", bad[:12].tolist(), "of", (np.abs(want).max(-1) > 0).sum(), "nonzero rows")
            rec = oracle.sample_index(g["shapes"], g["starts"], g["loc"], g["value"].shape[1])
            print("  records [h_low,w_low,flags]:", rec[0, 0, 0].reshape(-1, 4)[:, :3].tolist())
            b, s, m = bad[0]
            print("  row", (b, s, m), "got", got[b, s, m, :4], "want", want[b, s, m, :4], "ratio", got[b, s, m, :4] / want[b, s, m, :4])
