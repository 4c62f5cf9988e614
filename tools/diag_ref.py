"""Diagnostic: ours vs reference CUDA op vs fp64 oracle, per quantity, with the worst sample located."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-groundingdino_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
from helpers import SWIN_T_800x1333, make_inputs
from oracle import oracle, ref_cuda
from msda_b200 import functions as F

ref = ref_cuda.load()
for dist, N in (("U", 1), ("I", 1), ("W", 2)):
    g = make_inputs(SWIN_T_800x1333, N=N, M=8, D=32, Lq=None, P=4, dist=dist, seed=31 if N == 1 else 3)
    t = lambda k: torch.from_numpy(g[k]).cuda()
    value, shapes, starts, loc, attn, gout = (t(k) for k in ("value", "shapes", "starts", "loc", "attn", "grad_out"))
    ours = [F.ms_deform_attn_forward(value, shapes, starts, loc, attn, 64)] + F.ms_deform_attn_backward(value, shapes, starts, loc, attn, gout, 64)
    o = ref.ms_deform_attn_forward(value, shapes, starts, loc, attn, 64)
    theirs = [o] + ref.ms_deform_attn_backward(value, shapes, starts, loc, attn, gout.view(o.shape), 64)
    truth = [oracle.forward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], np.float64)] + list(
        oracle.backward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], g["grad_out"], np.float64))
    o32 = list(oracle.backward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], g["grad_out"], np.float32))
    names = ("out", "gv", "gl", "ga")
    for i, n in enumerate(names):
        a = ours[i].double().cpu().numpy().reshape(truth[i].shape); b = theirs[i].double().cpu().numpy().reshape(truth[i].shape)
        print(f"{dist} N={N} {n}: ours-truth {oracle.normwise(a, truth[i]):.2e}  ref-truth {oracle.normwise(b, truth[i]):.2e}  ours-ref {oracle.normwise(a, b):.2e}"
              + (f"  oracle32-ref {oracle.normwise(o32[i-1], b):.2e}" if i else ""))
        if n == "gl":
            d = np.abs(a - b); idx = np.unravel_index(d.argmax(), d.shape)
            print("   worst ours-ref at", idx, "ours", a[idx], "ref", b[idx], "truth", truth[i][idx], "loc", g["loc"][idx[:-1]], "max|gl|", np.abs(b).max())
            d = np.abs(a - truth[i]); idx = np.unravel_index(d.argmax(), d.shape)
            print("   worst ours-truth at", idx, "ours", a[idx], "ref", b[idx], "truth", truth[i][idx], "loc", g["loc"][idx[:-1]])
