#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own oracle function.

Run in the authoring container only (needs /root/reference, which is absent on the GPU box):

    python tests/golden/make_golden.py

It imports ``ms_deform_attn_core_pytorch`` from
/root/reference/models/GroundingDINO/ops/functions/ms_deform_attn_func.py:41-61 (the module
hard-imports the compiled extension at :18, so an empty stand-in named
``MultiScaleDeformableAttention`` is registered first; the oracle function itself never touches
it) and records, per case: the fp32 inputs, the fp64 outputs and autograd gradients ("truth"), and
the reference's own fp32 outputs/gradients.  Nothing of the reference's source is copied: only
numbers it produced.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_FUNC = "/root/reference/models/GroundingDINO/ops/functions/ms_deform_attn_func.py"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def load_reference_oracle():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("_ref_msda_func", REF_FUNC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def starts_of(shapes):
    areas = [h * w for h, w in shapes]
    return [0] + list(np.cumsum(areas)[:-1])


def run_case(ref, name, shapes, N, M, D, Lq, P, seed, loc_mode, value_scale=1.0, bf16_value=False):
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g) * value_scale
    if bf16_value:
        value = value.to(torch.bfloat16).float()
    if loc_mode == "unit":            # U[0,1) like OPS/test.py:34
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    elif loc_mode == "wide":          # crosses every border: U[-0.25, 1.25)
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.5 - 0.25
    elif loc_mode == "pixel":         # exact pixel centres / edges / corners of each level
        loc = torch.empty(N, Lq, M, L, P, 2)
        for l, (h, w) in enumerate(shapes):
            ix = torch.randint(-1, 2 * w + 2, (N, Lq, M, P), generator=g).float() * 0.5   # half-pixel lattice
            iy = torch.randint(-1, 2 * h + 2, (N, Lq, M, P), generator=g).float() * 0.5
            loc[:, :, :, l, :, 0] = ix / w
            loc[:, :, :, l, :, 1] = iy / h
    else:
        raise ValueError(loc_mode)
    attn = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    grad_out = torch.randn(N, Lq, M * D, generator=g)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)

    rec = dict(shapes=np.asarray(shapes, dtype=np.int64), starts=np.asarray(starts_of(shapes), dtype=np.int64),
               value=value.numpy(), loc=loc.numpy(), attn=attn.numpy(), grad_out=grad_out.numpy())
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        v = value.to(dt).requires_grad_(True)
        s = loc.to(dt).requires_grad_(True)
        a = attn.to(dt).requires_grad_(True)
        out = ref(v, shapes_t, s, a)
        out.backward(grad_out.to(dt))
        rec[f"out_{tag}"] = out.detach().numpy()
        rec[f"gv_{tag}"] = v.grad.numpy()
        rec[f"gl_{tag}"] = s.grad.numpy()
        rec[f"ga_{tag}"] = a.grad.numpy()
    np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), **rec)
    print(f"{name}: S={S} out|max|={np.abs(rec['out_f64']).max():.3e}")


def main():
    ref = load_reference_oracle()
    # (1) the reference test's own fixture (OPS/test.py:21-28): N,M=1,2  Lq,L,P=2,2,2  shapes (6,4),(3,2),
    #     for every channel count its gradcheck loop visits (:85) plus its forward D=2.
    for D in (2, 30, 32, 64, 71):
        run_case(ref, f"testpy_D{D}", [(6, 4), (3, 2)], N=1, M=2, D=D, Lq=2, P=2, seed=3,
                 loc_mode="unit", value_scale=0.01)
    # (2) a small 4-level pyramid at the model's head geometry, samples crossing all borders
    run_case(ref, "pyr4_wide_D32", [(12, 20), (6, 10), (3, 5), (2, 3)], N=2, M=2, D=32, Lq=37, P=4,
             seed=11, loc_mode="wide")
    # (3) 5 levels, 8 points, 16 heads of 16 channels (config 5 of BASELINE.json)
    run_case(ref, "pyr5_P8_D16", [(9, 14), (5, 7), (3, 4), (2, 2), (1, 1)], N=1, M=4, D=16, Lq=19, P=8,
             seed=12, loc_mode="wide")
    # (4) exact pixel-lattice coordinates (the default-init hazard of SURVEY.md section 7.3)
    run_case(ref, "pyr3_pixel_D32", [(8, 8), (4, 4), (2, 2)], N=1, M=2, D=32, Lq=23, P=4,
             seed=13, loc_mode="pixel")
    # (5) bf16-representable value, for the bf16-value / fp32-coordinates contract
    run_case(ref, "pyr4_bf16_D32", [(12, 20), (6, 10), (3, 5), (2, 3)], N=1, M=2, D=32, Lq=29, P=4,
             seed=14, loc_mode="unit", bf16_value=True)
    # (6) one level, one point, odd channel count
    run_case(ref, "lvl1_P1_D7", [(5, 3)], N=3, M=3, D=7, Lq=4, P=1, seed=15, loc_mode="wide")


if __name__ == "__main__":
    main()
