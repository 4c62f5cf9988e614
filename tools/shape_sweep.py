"""BASELINE.json configs[4]: MSDeformAttn shape sweep -- levels 1-5, points 4/8, heads 8/16 (d_model 256), batch
1-32, fp32 and bf16 value, atomic vs deterministic backward.  Kernel times with CUDA events (median of `iters`),
inputs of distribution U, rotating 3 buffer sets.  Prints a markdown table (committed as profiles/r01_shape_sweep.md)."""
import argparse, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-groundingdino_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import numpy as np, torch
from msda_b200 import functions as F

PYRAMID = [(100, 167), (50, 84), (25, 42), (13, 21), (7, 11)]
VARIANT = int(os.environ.get("MSDA_SWEEP_VARIANT", "0"), 0)      # include/msda_b200.h variant bits, for A/B runs

def case(L, P, M, N, dtype, mode, iters=5, sets=3):
    D = 256 // M
    shapes = PYRAMID[:L]
    S = sum(h * w for h, w in shapes)
    dev = torch.device("cuda")
    sh = torch.tensor(shapes, dtype=torch.long, device=dev)
    st = torch.cat((sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]))
    g = torch.Generator(device=dev).manual_seed(L * 100 + P)
    bufs = []
    for _ in range(sets):
        v = torch.randn(N, S, M, D, device=dev, generator=g).to(dtype)
        loc = torch.rand(N, S, M, L, P, 2, device=dev, generator=g)
        aw = torch.softmax(torch.randn(N, S, M, L * P, device=dev, generator=g), -1).view(N, S, M, L, P)
        go = torch.randn(N, S, M * D, device=dev, generator=g).to(dtype)
        bufs.append((v, loc, aw, go))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf, tb = [], []
    for it in range(iters + 1):
        for v, loc, aw, go in bufs:
            a, b, c = ev(), ev(), ev()
            a.record(); F.ms_deform_attn_forward(v, sh, st, loc, aw, 64, variant=VARIANT); b.record()
            F.ms_deform_attn_backward(v, sh, st, loc, aw, go, 64, mode=mode, variant=VARIANT); c.record()
            torch.cuda.synchronize()
            if it:
                tf.append(a.elapsed_time(b)); tb.append(b.elapsed_time(c))
    e_v = 2 if dtype == torch.bfloat16 else 4
    bf = N * (S * M * D * e_v + S * M * L * P * 12 + S * M * D * e_v)
    bb = N * (2 * S * M * D * e_v + S * M * D * e_v + 2 * S * M * L * P * 12)
    f, b = statistics.median(tf), statistics.median(tb)
    return f, b, bf / f / 1e6, bb / b / 1e6

if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--quick", action="store_true"); a = ap.parse_args()
    print("| L | P | heads x D | batch | value | backward | fwd ms | bwd ms (incl. zero/narrow) | fwd GB/s | bwd GB/s |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    cases = []
    for L in (1, 2, 3, 4, 5):
        cases += [(L, 4, 8, 4, torch.float32, 0), (L, 4, 8, 4, torch.bfloat16, 0)]
    for P in (8,):
        cases += [(4, P, 8, 4, torch.float32, 0), (4, P, 8, 4, torch.bfloat16, 0)]
    for M in (16,):
        cases += [(4, 4, M, 4, torch.float32, 0), (4, 4, M, 4, torch.bfloat16, 0)]
    for N in (1, 2, 4, 8, 16, 32):
        cases += [(4, 4, 8, N, torch.float32, 0), (4, 4, 8, N, torch.bfloat16, 0)]
    cases += [(4, 4, 8, 4, torch.float32, 1), (4, 4, 8, 4, torch.bfloat16, 1)]
    if a.quick: cases = cases[:6]
    for L, P, M, N, dt, mode in cases:
        f, b, gf, gb = case(L, P, M, N, dt, mode)
        print(f"| {L} | {P} | {M} x {256 // M} | {N} | {'bf16' if dt == torch.bfloat16 else 'fp32'} | {'deterministic' if mode else 'atomic'} | {f:.3f} | {b:.3f} | {gf:.0f} | {gb:.0f} |", flush=True)
        torch.cuda.empty_cache()
