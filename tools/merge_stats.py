"""How many grad_value row-reductions survive different aggregation scopes (CPU, oracle records).
Counts CELLS (each cell = up to 4 neighbour rows) per (query, head) item for the Swin-T encoder shape."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")): sys.path.insert(0, p)
from helpers import SWIN_T_800x1333, make_inputs, encoder_reference_points
from oracle import oracle

def stats(name, g):
    shapes = g["shapes"]; S = g["dims"][1]
    rec = oracle.sample_index(g["shapes"], g["starts"], g["loc"], S)[0]          # (Lq, M, L, P, 4)
    Lq, M, L, P, _ = rec.shape
    inside = rec[..., 2] & 1
    key = (rec[..., 0].astype(np.int64) + 2) * 100000 + (rec[..., 1] + 2)         # cell id per sample
    key = np.where(inside > 0, key, -1)
    items = Lq * M
    s0 = (inside > 0).sum() / items
    same_prev = np.zeros_like(key, dtype=bool); same_prev[..., 1:] = (key[..., 1:] == key[..., :-1]) & (key[..., 1:] >= 0)
    s1 = ((inside > 0) & ~same_prev).sum() / items
    ks = np.sort(key, axis=-1); s2 = ((ks[..., 0] >= 0).sum() + ((ks[..., 1:] != ks[..., :-1]) & (ks[..., 1:] >= 0)).sum()) / items
    # tile scopes: queries of level 0 only (75 % of all), arranged as H0 x W0
    H0, W0 = int(shapes[0][0]), int(shapes[0][1])
    k0 = key[: H0 * W0].reshape(H0, W0, M, L, P)
    def scope(th, tw):
        tot = 0; n = 0
        for y in range(0, H0 - th + 1, th):
            for x in range(0, W0 - tw + 1, tw):
                blk = k0[y:y + th, x:x + tw].transpose(2, 3, 0, 1, 4).reshape(M, L, -1)     # (M, L, th*tw*P)
                for m in range(M):
                    for l in range(L):
                        u = np.unique(blk[m, l]); tot += (u >= 0).sum()
                n += th * tw * M
        return tot / n
    print(f"{name}: cells/item  none {s0:.1f} | consecutive-same-cell {s1:.1f} | distinct per (item,level) {s2:.1f} | "
          f"1x8 queries {scope(1, 8):.1f} | 2x4 {scope(2, 4):.1f} | 8x8 tile {scope(8, 8):.2f} | 16x16 {scope(16, 16):.2f}")

gI = make_inputs(SWIN_T_800x1333, N=1, M=8, D=1, Lq=None, P=4, dist="I", seed=1)
stats("default-init (I)", gI)
# "trained-like": default-init offsets plus N(0, sigma px) jitter per sample
for sigma in (0.5, 2.0):
    g = make_inputs(SWIN_T_800x1333, N=1, M=8, D=1, Lq=None, P=4, dist="I", seed=1)
    rng = np.random.default_rng(5)
    wh = np.asarray([(w, h) for h, w in SWIN_T_800x1333], dtype=np.float32)
    g["loc"] = (g["loc"] + rng.standard_normal(g["loc"].shape).astype(np.float32) * sigma / wh[None, None, None, :, None, :]).astype(np.float32)
    stats(f"I + N(0,{sigma}px) jitter", g)
gU = make_inputs(SWIN_T_800x1333, N=1, M=8, D=1, Lq=None, P=4, dist="U", seed=1)
stats("uniform (U)", gU)
