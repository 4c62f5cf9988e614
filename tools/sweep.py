"""Kernel-level sweep: forward / backward-kernel time per layer for (config, distribution, variant).
Uses bench.py's buffers (2-3 rotating layers so inputs never sit in L2).  Prints one line per case."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
from msda_b200 import _lib

def time_case(cfg_name, dist, variant, iters=10, nl=3):
    cfg = dict(bench.CONFIGS[cfg_name]); cfg["layers"] = nl
    dev = torch.device("cuda", 0)
    layers = [bench.OurLayer(torch, _lib, bench.make_layer_inputs(torch, cfg, dist, 7 + l, dev), variant) for l in range(nl)]
    for lay in layers:
        lay.forward(); lay.backward()
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    f0, f1 = ev(), ev()
    f0.record()
    for _ in range(iters):
        for lay in layers: lay.forward()
    f1.record()
    bw = [(ev(), ev()) for _ in range(iters * nl)]
    k = 0
    for _ in range(iters):
        for lay in layers:
            lay.backward(*bw[k]); k += 1
    torch.cuda.synchronize()
    fwd = f0.elapsed_time(f1) / (iters * nl)
    bwd = sum(a.elapsed_time(b) for a, b in bw) / len(bw)
    N, S, M, D, L, Lq, P = layers[0].i["dims"]
    e_v = 2 if cfg["dtype"] == "bf16" else 4
    bf, bb = bench.algorithmic_bytes(N, S, L, Lq, e_v)
    print(f"{cfg_name} dist={dist} variant={variant:#x}: fwd {fwd*1e3:8.1f} us ({bf/fwd/1e6:7.0f} GB/s)  bwd-kernel {bwd*1e3:8.1f} us ({bb/bwd/1e6:7.0f} GB/s)", flush=True)
    del layers
    torch.cuda.empty_cache()

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg2:I:0,cfg2:I:0x400,cfg2:I:0x800,cfg2:U:0,cfg2:U:0x800,cfg1:I:0,cfg1:U:0,cfg3:I:0,cfg4enc:I:0")
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    for c in a.cases.split(","):
        name, dist, var = c.split(":")
        time_case(name, dist, int(var, 0), iters=a.iters)
