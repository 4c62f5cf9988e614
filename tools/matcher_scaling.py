"""How the assignment kernel scales with the number of targets (one problem = one CTA): Gaussian costs, 900 predictions.
Prints a markdown table: device time per call (CUDA events, 28 problems in one launch) vs scipy on the host for ONE problem.
Not a bench line; bench.py --workload matcher is (the oracle-side timing there lives in bench.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-groundingdino_b200")): sys.path.insert(0, p)
import numpy as np, torch
from matcher_b200 import hungarian_assign

def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    try:
        from scipy.optimize import linear_sum_assignment
    except Exception:
        linear_sum_assignment = None
    print("| targets per problem | device ms, 28 problems in one launch | scipy ms, ONE problem on the host |")
    print("|---|---|---|")
    for T in (1, 10, 30, 100, 300, 900):
        blocks = [torch.from_numpy(rng.standard_normal((900, T)).astype(np.float32)).to(dev) for _ in range(28)]
        for _ in range(3):
            hungarian_assign(blocks)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            pairs, status = hungarian_assign(blocks)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        host = float("nan")
        if linear_sum_assignment is not None:
            c = blocks[0].cpu().numpy()
            t0 = time.perf_counter()
            for _ in range(3):
                i, j = linear_sum_assignment(c)
            host = (time.perf_counter() - t0) / 3 * 1e3
            assert np.array_equal(pairs[0][0].cpu().numpy(), i) and np.array_equal(pairs[0][1].cpu().numpy(), j)
        print(f"| {T} | {ms:.3f} | {host:.3f} |")

if __name__ == "__main__":
    main()
