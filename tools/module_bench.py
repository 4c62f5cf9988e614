"""Module-level timing (row a1 of the scope table): MSDeformAttn.forward + backward at the cfg-2 encoder shape
(bs 4, 22,223 tokens, d_model 256, bf16 autocast) and the cfg-3 decoder shape (900 queries, bs 8), for
  fused    -- softmax/locations AND the input_padding_mask inside the kernels (fuse_pre_op, mask_in_kernel: scope rows f1, f4)
  fused+fill -- the same with the reference's value.masked_fill pass kept (mask_in_kernel=False)
  unfused  -- the reference's op-by-op graph around our op
  ref-op   -- the same graph around the reference's own CUDA op (fp32, with the casts its AMP branch does)
CUDA events, median of `iters` after warm-up; includes the four Linear layers (cuBLAS) of the module."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-groundingdino_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import torch
from msda_b200 import MSDeformAttn
from msda_b200 import functions as F
from oracle import ref_cuda

SHAPES = [(100, 167), (50, 84), (25, 42), (13, 21)]


class RefOpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, sh, st, loc, aw, step):
        ref = ref_cuda.load()
        ctx.save_for_backward(value, sh, st, loc, aw)
        return ref.ms_deform_attn_forward(value, sh, st, loc, aw, step)

    @staticmethod
    def backward(ctx, g):
        ref = ref_cuda.load()
        value, sh, st, loc, aw = ctx.saved_tensors
        gv, gl, ga = ref.ms_deform_attn_backward(value, sh, st, loc, aw, g.contiguous(), 64)
        return gv, None, None, gl, ga, None


def run(mode, N, Lq, ref_dim, iters=10):
    """A non-trivial input_padding_mask (10 % of the value rows) is passed in every mode."""
    torch.manual_seed(0)
    dev = "cuda"
    sh = torch.tensor(SHAPES, dtype=torch.long, device=dev)
    st = torch.cat((sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]))
    S = int(sh.prod(1).sum())
    Lq = Lq or S
    mod = MSDeformAttn(256, 4, 8, 4).to(dev)
    mod.fuse_pre_op = mode in ("fused", "fused+fill")
    mod.mask_in_kernel = mode != "fused+fill"
    pad = torch.rand(N, S, device=dev) < 0.1
    q = torch.randn(N, Lq, 256, device=dev)
    src = torch.randn(N, S, 256, device=dev, requires_grad=True)
    refp = torch.rand(N, Lq, 4, ref_dim, device=dev)
    if mode == "ref-op":
        orig = F.MSDeformAttnFunction
        import msda_b200.modules as mm

        class Shim:
            @staticmethod
            def apply(value, sh_, st_, loc, aw, step):   # the reference's AMP branch: widen, run, narrow
                return RefOpFunction.apply(value.float(), sh_, st_, loc.float(), aw.float(), step).to(value.dtype)
        mm.MSDeformAttnFunction = Shim
    ts = []
    for it in range(iters + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = mod(q, refp, src, sh, st, pad)
        out.float().square().mean().backward()
        b.record(); torch.cuda.synchronize()
        if it >= 3: ts.append(a.elapsed_time(b))
    if mode == "ref-op":
        mm.MSDeformAttnFunction = orig
    return statistics.median(ts)


if __name__ == "__main__":
    print("| shape (padding mask on 10 % of the rows) | fused, mask in kernel ms | fused + masked_fill ms | unfused ms | around reference CUDA op ms |")
    print("|---|---|---|---|---|")
    for name, N, Lq, rd in (("cfg-2 encoder self-attention, bs 4", 4, None, 2), ("cfg-3 decoder cross-attention, 900 queries, bs 8", 8, 900, 4)):
        r = [run(m, N, Lq, rd) for m in ("fused", "fused+fill", "unfused")]
        r.append(run("ref-op", N, Lq, rd) if ref_cuda.available() else float("nan"))
        print(f"| {name} | {r[0]:.2f} | {r[1]:.2f} | {r[2]:.2f} | {r[3]:.2f} |", flush=True)
